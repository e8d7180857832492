"""Train-mode forward + backward of one ``TransformerFusion`` call as a SEQUENCE of primitive ops (BASELINE config 5;
reference ``train.py:96-135`` differentiates ``fusion.py:52-188`` with autograd).

This file only orders primitives - it holds no arithmetic of its own.  ``ops`` is the object that runs them:
``cfpnet_b200.train.CudaOps`` in the product (every method one or two libcfp kernels, fp32 as the reference trains).
The sequencing is written once against that interface so that it can be exercised without a GPU by handing it an
object with the same methods (the tests do that with plain torch ops and hold the result to the reference's own
``.train()`` gradients, tests/test_train_seq.py); the product never does.

Layout: every activation is a token-major fp32 matrix ``[rows, C]``.  Regrouping (LSA windows, hist2image zone canvas,
DAPM inside / outside sets, conv taps, the strided sr conv) is ``gather_rows`` / ``scatter_add_rows`` over host-built
int32 index vectors (``-1`` = a zero row: padding cells, cells outside the image) - the masks of ``fusion.py:103-120``
are never materialised here either.  Convolutions are tap sums of gathered rows times ``[Cout, Cin]`` slices, so their
input and weight gradients are the same two primitives transposed.

Formulas: the closed-form backward of DESIGN.md section 8 (linear attention in the two-phase state form, LayerNorm,
train-mode BatchNorm, the layer compositions), stage by stage.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

ATTN_EPS, LN_EPS = 1e-6, 1e-5


def sub(P: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in P.items() if k.startswith(prefix)}


def _pre(prefix, g):
    return {prefix + k: v for k, v in g.items()}


# ---------------------------------------------------------------------------------------------- index vectors (host ints)
class Indexer:
    """int32 index vectors of one (B, H, W, geometry, ws) configuration, built once on the host and kept on the device."""

    def __init__(self, B, H, W, g, ws, device):
        self.B, self.H, self.W, self.N = B, H, W, H * W
        dev = device
        base = (torch.arange(B, dtype=torch.int64) * self.N).view(B, 1)
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        tok = (yy * W + xx).reshape(-1)

        def finish(t):                                   # [B, rows] token index (or -1) -> flat int32 on the device
            t = torch.where(t >= 0, t + base, torch.full_like(t, -1))
            return t.reshape(-1).to(torch.int32).to(dev)

        # 3x3 conv taps (zero padding 1): tap (dy, dx) of output pixel p reads pixel p + (dy-1, dx-1)
        self.conv_tap = []
        for dy in range(3):
            for dx in range(3):
                y, x = yy + dy - 1, xx + dx - 1
                ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
                self.conv_tap.append(finish(torch.where(ok, y * W + x, torch.full_like(y, -1)).reshape(1, -1).repeat(B, 1)))
        # the nine taps of every output pixel side by side ([rows][9], row-major): one gather builds the im2col matrix
        self.conv_col = torch.stack(self.conv_tap, dim=1).reshape(-1).contiguous()
        # LSA windows over the zero-padded map (transformer.py:94-104)
        self.ws = ws
        if ws:
            nwy, nwx = -(-H // ws), -(-W // ws)
            wy, wx, iy, ix = torch.meshgrid(torch.arange(nwy), torch.arange(nwx), torch.arange(ws), torch.arange(ws), indexing="ij")
            y, x = wy * ws + iy, wx * ws + ix
            ok = (y < H) & (x < W)
            win = torch.where(ok, y * W + x, torch.full_like(y, -1)).reshape(1, -1)
            self.nwin = nwy * nwx
            self.win = finish(win.repeat(B, 1))
            inv = torch.full((self.N,), -1, dtype=torch.int64)
            flat = win.reshape(-1)
            inv[flat[flat >= 0]] = torch.arange(flat.numel())[flat >= 0]
            rows_per_frame = self.nwin * ws * ws
            self.win_inv = (inv.view(1, -1) + (torch.arange(B, dtype=torch.int64) * rows_per_frame).view(B, 1)).reshape(-1).to(torch.int32).to(dev)
            # GSA sr conv (kernel = stride = ws, no padding, transformer.py:144)
            nsy, nsx = H // ws, W // ws
            self.Ns = nsy * nsx
            sy_, sx_ = torch.meshgrid(torch.arange(nsy), torch.arange(nsx), indexing="ij")
            self.sr_tap = [finish(((sy_ * ws + dy) * W + sx_ * ws + dx).reshape(1, -1).repeat(B, 1)) for dy in range(ws) for dx in range(ws)]
            self.sr_col = torch.stack(self.sr_tap, dim=1).reshape(-1).contiguous()          # [B*Ns][ws*ws]
        # zone geometry (fusion.py:67-84,104; no resize branch in training)
        self.g = g
        if g is not None:
            zn, p1, p2 = g.zone_num, g.p1, g.p2
            zy, zx, py, px = torch.meshgrid(torch.arange(zn), torch.arange(zn), torch.arange(p1), torch.arange(p2), indexing="ij")
            y, x = g.sy_wo + zy * p1 + py, g.sx_wo + zx * p2 + px
            ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
            self.zone = finish(torch.where(ok, y * W + x, torch.full_like(y, -1)).reshape(1, -1).repeat(B, 1))
            inside = torch.zeros(H, W, dtype=torch.bool)
            inside[g.ry0:g.ry1, g.rx0:g.rx1] = True
            inside = inside.reshape(-1)
            self.Ni, self.No = int(inside.sum()), int((~inside).sum())
            self.inside = finish(tok[inside].view(1, -1).repeat(B, 1))
            self.outside = finish(tok[~inside].view(1, -1).repeat(B, 1))


# ---------------------------------------------------------------------------------------------- linear attention
def attn_fwd(ops, q, k, v, G, L, S, nh):
    """attention.py:31-49 with the reference's /S ... *S guard cancelled.  q [G*L, C], k / v [G*S, C] -> msg, saved."""
    Q, K = ops.ew(q, None, "elu1"), ops.ew(k, None, "elu1")
    KV, Ks = ops.attn_reduce(K, v, None, G, S, nh)
    den = ops.head_dot(Q, Ks, nh, L, ATTN_EPS)                      # [G*L, nh] = Q_h . Ks_h + eps
    att = ops.head_scale(ops.attn_apply(Q, KV, G, L, nh, False), den, nh, True)
    return att, (q, k, v, Q, K, KV, Ks, den, att)


def attn_bwd(ops, saved, datt, G, L, S, nh):
    q, k, v, Q, K, KV, Ks, den, att = saved
    dnum = ops.head_scale(datt, den, nh, True)
    dden = ops.ew(ops.head_dot(datt, att, nh, 0, 0.0), den, "neg_div")          # -(dmsg . msg) / den
    dQ = ops.attn_apply(dnum, KV, G, L, nh, True)
    ops.head_axpy(dQ, dden, Ks, nh, L)                                            # + dden * Ks
    dKV, dKs = ops.attn_reduce(Q, dnum, dden, G, L, nh)
    dK = ops.attn_apply(v, dKV, G, S, nh, True)
    ops.group_add(dK, dKs, S)
    dV = ops.attn_apply(K, dKV, G, S, nh, False)
    return ops.ew(dQ, q, "elu1_grad_mul"), ops.ew(dK, k, "elu1_grad_mul"), dV


# ---------------------------------------------------------------------------------------------- LoFTR layer
def loftr_fwd(ops, P, x, src, G, L, S, nh):
    """transformer.py:41-71: out = x + LN2(MLP([x | LN1(merge(attention))]))."""
    C = x.shape[1]
    q, k, v = ops.linear(x, P["q_proj.weight"]), ops.linear(src, P["k_proj.weight"]), ops.linear(src, P["v_proj.weight"])
    att, asaved = attn_fwd(ops, q, k, v, G, L, S, nh)
    m0 = ops.linear(att, P["merge.weight"])
    m1 = ops.ln_fwd(m0, P["norm1.weight"], P["norm1.bias"], LN_EPS)
    W1 = P["mlp.0.weight"]
    W1a, W1b = W1[:, :C].contiguous(), W1[:, C:].contiguous()
    h0 = ops.linear(x, W1a)
    ops.linear(m1, W1b, acc=h0)
    m2 = ops.linear(ops.ew(h0, None, "relu"), P["mlp.2.weight"])
    out = ops.ew(x, ops.ln_fwd(m2, P["norm2.weight"], P["norm2.bias"], LN_EPS), "add")
    return out, (x, src, asaved, m0, m1, h0, m2, W1a, W1b, (G, L, S, nh))


def loftr_bwd(ops, P, saved, dout):
    x, src, asaved, m0, m1, h0, m2, W1a, W1b, (G, L, S, nh) = saved
    att = asaved[-1]
    g = {}
    dm2, g["norm2.weight"], g["norm2.bias"] = ops.ln_bwd(m2, P["norm2.weight"], dout, LN_EPS)
    h = ops.ew(h0, None, "relu")
    g["mlp.2.weight"] = ops.linear_dw(dm2, h)
    dh0 = ops.ew(ops.linear_dx(dm2, P["mlp.2.weight"]), h0, "relu_mask")
    g["mlp.0.weight"] = torch.cat([ops.linear_dw(dh0, x), ops.linear_dw(dh0, m1)], dim=1)
    dx = ops.ew(dout, ops.linear_dx(dh0, W1a), "add")
    dm0, g["norm1.weight"], g["norm1.bias"] = ops.ln_bwd(m0, P["norm1.weight"], ops.linear_dx(dh0, W1b), LN_EPS)
    g["merge.weight"] = ops.linear_dw(dm0, att)
    dq, dk, dv = attn_bwd(ops, asaved, ops.linear_dx(dm0, P["merge.weight"]), G, L, S, nh)
    g["q_proj.weight"], g["k_proj.weight"], g["v_proj.weight"] = ops.linear_dw(dq, x), ops.linear_dw(dk, src), ops.linear_dw(dv, src)
    dx = ops.ew(dx, ops.linear_dx(dq, P["q_proj.weight"]), "add")
    dsrc = ops.ew(ops.linear_dx(dk, P["k_proj.weight"]), ops.linear_dx(dv, P["v_proj.weight"]), "add")
    return dx, dsrc, g


# ---------------------------------------------------------------------------------------------- layers
def h2i_fwd(ops, P, feat, ztok, zmask, ix: Indexer):
    """fusion.py:132-157 (change_embedding, no resize): zone cells += mask * loftr(zone cells, zone tokens)."""
    g = ix.g
    G, L = ix.B * g.zone_num ** 2, g.p1 * g.p2
    S = ztok.shape[0] // G
    xz = ops.gather_rows(feat, ix.zone)
    oz, ls = loftr_fwd(ops, P, xz, ztok, G, L, S, 4)
    out = ops.scatter_add_rows(ops.group_scale(oz, zmask, L), ix.zone, feat)          # feat + scatter
    return out, (ls, G, L)


def h2i_bwd(ops, P, saved, zmask, ix: Indexer, dout):
    ls, G, L = saved
    dt = ops.group_scale(ops.gather_rows(dout, ix.zone), zmask, L)
    dxz, dztok, g = loftr_bwd(ops, P, ls, dt)
    return ops.scatter_add_rows(dxz, ix.zone, dout), dztok, g


def lsa_fwd(ops, P, x, ix: Indexer):
    G, L = ix.B * ix.nwin, ix.ws * ix.ws
    win = ops.gather_rows(x, ix.win)
    ow, ls = loftr_fwd(ops, sub(P, "encoder_layer."), win, win, G, L, L, 8)
    return ops.gather_rows(ow, ix.win_inv), ls


def lsa_bwd(ops, P, ls, ix: Indexer, dout):
    dq, dsrc, g = loftr_bwd(ops, sub(P, "encoder_layer."), ls, ops.gather_rows(dout, ix.win))
    return ops.gather_rows(ops.ew(dq, dsrc, "add"), ix.win_inv), _pre("encoder_layer.", g)


def _sr_matrix(P):
    """sr.weight [C, C, ws, ws] -> [C, ws*ws*C] with the columns ordered (tap, cin) like the im2col rows below."""
    w = P["sr.weight"]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def gsa_fwd(ops, P, x, ix: Indexer):
    """transformer.py:138-150.  The kernel = stride = ws sub-sampling conv is ONE gather (the ws*ws taps of every
    sub-sampled token side by side: an im2col matrix [B*Ns, ws*ws*C]) and ONE product with the [C, ws*ws*C] weight."""
    C = x.shape[1]
    col = ops.gather_rows(x, ix.sr_col).view(ix.B * ix.Ns, -1)
    s0 = ops.linear(col, _sr_matrix(P), P["sr.bias"])
    s1 = ops.ln_fwd(s0, P["norm.weight"], P["norm.bias"], LN_EPS)
    out, ls = loftr_fwd(ops, sub(P, "encoder_layer."), x, s1, ix.B, ix.N, ix.Ns, 8)
    return out, (ls, s0, x)


def gsa_bwd(ops, P, saved, ix: Indexer, dout):
    ls, s0, x = saved
    dx, ds1, g = loftr_bwd(ops, sub(P, "encoder_layer."), ls, dout)
    grads = _pre("encoder_layer.", g)
    ds0, grads["norm.weight"], grads["norm.bias"] = ops.ln_bwd(s0, P["norm.weight"], ds1, LN_EPS)
    grads["sr.bias"] = ops.colsum(ds0)
    C, ws = x.shape[1], ix.ws
    col = ops.gather_rows(x, ix.sr_col).view(ix.B * ix.Ns, -1)                      # recomputed, not kept
    grads["sr.weight"] = ops.linear_dw(ds0, col).view(C, ws, ws, C).permute(0, 3, 1, 2).contiguous()
    dcol = ops.linear_dx(ds0, _sr_matrix(P)).view(-1, C)                           # [B*Ns*ws*ws, C]: the windows do not overlap
    return ops.scatter_add_rows(dcol, ix.sr_col, dx), grads


def _conv_matrix(w):
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin], columns ordered (tap, cin)."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _conv3x3(ops, srcs, weights, ix: Indexer):
    """3x3 conv, zero padding 1, of the concatenated sources: per source ONE gather (the nine taps of every pixel side by
    side, [rows, 9*Cin]) and ONE product with the [Cout, 9*Cin] weight;  weights[i] [Cout, Cin_i, 3, 3]."""
    out = None
    for s_, w in zip(srcs, weights):
        col = ops.gather_rows(s_, ix.conv_col).view(s_.shape[0], -1)
        out = ops.linear(col, _conv_matrix(w)) if out is None else ops.linear(col, _conv_matrix(w), acc=out)
    return out


def _conv3x3_bwd(ops, srcs, weights, ix: Indexer, dy):
    """Returns ([d src_i], [d weights_i]).  dW = dy^T x im2col (recomputed); d src = the adjoint of the im2col gather
    (scatter-add of dy x W over the same index matrix: a pixel collects from its nine neighbours)."""
    dsrc, dws = [], []
    for s_, w in zip(srcs, weights):
        cin = w.shape[1]
        col = ops.gather_rows(s_, ix.conv_col).view(s_.shape[0], -1)
        dws.append(ops.linear_dw(dy, col).view(w.shape[0], 3, 3, cin).permute(0, 3, 1, 2).contiguous())
        dcol = ops.linear_dx(dy, _conv_matrix(w)).view(-1, cin)
        dsrc.append(ops.scatter_add_rows(dcol, ix.conv_col, ops.zeros_like(s_)))
    return dsrc, dws


def dapm_fwd(ops, P, bn1, bn2, feat, ix: Indexer):
    """transformer.py:204-248 in train mode: out = feat + BN2(conv2(BN1(conv1([feat | message map]))))."""
    C = feat.shape[1]
    fin, fout = ops.gather_rows(feat, ix.inside), ops.gather_rows(feat, ix.outside)
    q, k, v = ops.linear(fout, P["q_proj.weight"]), ops.linear(fin, P["k_proj.weight"]), ops.linear(fin, P["v_proj.weight"])
    att, asaved = attn_fwd(ops, q, k, v, ix.B, ix.No, ix.Ni, 4)
    tmp = ops.scatter_add_rows(att, ix.outside, ops.zeros_like(feat))
    W1 = P["conv1.weight"]
    W1a, W1b = W1[:, :C].contiguous(), W1[:, C:].contiguous()
    c1 = _conv3x3(ops, [feat, tmp], [W1a, W1b], ix)
    m1, mean1, rstd1 = ops.bn_fwd(c1, bn1, False)
    c2 = _conv3x3(ops, [m1], [P["conv2.weight"]], ix)
    y, mean2, rstd2 = ops.bn_fwd(c2, bn2, False)
    return ops.ew(feat, y, "add"), (feat, fin, fout, asaved, tmp, c1, mean1, rstd1, m1, c2, mean2, rstd2, W1a, W1b)


def dapm_bwd(ops, P, bn1, bn2, saved, ix: Indexer, dout):
    feat, fin, fout, asaved, tmp, c1, mean1, rstd1, m1, c2, mean2, rstd2, W1a, W1b = saved
    g = {}
    dc2, g["bn2.weight"], g["bn2.bias"] = ops.bn_bwd(dout, c2, mean2, rstd2, bn2, False)
    (dm1,), (g["conv2.weight"],) = _conv3x3_bwd(ops, [m1], [P["conv2.weight"]], ix, dc2)
    dc1, g["bn1.weight"], g["bn1.bias"] = ops.bn_bwd(dm1, c1, mean1, rstd1, bn1, False)
    (dfeat_c, dtmp), (gw1a, gw1b) = _conv3x3_bwd(ops, [feat, tmp], [W1a, W1b], ix, dc1)
    g["conv1.weight"] = torch.cat([gw1a, gw1b], dim=1)
    dfeat = ops.ew(dout, dfeat_c, "add")
    dq, dk, dv = attn_bwd(ops, asaved, ops.gather_rows(dtmp, ix.outside), ix.B, ix.No, ix.Ni, 4)
    g["q_proj.weight"], g["k_proj.weight"], g["v_proj.weight"] = ops.linear_dw(dq, fout), ops.linear_dw(dk, fin), ops.linear_dw(dv, fin)
    dfeat = ops.scatter_add_rows(ops.linear_dx(dq, P["q_proj.weight"]), ix.outside, dfeat)
    din = ops.ew(ops.linear_dx(dk, P["k_proj.weight"]), ops.linear_dx(dv, P["v_proj.weight"]), "add")
    return ops.scatter_add_rows(din, ix.inside, dfeat), g


# ---------------------------------------------------------------------------------------------- the whole call
def fusion_fwd(ops, mod, P, x_nchw, feat1, zmask, ix: Indexer, oy, ox):
    """fusion.py:52-188 in train mode.  Returns (out NCHW, saved)."""
    B, C, H, W = x_nchw.shape
    feat = ops.posenc_tokens(x_nchw, P["positional_encodings"], mod.max_resolution, oy, ox)        # [B*N, C]
    S = feat1.shape[2]
    pos2_rows = ops.gather_rows(P["positional_encodings2"], ops.index_mod(feat1.shape[0] * feat1.shape[1] * S, S, feat.device))
    ztok = ops.ew(feat1.reshape(-1, C), pos2_rows, "add")
    saved = []
    for i, name in enumerate(mod.layer_names):
        Pl = sub(P, f"layers.{i}.")
        layer = mod.layers[i]
        if name == "hist2image":
            feat, s_ = h2i_fwd(ops, Pl, feat, ztok, zmask, ix)
        elif name == "image":
            mid, s1 = lsa_fwd(ops, sub(Pl, "lga."), feat, ix)
            feat, s2 = gsa_fwd(ops, sub(Pl, "gsa."), mid, ix)
            s_ = (s1, s2)
        elif name == "combine1":
            tp = layer.transformer_path
            mid, s1 = dapm_fwd(ops, sub(Pl, "transformer_path."), tp.bn1, tp.bn2, feat, ix)
            feat, s2 = ops.lkpm_fwd(layer.large_kernel_path, mid, B, H, W)
            s_ = (s1, s2)
        else:
            raise NotImplementedError(name)
        saved.append(s_)
    return ops.tokens_to_nchw(feat, B, C, H, W), (saved, ztok)


def fusion_bwd(ops, mod, P, saved_all, zmask, ix: Indexer, oy, ox, dout_nchw, feat1_shape):
    """Returns (dx NCHW, dfeat1, {state_dict name: gradient})."""
    saved, ztok = saved_all
    B, C, H, W = dout_nchw.shape
    d = ops.nchw_to_tokens(dout_nchw)
    grads: Dict[str, torch.Tensor] = {}
    dztok = None
    for i in reversed(range(len(mod.layer_names))):
        name, Pl, layer, s_ = mod.layer_names[i], sub(P, f"layers.{i}."), mod.layers[i], saved[i]
        if name == "hist2image":
            d, dz, gl = h2i_bwd(ops, Pl, s_, zmask, ix, d)
            dztok = dz if dztok is None else ops.ew(dztok, dz, "add")
        elif name == "image":
            d, g2 = gsa_bwd(ops, sub(Pl, "gsa."), s_[1], ix, d)
            d, g1 = lsa_bwd(ops, sub(Pl, "lga."), s_[0], ix, d)
            gl = {**_pre("lga.", g1), **_pre("gsa.", g2)}
        else:
            tp = layer.transformer_path
            d, g2 = ops.lkpm_bwd(layer.large_kernel_path, s_[1], d, B, H, W)
            d, g1 = dapm_bwd(ops, sub(Pl, "transformer_path."), tp.bn1, tp.bn2, s_[0], ix, d)
            gl = {**_pre("transformer_path.", g1), **_pre("large_kernel_path.", g2)}
        grads.update(_pre(f"layers.{i}.", gl))
    mh, mw = mod.max_resolution
    dpos = torch.zeros(mh, mw, C, device=d.device, dtype=d.dtype)
    dpos[oy:oy + H, ox:ox + W] = ops.colsum(d.view(B, H * W * C)).view(H, W, C)
    grads["positional_encodings"] = dpos.view(mh * mw, C)
    dfeat1 = None
    if dztok is not None:
        Bf, Z, S, _ = feat1_shape
        grads["positional_encodings2"] = ops.colsum(dztok.view(Bf * Z, S * C)).view(S, C)
        dfeat1 = dztok.view(feat1_shape)
    return ops.tokens_to_nchw(d, B, C, H, W), dfeat1, grads
