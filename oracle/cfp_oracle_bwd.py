"""Explicit backward restatements for the training row of the scope table.

TEST INFRASTRUCTURE ONLY (same rules as ``cfp_oracle.py``: imported by ``tests/`` only, never by the product).

``cfp_oracle.py`` in train mode gives gradients through autograd - enough to CHECK a backward kernel, not to write
one.  This file states the backward of the path's non-trivial pieces in closed form, in the same two-phase shape the
forward kernels have (per-group attention state, then per-row work), so that a fused backward kernel has a formula to
follow line by line.  Every function is pinned against autograd over the forward restatement
(``tests/test_oracle_bwd.py``, float64, 1e-10), which in turn is pinned on the reference's own train-mode forward +
backward (``tests/test_oracle_train_golden.py``).

Linear attention (src/models/attention.py:31-49), per group n and head h, with the reference's ``/S ... *S`` guard
cancelled (it cancels exactly):

    forward   K = elu(k)+1, Q = elu(q)+1                       (feature map, attention.py:31-32)
              KV  = sum_s K_s^T V_s   [d x d],   Ks = sum_s K_s   [d]        <- phase 1 (source rows -> state)
              den_l = Q_l . Ks + eps,   msg_l = (Q_l KV) / den_l              <- phase 2 (query rows)

    backward  given dmsg_l:
              dnum_l = dmsg_l / den_l,      dden_l = -(dmsg_l . msg_l) / den_l
              dQ_l   = dnum_l KV^T + dden_l Ks                               <- phase 2' (query rows, needs KV, Ks)
              dKV    = sum_l Q_l^T dnum_l,  dKs = sum_l dden_l Q_l           <- ... which also REDUCES a state gradient
              dK_s   = V_s dKV^T + dKs,     dV_s = K_s dKV                   <- phase 1' (source rows, needs dKV, dKs)
              dq = dQ * elu'(q),  dk = dK * elu'(k),   elu'(t) = 1 (t > 0) else exp(t) = elu(t)+1

So the backward is the forward's mirror image: a row pass over the QUERY rows that produces dQ and accumulates a
per-group [d x d] + [d] state gradient (the role kv_state plays in the forward, with the same atomics / run logic),
followed by a row pass over the SOURCE rows.  Both are the same GEMM-shaped tiles as the forward kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Tuple

import torch
import torch.nn.functional as F

from .cfp_oracle import ATTN_EPS, BN_EPS, LKPM_LN_EPS, LN_EPS, linear_attention

Tensor = torch.Tensor


def elu1_grad(t: Tensor) -> Tensor:
    """d/dt (elu(t) + 1)"""
    return torch.where(t > 0, torch.ones_like(t), torch.exp(t))


def linear_attention_bwd(q: Tensor, k: Tensor, v: Tensor, nhead: int, dmsg: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Gradients of ``cfp_oracle.linear_attention(q, k, v, nhead)`` w.r.t. q [n,L,C], k, v [n,S,C] given dmsg [n,L,C]."""
    n, L, C = q.shape
    S = k.shape[1]
    d = C // nhead
    Q = (F.elu(q) + 1).view(n, L, nhead, d)
    K = (F.elu(k) + 1).view(n, S, nhead, d)
    V = v.view(n, S, nhead, d)
    # forward state (phase 1) and query-side quantities (phase 2)
    KV = torch.einsum("nshd,nshv->nhdv", K, V)
    Ks = K.sum(dim=1)                                                   # [n,h,d]
    den = torch.einsum("nlhd,nhd->nlh", Q, Ks) + ATTN_EPS               # [n,L,h]
    msg = torch.einsum("nlhd,nhdv->nlhv", Q, KV) / den.unsqueeze(-1)
    g = dmsg.view(n, L, nhead, d)
    # phase 2': query rows
    dnum = g / den.unsqueeze(-1)
    dden = -(g * msg).sum(-1) / den                                     # [n,L,h]
    dQ = torch.einsum("nlhv,nhdv->nlhd", dnum, KV) + dden.unsqueeze(-1) * Ks.unsqueeze(1)
    dKV = torch.einsum("nlhd,nlhv->nhdv", Q, dnum)                      # state gradient, reduced over the query rows
    dKs = torch.einsum("nlh,nlhd->nhd", dden, Q)
    # phase 1': source rows
    dK = torch.einsum("nshv,nhdv->nshd", V, dKV) + dKs.unsqueeze(1)
    dV = torch.einsum("nshd,nhdv->nshv", K, dKV)
    dq = dQ.reshape(n, L, C) * elu1_grad(q)
    dk = dK.reshape(n, S, C) * elu1_grad(k)
    return dq, dk, dV.reshape(n, S, C)


def layer_norm_bwd(x: Tensor, w: Tensor, dy: Tensor, eps: float = LN_EPS) -> Tuple[Tensor, Tensor, Tensor]:
    """LayerNorm over the last dim: returns (dx, dweight, dbias).  One row = one thread in the chain kernels:
    dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w."""
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (x - mean) * rstd
    g = dy * w
    dx = rstd * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    red = tuple(range(x.dim() - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


def bn_train_bwd(x: Tensor, w: Tensor, dy: Tensor, dim: int, eps: float = BN_EPS) -> Tuple[Tensor, Tensor, Tensor]:
    """Train-mode BatchNorm (batch statistics over every dim but ``dim``): returns (dx, dweight, dbias).
    Two grid-wide reductions (sum dy, sum dy * xhat per channel), then an elementwise pass - the backward twin of the
    forward's (mean, var) reductions.  Note dbias of a conv / linear bias IN FRONT of this BN is exactly zero."""
    dims = [i for i in range(x.dim()) if i != dim]
    shape = [1] * x.dim()
    shape[dim] = -1
    n = x.numel() // x.shape[dim]
    mean = x.mean(dims).view(shape)
    rstd = torch.rsqrt(x.var(dims, unbiased=False) + eps).view(shape)
    xhat = (x - mean) * rstd
    s1 = dy.sum(dims).view(shape)
    s2 = (dy * xhat).sum(dims).view(shape)
    dx = (w.view(shape) * rstd / n) * (n * dy - s1 - xhat * s2)
    return dx, s2.reshape(-1), s1.reshape(-1)


def gelu_erf_grad(t: Tensor) -> Tensor:
    """d/dt of the erf-form GELU the reference uses (convnext.py:33, nn.GELU())."""
    return 0.5 * (1 + torch.erf(t / math.sqrt(2.0))) + t * torch.exp(-0.5 * t * t) / math.sqrt(2.0 * math.pi)


def loftr_layer_bwd(p: Mapping, x: Tensor, source: Tensor, nhead: int, dout: Tensor) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """Backward of ``cfp_oracle.loftr_layer`` (src/models/transformer.py:41-71), stage by stage in the order a fused
    chain kernel would walk it (out -> LN2 -> W2 -> relu -> W1 -> [x | LN1 -> merge -> attention -> q]; source side:
    k / v projections).  Returns (dx, dsource, {parameter name: gradient})."""
    C = x.shape[-1]
    Wq, Wk, Wv, Wm = p["q_proj.weight"], p["k_proj.weight"], p["v_proj.weight"], p["merge.weight"]
    W1, W2 = p["mlp.0.weight"], p["mlp.2.weight"]
    # ---- forward, keeping what the backward needs (a kernel would recompute most of it from x and the state)
    q, k, v = x @ Wq.t(), source @ Wk.t(), source @ Wv.t()
    att = linear_attention(q, k, v, nhead)
    m0 = att @ Wm.t()
    m1 = F.layer_norm(m0, (C,), p["norm1.weight"], p["norm1.bias"], LN_EPS)
    cat = torch.cat([x, m1], dim=-1)
    h0 = cat @ W1.t()
    h = torch.relu(h0)
    m2 = h @ W2.t()
    # out = x + LN2(m2)
    grads: Dict[str, Tensor] = {}
    flat = lambda t: t.reshape(-1, t.shape[-1])          # noqa: E731
    dx = dout.clone()
    dm2, grads["norm2.weight"], grads["norm2.bias"] = layer_norm_bwd(m2, p["norm2.weight"], dout)
    grads["mlp.2.weight"] = flat(dm2).t() @ flat(h)
    dh0 = (dm2 @ W2) * (h0 > 0)
    grads["mlp.0.weight"] = flat(dh0).t() @ flat(cat)
    dcat = dh0 @ W1
    dx = dx + dcat[..., :C]
    dm0, grads["norm1.weight"], grads["norm1.bias"] = layer_norm_bwd(m0, p["norm1.weight"], dcat[..., C:])
    grads["merge.weight"] = flat(dm0).t() @ flat(att)
    dq, dk, dv = linear_attention_bwd(q, k, v, nhead, dm0 @ Wm)
    grads["q_proj.weight"] = flat(dq).t() @ flat(x)
    grads["k_proj.weight"] = flat(dk).t() @ flat(source)
    grads["v_proj.weight"] = flat(dv).t() @ flat(source)
    dx = dx + dq @ Wq
    dsource = dk @ Wk + dv @ Wv
    return dx, dsource, grads


def depthwise_conv_bwd(x: Tensor, w: Tensor, dy: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Backward of the k x k depthwise conv of LKPM (convnext.py:29, zero padding (k-1)/2, stride 1): (dx, dw, dbias).

    dx is the SAME depthwise conv applied to dy with the kernel flipped in both axes - so the forward's Toeplitz
    tensor-core kernel serves it unchanged, fed with flipped taps.  dw[c][i][j] = sum_{b,y,x} dy[b,c,y,x] *
    xpad[b,c,y+i,x+j]: k*k correlations per channel plane (a [k*k x H*W] x [H*W] contraction per (b, c))."""
    k = w.shape[-1]
    pad = (k - 1) // 2
    C = x.shape[1]
    dx = F.conv2d(dy, torch.flip(w, dims=(2, 3)), padding=pad, groups=C)
    xp = F.pad(x, (pad, pad, pad, pad))
    H, W = x.shape[2], x.shape[3]
    dw = torch.stack([torch.stack([(dy * xp[:, :, i:i + H, j:j + W]).sum(dim=(0, 2, 3)) for j in range(k)], dim=-1)
                      for i in range(k)], dim=-2)                                 # [C, k, k]
    return dx, dw.unsqueeze(1), dy.sum(dim=(0, 2, 3))


def lkpm_bwd(p: Mapping, feat0: Tensor, H: int, W: int, dout: Tensor) -> Tuple[Tensor, Dict[str, Tensor]]:
    """Backward of ``cfp_oracle.lkpm`` in TRAIN mode (convnext.py:42-58 with batch-statistics BN): returns
    (dfeat0, {parameter name: gradient}).  Token-major in / out like the forward."""
    B, N, C = feat0.shape
    w, bias = p["dwconv2.weight"], p["dwconv2.bias"]
    k = w.shape[-1]
    m = feat0.transpose(1, 2).reshape(B, C, H, W)
    # ---- forward (train mode)
    y0 = F.conv2d(m, w, bias, padding=(k - 1) // 2, groups=C)
    dims = (0, 2, 3)
    mean, var = y0.mean(dims, keepdim=True), y0.var(dims, unbiased=False, keepdim=True)
    y1 = (y0 - mean) * torch.rsqrt(var + BN_EPS) * p["bn1.weight"].view(1, C, 1, 1) + p["bn1.bias"].view(1, C, 1, 1)
    y2 = torch.relu(y1).reshape(B, C, N).transpose(1, 2)                        # tokens
    y3 = F.layer_norm(y2, (C,), p["norm.weight"], p["norm.bias"], LKPM_LN_EPS)
    h0 = y3 @ p["pwconv1.weight"].t() + p["pwconv1.bias"]
    h = F.gelu(h0)
    # out = feat0 + h W2^T + b2
    flat = lambda t: t.reshape(-1, t.shape[-1])          # noqa: E731
    g: Dict[str, Tensor] = {}
    g["pwconv2.bias"] = flat(dout).sum(0)
    g["pwconv2.weight"] = flat(dout).t() @ flat(h)
    dh0 = (dout @ p["pwconv2.weight"]) * gelu_erf_grad(h0)
    g["pwconv1.bias"] = flat(dh0).sum(0)
    g["pwconv1.weight"] = flat(dh0).t() @ flat(y3)
    dy2, g["norm.weight"], g["norm.bias"] = layer_norm_bwd(y2, p["norm.weight"], dh0 @ p["pwconv1.weight"], LKPM_LN_EPS)
    dy1 = (dy2 * (y2 > 0)).transpose(1, 2).reshape(B, C, H, W)
    dy0, g["bn1.weight"], g["bn1.bias"] = bn_train_bwd(y0, p["bn1.weight"], dy1, 1)
    dm, g["dwconv2.weight"], g["dwconv2.bias"] = depthwise_conv_bwd(m, w, dy0)
    return dout + dm.reshape(B, C, N).transpose(1, 2), g


# ----------------------------------------------------------------------------------------------
# Whole layers and the whole TransformerFusion call, TRAIN mode (batch-statistics BN).  Each function recomputes the
# forward internals it needs from the layer's input (what a fused backward kernel does too) and returns the gradient
# of the layer input plus {parameter name: gradient} with names relative to the layer's prefix.
# ----------------------------------------------------------------------------------------------
from torch.nn.grad import conv2d_input, conv2d_weight  # noqa: E402  (explicit adjoints of conv2d, no autograd graph)

from .cfp_oracle import dapm as _dapm_fwd, hist2image as _h2i_fwd, loftr_layer, lkpm as _lkpm_fwd, lsa as _lsa_fwd  # noqa: E402
from .cfp_oracle import gsa as _gsa_fwd, sub, twins_window_size, zone_geometry  # noqa: E402


def _prefixed(prefix: str, grads: Mapping[str, Tensor]) -> Dict[str, Tensor]:
    return {prefix + k: v for k, v in grads.items()}


def _conv_bn_train(x: Tensor, w: Tensor, p: Mapping, bn: str, stride: int = 1, padding: int = 1):
    y0 = F.conv2d(x, w, padding=padding, stride=stride)
    dims = (0, 2, 3)
    rstd = torch.rsqrt(y0.var(dims, unbiased=False, keepdim=True) + BN_EPS)
    C = y0.shape[1]
    return y0, (y0 - y0.mean(dims, keepdim=True)) * rstd * p[bn + ".weight"].view(1, C, 1, 1) + p[bn + ".bias"].view(1, C, 1, 1)


def lsa_bwd(p: Mapping, x: Tensor, H: int, W: int, ws: int, dout: Tensor):
    """LSA (transformer.py:89-116): the window regroup is a permutation of the zero-padded map, its adjoint the inverse
    permutation followed by the crop; queries and keys are the same tokens, so both gradients land on the windows."""
    B, N, C = x.shape
    pb, pr = (ws - H % ws) % ws, (ws - W % ws) % ws
    Hp, Wp = H + pb, W + pr
    nh, nw = Hp // ws, Wp // ws

    def to_win(t):
        t = F.pad(t.view(B, H, W, C), (0, 0, 0, pr, 0, pb))
        return t.view(B, nh, ws, nw, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nh * nw, ws * ws, C)

    def from_win(t):
        t = t.view(B, nh, nw, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
        return t[:, :H, :W].reshape(B, N, C)

    win = to_win(x)
    dq, dsrc, g = loftr_layer_bwd(sub(p, "encoder_layer."), win, win, 8, to_win(dout))   # padded output cells: zero cotangent
    return from_win(dq + dsrc), _prefixed("encoder_layer.", g)


def gsa_bwd(p: Mapping, x: Tensor, H: int, W: int, ws: int, dout: Tensor):
    """GSA (transformer.py:138-150): source = LN(sr(x)); the strided conv's adjoints are conv2d_input / conv2d_weight."""
    B, N, C = x.shape
    m = x.transpose(1, 2).reshape(B, C, H, W)
    s0 = F.conv2d(m, p["sr.weight"], p["sr.bias"], stride=ws)
    hs, wsz = s0.shape[2], s0.shape[3]
    s0t = s0.reshape(B, C, -1).transpose(1, 2)
    s1 = F.layer_norm(s0t, (C,), p["norm.weight"], p["norm.bias"], LN_EPS)
    dx, ds1, g = loftr_layer_bwd(sub(p, "encoder_layer."), x, s1, 8, dout)
    grads = _prefixed("encoder_layer.", g)
    ds0t, grads["norm.weight"], grads["norm.bias"] = layer_norm_bwd(s0t, p["norm.weight"], ds1)
    ds0 = ds0t.transpose(1, 2).reshape(B, C, hs, wsz)
    grads["sr.bias"] = ds0.sum(dim=(0, 2, 3))
    grads["sr.weight"] = conv2d_weight(m, p["sr.weight"].shape, ds0, stride=ws)
    dm = conv2d_input(m.shape, p["sr.weight"], ds0, stride=ws)
    return dx + dm.reshape(B, C, N).transpose(1, 2), grads


def twins_bwd(p: Mapping, x: Tensor, H: int, W: int, ws: int, dout: Tensor):
    y = _lsa_fwd(sub(p, "lga."), x, H, W, ws)
    dy, g2 = gsa_bwd(sub(p, "gsa."), y, H, W, ws, dout)
    dx, g1 = lsa_bwd(sub(p, "lga."), x, H, W, ws, dy)
    return dx, {**_prefixed("lga.", g1), **_prefixed("gsa.", g2)}


def dapm_bwd(p: Mapping, feat0: Tensor, g: Mapping[str, int], H: int, W: int, dout: Tensor, nhead: int = 4):
    """DAPM (transformer.py:204-248), train-mode BN: out = feat0 + BN2(conv2(BN1(conv1([feat0 | msg map])))), msg map =
    attention of the outside tokens over the inside tokens, zero inside the zone rectangle."""
    B, N, C = feat0.shape
    inside = torch.zeros(H, W, dtype=torch.bool)
    inside[g["ry0"]:g["ry1"], g["rx0"]:g["rx1"]] = True
    inside = inside.reshape(-1)
    fin, fout = feat0[:, inside], feat0[:, ~inside]
    Wq, Wk, Wv = p["q_proj.weight"], p["k_proj.weight"], p["v_proj.weight"]
    q, k, v = fout @ Wq.t(), fin @ Wk.t(), fin @ Wv.t()
    tmp = torch.zeros_like(feat0)
    tmp[:, ~inside] = linear_attention(q, k, v, nhead)
    m0 = torch.cat([feat0, tmp], dim=2).transpose(1, 2).reshape(B, 2 * C, H, W)
    c1, m1 = _conv_bn_train(m0, p["conv1.weight"], p, "bn1")
    c2, _ = _conv_bn_train(m1, p["conv2.weight"], p, "bn2")
    grads: Dict[str, Tensor] = {}
    dy = dout.transpose(1, 2).reshape(B, C, H, W)
    dc2, grads["bn2.weight"], grads["bn2.bias"] = bn_train_bwd(c2, p["bn2.weight"], dy, 1)
    grads["conv2.weight"] = conv2d_weight(m1, p["conv2.weight"].shape, dc2, padding=1)
    dm1 = conv2d_input(m1.shape, p["conv2.weight"], dc2, padding=1)
    dc1, grads["bn1.weight"], grads["bn1.bias"] = bn_train_bwd(c1, p["bn1.weight"], dm1, 1)
    grads["conv1.weight"] = conv2d_weight(m0, p["conv1.weight"].shape, dc1, padding=1)
    dm0 = conv2d_input(m0.shape, p["conv1.weight"], dc1, padding=1).reshape(B, 2 * C, N).transpose(1, 2)
    dfeat = dout + dm0[..., :C]
    dq, dk, dv = linear_attention_bwd(q, k, v, nhead, dm0[..., C:][:, ~inside])       # only outside cells carry a message
    flat = lambda t: t.reshape(-1, t.shape[-1])          # noqa: E731
    grads["q_proj.weight"] = flat(dq).t() @ flat(fout)
    grads["k_proj.weight"] = flat(dk).t() @ flat(fin)
    grads["v_proj.weight"] = flat(dv).t() @ flat(fin)
    dfeat = dfeat.clone()
    dfeat[:, ~inside] += dq @ Wq
    dfeat[:, inside] += dk @ Wk + dv @ Wv
    return dfeat, grads


def combine1_bwd(p: Mapping, feat0: Tensor, g, H: int, W: int, dout: Tensor):
    y = _dapm_fwd(sub(p, "transformer_path."), feat0, g, H, W, bn_stats={})
    dy, g2 = lkpm_bwd(sub(p, "large_kernel_path."), y, H, W, dout)
    dx, g1 = dapm_bwd(sub(p, "transformer_path."), feat0, g, H, W, dy)
    return dx, {**_prefixed("transformer_path.", g1), **_prefixed("large_kernel_path.", g2)}


def bilinear_matrix(n_in: int, n_out: int, dtype=torch.float64) -> Tensor:
    """[n_out, n_in] matrix of F.interpolate(mode="bilinear", align_corners=True) along one axis: the resize is
    separable, out = Wy in Wx^T per channel, so its adjoint is Wy^T dout Wx - the same two small GEMMs transposed."""
    m = torch.zeros(n_out, n_in, dtype=dtype)
    for o in range(n_out):
        f = o * ((n_in - 1) / (n_out - 1)) if n_out > 1 else 0.0
        i0 = min(int(f), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        m[o, i0] += 1.0 - (f - i0)
        m[o, i1] += f - i0
    return m


def hist2image_bwd(p: Mapping, feat0: Tensor, ztok: Tensor, mask: Tensor, g: Mapping[str, int], H: int, W: int, dout: Tensor):
    """hist2image (fusion.py:132-157) with change_embedding:  out = feat0, and on the zone rectangle
    out += crop(resize_back(mask * loftr(resize(canvas cells), zone tokens)))  where the canvas is cut from feat0 itself
    and the two resizes exist only in the resize branch (fusion.py:140-141,146-149).  Returns (dfeat0, dztok, grads)."""
    B, N, C = feat0.shape
    zn, p1, p2 = g["zone_num"], g["p1"], g["p2"]
    tzh, tzw = g["tzh"], g["tzw"]
    oh, ow = zn * p1, zn * p2
    top, left = max(-g["sy_wo"], 0), max(-g["sx_wo"], 0)
    hh, ww = g["ry1"] - g["ry0"], g["rx1"] - g["rx0"]
    rect = (slice(None), slice(g["ry0"], g["ry1"]), slice(g["rx0"], g["rx1"]))
    cv = (slice(None), slice(top, top + hh), slice(left, left + ww))
    interp = bool(g["interpolate"])
    if interp:
        Wy, Wx = bilinear_matrix(tzh, oh, feat0.dtype), bilinear_matrix(tzw, ow, feat0.dtype)        # canvas -> zones
        Vy, Vx = bilinear_matrix(oh, tzh, feat0.dtype), bilinear_matrix(ow, tzw, feat0.dtype)        # zones -> canvas

    def to_canvas(t_rect):                    # in-image rectangle -> zero-padded [B,tzh,tzw,C] canvas
        canvas = t_rect.new_zeros(B, tzh, tzw, C)
        canvas[cv] = t_rect
        return canvas

    def zones(canvas_oh_ow):                  # [B,oh,ow,C] -> [(B zn zn), p1 p2, C]
        return canvas_oh_ow.view(B, zn, p1, zn, p2, C).permute(0, 1, 3, 2, 4, 5).reshape(B * zn * zn, p1 * p2, C)

    def unzones(t):
        return t.view(B, zn, zn, p1, p2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, oh, ow, C)

    f = feat0.view(B, H, W, C)
    canvas = to_canvas(f[rect])
    xz = zones(torch.einsum("oy,byxc,px->bopc", Wy, canvas, Wx) if interp else canvas)
    mz = mask.reshape(B * zn * zn, 1, 1).to(feat0.dtype)
    # cotangent of the layer output: rectangle -> canvas (zero outside the image) -> adjoint of the resize back -> zones
    dcan = to_canvas(dout.view(B, H, W, C)[rect])
    dt = zones(torch.einsum("yo,byxc,xp->bopc", Vy, dcan, Vx) if interp else dcan) * mz    # invalid zones: zeroed, residual included
    dxz, dztok, grads = loftr_layer_bwd(p, xz, ztok, 4, dt)
    dcv = unzones(dxz)
    if interp:
        dcv = torch.einsum("oy,bopc,px->byxc", Wy, dcv, Wx)                                   # adjoint of the resize to zones
    dfeat = dout.clone().view(B, H, W, C)
    dfeat[rect] += dcv[cv]
    return dfeat.view(B, N, C), dztok, grads


def pointnet_block_bwd(p: Mapping, x: Tensor, dout: Tensor):
    """3 x [pointwise linear + train-mode BN + ReLU] (encoder.py:17-24)."""
    acts = [x]
    pre = []
    for i in (1, 2, 3):
        y0 = F.linear(acts[-1], p[f"conv{i}.weight"][:, :, 0], p[f"conv{i}.bias"])
        dims = list(range(y0.dim() - 1))
        y1 = (y0 - y0.mean(dims)) * torch.rsqrt(y0.var(dims, unbiased=False) + BN_EPS) * p[f"bn{i}.weight"] + p[f"bn{i}.bias"]
        pre.append((y0, y1))
        acts.append(torch.relu(y1))
    grads: Dict[str, Tensor] = {}
    d = dout
    flat = lambda t: t.reshape(-1, t.shape[-1])          # noqa: E731
    for i in (3, 2, 1):
        y0, y1 = pre[i - 1]
        d = d * (y1 > 0)
        d, grads[f"bn{i}.weight"], grads[f"bn{i}.bias"] = bn_train_bwd(y0, p[f"bn{i}.weight"], d, y0.dim() - 1)
        grads[f"conv{i}.bias"] = flat(d).sum(0)
        grads[f"conv{i}.weight"] = (flat(d).t() @ flat(acts[i - 1])).unsqueeze(-1)
        d = d @ p[f"conv{i}.weight"][:, :, 0]
    return d, grads


def hist_encoder_bwd(sd: Mapping, hist: Tensor, douts):
    """Backward of ``cfp_oracle.hist_encoder`` in train mode; ``douts`` = cotangents of the three outputs (None = unused)."""
    x = hist.unsqueeze(-1)
    ins = []
    for i in (1, 2, 3):
        ins.append(x)
        x = _pointnet_fwd_train(sub(sd, f"hist_extractor{i}.pointnet_encoder."), x)
    grads: Dict[str, Tensor] = {}
    d = None
    for i in (3, 2, 1):
        if douts[i - 1] is not None:
            d = douts[i - 1] if d is None else d + douts[i - 1]
        if d is None:
            continue
        d, g = pointnet_block_bwd(sub(sd, f"hist_extractor{i}.pointnet_encoder."), ins[i - 1], d)
        grads.update(_prefixed(f"hist_extractor{i}.pointnet_encoder.", g))
    return d.squeeze(-1), grads


def _pointnet_fwd_train(p: Mapping, x: Tensor) -> Tensor:
    from .cfp_oracle import pointnet_block
    return pointnet_block(p, x, bn_stats={})


def transformer_fusion_bwd(sd: Mapping, layer_names, max_res, x: Tensor, feat1: Tensor, mask: Tensor, patch_info: dict,
                           offsets, dout: Tensor):
    """Backward of one ``TransformerFusion`` call in train mode (fusion.py:52-188; change_embedding, every geometry branch).
    Returns (dx [B,C,H,W], dfeat1 [B,Z,S,C], {state_dict name: gradient})."""
    B, C, H, W = x.shape
    g = zone_geometry(patch_info, max_res[1], H, W)
    oy, ox = offsets
    pos = sd["positional_encodings"].view(max_res[0], max_res[1], C)[oy:oy + H, ox:ox + W]
    feat = (x.permute(0, 2, 3, 1) + pos).reshape(B, H * W, C)
    ztok = (feat1 + sd["positional_encodings2"]).reshape(-1, feat1.shape[2], C)
    ws = twins_window_size(max_res)
    inputs = []
    for i, name in enumerate(layer_names):                                   # forward, keeping every layer's input
        p = sub(sd, f"layers.{i}.")
        inputs.append(feat)
        if name == "image":
            feat = _gsa_fwd(sub(p, "gsa."), _lsa_fwd(sub(p, "lga."), feat, H, W, ws), H, W, ws)
        elif name == "hist2image":
            feat = _h2i_fwd(p, feat, feat, ztok, mask, g, H, W)
        elif name == "combine1":
            feat = _lkpm_fwd(sub(p, "large_kernel_path."),
                             _dapm_fwd(sub(p, "transformer_path."), feat, g, H, W, bn_stats={}), H, W, bn_stats={})
        else:
            raise NotImplementedError(name)
    grads: Dict[str, Tensor] = {}
    d = dout.permute(0, 2, 3, 1).reshape(B, H * W, C)
    dztok = torch.zeros_like(ztok)
    for i in reversed(range(len(layer_names))):
        p, name = sub(sd, f"layers.{i}."), layer_names[i]
        if name == "image":
            d, gl = twins_bwd(p, inputs[i], H, W, ws, d)
        elif name == "hist2image":
            d, dz, gl = hist2image_bwd(p, inputs[i], ztok, mask, g, H, W, d)
            dztok = dztok + dz
        else:
            d, gl = combine1_bwd(p, inputs[i], g, H, W, d)
        grads.update(_prefixed(f"layers.{i}.", gl))
    dmap = d.view(B, H, W, C)
    dpos = torch.zeros(max_res[0], max_res[1], C, dtype=x.dtype)
    dpos[oy:oy + H, ox:ox + W] = dmap.sum(0)
    grads["positional_encodings"] = dpos.view(max_res[0] * max_res[1], C)
    dfeat1 = dztok.view(feat1.shape)
    grads["positional_encodings2"] = dfeat1.sum(dim=(0, 1))
    return dmap.permute(0, 3, 1, 2).contiguous(), dfeat1, grads


def silog_loss_bwd(pred: Tensor, target: Tensor, mask: Tensor) -> Tensor:
    """d silog_loss / d pred in closed form (src/loss.py:9-19; pred [B,1,h,w], target / mask [B,1,H,W]):
    with n masked elements, g = log(up(pred)) - log(target), D = var_unbiased(g) + 0.15 mean(g)^2, L = 10 sqrt(D):
        dL/dg_i = (5 / sqrt(D)) * ( 2 (g_i - mean) / (n - 1) + 0.3 mean / n ),   dL/dup_i = dL/dg_i / up_i,
    scattered through the mask and pulled back through the separable bilinear resize (Wy^T . Wx)."""
    B, _, h, w = pred.shape
    H, W = target.shape[-2:]
    Wy, Wx = bilinear_matrix(h, H, pred.dtype), bilinear_matrix(w, W, pred.dtype)
    up = torch.einsum("oy,bcyx,px->bcop", Wy, pred, Wx)
    g = torch.log(up[mask]) - torch.log(target[mask])
    n = g.numel()
    mean = g.mean()
    D = g.var() + 0.15 * mean * mean
    dg = (5.0 / torch.sqrt(D)) * (2.0 * (g - mean) / (n - 1) + 0.3 * mean / n)
    dup = torch.zeros_like(up)
    dup[mask] = dg / up[mask]
    return torch.einsum("oy,bcop,px->bcyx", Wy, dup, Wx)
