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
