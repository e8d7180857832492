"""Closed-form backward restatements (oracle/cfp_oracle_bwd.py) against autograd over the forward restatement
(oracle/cfp_oracle.py, itself pinned on the reference's train-mode forward + backward).  float64, CPU only."""
import pytest
import torch
import torch.nn.functional as F

from oracle import cfp_oracle as O
from oracle import cfp_oracle_bwd as OB

TOL = 1e-10


def _close(a, b):
    return float((a - b).norm()) <= TOL * max(float(b.norm()), 1.0)


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


@pytest.mark.parametrize("n,L,S,C,nhead", [(3, 16, 16, 32, 4), (2, 9, 16, 128, 4), (2, 36, 36, 64, 8), (1, 50, 7, 32, 8)])
def test_linear_attention_backward(n, L, S, C, nhead):
    q, k, v = (_rand(n, r, C, seed=i).requires_grad_(True) for i, r in enumerate((L, S, S)))
    dmsg = _rand(n, L, C, seed=9)
    O.linear_attention(q, k, v, nhead).backward(dmsg)
    dq, dk, dv = OB.linear_attention_bwd(q.detach(), k.detach(), v.detach(), nhead, dmsg)
    assert _close(dq, q.grad) and _close(dk, k.grad) and _close(dv, v.grad)


def test_layer_norm_backward():
    x, w, b = _rand(5, 7, 64, seed=1).requires_grad_(True), (_rand(64, seed=2) + 2).requires_grad_(True), _rand(64, seed=3).requires_grad_(True)
    dy = _rand(5, 7, 64, seed=4)
    F.layer_norm(x, (64,), w, b, O.LN_EPS).backward(dy)
    dx, dw, db = OB.layer_norm_bwd(x.detach(), w.detach(), dy)
    assert _close(dx, x.grad) and _close(dw, w.grad) and _close(db, b.grad)


@pytest.mark.parametrize("shape,dim", [((2, 32, 6, 5), 1), ((3, 8, 16, 64), 3)])
def test_batchnorm_train_backward(shape, dim):
    x = _rand(*shape, seed=1).requires_grad_(True)
    Cn = shape[dim]
    w, b = (_rand(Cn, seed=2) + 2).requires_grad_(True), _rand(Cn, seed=3).requires_grad_(True)
    p = {"bn.weight": w, "bn.bias": b, "bn.running_mean": torch.zeros(Cn, dtype=torch.float64),
         "bn.running_var": torch.ones(Cn, dtype=torch.float64), "bn.num_batches_tracked": torch.tensor(0)}
    dy = _rand(*shape, seed=4)
    O._bn(x, p, "bn", dim, bn_stats={}).backward(dy)
    dx, dw, db = OB.bn_train_bwd(x.detach(), w.detach(), dy, dim)
    assert _close(dx, x.grad) and _close(dw, w.grad) and _close(db, b.grad)
    # a bias added in front of a train-mode BN gets an exactly-zero gradient (dx sums to zero per channel)
    dims = [i for i in range(len(shape)) if i != dim]
    assert float(dx.sum(dims).abs().max()) <= 1e-10


def test_gelu_erf_derivative():
    t = _rand(1000, seed=5, scale=3.0).requires_grad_(True)
    F.gelu(t).sum().backward()
    assert _close(OB.gelu_erf_grad(t.detach()), t.grad)


@pytest.mark.parametrize("C,nhead,L,S", [(32, 4, 36, 16), (64, 8, 81, 81)])
def test_loftr_layer_backward(C, nhead, L, S):
    names = {"q_proj.weight": (C, C), "k_proj.weight": (C, C), "v_proj.weight": (C, C), "merge.weight": (C, C),
             "mlp.0.weight": (2 * C, 2 * C), "mlp.2.weight": (C, 2 * C), "norm1.weight": (C,), "norm1.bias": (C,),
             "norm2.weight": (C,), "norm2.bias": (C,)}
    p = {}
    for i, (k, shp) in enumerate(names.items()):
        t = _rand(*shp, seed=20 + i, scale=0.2)
        p[k] = (t + 1.0 if k.endswith("norm1.weight") or k.endswith("norm2.weight") else t).requires_grad_(True)
    x, src = _rand(4, L, C, seed=1).requires_grad_(True), _rand(4, S, C, seed=2).requires_grad_(True)
    dout = _rand(4, L, C, seed=3)
    O.loftr_layer(p, x, src, nhead).backward(dout)
    with torch.no_grad():
        dx, dsrc, grads = OB.loftr_layer_bwd({k: v.detach() for k, v in p.items()}, x.detach(), src.detach(), nhead, dout)
    assert _close(dx, x.grad) and _close(dsrc, src.grad)
    assert set(grads) == set(names)
    for k, g in grads.items():
        assert _close(g, p[k].grad), k


@pytest.mark.parametrize("C,k,H,W", [(32, 31, 20, 24), (64, 15, 13, 17), (128, 7, 9, 10)])
def test_lkpm_train_backward(C, k, H, W):
    """explicit LKPM backward (depthwise conv via the flipped-kernel forward conv, train-mode BN, LN, erf GELU) vs
    autograd over the train-mode forward restatement"""
    shapes = {"dwconv2.weight": (C, 1, k, k), "dwconv2.bias": (C,), "bn1.weight": (C,), "bn1.bias": (C,),
              "norm.weight": (C,), "norm.bias": (C,), "pwconv1.weight": (4 * C, C), "pwconv1.bias": (4 * C,),
              "pwconv2.weight": (C, 4 * C), "pwconv2.bias": (C,)}
    p = {}
    for i, (name, shp) in enumerate(shapes.items()):
        t = _rand(*shp, seed=40 + i, scale=0.1)
        p[name] = (t + 1.0 if name in ("bn1.weight", "norm.weight") else t).requires_grad_(True)
    p["bn1.running_mean"], p["bn1.running_var"] = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    p["bn1.num_batches_tracked"] = torch.tensor(0)
    feat0 = _rand(2, H * W, C, seed=1).requires_grad_(True)
    dout = _rand(2, H * W, C, seed=2)
    O.lkpm(p, feat0, H, W, bn_stats={}).backward(dout)
    with torch.no_grad():
        dfeat, grads = OB.lkpm_bwd({n: v.detach() for n, v in p.items()}, feat0.detach(), H, W, dout)
    assert _close(dfeat, feat0.grad)
    assert set(grads) == set(shapes)
    for name, g in grads.items():
        if name == "dwconv2.bias":           # exactly zero in front of a train-mode BN: both sides hold rounding noise
            assert float(g.abs().max()) <= 1e-9 and float(p[name].grad.abs().max()) <= 1e-9
            continue
        assert g.shape == p[name].grad.shape and _close(g, p[name].grad), name


@pytest.mark.parametrize("geom,level", [("G416z6", 3), ("G480", 3), ("G480pad", 3), ("G480pad", 2)])
def test_hist2image_backward_all_branches(geom, level):
    """explicit hist2image backward (gather / scatter adjoints, invalid-zone zeroing, and in the resize branch the two
    separable bilinear resizes as matrices) against autograd over the forward restatement: no-resize, resize, pad"""
    from cfpnet_b200 import synth
    C, _, max_res, _ = synth.LEVELS[level]
    inp = synth.make_inputs(geom, 2, seed=3, levels=(level,))
    H, W = synth.level_hw(geom, level)
    g = O.zone_geometry(inp["patch_info"], max_res[1], H, W)
    Z = g["zone_num"] ** 2
    names = {"q_proj.weight": (C, C), "k_proj.weight": (C, C), "v_proj.weight": (C, C), "merge.weight": (C, C),
             "mlp.0.weight": (2 * C, 2 * C), "mlp.2.weight": (C, 2 * C), "norm1.weight": (C,), "norm1.bias": (C,),
             "norm2.weight": (C,), "norm2.bias": (C,)}
    p = {}
    for i, (k, shp) in enumerate(names.items()):
        t = _rand(*shp, seed=60 + i, scale=0.2)
        p[k] = (t + 1.0 if k in ("norm1.weight", "norm2.weight") else t).requires_grad_(True)
    feat0 = _rand(2, H * W, C, seed=1).requires_grad_(True)
    ztok = _rand(2 * Z, 16, C, seed=2).requires_grad_(True)
    dout = _rand(2, H * W, C, seed=3)
    O.hist2image(p, feat0, feat0, ztok, inp["mask"], g, H, W).backward(dout)
    with torch.no_grad():
        dfeat, dz, grads = OB.hist2image_bwd({k: v.detach() for k, v in p.items()}, feat0.detach(), ztok.detach(),
                                             inp["mask"], g, H, W, dout)
    assert _close(dfeat, feat0.grad) and _close(dz, ztok.grad)
    for k, gr in grads.items():
        assert _close(gr, p[k].grad), k


def test_bilinear_matrix_is_the_align_corners_resize():
    x = _rand(1, 3, 28, 28, seed=1)
    Wy, Wx = OB.bilinear_matrix(28, 32), OB.bilinear_matrix(28, 32)
    want = F.interpolate(x, size=[32, 32], mode="bilinear", align_corners=True)
    assert _close(torch.einsum("oy,bcyx,px->bcop", Wy, x, Wx), want)


def test_silog_loss_and_gradient_match_reference():
    """f4 oracle step: the training loss (src/loss.py:9-19) and its closed-form gradient against the reference's value
    and autograd gradient (tests/golden/silog_loss.npz, tools/make_golden_loss.py)"""
    import os
    import numpy as np
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "silog_loss.npz"))
    pred, target, mask = (torch.from_numpy(z[k]) for k in ("pred", "target", "mask"))
    assert abs(float(O.silog_loss(pred, target, mask)) - float(z["loss"])) <= 1e-12
    assert _close(OB.silog_loss_bwd(pred, target, mask), torch.from_numpy(z["grad"]))
