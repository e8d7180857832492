"""A whole ``TransformerFusion`` call in TRAIN mode on the GPU (BASELINE config 5) against the REFERENCE's own ``.train()``
forward + backward (tests/golden/train_*.npz: the reference's modules under autograd in float64,
tools/make_golden_train.py): output, gradient w.r.t. x, every parameter gradient (norm + 48 seeded probes), the set of
parameters the reference never reaches (grad must stay None), BatchNorm running buffers after the step, and the gradient
handed back to the histogram encoder (pushed through the oracle's encoder backward to the reference's grad_hist).

The CUDA path: drop-in module in ``.train()`` -> ``FusionTrainFn`` (torch.autograd.Function) -> the op sequence of
cfpnet_b200/train_seq.py on ``CudaOps`` -> cfp_tr_* kernels (fp32, as the reference trains).  Tolerance 1e-3 (rel-L2 /
relative norm) per tensor; fp32 against a float64 reference."""
import os

import numpy as np
import pytest
import torch

import cfpnet_b200
from cfpnet_b200 import synth
from cfpnet_b200.config import args
from helpers import GOLDEN, ref_keys, rel_l2
from test_oracle_train_golden import _probe_index

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-3
CASES = ["G416z6_L3_B2", "G416_L2_B2", "G416z6_L1_B1"]


def _check_map(z, key, t, what):
    """a map of the fixture: stored whole, or (above 100 k elements) as seeded samples + per-channel sums"""
    assert tuple(t.shape) == tuple(int(v) for v in z[key + "_shape"]), what
    if key in z.files:
        assert rel_l2(t, torch.from_numpy(z[key])) <= TOL, what
        return
    idx = torch.from_numpy(z[key + "_idx"])
    err = rel_l2(t.reshape(-1)[idx], torch.from_numpy(z[key + "_sample"]))
    assert err <= TOL, f"{what}: samples rel-L2 {err:.3e}"
    err = rel_l2(t.sum(dim=(2, 3)), torch.from_numpy(z[key + "_perchan"]))
    assert err <= TOL, f"{what}: per-channel sums rel-L2 {err:.3e}"


@pytest.mark.parametrize("tag", CASES)
def test_train_mode_fusion_matches_reference_gradients(tag):
    from oracle import cfp_oracle as O
    from oracle import cfp_oracle_bwd as OB
    z = np.load(os.path.join(GOLDEN, f"train_{tag}.npz"))
    geometry, level, batch, layers = [str(v) for v in z["meta"]]
    level, batch, layers = int(level), int(batch), tuple(layers.split(","))
    C, _, max_res, lk = synth.LEVELS[level]
    saved_layers = list(args.attention_layer)
    try:
        args.attention_layer = list(layers)
        mod = cfpnet_b200.TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1])
    finally:
        args.attention_layer = saved_layers
    mod.load_state_dict(synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{level}"], seed=level), strict=True)
    mod = mod.to(DEV).train()
    hsd = {k: (v.double() if v.is_floating_point() else v) for k, v in synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0).items()}
    inp = synth.make_inputs(geometry, batch, seed=1, levels=(level,))
    with torch.no_grad():
        feats = O.hist_encoder(hsd, inp["hist_data"].double(), bn_stats={})      # train-mode encoder (test infrastructure)
    feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[C].float().to(DEV).requires_grad_(True)
    x = inp[f"x{level}"].float().to(DEV).requires_grad_(True)
    torch.manual_seed(2)                                                          # the crop draws of the fixture
    out = mod(x, feat1, mask=inp["mask"].to(DEV), patch_info=inp["patch_info"], rect_data=inp["rect_data"], rgb=None)
    ct = torch.randn(out.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float64)
    out.backward(ct.float().to(DEV))
    torch.cuda.synchronize()

    _check_map(z, "out", out.detach().double().cpu(), tag + " out")
    _check_map(z, "grad_x", x.grad.double().cpu(), tag + " grad_x")
    named = dict(mod.named_parameters())
    checked, worst = 0, (0.0, "")
    names = [str(n) for n in z["param_names"]]
    for i, full in enumerate(names):
        scope, name = full.split(".", 1)
        if scope != "fusion":
            continue
        p = named[name]
        if not bool(z["param_has_grad"][i]):
            assert p.grad is None, f"{full}: never used by the reference, must keep grad None"
            continue
        assert p.grad is not None, f"{full}: received no gradient"
        got = p.grad.double().cpu()
        norm = float(z["param_grad_norm"][i])
        if norm < 1e-9 * max(1.0, got.numel() ** 0.5):
            # a conv bias in front of a batch-statistics BatchNorm has an exactly-zero gradient (the reference's autograd
            # leaves 1e-12 rounding noise there): hold the fp32 sum against the scale of the layer's weight gradient
            sib = float(z["param_grad_norm"][names.index(full.rsplit(".", 1)[0] + ".weight")])
            assert float(got.norm()) <= TOL * sib + 1e-4, f"{full} should be ~0, got |g| = {float(got.norm()):.3e} (weight gradient {sib:.3e})"
            checked += 1
            continue
        assert abs(float(got.norm()) - norm) <= TOL * norm + 1e-7, (full, float(got.norm()), norm)
        probe = got.reshape(-1)[_probe_index(full, got.numel())]
        want = torch.from_numpy(z["param_grad_probe"][i])
        # 48 probes of a tensor: compare on the tensor's scale (|g| / sqrt(numel) per element), not on the probes' own
        err = float((probe - want).norm()) / max(float(want.norm()), norm * (48.0 / got.numel()) ** 0.5, 1e-30)
        worst = max(worst, (err, full))
        assert err <= 5 * TOL, f"{full}: probe error {err:.3e}"
        checked += 1
    assert checked >= 100, checked
    for bname in (str(b) for b in z["buffer_names"]):
        scope, name = bname.split(".", 1)
        if scope != "fusion":
            continue
        want = torch.from_numpy(np.asarray(z["buf:" + bname]))
        got = dict(mod.named_buffers())[name]
        if want.dtype == torch.int64:
            assert int(got) == int(want), bname
        else:
            assert rel_l2(got.double().cpu(), want) <= TOL, f"buffer {bname}: {rel_l2(got.double().cpu(), want):.3e}"
    douts = [None, None, None]
    douts[{32: 0, 64: 1, 128: 2}[C]] = feat1.grad.double().cpu()
    dhist, _ = OB.hist_encoder_bwd(hsd, inp["hist_data"].double(), douts)
    assert rel_l2(dhist, torch.from_numpy(z["grad_hist"])) <= TOL
    print(f"{tag}: {checked} parameter gradients within tolerance, worst probe error {worst[0]:.2e} ({worst[1]})")


def test_train_mode_refuses_what_it_does_not_serve():
    mod = cfpnet_b200.TransformerFusion(128, [30, 40], large_kernel=7, patch_size=16).to(DEV).train()
    inp = synth.make_inputs("G480", 1, levels=(3,))                               # 480x640: the bilinear-resize branch
    with pytest.raises(NotImplementedError):
        mod(inp["x3"].to(DEV), torch.zeros(1, 64, 16, 128, device=DEV), mask=inp["mask"].to(DEV), patch_info=inp["patch_info"],
            rect_data=inp["rect_data"], rgb=None)
