"""Training row of the scope table, oracle step: pin the TRAIN-mode restatement
(BatchNorm batch statistics + autograd over oracle/cfp_oracle.py) on one forward + backward of the REFERENCE modules
in ``.train()`` mode (tests/golden/train_*.npz, made by tools/make_golden_train.py from /root/reference).
CPU only.  No backward CUDA kernel exists yet - this is the checker they will be held to."""
import os

import numpy as np
import pytest
import torch

from cfpnet_b200 import synth
from helpers import GOLDEN, ref_keys, rel_l2
from oracle import cfp_oracle as O

CASES = ["G416z6_L3_B2", "G416_L2_B2", "G416z6_L1_B1", "G480_L3_B1"]
TOL = 2e-6          # float64 restatement vs float32 copy of the float64 reference run
RTOL, ATOL = 1e-8, 1e-9     # parameter gradients are stored in float64: |diff| <= RTOL * |grad| + ATOL


def _probe_index(name, numel, n=48):       # same seeded positions as tools/make_golden_train.py
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000000007
    return torch.randint(0, numel, (n,), generator=torch.Generator().manual_seed(h % (2 ** 31)))


def _leaves(sd):
    """float parameters -> float64 leaves that record gradients; buffers stay plain tensors"""
    out = {}
    for k, v in sd.items():
        is_buf = k.endswith(("running_mean", "running_var", "num_batches_tracked"))
        out[k] = v.double().requires_grad_(True) if (v.is_floating_point() and not is_buf) else v
    return out


def _check_map(z, key, t, what):
    assert tuple(t.shape) == tuple(int(v) for v in z[key + "_shape"]), what
    if key in z.files:
        assert rel_l2(t, torch.from_numpy(z[key])) <= TOL, what
        return
    idx = torch.from_numpy(z[key + "_idx"])
    assert rel_l2(t.reshape(-1)[idx], torch.from_numpy(z[key + "_sample"])) <= TOL, what
    assert rel_l2(t.sum(dim=(2, 3)), torch.from_numpy(z[key + "_perchan"])) <= TOL, what


@pytest.fixture(scope="module", params=CASES)
def step(request):
    tag = request.param
    z = np.load(os.path.join(GOLDEN, f"train_{tag}.npz"))
    geometry, level, batch, layers = [str(v) for v in z["meta"]]
    level, batch, layers = int(level), int(batch), tuple(layers.split(","))
    C, _, max_res, _ = synth.LEVELS[level]
    sd = _leaves(synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{level}"], seed=level))
    hsd = _leaves(synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0))
    inp = synth.make_inputs(geometry, batch, seed=1, levels=(level,))
    x = inp[f"x{level}"].double().requires_grad_(True)
    hist = inp["hist_data"].double().requires_grad_(True)
    stats, hstats = {}, {}
    feats = O.hist_encoder(hsd, hist, bn_stats=hstats)
    feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[C]
    torch.manual_seed(2)                    # the reference run drew its positional-encoding crop after this seed
    out = O.transformer_fusion(sd, layers, max_res, x, feat1, inp["mask"], inp["patch_info"], bn_stats=stats)
    ct = torch.randn(out.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float64)
    (out * ct).sum().backward()
    return dict(tag=tag, z=z, out=out.detach(), x=x, hist=hist, sd=sd, hsd=hsd, stats=stats, hstats=hstats)


def test_train_forward_matches_reference(step):
    _check_map(step["z"], "out", step["out"], step["tag"])


def test_input_gradients_match_reference(step):
    _check_map(step["z"], "grad_x", step["x"].grad, step["tag"] + " grad_x")
    assert rel_l2(step["hist"].grad, torch.from_numpy(step["z"]["grad_hist"])) <= TOL


def test_parameter_gradients_match_reference(step):
    z = step["z"]
    for i, full in enumerate(str(n) for n in z["param_names"]):
        scope, name = full.split(".", 1)
        p = (step["sd"] if scope == "fusion" else step["hsd"])[name]
        if not bool(z["param_has_grad"][i]):
            # registered but never used by the reference's forward: no gradient, stays out of the allreduce
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, full
            continue
        assert p.grad is not None, full
        norm = float(z["param_grad_norm"][i])
        # a bias in front of a train-mode BatchNorm (dwconv2.bias, the encoder's conv biases) has an exactly-zero
        # gradient: both sides hold rounding noise of ~1e-12 there, hence the absolute term
        tol = RTOL * norm + ATOL
        assert abs(float(p.grad.norm()) - norm) <= tol, full
        assert abs(float(p.grad.sum()) - float(z["param_grad_sum"][i])) <= tol * p.numel() ** 0.5, full
        probe = p.grad.reshape(-1)[_probe_index(full, p.numel())]
        err = float((probe - torch.from_numpy(z["param_grad_probe"][i])).norm())
        assert err <= tol, (full, err, tol)
    assert int(z["param_has_grad"].sum()) >= 100


def test_unused_parameters_are_the_ones_the_reference_never_touches(step):
    """the never-used tensors (DAPM's merge / mlp / norms, LKPM's conv1, at L2 also the third
    histogram extractor) get no gradient; everything on the path does."""
    z = step["z"]
    unused = {str(n) for n, h in zip(z["param_names"], z["param_has_grad"]) if not h}
    assert any("transformer_path.merge" in n for n in unused)
    assert any("large_kernel_path.conv1" in n for n in unused)
    assert not any(".q_proj." in n or "dwconv2" in n or "pwconv" in n for n in unused)
    if step["tag"].startswith("G416_L2"):
        assert any(n.startswith("hist.hist_extractor3.") for n in unused)


def test_batchnorm_buffers_after_the_step_match_reference(step):
    z = step["z"]
    touched = 0
    for full in (str(n) for n in z["buffer_names"]):
        scope, name = full.split(".", 1)
        sd, stats = (step["sd"], step["stats"]) if scope == "fusion" else (step["hsd"], step["hstats"])
        want = torch.from_numpy(np.asarray(z["buf:" + full]))
        got = stats.get(name, sd[name])
        touched += name in stats
        if want.is_floating_point():
            assert rel_l2(got, want) <= 1e-12, full
        else:
            assert int(got) == int(want), full
    assert touched >= 3 * 6            # at least DAPM bn1/bn2 + LKPM bn1 of both combine1 layers... and the encoder's


def test_explicit_backward_chain_matches_reference(step):
    """The closed-form backward of the WHOLE call (oracle/cfp_oracle_bwd.py: posenc, hist2image, DAPM, LKPM, LSA, GSA,
    histogram encoder - no autograd anywhere) against the gradients the REFERENCE computed in .train() mode."""
    from oracle import cfp_oracle_bwd as OB
    z = step["z"]
    geometry, level, batch, layers = [str(v) for v in z["meta"]]
    level, batch, layers = int(level), int(batch), tuple(layers.split(","))
    C, _, max_res, _ = synth.LEVELS[level]
    sd = {k: (v.double() if v.is_floating_point() else v)
          for k, v in synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{level}"], seed=level).items()}
    hsd = {k: (v.double() if v.is_floating_point() else v)
           for k, v in synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0).items()}
    inp = synth.make_inputs(geometry, batch, seed=1, levels=(level,))
    x, hist = inp[f"x{level}"].double(), inp["hist_data"].double()
    H, W = x.shape[2], x.shape[3]
    torch.manual_seed(2)
    offsets = O.draw_posenc_offsets(max_res, H, W)
    ct = torch.randn(x.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float64)
    with torch.no_grad():
        feats = O.hist_encoder(hsd, hist, bn_stats={})
        slot = {32: 0, 64: 1, 128: 2}[C]
        dx, dfeat1, grads = OB.transformer_fusion_bwd(sd, layers, max_res, x, feats[slot], inp["mask"], inp["patch_info"],
                                                      offsets, ct)
        douts = [None, None, None]
        douts[slot] = dfeat1
        dhist, hgrads = OB.hist_encoder_bwd(hsd, hist, douts)
    _check_map(z, "grad_x", dx, step["tag"] + " explicit grad_x")
    assert rel_l2(dhist, torch.from_numpy(z["grad_hist"])) <= TOL
    checked = 0
    for i, full in enumerate(str(n) for n in z["param_names"]):
        scope, name = full.split(".", 1)
        got = (grads if scope == "fusion" else hgrads).get(name)
        if not bool(z["param_has_grad"][i]):
            assert got is None, full                         # never used by the forward: no gradient is produced
            continue
        assert got is not None, full
        norm = float(z["param_grad_norm"][i])
        tol = RTOL * norm + ATOL
        assert abs(float(got.norm()) - norm) <= tol, full
        probe = got.reshape(-1)[_probe_index(full, got.numel())]
        assert float((probe - torch.from_numpy(z["param_grad_probe"][i])).norm()) <= tol, full
        checked += 1
    assert checked == int(z["param_has_grad"].sum())
