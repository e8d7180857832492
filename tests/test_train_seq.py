"""The op sequencing of the training step (cfpnet_b200/train_seq.py) run with plain torch ops (tests/torch_ops.py, float64,
CPU) against the REFERENCE's own ``.train()`` forward + backward of a whole TransformerFusion call
(tests/golden/train_*.npz): output, input gradients and every parameter gradient.  What this pins is the ORDER and the
index vectors - the CUDA primitives themselves are held to torch ops one by one in tests/test_gpu_train.py."""
import os

import numpy as np
import pytest
import torch

import cfpnet_b200
from cfpnet_b200 import synth, train_seq as TS
from cfpnet_b200.config import args
from cfpnet_b200.geometry import zone_geometry
from helpers import GOLDEN, ref_keys, rel_l2
from test_oracle_train_golden import _check_map, _probe_index
from torch_ops import TorchOps

CASES = ["G416z6_L3_B2", "G416_L2_B2", "G416z6_L1_B1"]            # (the 480x640 case is the resize branch: not served in training)


@pytest.mark.parametrize("tag", CASES)
def test_sequencing_matches_reference_gradients(tag):
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
    sd = synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{level}"], seed=level)
    mod.load_state_dict(sd, strict=True)
    mod = mod.double()
    P = {k: v.detach() for k, v in mod.state_dict().items()}
    from oracle import cfp_oracle as O
    hsd = {k: (v.double() if v.is_floating_point() else v) for k, v in synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0).items()}
    inp = synth.make_inputs(geometry, batch, seed=1, levels=(level,))
    x = inp[f"x{level}"].double()
    with torch.no_grad():
        feats = O.hist_encoder(hsd, inp["hist_data"].double(), bn_stats={})
    feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[C]
    B, _, H, W = x.shape
    g = zone_geometry(inp["patch_info"], max_res[1], H, W)
    ix = TS.Indexer(B, H, W, g, mod.ws, "cpu")
    torch.manual_seed(2)
    oy, ox = mod.draw_crop(H, W)
    ops = TorchOps()
    zmask = inp["mask"].reshape(-1).double()
    with torch.no_grad():
        out, saved = TS.fusion_fwd(ops, mod, P, x, feat1, zmask, ix, oy, ox)
        ct = torch.randn(out.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float64)
        dx, dfeat1, grads = TS.fusion_bwd(ops, mod, P, saved, zmask, ix, oy, ox, ct, tuple(feat1.shape))
    _check_map(z, "out", out, tag)
    _check_map(z, "grad_x", dx, tag + " grad_x")
    checked = 0
    for i, full in enumerate(str(n) for n in z["param_names"]):
        scope, name = full.split(".", 1)
        if scope != "fusion":
            continue
        got = grads.get(name)
        if not bool(z["param_has_grad"][i]):
            assert got is None, full
            continue
        assert got is not None, full
        norm = float(z["param_grad_norm"][i])
        tol = 1e-8 * norm + 1e-9
        assert abs(float(got.norm()) - norm) <= tol, (full, float(got.norm()), norm)
        probe = got.reshape(-1)[_probe_index(full, got.numel())]
        assert float((probe - torch.from_numpy(z["param_grad_probe"][i])).norm()) <= tol, full
        checked += 1
    assert checked >= 100
    # the gradient handed back to the histogram encoder: through the oracle's encoder backward it must give the reference's grad_hist
    from oracle import cfp_oracle_bwd as OB
    douts = [None, None, None]
    douts[{32: 0, 64: 1, 128: 2}[C]] = dfeat1
    dhist, _ = OB.hist_encoder_bwd(hsd, inp["hist_data"].double(), douts)
    assert rel_l2(dhist, torch.from_numpy(z["grad_hist"])) <= 2e-6
