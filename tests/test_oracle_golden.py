"""Pin the CPU oracle (oracle/cfp_oracle.py) against the committed outputs of the
REFERENCE modules (tests/golden, made by tools/make_golden.py from /root/reference).
CPU only; nothing here touches the CUDA library."""
import os

import numpy as np
import pytest
import torch

from cfpnet_b200 import synth
from helpers import FUSION_CASES as _BASE_CASES, FUSION_CASES_Z6, FusionCase, GOLDEN, ref_keys, rel_l2

FUSION_CASES = _BASE_CASES + FUSION_CASES_Z6
from oracle import cfp_oracle as O


def test_hist_encoder_matches_reference():
    z = np.load(os.path.join(GOLDEN, "hist_encoder_B2.npz"))
    sd = synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0)
    inp = synth.make_inputs("G416", 2, seed=1, levels=())
    for dt, tol in ((torch.float64, 1e-6), (torch.float32, 1e-5)):
        outs = O.hist_encoder(sd, inp["hist_data"].to(dt))
        for c, o in zip((32, 64, 128), outs):
            assert rel_l2(o, torch.from_numpy(z[f"out{c}"])) <= tol


@pytest.mark.parametrize("tag", FUSION_CASES)
def test_geometry_and_masks_bit_exact(tag):
    case = FusionCase(tag)
    inp = case.inputs()
    H, W = synth.level_hw(case.geometry, case.level)
    # oracle's own patch_info restatement == product host logic == what the reference produced
    for b in range(case.batch):
        pi = O.patch_info_from_rects(inp["rect_data"][b])
        for cps in (4, 8, 16):
            for k in pi[cps]:
                assert torch.equal(pi[cps][k], inp["patch_info"][cps][k][b]), (cps, k)
    g = O.zone_geometry(inp["patch_info"], case.max_res[1], H, W)
    ref = case.geo
    for mine, theirs in (("pad_h", "pad_height"), ("pad_w", "pad_width"), ("p1", "p1"), ("p2", "p2"),
                         ("sy_wo", "sy_wo_pad"), ("sx_wo", "sx_wo_pad"), ("ey_wo", "ey_wo_pad"),
                         ("ex_wo", "ex_wo_pad"), ("sy", "sy"), ("ey", "ey"), ("sx", "sx"), ("ex", "ex"),
                         ("tzh", "tzh"), ("tzw", "tzw"), ("interpolate", "interpolate"),
                         ("zone_num", "zone_num")):
        assert g[mine] == ref[theirs], (mine, g[mine], ref[theirs])
    zm, hm, pm = O.zone_masks(g, inp["mask"], case.batch, H, W, case.C)
    assert np.array_equal(zm[:, :, 0].numpy(), case.mask_bits("zone_mask"))
    assert np.array_equal(hm[:, :, 0].numpy(), case.mask_bits("hist_mask"))
    assert np.array_equal(pm.view(case.batch, g["tzh"], g["tzw"], case.C)[..., 0].numpy(), case.mask_bits("pad_mask"))
    # the channel repeat is a pure broadcast
    assert bool((zm == zm[:, :, :1]).all()) and bool((hm == hm[:, :, :1]).all())


@pytest.mark.parametrize("tag", FUSION_CASES)
def test_fusion_matches_reference(tag):
    case = FusionCase(tag)
    inp = case.inputs()
    sd, hsd = case.state_dict(), case.hist_state_dict()
    for dt, tol in ((torch.float64, 2e-6), (torch.float32, 1e-3)):
        feats = O.hist_encoder(hsd, inp["hist_data"].to(dt))
        feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[case.C]
        out = O.transformer_fusion(sd, case.layers, case.max_res, inp[f"x{case.level}"].to(dt), feat1,
                                   inp["mask"], inp["patch_info"], offsets=case.offsets,
                                   change_embedding=case.change_embedding, no_skip_inside=case.no_skip_inside)
        case.check_output(out, tol, f"oracle {dt} {tag}")


def test_oracle_draws_offsets_like_reference():
    """fusion.py:87-91: y then x from the global CPU generator, only for dims smaller than the table."""
    case = FusionCase("G416_L3_B2")
    torch.manual_seed(2)
    assert O.draw_posenc_offsets(case.max_res, 26, 34) == case.offsets
    case = FusionCase("G480_L3_B1")
    torch.manual_seed(2)
    before = torch.get_rng_state()
    assert O.draw_posenc_offsets(case.max_res, 30, 40) == (0, 0)
    assert torch.equal(before, torch.get_rng_state())      # nothing drawn


@pytest.mark.parametrize("zn", [8, 6])
def test_zone_depth_samples_match_reference(zn):
    """f3 (input side, oracle step): (mu, sigma, mask) -> 16 depth samples per zone, bit-exact in the uniform mode the
    configs use (--sample_uniform), 1e-6 in the quantile mode (erfinv vs Normal.icdf)."""
    z = np.load(os.path.join(GOLDEN, "input_side.npz"))
    hist, mask = torch.from_numpy(z[f"hist_z{zn}"]), torch.from_numpy(z[f"mask_z{zn}"])
    got = O.sample_points_from_hist(hist, mask, 16, True)
    assert got.dtype == torch.float32 and np.array_equal(got.numpy(), z[f"samples_z{zn}_uniform"])
    assert float(got[~mask].abs().max()) == 0.0
    got = O.sample_points_from_hist(hist, mask, 16, False)
    assert np.allclose(got.numpy(), z[f"samples_z{zn}_icdf"], rtol=1e-6, atol=1e-6)


def test_synthetic_histograms_follow_the_sampling_rule():
    """cfpnet_b200.synth builds hist_data as the reference's uniform sampling of (mu, sigma): even grid over mu +- 3 sigma,
    zeros for invalid zones (what the bench and every fixture feed the histogram encoder)."""
    inp = synth.make_inputs("G416", 2, seed=1, levels=())
    h, m = inp["hist_data"], inp["mask"]
    assert h.shape == (2, 64, 16) and float(h[~m].abs().max()) == 0.0
    v = h[m]
    step = v[:, 1:] - v[:, :-1]
    assert float((step - step.mean(dim=1, keepdim=True)).abs().max()) <= 1e-5            # even grid
    mu, sigma = v.mean(dim=1), (v[:, -1] - v[:, 0]) / 6.0
    assert float(mu.min()) >= 0.3 - 1e-4 and float(mu.max()) <= 4.0 + 1e-4 and float(sigma.min()) >= 0.01 - 1e-4 \
        and float(sigma.max()) <= 0.2 + 1e-4
