"""Input side (f3) and loss / metrics (f4) kernels against the REFERENCE's outputs (fixtures generated from
/root/reference: tests/golden/input_side.npz, zone_hist.npz, silog_loss.npz, metrics.npz).  Integer results (histogram
counts via the validity mask, the zone rectangles) and the uniform samples are bit-exact; float results within 1e-6
(float64 moments cast to fp32) / 1e-5 (loss, gradient, metrics: fp32 logs against the float64 reference)."""
import os
import types

import numpy as np
import pytest
import torch

from cfpnet_b200 import inputs, loss, synth
from helpers import GOLDEN, rel_l2
from oracle import cfp_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("zn", [8, 6])
def test_zone_samples_bit_exact(zn):
    z = np.load(os.path.join(GOLDEN, "input_side.npz"))
    hist = torch.from_numpy(z[f"hist_z{zn}"]).to(DEV)
    mask = torch.from_numpy(z[f"mask_z{zn}"]).to(DEV)
    for uniform in (True, False):
        cfg = types.SimpleNamespace(zone_sample_num=16, sample_uniform=uniform)
        # batched: the same zones twice, second copy with every zone invalid
        h2 = torch.stack([hist, hist])
        m2 = torch.stack([mask, torch.zeros_like(mask)])
        got = inputs.sample_point_from_hist_parallel(h2, m2, cfg).cpu()
        want = torch.from_numpy(z[f"samples_z{zn}_{'uniform' if uniform else 'icdf'}"])
        assert got.shape == (2, zn * zn, 16)
        assert torch.equal(got[1], torch.zeros_like(want))
        if uniform:
            assert torch.equal(got[0], want), "uniform samples must be bit-identical to the reference's"
        else:
            assert torch.allclose(got[0], want, rtol=2e-7, atol=0), float((got[0] - want).abs().max())


CASES = [("eval480", 480, 640, "online_eval", 8, False), ("train416", 416, 544, "train", 6, False),
         ("eval480_rand", 480, 640, "online_eval", 8, True)]


@pytest.mark.parametrize("name,h,w,mode,zn,rand", CASES)
def test_zone_hist_matches_reference(name, h, w, mode, zn, rand):
    z = np.load(os.path.join(GOLDEN, "zone_hist.npz"))
    dep = synth.synthetic_depth_map(h, w, len(name))
    cfg = types.SimpleNamespace(mode=mode, train_zone_num=zn, train_zone_random_offset=0, simu_max_distance=4.0,
                                random_simu_max_d=rand, simu_max_d=4.0, simu_min_d=3.0)
    np.random.seed(5)                       # the reference draws its maximum distance from numpy's global generator
    batch = torch.stack([dep, dep.flip(1)]).unsqueeze(1).to(DEV)           # frame 1: mirrored (a different frame in the batch)
    fh, fr, mask, hist = inputs.get_hist_parallel(torch.zeros(2, 3, h, w), batch, cfg, return_hist=True)
    assert np.array_equal(mask[0].cpu().numpy(), z[f"{name}_mask"])
    assert np.array_equal(fr.numpy(), z[f"{name}_fr"])
    want = torch.from_numpy(z[f"{name}_fh"]).float()
    assert torch.allclose(fh[0].cpu(), want, rtol=1e-6, atol=1e-12), float((fh[0].cpu() - want).abs().max())
    # integer histogram of both frames against the oracle's (torch.histc on the CPU), bit-exact
    p = 64 if mode == "train" else 56
    sy, sx = int((h - p * zn) / 2), int((w - p * zn) / 2)
    for b in range(2):
        _, omask, ohist = O.zone_hist_params(batch[b, 0].cpu(), sy, sx, p, p, zn, float(z[f"{name}_maxd"]))
        assert torch.equal(hist[b].cpu().float(), ohist), f"frame {b}: histogram counts differ"
        assert torch.equal(mask[b].cpu(), omask)


def test_zone_hist_bin_edges_bit_exact():
    """Depths exactly on / next to the 4 cm bin edges, for a maximum distance that is not a power of two: the bin index
    must be torch.histc's."""
    maxd = 3.37
    bins = int(maxd / 0.04)
    edges = torch.arange(0, bins + 1, dtype=torch.float64) * (maxd / bins)
    vals = torch.cat([(edges + d).float() for d in (-1e-7, 0.0, 1e-7, 3e-7)])
    # every value fills one 56-px row segment of one zone (56 copies: 36 survive the "- 20" of dataloader.py:111)
    dep = torch.full((56 * 8, 56 * 8), 1.0)
    dep.view(-1)[: vals.numel() * 56] = vals.repeat_interleave(56)
    cfg = types.SimpleNamespace(mode="online_eval", train_zone_num=8, train_zone_random_offset=0, simu_max_distance=maxd,
                                random_simu_max_d=False)
    big = dep.unsqueeze(0).unsqueeze(0).to(DEV)
    _, _, _, hist = inputs.get_hist_parallel(torch.zeros(1, 3, 448, 448), big, cfg, return_hist=True)
    _, _, ohist = O.zone_hist_params(dep, 0, 0, 56, 56, 8, maxd)
    assert torch.equal(hist[0].cpu().float(), ohist)


def test_silog_loss_and_gradient_match_reference():
    z = np.load(os.path.join(GOLDEN, "silog_loss.npz"))
    pred = torch.from_numpy(z["pred"]).float().to(DEV).requires_grad_(True)
    target = torch.from_numpy(z["target"]).float().to(DEV)
    mask = torch.from_numpy(z["mask"]).to(DEV)
    crit = loss.SILogLoss()
    val = crit(pred, target, mask=mask, interpolate=True)
    val.backward()
    torch.cuda.synchronize()
    assert abs(float(val) - float(z["loss"])) <= 1e-5 * float(z["loss"])
    assert rel_l2(pred.grad, torch.from_numpy(z["grad"])) <= 1e-4
    # same-size branch (interpolate=False) against the oracle restatement (fp64 autograd)
    p2 = (torch.rand(2, 1, 40, 50, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 3 + 0.4).requires_grad_(True)
    t2 = torch.rand(2, 1, 40, 50, dtype=torch.float64, generator=torch.Generator().manual_seed(2)) * 3 + 0.4
    want = O.silog_loss(p2, t2, None, interpolate=False)
    want.backward()
    p2c = p2.detach().float().to(DEV).requires_grad_(True)
    got = crit(p2c, t2.float().to(DEV), mask=None, interpolate=False)
    got.backward()
    assert abs(float(got) - float(want)) <= 1e-5 * float(want)
    assert rel_l2(p2c.grad, p2.grad) <= 1e-4


def test_depth_metrics_match_reference():
    z = np.load(os.path.join(GOLDEN, "metrics.npz"))
    m = loss.compute_errors(torch.from_numpy(z["gt"]).to(DEV), torch.from_numpy(z["pred"]).to(DEV), torch.from_numpy(z["valid"]).to(DEV))
    assert m["n"] == int(z["valid"].sum())
    for k in loss.METRIC_NAMES:
        assert abs(m[k] - float(z["m_" + k])) <= 1e-6 * max(1.0, abs(float(z["m_" + k]))), (k, m[k], float(z["m_" + k]))


def test_io_entry_points_refuse_cpu_tensors():
    from cfpnet_b200 import _lib
    cfg = types.SimpleNamespace(zone_sample_num=16, sample_uniform=True)
    with pytest.raises(_lib.CfpError, match="no CPU implementation"):
        inputs.sample_point_from_hist_parallel(torch.zeros(4, 2), torch.ones(4, dtype=torch.bool), cfg)
    with pytest.raises(_lib.CfpError, match="no CPU implementation"):
        loss.SILogLoss()(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 8, 8))
