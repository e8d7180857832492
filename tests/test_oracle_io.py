"""f3 / f4 oracle steps pinned on the REFERENCE's own outputs (tests/golden/zone_hist.npz, metrics.npz - generated from
/root/reference by tools/make_golden_io.py): get_hist_parallel and compute_errors."""
import os

import numpy as np
import torch

from cfpnet_b200 import synth
from helpers import GOLDEN
from oracle import cfp_oracle as O

CASES = [("eval480", 480, 640, 56, 8), ("train416", 416, 544, 64, 6), ("eval480_rand", 480, 640, 56, 8)]


def test_zone_hist_params_match_reference():
    z = np.load(os.path.join(GOLDEN, "zone_hist.npz"))
    for name, h, w, p, zn in CASES:
        dep = synth.synthetic_depth_map(h, w, len(name))
        sy, sx = int((h - p * zn) / 2), int((w - p * zn) / 2)
        fh, mask, _ = O.zone_hist_params(dep, sy, sx, p, p, zn, float(z[f"{name}_maxd"]))
        assert np.array_equal(mask.numpy(), z[f"{name}_mask"]), name
        assert np.allclose(fh.numpy(), z[f"{name}_fh"], rtol=1e-12, atol=1e-15), name
        fr = z[f"{name}_fr"]
        assert fr[0].tolist() == [sy, sx, sy + p, sx + p] and fr[-1].tolist() == [sy + (zn - 1) * p, sx + (zn - 1) * p, sy + zn * p, sx + zn * p]
        assert 0 < int(mask.sum()) < mask.numel(), "the fixture must contain valid and invalid zones"


def test_depth_metrics_match_reference():
    z = np.load(os.path.join(GOLDEN, "metrics.npz"))
    v = torch.from_numpy(z["valid"])
    m = O.depth_metrics(torch.from_numpy(z["gt"])[v], torch.from_numpy(z["pred"])[v])
    for k, val in m.items():
        assert abs(val - float(z["m_" + k])) <= 1e-12 * max(1.0, abs(val)), k
