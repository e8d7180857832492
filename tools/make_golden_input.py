"""tests/golden/input_side.npz from the REFERENCE's ``sample_point_from_hist_parallel`` (container only;
src/utils/dataloader.py:65-81): (mu, sigma) per zone + validity mask -> the 16 depth samples the histogram encoder
consumes, for both sampling modes and both zone grids.  Pins ``oracle.cfp_oracle.sample_points_from_hist``."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ref_import import import_reference  # noqa: E402

ref = import_reference()
fn = ref["dataloader"].sample_point_from_hist_parallel
rec = {}
for zn in (8, 6):
    g = torch.Generator().manual_seed(zn)
    Z = zn * zn
    hist = torch.stack([torch.rand(Z, generator=g) * 3.7 + 0.3, torch.rand(Z, generator=g) * 0.19 + 0.01], dim=1)
    mask = torch.rand(Z, generator=g) < 0.8
    rec[f"hist_z{zn}"], rec[f"mask_z{zn}"] = hist.numpy(), mask.numpy()
    for uniform in (True, False):
        cfg = types.SimpleNamespace(zone_sample_num=16, sample_uniform=uniform)
        rec[f"samples_z{zn}_{'uniform' if uniform else 'icdf'}"] = fn(hist, mask, cfg).numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "input_side.npz"), **rec)
print({k: v.shape for k, v in rec.items()})
