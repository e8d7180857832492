"""tests/golden/zone_hist.npz and metrics.npz from the REFERENCE (container only): ``get_hist_parallel``
(src/utils/dataloader.py:84-134) on synthetic depth maps for the evaluation geometry (8x8 zones of 56 px at 480x640) and
the training geometry (6x6 zones of 64 px at 416x544), with the fixed and with a drawn maximum distance; and
``compute_errors`` (src/utils/metrics.py:4-24).  Pins oracle.cfp_oracle.zone_hist_params / depth_metrics and, through
the fixtures, the CUDA kernels cfp_zone_hist / cfp_depth_metrics."""
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
fn = ref["dataloader"].get_hist_parallel
from src.utils.metrics import compute_errors  # noqa: E402


from cfpnet_b200.synth import synthetic_depth_map as depth_map  # noqa: E402

rec = {}
cases = [("eval480", 480, 640, "online_eval", 8, False), ("train416", 416, 544, "train", 6, False),
         ("eval480_rand", 480, 640, "online_eval", 8, True)]
for name, h, w, mode, zn, rand in cases:
    dep = depth_map(h, w, len(name))
    cfg = types.SimpleNamespace(mode=mode, train_zone_num=zn, train_zone_random_offset=0, simu_max_distance=4.0,
                                random_simu_max_d=rand, simu_max_d=4.0, simu_min_d=3.0)
    np.random.seed(5)
    fh, fr, mask = fn(torch.zeros(3, h, w), dep.unsqueeze(0), cfg)
    np.random.seed(5)
    maxd = float(np.random.uniform(low=3.0, high=4.0, size=1)[0]) if rand else 4.0
    rec[f"{name}_fh"], rec[f"{name}_fr"], rec[f"{name}_mask"] = fh.numpy(), fr.numpy(), mask.numpy()      # the depth map is synth.synthetic_depth_map(h, w, len(name))
    rec[f"{name}_maxd"] = np.float64(maxd)
    print(name, "valid zones", int(mask.sum()), "of", mask.numel(), "max_distance", maxd)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "zone_hist.npz"), **rec)

g = torch.Generator().manual_seed(3)
gt = (torch.rand(40000, generator=g) * 8 + 0.2).numpy().astype(np.float32)
pred = (gt * np.exp(0.25 * torch.randn(40000, generator=g).numpy())).astype(np.float32)
valid = (torch.rand(40000, generator=g) < 0.7).numpy()
m = compute_errors(gt[valid].astype(np.float64), pred[valid].astype(np.float64))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), gt=gt, pred=pred, valid=valid,
                    **{"m_" + k: np.float64(v) for k, v in m.items()})
print({k: float(v) for k, v in m.items()})
