"""tests/golden/silog_loss.npz from the REFERENCE's ``SILogLoss`` (src/loss.py, container only): loss value and the
gradient w.r.t. the prediction for a half-resolution prediction, a masked full-resolution target."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ref_import import import_reference  # noqa: E402

import_reference()
from src.loss import SILogLoss  # noqa: E402

g = torch.Generator().manual_seed(0)
pred = (torch.rand(2, 1, 52, 68, generator=g, dtype=torch.float64) * 4 + 0.3).requires_grad_(True)
target = torch.rand(2, 1, 104, 136, generator=g, dtype=torch.float64) * 4 + 0.3
mask = torch.rand(2, 1, 104, 136, generator=g) < 0.7
loss = SILogLoss()(pred, target, mask=mask, interpolate=True)
loss.backward()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "silog_loss.npz"), pred=pred.detach().numpy(), target=target.numpy(),
                    mask=mask.numpy(), loss=loss.detach().numpy(), grad=pred.grad.numpy())
print(float(loss.detach()), float(pred.grad.norm()))
