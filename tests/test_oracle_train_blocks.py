"""The oracle's closed-form backward of the histogram encoder and LKPM (oracle/cfp_oracle_bwd.py: hist_encoder_bwd,
lkpm_bwd - the formulas cfpnet_b200/train.py sequences as kernels) pinned on the REFERENCE's own train-mode gradients
(tests/golden/trainblk_*.npz, generated from /root/reference by tools/make_golden_train_blocks.py)."""
import os

import numpy as np
import pytest
import torch

from cfpnet_b200 import synth
from helpers import GOLDEN, rel_l2
from oracle import cfp_oracle_bwd as OB


def test_hist_encoder_bwd_matches_reference_gradients():
    z = np.load(os.path.join(GOLDEN, "trainblk_hist_B2.npz"))
    names = [str(n) for n in z["param_names"]]
    shapes = {n: z["grad:" + n].shape for n in names}
    for n in z.files:
        if n.startswith("buf:"):
            shapes[n[4:]] = np.asarray(z[n]).shape
    sd = {k: v.double() if v.is_floating_point() else v for k, v in synth.synthetic_state_dict(shapes, seed=0).items()}
    inp = synth.make_inputs("G416", 2, seed=1, levels=())
    g = torch.Generator().manual_seed(77)
    cts = [torch.randn(2, 64, 16, c, generator=g, dtype=torch.float64) for c in (32, 64, 128)]
    dhist, grads = OB.hist_encoder_bwd(sd, inp["hist_data"].double(), cts)
    assert rel_l2(dhist, torch.from_numpy(z["grad_hist"])) <= 1e-5
    for n in names:
        want = torch.from_numpy(z["grad:" + n])
        if float(want.norm()) < 1e-9:
            assert float(grads[n].norm()) < 1e-6
        else:
            assert rel_l2(grads[n], want) <= 1e-5, n


@pytest.mark.parametrize("level,batch", [(3, 2), (2, 2), (1, 1)])
def test_lkpm_bwd_matches_reference_gradients(level, batch):
    z = np.load(os.path.join(GOLDEN, f"trainblk_lkpm_L{level}_B{batch}.npz"))
    C, _, _, k = synth.LEVELS[level]
    H, W = synth.level_hw("G416", level)
    names = [str(n) for n in z["param_names"]]
    has = dict(zip(names, [bool(v) for v in z["param_has_grad"]]))
    shapes = {n: (z["grad:" + n].shape if has[n] else (C, 2 * C, 1, 1)) for n in names}
    for n in z.files:
        if n.startswith("buf:"):
            shapes[n[4:]] = np.asarray(z[n]).shape
    sd = {kk: v.double() if v.is_floating_point() else v for kk, v in synth.synthetic_state_dict(shapes, seed=3).items()}
    x = torch.randn(batch, C, H, W, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    ct = torch.randn(batch, C, H, W, generator=torch.Generator().manual_seed(78), dtype=torch.float64)
    tok = lambda t: t.reshape(batch, C, H * W).transpose(1, 2).contiguous()      # noqa: E731
    dx, grads = OB.lkpm_bwd(sd, tok(x), H, W, tok(ct))
    dx = dx.transpose(1, 2).reshape(batch, C, H, W)
    if "grad_x" in z.files:
        assert rel_l2(dx, torch.from_numpy(z["grad_x"])) <= 1e-5
    else:
        assert rel_l2(dx.reshape(-1)[torch.from_numpy(z["grad_x_idx"])], torch.from_numpy(z["grad_x_sample"])) <= 1e-5
    for n in names:
        if not has[n]:
            assert n not in grads
            continue
        want = torch.from_numpy(z["grad:" + n])
        if float(want.norm()) < 1e-9:
            assert float(grads[n].norm()) < 1e-6
        else:
            assert rel_l2(grads[n].reshape(want.shape), want) <= 1e-5, n
