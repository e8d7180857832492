"""Shared test plumbing: golden fixtures, synthetic cases, error metrics."""
import json
import os

import numpy as np
import torch

from cfpnet_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FUSION_CASES = [
    "G416_L3_B2", "G416_L2_B1", "G416_L1_B1", "G480_L3_B1", "G480pad_L3_B1", "G480pad_L2_B1",
    "G416_L3_B1_baseline", "G416_L3_B1_noskip", "G416_L3_B1_keepemb",
    # round 2: every shape a bench configuration runs (BASELINE.json configs[1] / [3]) at every level it touches
    "G480_L2_B1", "G480_L1_B1", "G480pad_L1_B1", "G416_L2_B1_baseline", "G416_L1_B1_baseline", "G416_L3_B16_baseline",
]
# 6x6 zones of 64 px - the reference's training layout (--train_zone_num 6).  Fixtures from the reference; oracle,
# geometry and (since the workspace-size fix, profiles/r1t_z6_gpu_check.log) the CUDA path are held to them.
FUSION_CASES_Z6 = ["G416z6_L3_B2", "G416z6_L2_B1", "G416z6_L1_B1"]


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def ref_keys():
    with open(os.path.join(GOLDEN, "state_dict_keys.json")) as fh:
        return json.load(fh)


class FusionCase:
    """One golden fusion fixture + everything needed to re-run it."""

    def __init__(self, tag):
        z = np.load(os.path.join(GOLDEN, f"fusion_{tag}.npz"))
        self.tag, self.z = tag, z
        geometry, level, batch, layers, chg, nsk = [str(v) for v in z["meta"]]
        self.geometry, self.level, self.batch = geometry, int(level), int(batch)
        self.layers = tuple(layers.split(","))
        self.change_embedding, self.no_skip_inside = bool(int(chg)), bool(int(nsk))
        self.C, _, self.max_res, self.large_kernel = synth.LEVELS[self.level]
        self.geo = dict(zip([str(k) for k in z["geo_keys"]], [int(v) for v in z["geo_vals"]]))
        self.offsets = (self.geo["offset_y"], self.geo["offset_x"])
        kind = "baseline" if self.layers == synth.BASELINE_LAYERS else "combine1"
        self.shapes = ref_keys()[f"fusion_{kind}_L{self.level}"]
        self.hist_shapes = ref_keys()["hist_encoder"]

    def inputs(self):
        return synth.make_inputs(self.geometry, self.batch, seed=1, levels=(self.level,))

    def state_dict(self):
        return synth.synthetic_state_dict(self.shapes, seed=self.level)

    def hist_state_dict(self):
        return synth.synthetic_state_dict(self.hist_shapes, seed=0)

    def mask_bits(self, name):
        shape = tuple(int(v) for v in self.z[f"{name}_shape"])
        n = int(np.prod(shape))
        return np.unpackbits(self.z[f"{name}_bits"])[:n].reshape(shape).astype(bool)

    def check_output(self, out: torch.Tensor, tol: float, what: str):
        """out [B,C,H,W] vs the stored reference output (full or sampled)."""
        out = out.detach().double().cpu()
        z = self.z
        assert tuple(out.shape) == tuple(int(v) for v in z["out_shape"]), what
        assert torch.isfinite(out).all(), what
        if "out_full" in z.files:
            err = rel_l2(out, torch.from_numpy(z["out_full"]))
            assert err <= tol, f"{what}: rel-L2 {err:.3e} > {tol:.1e}"
            return err
        idx = torch.from_numpy(z["out_idx"])
        e1 = rel_l2(out.reshape(-1)[idx], torch.from_numpy(z["out_sample"]))
        e2 = rel_l2(out.sum(dim=1), torch.from_numpy(z["out_chansum"]))
        e3 = rel_l2(out.sum(dim=(2, 3)), torch.from_numpy(z["out_perchan"]))
        # sums over C (or H*W) elements shrink relative errors of independent noise but keep
        # systematic ones; same tolerance on all three
        assert e1 <= tol and e2 <= tol and e3 <= tol, \
            f"{what}: rel-L2 sample {e1:.3e} chansum {e2:.3e} perchan {e3:.3e} > {tol:.1e}"
        return max(e1, e2, e3)
