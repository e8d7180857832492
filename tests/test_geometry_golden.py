"""Zone geometry and masks against 216 layouts run through the REFERENCE (tests/golden/geometry_cases.json, made by
tools/make_golden_geometry.py): 8x8 and 6x6 grids, random zone sizes / offsets, zones leaving the canvas, sizes that
trigger the bilinear-resize branch, both input sizes, all three levels.  Integers and masks must be bit-exact; layouts
on which the reference's own forward raises must be refused by the product's geometry check."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from cfpnet_b200 import geometry, synth
from oracle import cfp_oracle as O

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "geometry_cases.json")))
NAMES = {"pad_height": "pad_h", "pad_width": "pad_w", "p1": "p1", "p2": "p2", "sy_wo_pad": "sy_wo", "sx_wo_pad": "sx_wo",
         "ey_wo_pad": "ey_wo", "ex_wo_pad": "ex_wo", "sy": "sy", "ey": "ey", "sx": "sx", "ex": "ex", "tzh": "tzh", "tzw": "tzw",
         "interpolate": "interpolate", "zone_num": "zone_num"}


def grid_rects(c):
    zn, px = c["zone_num"], c["px"]
    ys = torch.arange(zn, dtype=torch.float32) * px + c["y0"]
    xs = torch.arange(zn, dtype=torch.float32) * px + c["x0"]
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([yy, xx, yy + px, xx + px], dim=-1).reshape(-1, 4)


def sha(mask_2d):
    return hashlib.sha1(np.packbits(np.asarray(mask_2d).astype(bool)).tobytes()).hexdigest()


def test_fixture_covers_every_branch():
    ok = [c for c in CASES if "geo" in c]
    clamped = 0
    for c in ok:
        _, stride, _, _ = synth.LEVELS[c["level"]]
        H, W = c["img"][0] // stride, c["img"][1] // stride
        q = c["geo"]
        clamped += q["sy"] < 0 or q["sx"] < 0 or q["ey"] > H + 2 * q["pad_height"] or q["ex"] > W + 2 * q["pad_width"]
    assert len(ok) - clamped >= 150                              # layouts the product serves
    assert len(CASES) >= 200 and len(ok) >= 190
    combos = {(c["geo"]["interpolate"], c["geo"]["pad_height"] > 0 or c["geo"]["pad_width"] > 0) for c in ok}
    assert combos == {(0, False), (0, True), (1, False), (1, True)}
    assert {c["zone_num"] for c in ok} == {6, 8} and {c["level"] for c in ok} == {1, 2, 3}


@pytest.mark.parametrize("i", range(len(CASES)))
def test_geometry_and_masks_bit_exact(i):
    c = CASES[i]
    _, stride, max_res, _ = synth.LEVELS[c["level"]]
    H, W = c["img"][0] // stride, c["img"][1] // stride
    rect = grid_rects(c)
    pi1 = geometry.patch_info_from_rect_data(rect)
    for cps in (4, 8, 16):                                       # the reference's own patch_info_from_rect_data
        for k, v in c["patch_info"][str(cps)].items():
            assert pi1[cps][k].tolist() == v, (cps, k)
    assert pi1["zone_num"] == c["zone_num"]
    g = geometry.zone_geometry(geometry.collate_patch_info([pi1]), max_res[1], H, W)
    if "reference_raises" in c:
        with pytest.raises(ValueError):
            geometry.check_geometry(g, H, W)
        return
    for ref_name, mine in NAMES.items():
        assert getattr(g, mine) == c["geo"][ref_name], ref_name
    leaves = g.sy < 0 or g.sx < 0 or g.ey > H + 2 * g.pad_h or g.ex > W + 2 * g.pad_w
    if leaves:
        # The canvas slice leaves the padded map (pads are computed against the 480x640 canvas, utils/dataloader.py:20-23,
        # the map is smaller).  The reference only survives this in the resize branch, where Python slicing silently
        # clamps the slice and F.interpolate stretches whatever is left; the product refuses the layout (documented
        # limitation) instead of reproducing the distortion.
        assert c["geo"]["interpolate"] == 1
        with pytest.raises(ValueError, match="leaves the padded"):
            geometry.check_geometry(g, H, W)
        return
    geometry.check_geometry(g, H, W)                             # the product accepts what the reference accepts
    Z = c["zone_num"] ** 2
    zm, _hm, pm = O.zone_masks(g.asdict(), torch.ones(1, Z, dtype=torch.bool), 1, H, W, 1)
    assert int(zm.sum()) == c["zone_mask_sum"] and sha(zm.reshape(H, W).numpy()) == c["zone_mask_sha1"]
    assert int(pm.sum()) == c["pad_mask_sum"] and sha(pm.reshape(g.tzh, g.tzw).numpy()) == c["pad_mask_sha1"]
