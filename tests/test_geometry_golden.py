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


# ------------------------------------------------------------------ the same integers through the C ABI
import ctypes  # noqa: E402

from cfpnet_b200 import _lib  # noqa: E402


def c_geometry(rects, max_width, H, W):
    """cfp_geometry_from_rects on a [B,Z,4] float32 array -> (return code, CfpGeom, message)"""
    from cfpnet_b200.build import build
    build()
    lib = _lib.load()
    r = np.ascontiguousarray(np.asarray(rects, dtype=np.float32))
    if r.ndim == 2:
        r = r[None]
    out = _lib.CfpGeom()
    rc = lib.cfp_geometry_from_rects(r.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), r.shape[0], r.shape[1], max_width, H, W,
                                     ctypes.byref(out))
    return rc, out, lib.cfp_last_error().decode()


C_FIELDS = {"pad_height": "pad_h", "pad_width": "pad_w", "p1": "p1", "p2": "p2", "sy_wo_pad": "sy_wo", "sx_wo_pad": "sx_wo",
            "ey_wo_pad": "ey_wo", "ex_wo_pad": "ex_wo", "tzh": "tzh", "tzw": "tzw", "interpolate": "interpolate",
            "zone_num": "zone_num"}


@pytest.mark.parametrize("i", range(len(CASES)))
def test_c_abi_geometry_bit_exact(i):
    """include/cfp.h::cfp_geometry_from_rects (plain C, host only) against the integers the REFERENCE computed, and its
    verdict (accept / refuse) against the Python host logic, on all 216 layouts."""
    c = CASES[i]
    _, stride, max_res, _ = synth.LEVELS[c["level"]]
    H, W = c["img"][0] // stride, c["img"][1] // stride
    rect = grid_rects(c)
    rc, out, msg = c_geometry(rect.numpy(), max_res[1], H, W)
    g = geometry.zone_geometry(geometry.collate_patch_info([geometry.patch_info_from_rect_data(rect)]), max_res[1], H, W)
    for name, _t in _lib.CfpGeom._fields_:                       # the C struct equals the Python dataclass field by field
        assert getattr(out, name) == getattr(g, name), name
    if "geo" in c:
        for ref_name, mine in C_FIELDS.items():
            assert getattr(out, mine) == c["geo"][ref_name], ref_name
    try:
        geometry.check_geometry(g, H, W)
        accepted = True
    except ValueError:
        accepted = False
    assert (rc == 0) == accepted, msg
    if "reference_raises" in c:
        assert rc != 0


def test_c_abi_geometry_takes_the_batch_extremes():
    """fusion.py:75-78: pads / patch sizes are the batch maximum, the canvas the union over the batch"""
    a = grid_rects({"zone_num": 8, "px": 48, "y0": 16, "x0": 80}).numpy()
    b = grid_rects({"zone_num": 8, "px": 40, "y0": 48, "x0": 112}).numpy()
    rc, both, _ = c_geometry(np.stack([a, b]), 40, 26, 34)
    _, ga, _ = c_geometry(a, 40, 26, 34)
    _, gb, _ = c_geometry(b, 40, 26, 34)
    assert both.p1 == max(ga.p1, gb.p1) and both.sy_wo == min(ga.sy_wo, gb.sy_wo) and both.ey_wo == max(ga.ey_wo, gb.ey_wo)
    pi = geometry.collate_patch_info([geometry.patch_info_from_rect_data(torch.from_numpy(r)) for r in (a, b)])
    g = geometry.zone_geometry(pi, 40, 26, 34)
    for name, _t in _lib.CfpGeom._fields_:
        assert getattr(both, name) == getattr(g, name), name


def test_c_abi_geometry_rejects_bad_arguments():
    a = grid_rects({"zone_num": 8, "px": 48, "y0": 16, "x0": 80}).numpy()
    rc, _, msg = c_geometry(a, 48, 26, 34)
    assert rc != 0 and "640" in msg
    rc, _, msg = c_geometry(a, 640, 26, 34)                      # cell size 1: no patch_info entry in the reference either
    assert rc != 0 and "cell size" in msg
    bad = a.copy()
    bad[5, 2] = np.nan
    rc, _, msg = c_geometry(bad, 40, 26, 34)
    assert rc != 0 and "rect_data[0][5]" in msg
