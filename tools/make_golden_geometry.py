"""Generate ``tests/golden/geometry_cases.json`` from the REFERENCE (container only).

    python tools/make_golden_geometry.py

Widens the bit-exact pin of SURVEY.md §8 a2 / a3 (integer zone geometry and the three masks) from the nine full fusion
fixtures to ~150 zone layouts: 8x8 and 6x6 (the reference's training layout) grids of square zones with random sizes and
offsets - including grids that leave the 480x640 canvas (pad_mask branch) and sizes that do not divide the feature
stride (bilinear-resize branch) - on 416x544 and 480x640 inputs, at all three decoder levels.  For every layout the
reference's own ``patch_info_from_rect_data`` and ``TransformerFusion.forward`` (tiny embedding, hist2image layer only:
the geometry does not depend on the channel count) are run and the geometry ints plus SHA-1 digests of the
materialised ``zone_mask`` / ``pad_mask`` bit patterns are recorded; layouts on which the reference itself raises are
recorded as such (the product must refuse them too).
"""
import hashlib
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ref_import import import_reference  # noqa: E402
from cfpnet_b200 import synth  # noqa: E402
import make_golden as MG  # noqa: E402  (run_captured, GEO_KEYS; its __main__ block does not run on import)

ref = MG.ref
args = MG.args
TransformerFusion = MG.TransformerFusion
C = 8


def grid_rects(y0, x0, px, zn):
    ys = torch.arange(zn, dtype=torch.float32) * px + y0
    xs = torch.arange(zn, dtype=torch.float32) * px + x0
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([yy, xx, yy + px, xx + px], dim=-1).reshape(-1, 4)


def digest(mask_2d: np.ndarray) -> str:
    return hashlib.sha1(np.packbits(mask_2d.astype(bool)).tobytes()).hexdigest()


def run_case(img_h, img_w, zn, px, y0, x0, level):
    _, stride, max_res, _ = synth.LEVELS[level]
    H, W = img_h // stride, img_w // stride
    args.attention_layer = ["hist2image"]
    args.change_embedding = True
    args.no_skip_inside = False
    mod = TransformerFusion(C, list(max_res), large_kernel=7, patch_size=640 // max_res[1]).eval()
    rect = grid_rects(y0, x0, px, zn)
    pi = synth.collate_patch_info([ref["dataloader"].patch_info_from_rect_data(rect)])
    rec = dict(img=[img_h, img_w], zone_num=zn, px=px, y0=y0, x0=x0, level=level,
               patch_info={str(cps): {k: pi[cps][k][0].tolist() for k in pi[cps]} for cps in (4, 8, 16)})
    x = torch.zeros(1, C, H, W)
    feat1 = torch.zeros(1, zn * zn, 16, C)
    mask = torch.ones(1, zn * zn, dtype=torch.bool)
    try:
        with torch.no_grad():
            torch.manual_seed(0)
            _, cap = MG.run_captured(mod, x, feat1, rect_data=rect[None], mask=mask, patch_info=pi, rgb=None)
    except Exception as e:                                  # the reference's own forward is not defined for this layout
        rec["reference_raises"] = type(e).__name__
        return rec
    rec["geo"] = {k: int(cap[k]) for k in MG.GEO_KEYS if not k.startswith("offset")}
    zm = MG.channel_const(cap["zone_mask"], C).reshape(H, W)
    pm = MG.channel_const(cap["pad_mask"], C).reshape(cap["tzh"], cap["tzw"])
    rec["zone_mask_sha1"], rec["zone_mask_sum"] = digest(zm), int(zm.sum())
    rec["pad_mask_sha1"], rec["pad_mask_sum"] = digest(pm), int(pm.sum())
    return rec


def main():
    rng = random.Random(7)
    cases = []
    for img_h, img_w in ((416, 544), (480, 640)):
        for zn in (8, 6):
            layouts = [(px, int((img_h - px * zn) / 2), int((img_w - px * zn) / 2)) for px in (40, 48, 56, 64)]   # centred
            for _ in range(14):                              # random size / offset, may leave the 480x640 canvas
                px = rng.choice([36, 40, 44, 48, 52, 56, 60, 64, 68, 72])
                layouts.append((px, rng.randrange(-24, max(img_h - px * zn + 24, -23)), rng.randrange(-24, max(img_w - px * zn + 24, -23))))
            for px, y0, x0 in layouts:
                for level in (3, 2, 1):
                    cases.append(run_case(img_h, img_w, zn, px, y0, x0, level))
    ok = sum("geo" in c for c in cases)
    print(f"{len(cases)} layouts x levels, {ok} with a defined reference forward, {len(cases) - ok} on which the reference raises")
    with open(os.path.join(ROOT, "tests", "golden", "geometry_cases.json"), "w") as fh:
        json.dump(cases, fh)


if __name__ == "__main__":
    main()
