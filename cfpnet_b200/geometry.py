"""Host-side integer zone geometry of the fusion path.

Pure Python/torch-CPU integer logic, bit-exact with the reference:

* :func:`patch_info_from_rect_data` mirrors the reference helper of the same
  name, ``src/utils/dataloader.py:13-40`` — the producer of the ``patch_info``
  kwarg every ``TransformerFusion.forward`` call receives.
* :class:`ZoneGeometry` holds the per-level integers ``TransformerFusion.forward``
  derives from ``patch_info`` (``src/models/fusion.py:67-84``) plus the clipped
  in-image zone rectangle of ``fusion.py:104``.  These few ints are all the CUDA
  kernels need: the reference's materialised ``zone_mask`` / ``hist_mask`` /
  ``pad_mask`` tensors (``fusion.py:103-120``) are pure functions of them and of
  the ``[B,Z]`` validity mask.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Sequence

import torch

REF_CANVAS_H, REF_CANVAS_W = 480, 640      # hard-coded in utils/dataloader.py:20-23


def patch_info_from_rect_data(rect_data: torch.Tensor) -> dict:
    """``rect_data`` [Z,4] float rows (y0,x0,y1,x1) -> per-frame patch_info dict
    keyed by conv patch size 4/8/16 (+ ``zone_num``).

    Truncation rules that matter for bit-exactness: the start/end cell indices
    are float divisions cast to int32 (truncate toward zero), the patch size is
    ``ceil(max_zone_px / cps)``, the pad is ``ceil(px_outside_480x640 / cps)``.
    """
    r = rect_data.detach().to("cpu", torch.float32)
    zone_num = int(math.sqrt(r.shape[0]))
    hgt = (r[..., 2] - r[..., 0]).max().to(torch.int32).item()
    wid = (r[..., 3] - r[..., 1]).max().to(torch.int32).item()
    over_h = int(max(r[..., 0].clamp(max=0).abs().max().item(),
                     (r[..., 2].clamp(min=REF_CANVAS_H) - REF_CANVAS_H).max().item()))
    over_w = int(max(r[..., 1].clamp(max=0).abs().max().item(),
                     (r[..., 3].clamp(min=REF_CANVAS_W) - REF_CANVAS_W).max().item()))
    info = {}
    for cps in (4, 8, 16):
        idx = [(r[..., 0] / cps).min(), (r[..., 1] / cps).min(),
               (r[..., 2] / cps).max(), (r[..., 3] / cps).max()]
        info[cps] = {
            "pad_size": torch.tensor([math.ceil(over_h / cps), math.ceil(over_w / cps)], dtype=torch.int),
            "patch_size": torch.tensor([math.ceil(hgt / cps), math.ceil(wid / cps)], dtype=torch.int),
            "index_wo_pad": torch.tensor([int(v.to(torch.int32)) for v in idx], dtype=torch.int),
        }
    info["zone_num"] = zone_num
    return info


def collate_patch_info(infos: Sequence[dict]) -> dict:
    """Batch per-frame dicts the way the reference's DataLoader (default
    collate) does: every tensor gains a leading B dim, ``zone_num`` -> [B]."""
    out = {cps: {k: torch.stack([f[cps][k] for f in infos]) for k in infos[0][cps]}
           for cps in (4, 8, 16)}
    out["zone_num"] = torch.tensor([f["zone_num"] for f in infos])
    return out


@dataclass(frozen=True)
class ZoneGeometry:
    """Per-level integers of one ``TransformerFusion.forward`` call."""
    zone_num: int
    pad_h: int
    pad_w: int
    p1: int
    p2: int
    sy_wo: int
    sx_wo: int
    ey_wo: int
    ex_wo: int
    sy: int
    ey: int
    sx: int
    ex: int
    tzh: int
    tzw: int
    interpolate: int
    ry0: int    # in-image zone rectangle, rows [ry0, ry1)
    ry1: int
    rx0: int
    rx1: int

    def asdict(self):
        return asdict(self)

    @property
    def n_inside(self):
        return (self.ry1 - self.ry0) * (self.rx1 - self.rx0)


def _host_ints(t: torch.Tensor):
    return [int(v) for v in t.detach().to("cpu").tolist()]


def zone_geometry(patch_info: dict, max_width: int, H: int, W: int) -> ZoneGeometry:
    """fusion.py:41,67-84.  One host read per tensor; no per-comparison syncs
    even when ``patch_info`` lives on the GPU (nn.DataParallel scatters it)."""
    cps = 640 / max_width                       # float key, hashes like the int key
    info = patch_info[cps]
    zn = int(patch_info["zone_num"][0])
    pad_h, pad_w = _host_ints(info["pad_size"].max(dim=0)[0])
    p1, p2 = _host_ints(info["patch_size"].max(dim=0)[0])
    lo = _host_ints(info["index_wo_pad"].min(dim=0)[0])
    hi = _host_ints(info["index_wo_pad"].max(dim=0)[0])
    sy_wo, sx_wo, ey_wo, ex_wo = lo[0], lo[1], hi[2], hi[3]
    sy, ey, sx, ex = sy_wo + pad_h, ey_wo + pad_h, sx_wo + pad_w, ex_wo + pad_w
    tzh, tzw = ey - sy, ex - sx
    clip = lambda v, top: min(max(v, 0), top)
    return ZoneGeometry(
        zone_num=zn, pad_h=pad_h, pad_w=pad_w, p1=p1, p2=p2,
        sy_wo=sy_wo, sx_wo=sx_wo, ey_wo=ey_wo, ex_wo=ex_wo,
        sy=sy, ey=ey, sx=sx, ex=ex, tzh=tzh, tzw=tzw,
        interpolate=int(tzh != p1 * zn or tzw != p2 * zn),
        ry0=clip(sy_wo, H), ry1=clip(ey_wo, H), rx0=clip(sx_wo, W), rx1=clip(ex_wo, W))


def check_geometry(g: ZoneGeometry, H: int, W: int) -> None:
    """Conditions under which the reference's own forward is well defined: the
    canvas slice must lie inside the padded map (``fusion.py:136-138``) and the
    number of in-image canvas cells must equal the zone-rectangle cell count
    (``feat0[zone_mask] += zone_feature[pad_mask]``, ``fusion.py:157``)."""
    if g.sy < 0 or g.sx < 0 or g.ey > H + 2 * g.pad_h or g.ex > W + 2 * g.pad_w:
        raise ValueError(f"zone canvas [{g.sy}:{g.ey},{g.sx}:{g.ex}] leaves the padded "
                         f"{H + 2 * g.pad_h}x{W + 2 * g.pad_w} map")
    if g.tzh <= 0 or g.tzw <= 0:
        raise ValueError("empty zone canvas")
    top, left = max(-g.sy_wo, 0), max(-g.sx_wo, 0)
    bot, right = max(g.ey_wo - H, 0), max(g.ex_wo - W, 0)
    if g.pad_h == 0 and g.pad_w == 0:
        top = left = bot = right = 0
    if (g.tzh - top - bot) != (g.ry1 - g.ry0) or (g.tzw - left - right) != (g.rx1 - g.rx0):
        raise ValueError("in-image canvas cells do not match the zone rectangle")
