"""Deterministic synthetic inputs and weights for the fusion path.

Shared by ``tests/``, ``tools/make_golden.py`` and ``bench.py`` so that every
party (reference modules in the build container, CPU oracle, CUDA path) sees
byte-identical inputs and weights without shipping them: everything is a pure
function of (name, shape, seed) on torch's CPU generator.

Input recipe = SURVEY.md §8d: decoder features ``x ~ N(0,1)`` per level, zone
samples ``linspace(mu-3s, mu+3s, 16)`` with ``mu~U(0.3,4)``, ``s~U(0.01,0.2)``
(the reference's ``sample_uniform`` path, ``src/utils/dataloader.py:74-79``),
zeros for invalid zones, ``mask ~ Bernoulli(0.8)``, zone rectangles on a
centred grid.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Sequence, Tuple

import torch

from .geometry import patch_info_from_rect_data, collate_patch_info

# name -> (image H, image W, zone px)   (SURVEY.md §8 geometries)
GEOMETRIES = {
    "G416": (416, 544, 48),
    "G480": (480, 640, 56),      # L3 takes the bilinear-resize branch
    "G480pad": (480, 640, 64),   # zones reach outside the image: pad_mask path
    "G416z6": (416, 544, 64),    # 6x6 zones of 64 px: the reference's training layout (--train_zone_num 6)
}
ZONE_NUM = {"G416z6": 6}         # zones per side where it is not 8
# level -> (C, downscale, max_resolution, large_kernel)   (decoder.py:82-94)
LEVELS = {
    3: (128, 16, (30, 40), 7),
    2: (64, 8, (60, 80), 15),
    1: (32, 4, (120, 160), 31),
}
COMBINE1_LAYERS = ("hist2image", "combine1", "image", "hist2image", "combine1", "image")
BASELINE_LAYERS = ("hist2image", "image", "hist2image", "image")


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def centred_rects(img_h: int, img_w: int, zone_px: int, zone_num: int = 8) -> torch.Tensor:
    """[Z,4] float (y0,x0,y1,x1): zone_num x zone_num grid of square zones
    centred in the image, row-major (what ``get_hist_parallel`` emits,
    ``src/utils/dataloader.py:101-123``)."""
    y0 = int((img_h - zone_px * zone_num) / 2)
    x0 = int((img_w - zone_px * zone_num) / 2)
    ys = torch.arange(zone_num, dtype=torch.float32) * zone_px + y0
    xs = torch.arange(zone_num, dtype=torch.float32) * zone_px + x0
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([yy, xx, yy + zone_px, xx + zone_px], dim=-1).reshape(-1, 4)


def level_hw(geometry: str, level: int) -> Tuple[int, int]:
    img_h, img_w, _ = GEOMETRIES[geometry]
    s = LEVELS[level][1]
    return img_h // s, img_w // s


def make_inputs(geometry: str, batch: int, seed: int = 1, levels: Sequence[int] = (3, 2, 1),
                dtype=torch.float32, zone_num: int | None = None, p_valid: float = 0.8) -> dict:
    """One synthetic batch: ``x{level}`` [B,C,h,w], ``hist_data`` [B,Z,16],
    ``mask`` [B,Z] bool, ``rect_data`` [B,Z,4], ``patch_info`` (collated)."""
    img_h, img_w, zone_px = GEOMETRIES[geometry]
    zone_num = ZONE_NUM.get(geometry, 8) if zone_num is None else zone_num
    Z = zone_num * zone_num
    out = {}
    for lv in levels:
        C = LEVELS[lv][0]
        h, w = level_hw(geometry, lv)
        out[f"x{lv}"] = torch.randn(batch, C, h, w, generator=_gen(seed, f"x{lv}")).to(dtype)
    g = _gen(seed, "hist")
    mu = torch.rand(batch, Z, generator=g) * 3.7 + 0.3
    sg = torch.rand(batch, Z, generator=g) * 0.19 + 0.01
    mask = torch.rand(batch, Z, generator=g) < p_valid
    t = torch.linspace(0, 1, 16)
    samples = (mu - 3 * sg).unsqueeze(-1) * (1 - t) + (mu + 3 * sg).unsqueeze(-1) * t
    out["hist_data"] = (samples * mask.unsqueeze(-1)).to(torch.float32)
    out["mask"] = mask
    rect = centred_rects(img_h, img_w, zone_px, zone_num)
    out["rect_data"] = rect.unsqueeze(0).repeat(batch, 1, 1)
    out["patch_info"] = collate_patch_info([patch_info_from_rect_data(rect)] * batch)
    return out


ENCODER_CHANNELS = (16, 40, 56, 136, 232)          # x_block0..4 at strides 2..32 (decoder.py:100-104)


def encoder_features(geometry: str, batch: int, seed: int = 1) -> list:
    """Synthetic stand-ins for the five image-encoder feature maps the decoder consumes (the timm backbone is
    third-party and out of scope): ``[B, c, H/s, W/s]`` ~ N(0, 1) for s = 2, 4, 8, 16, 32."""
    img_h, img_w, _ = GEOMETRIES[geometry]
    g = _gen(seed, "encoder_features")
    return [torch.randn(batch, c, img_h // s, img_w // s, generator=g) for c, s in zip(ENCODER_CHANNELS, (2, 4, 8, 16, 32))]


def synthetic_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic weights for any module given its state_dict shapes.

    Magnitudes follow ``Deltar._reset_parameters`` (``deltar.py:23-32``):
    kaiming-normal fan_out for conv/linear weights, ``trunc_normal(std=0.2)``
    for the positional tables (``fusion.py:22-23``) — but norm layers and BN
    running statistics are *randomised* (not 1/0) so that BN folding, affine
    LayerNorm and biases are actually exercised by the parity tests.
    """
    sd = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        parent = name.rsplit(".", 2)[-2] if name.count(".") else ""
        is_norm = parent.startswith(("bn", "norm"))
        if leaf == "num_batches_tracked":
            t = torch.tensor(7, dtype=torch.long)
        elif name.endswith("positional_encodings") or name.endswith("positional_encodings2"):
            t = (torch.randn(shape, generator=g) * 0.2).clamp_(-0.4, 0.4)
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.2
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "weight" and (is_norm or len(shape) == 1):     # norm layers, also unnamed ones in nn.Sequential
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight" and len(shape) >= 2:
            fan_out = shape[0] * int(math.prod(shape[2:]))
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        else:
            raise KeyError(f"no synthetic rule for {name} {shape}")
        sd[name] = t
    return sd


def synthetic_depth_map(h: int, w: int, seed: int) -> torch.Tensor:
    """Piecewise-smooth synthetic depth map [h,w] in metres for the input-side fixtures: a tilted plane with noise, boxes
    at other depths (zones with two clusters), a band of invalid (zero) pixels and a far corner beyond the sensor range
    (zones with no signal)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    d = 0.6 + 2.8 * yy + 0.5 * xx + 0.02 * torch.randn(h, w, generator=g)
    for _ in range(14):
        y0, x0 = int(torch.randint(0, h - 40, [1], generator=g)), int(torch.randint(0, w - 40, [1], generator=g))
        hh, ww = int(torch.randint(20, 90, [1], generator=g)), int(torch.randint(20, 90, [1], generator=g))
        d[y0:y0 + hh, x0:x0 + ww] = float(torch.rand(1, generator=g) * 5.0 + 0.2)
    d[h // 3: h // 3 + 60, : w // 2] = 0.0
    d[:130, w - 200:] = 7.5
    return d.float()
