"""Input side of the path on the GPU: the two functions the reference's dataloader runs on the host CPU
per frame - ``get_hist_parallel`` and ``sample_point_from_hist_parallel`` (``src/utils/dataloader.py:84-134, 65-81``) -
as batched libcfp kernels.  Same argument meaning as the reference (a ``config`` object with the reference's field names),
batched over frames; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import random
from typing import Tuple

import numpy as np
import torch

from . import _lib


def _bin_centres(max_distance: float) -> Tuple[int, torch.Tensor]:
    """(bins, centres float64 [bins]) exactly as dataloader.py:92,103,120 forms them: ``range_margin`` is a float64
    ``np.arange`` in 4 cm steps, the upper edges pass through a float32 ``torch.Tensor`` before the average."""
    range_margin = list(np.arange(0, max_distance + 1e-9, 0.04))
    bins = int(max_distance / 0.04)
    dist = (torch.tensor(range_margin[1:], dtype=torch.float32).double() + torch.tensor(range_margin[:-1], dtype=torch.float64)) / 2
    if dist.numel() != bins:
        raise ValueError(f"max_distance {max_distance}: {dist.numel()} bin centres for {bins} bins (the reference fails to "
                         "broadcast here too)")
    return bins, dist


def get_hist_parallel(rgb: torch.Tensor, dep: torch.Tensor, config, return_hist: bool = False):
    """Batched ``get_hist_parallel``: ``rgb`` [B,3,H,W] (only its size is used, as in the reference), ``dep`` [B,1,H,W] or
    [B,H,W] depth in metres on a CUDA device.  Returns ``(fh [B,Z,2] (mu, sigma), fr [Z,4] zone rectangles
    (y0, x0, y1, x1), mask [B,Z] bool)`` (+ the surviving histogram counts [B,Z,bins] when ``return_hist``).  The random
    draws (``random_simu_max_d``, ``train_zone_random_offset``) come from the same host generators, once per call."""
    _lib.require_cuda(dep, "dep")
    height, width = rgb.shape[-2], rgb.shape[-1]
    d = dep.detach()
    if d.dim() == 4:
        d = d[:, 0]
    d = d.float().contiguous()
    B = d.shape[0]
    if getattr(config, "random_simu_max_d", False):
        max_distance = float(np.random.uniform(low=config.simu_min_d, high=config.simu_max_d, size=1)[0])
    else:
        max_distance = float(config.simu_max_distance)
    ph, pw = (64, 64) if config.mode == "train" else (56, 56)
    offset = 0
    if config.train_zone_random_offset > 0:
        offset = random.randint(-config.train_zone_random_offset, config.train_zone_random_offset)
    zn = config.train_zone_num if config.mode == "train" else 8
    sy = int((height - ph * zn) / 2) + offset
    sx = int((width - pw * zn) / 2) + offset
    bins, centres = _bin_centres(max_distance)
    dev = d.device
    centres = centres.to(dev)
    fh = torch.empty(B, zn * zn, 2, device=dev, dtype=torch.float32)
    mask = torch.empty(B, zn * zn, device=dev, dtype=torch.uint8)
    hist = torch.empty(B, zn * zn, bins, device=dev, dtype=torch.int32) if return_hist else None
    with torch.cuda.device(dev):
        _lib.call("cfp_zone_hist", d.data_ptr(), B, d.shape[1], d.shape[2], sy, sx, ph, pw, zn, bins, C.c_float(max_distance),
                  centres.data_ptr(), fh.data_ptr(), mask.data_ptr(), _lib.ptr(hist), _lib.stream_ptr())
    ys = torch.arange(sy, sy + ph * zn, ph, dtype=torch.float32).repeat_interleave(zn)
    xs = torch.arange(sx, sx + pw * zn, pw, dtype=torch.float32).repeat(zn)
    fr = torch.stack([ys, xs, ys + ph, xs + pw], dim=1)
    out = (fh, fr, mask.bool())
    return out + (hist,) if return_hist else out


_TABLES = {}


def _sample_tables(S: int, uniform: bool, dev):
    key = (S, uniform, str(dev))
    if key not in _TABLES:
        if uniform:                                     # tensor_linspace (dataloader.py:43-58): two float32 ramps
            w0, w1 = torch.linspace(1, 0, steps=S), torch.linspace(0, 1, steps=S)
        else:                                           # dataloader.py:70-73: ppf grid -> erfinv(2 p - 1)
            delta = 1e-3
            ppf = torch.Tensor(np.arange(delta, 1, (1 - 2 * delta) / (S - 1)).tolist())
            if ppf.numel() != S:
                raise ValueError(f"zone_sample_num {S}: the reference's ppf grid has {ppf.numel()} points")
            w0, w1 = torch.erfinv(2 * ppf - 1), torch.zeros(S)
        _TABLES[key] = (w0.float().to(dev), w1.float().to(dev))
    return _TABLES[key]


def sample_point_from_hist_parallel(hist_data: torch.Tensor, mask: torch.Tensor, config) -> torch.Tensor:
    """``hist_data`` [..., Z, 2] (mu, sigma), ``mask`` [..., Z] bool on a CUDA device -> [..., Z, zone_sample_num] float32
    depth samples, zeros for invalid zones (the tensor ``HistogramEncoder`` consumes)."""
    _lib.require_cuda(hist_data, "hist_data")
    S = int(config.zone_sample_num)
    fh = hist_data.detach().float().contiguous()
    m = mask.detach().to(device=fh.device, dtype=torch.uint8).contiguous()
    if fh.shape[:-1] != m.shape or fh.shape[-1] != 2:
        raise ValueError(f"hist_data {tuple(fh.shape)} / mask {tuple(m.shape)}: expected [..., Z, 2] and [..., Z]")
    out = torch.empty(*m.shape, S, device=fh.device, dtype=torch.float32)
    w0, w1 = _sample_tables(S, bool(config.sample_uniform), fh.device)
    with torch.cuda.device(fh.device):
        _lib.call("cfp_zone_samples", fh.data_ptr(), m.data_ptr(), out.data_ptr(), m.numel(), S, w0.data_ptr(), w1.data_ptr(),
                  0 if config.sample_uniform else 1, _lib.stream_ptr())
    return out
