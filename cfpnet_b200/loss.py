"""Loss and metrics of the path on the GPU: ``SILogLoss`` (``src/loss.py:9-19``) as a differentiable
drop-in over libcfp's ``cfp_silog_fwd`` / ``cfp_silog_bwd``, and ``compute_errors`` (``src/utils/metrics.py:4-24``) as
one reduction launch on device tensors.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .headers import SILOG_SCRATCH_DOUBLES, METRICS_SCRATCH_DOUBLES


def _mask_u8(mask, like):
    if mask is None:
        return None
    return mask.detach().to(device=like.device, dtype=torch.uint8).contiguous()


class _SILogFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, interpolate):
        B, _, h, w = pred.shape
        H, W = target.shape[-2:]
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        m = _mask_u8(mask, p)
        scratch = torch.zeros(SILOG_SCRATCH_DOUBLES, device=p.device, dtype=torch.float64)
        loss = torch.empty((), device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            _lib.call("cfp_silog_fwd", p.data_ptr(), t.data_ptr(), _lib.ptr(m), B, h, w, H, W, int(bool(interpolate)),
                      scratch.data_ptr(), loss.data_ptr(), _lib.stream_ptr())
        ctx.saved = (p, t, m, scratch, bool(interpolate), pred.dtype)
        return loss

    @staticmethod
    def backward(ctx, gout):
        p, t, m, scratch, interpolate, dt = ctx.saved
        B, _, h, w = p.shape
        H, W = t.shape[-2:]
        grad = torch.empty_like(p)
        g = float(gout)            # the reference's loss is the scalar at the root of the graph: grad_out is 1 (one host read)
        with torch.cuda.device(p.device):
            _lib.call("cfp_silog_bwd", p.data_ptr(), t.data_ptr(), _lib.ptr(m), B, h, w, H, W, int(interpolate),
                      scratch.data_ptr(), C.c_float(g), grad.data_ptr(), _lib.stream_ptr())
        return grad.to(dt), None, None, None


class SILogLoss(nn.Module):
    """Drop-in for ``src/loss.py: SILogLoss``: ``forward(input [B,1,h,w], target [B,1,H,W], mask=None, interpolate=True)``."""

    def __init__(self):
        super().__init__()
        self.name = "SILog"

    def forward(self, input, target, mask=None, interpolate=True):
        _lib.require_cuda(input, "input")
        if input.dim() != 4 or target.dim() != 4 or input.shape[1] != 1 or target.shape[1] != 1 or input.shape[0] != target.shape[0]:
            raise ValueError(f"SILogLoss serves [B,1,h,w] predictions and [B,1,H,W] targets, got {tuple(input.shape)} / {tuple(target.shape)}")
        if mask is not None and mask.shape != target.shape:
            raise ValueError("mask must have the target's shape")
        if not interpolate and input.shape != target.shape:
            raise ValueError("without interpolation the prediction must have the target's size")
        return _SILogFn.apply(input, target, mask, interpolate)


METRIC_NAMES = ("a1", "a2", "a3", "abs_rel", "rmse", "log_10", "rmse_log", "silog", "sq_rel")


def compute_errors(gt: torch.Tensor, pred: torch.Tensor, valid: torch.Tensor | None = None) -> dict:
    """``compute_errors(gt, pred)`` of the reference on device tensors; ``valid`` (bool, same shape) selects the pixels -
    the reference's callers index ``gt[valid]`` / ``pred[valid]`` on the host first.  Returns the reference's dict
    (Python floats; this read is the call's one synchronisation) plus ``n``."""
    _lib.require_cuda(pred, "pred")
    g = gt.detach().float().contiguous()
    p = pred.detach().float().contiguous()
    if g.shape != p.shape:
        raise ValueError("gt and pred must have the same shape")
    v = _mask_u8(valid, p)
    scratch = torch.zeros(METRICS_SCRATCH_DOUBLES, device=p.device, dtype=torch.float64)
    out = torch.empty(10, device=p.device, dtype=torch.float64)
    with torch.cuda.device(p.device):
        _lib.call("cfp_depth_metrics", g.data_ptr(), p.data_ptr(), _lib.ptr(v), p.numel(), scratch.data_ptr(), out.data_ptr(),
                  _lib.stream_ptr())
    vals = out.cpu().tolist()
    res = dict(zip(METRIC_NAMES, vals[:9]))
    res["n"] = int(vals[9])
    return res
