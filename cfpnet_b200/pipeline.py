"""The fusion path as one callable: histogram encoder + the three ``TransformerFusion``
calls of the reference decoder (``cross_atten3`` at 1/16, ``cross_atten2`` at 1/8,
``cross_atten1`` at 1/4; ``src/models/decoder.py:90-94,111,116,121``, ``deltar.py:40``).

One *frame* of BASELINE.json's metric is exactly this work for one image.  The
class only wires the drop-in modules together the way ``Deltar`` / ``Decoder`` do and
adds the host<->device staging used for the end-to-end measurement.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn as nn

from .config import args
from .encoder import HistogramEncoder
from .fusion import TransformerFusion
from .synth import LEVELS


class FusionPath(nn.Module):
    def __init__(self, layer_names: Sequence[str] | None = None):
        super().__init__()
        if layer_names is not None:
            args.attention_layer = list(layer_names)
        self.hist_encoder = HistogramEncoder()
        # same constructor arguments as decoder.py:92-94
        self.cross_atten1 = TransformerFusion(embedding_dim=32, max_resolution=[120, 160], large_kernel=31, patch_size=16)
        self.cross_atten2 = TransformerFusion(embedding_dim=64, max_resolution=[60, 80], large_kernel=15, patch_size=8)
        self.cross_atten3 = TransformerFusion(embedding_dim=128, max_resolution=[30, 40], large_kernel=7, patch_size=4)
        self._pinned_out = None

    def set_dtype(self, dtype: torch.dtype) -> "FusionPath":
        self.to(dtype)
        self.hist_encoder.out_dtype = dtype
        return self

    def forward(self, x3, x2, x1, hist_data, mask, patch_info, rect_data=None) -> List[torch.Tensor]:
        """x3/x2/x1: decoder features [B,128,h/16,w/16], [B,64,h/8,w/8], [B,32,h/4,w/4];
        hist_data [B,Z,S]; mask [B,Z] bool.  Returns the fused maps in call order."""
        f32, f64, f128 = self.hist_encoder(hist_data.unsqueeze(-1))
        kw = dict(rect_data=rect_data, mask=mask, patch_info=patch_info, rgb=None)
        return [self.cross_atten3(x3, f128, **kw), self.cross_atten2(x2, f64, **kw), self.cross_atten1(x1, f32, **kw)]

    # ------------------------------------------------------------------ host-buffer entry
    def forward_host(self, host: Dict[str, torch.Tensor], patch_info, device) -> List[torch.Tensor]:
        """End-to-end call with HOST buffers: pinned inputs are copied to the device, the path
        runs, and the three fused maps are read back into pinned host memory."""
        dev = {k: host[k].to(device, non_blocking=True) for k in ("x3", "x2", "x1", "hist_data", "mask")}
        outs = self.forward(dev["x3"], dev["x2"], dev["x1"], dev["hist_data"], dev["mask"], patch_info)
        if self._pinned_out is None or any(p.shape != o.shape or p.dtype != o.dtype
                                           for p, o in zip(self._pinned_out, outs)):
            self._pinned_out = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
        for p, o in zip(self._pinned_out, outs):
            p.copy_(o, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return self._pinned_out
