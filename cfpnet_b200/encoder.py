"""Drop-in for the reference's ``HistogramEncoder`` (``src/models/encoder.py:6-50``).

Same constructor, same ``forward(hist_data: [B,Z,N,1]) -> [f32, f64, f128]``,
same parameter names/shapes (``hist_extractor{1,2,3}.pointnet_encoder.{conv,bn}{1,2,3}.*``)
so reference checkpoints load with ``strict=True``.  The arithmetic runs in one
hand-written sm_100a kernel (``csrc/k_hist.cu``) behind ``cfp_hist_encoder_fwd``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .packing import PackCache, Packer, fold_bn, host, relocate, umma_block


class PointNetEncoder(nn.Module):
    """Parameter container of encoder.py:6-24 (three Conv1d(k=1) + BatchNorm1d)."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channel, out_channel, 1)
        self.conv2 = nn.Conv1d(out_channel, out_channel, 1)
        self.conv3 = nn.Conv1d(out_channel, out_channel, 1)
        self.bn1 = nn.BatchNorm1d(out_channel)
        self.bn2 = nn.BatchNorm1d(out_channel)
        self.bn3 = nn.BatchNorm1d(out_channel)

    def packed_stages(self):
        """[(w_t [Cin,Cout], b [Cout])] * 3 with eval-mode BN folded in."""
        out = []
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3)):
            scale, shift = fold_bn(bn)
            w = host(conv.weight)[:, :, 0] * scale[:, None]
            out.append((w.t().contiguous(), (host(conv.bias) * scale + shift).contiguous()))
        return out


class HistExtractor(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.pointnet_encoder = PointNetEncoder(in_channel, out_channel)


class HistogramEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        channels = [32, 64, 128]
        self.hist_extractor1 = HistExtractor(in_channel=1, out_channel=channels[0])
        self.hist_extractor2 = HistExtractor(in_channel=channels[0], out_channel=channels[1])
        self.hist_extractor3 = HistExtractor(in_channel=channels[1], out_channel=channels[2])
        self.out_dtype = None            # None -> float32; set to torch.bfloat16 for the bf16 path
        self._cache = PackCache()

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()      # replicas own their pack cache (see TransformerFusion)
        replica._cache = PackCache()
        return replica

    def _pack(self):
        stages = []
        for ex in (self.hist_extractor1, self.hist_extractor2, self.hist_extractor3):
            stages += ex.pointnet_encoder.packed_stages()
        keep = Packer()
        w = _lib.CfpHistW()
        for i, (wt, b) in enumerate(stages):
            w.w_t[i], w.b[i] = keep.ref(wt), keep.ref(b)
        # tensor-core path: stages 1..8 as bf16 UMMA blocks ([Cout,Cin] weight -> [Cin/8][Cout][8]), concatenated
        tc = torch.cat([umma_block(wt.t()).reshape(-1) for wt, _ in stages[1:]]).contiguous()
        w.tc = keep.ref(tc)
        buf = keep.upload(self.hist_extractor1.pointnet_encoder.conv1.weight.device)    # one host->device copy
        relocate(w, buf.data_ptr())
        return w, (stages, tc, buf)

    def forward(self, hist_data):
        _lib.require_cuda(hist_data, "hist_data")
        B, Z, N, D = hist_data.shape
        if D != 1:
            raise ValueError("hist_data must be [B,Z,N,1] (deltar.py:40)")
        if self.training:
            # train mode (train.py:75): BatchNorm on batch statistics, differentiable - fp32 training kernels behind an
            # autograd Function (cfpnet_b200/train.py); returns the three token tensors as a list like the eval path
            from .train import HistEncoderTrainFn
            return list(HistEncoderTrainFn.apply(self, hist_data, *self.parameters()))
        dt = self.out_dtype or (hist_data.dtype if hist_data.dtype == torch.bfloat16 else torch.float32)
        w, _keep = self._cache.get(self, self._pack)
        x = hist_data.detach().reshape(-1).float().contiguous()
        rows = x.numel()
        outs = [torch.empty(B, Z, N, c, device=x.device, dtype=dt) for c in (32, 64, 128)]
        with torch.cuda.device(x.device):
            _lib.call("cfp_hist_encoder_fwd", x.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(),
                      outs[2].data_ptr(), rows, C.byref(w), _lib.dtype_code(dt), _lib.stream_ptr())
        return outs
