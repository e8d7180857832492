"""Parameter containers + weight packing for the layers inside ``TransformerFusion``.

Class and parameter names follow the reference (``src/models/transformer.py``,
``src/models/convnext.py``) so that ``state_dict`` keys are identical, including
the parameters the reference constructs but never uses (kept registered, never
given a gradient: SURVEY.md §2).  The classes hold no torch arithmetic: their
``pack()`` methods emit the C structs of include/cfp.h and the launch itself is
``TransformerFusion.forward`` -> libcfp.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .packing import Packer, fold_bn, host, linear_t, umma_block


class LinearAttention(nn.Module):
    """Parameter-free; kept so module trees print like the reference (attention.py:10-14)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps


class LoFTREncoderLayer(nn.Module):
    """transformer.py:14-39."""

    def __init__(self, d_model, nhead):
        super().__init__()
        self.dim = d_model // nhead
        self.nhead = nhead
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.attention = LinearAttention()
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(
            nn.Linear(d_model * 2, d_model * 2, bias=False),
            nn.ReLU(True),
            nn.Linear(d_model * 2, d_model, bias=False),
        )
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def pack(self, keep: Packer) -> _lib.CfpLoftrW:
        t = dict(
            wq_t=linear_t(self.q_proj.weight),
            wkv_t=torch.cat([linear_t(self.k_proj.weight), linear_t(self.v_proj.weight)], dim=1).contiguous(),
            wm_t=linear_t(self.merge.weight),
            w1_t=linear_t(self.mlp[0].weight),
            w2_t=linear_t(self.mlp[2].weight),
            ln1_g=host(self.norm1.weight).contiguous(),
            ln1_b=host(self.norm1.bias).contiguous(),
            ln2_g=host(self.norm2.weight).contiguous(),
            ln2_b=host(self.norm2.bias).contiguous(),
        )
        C = self.q_proj.weight.shape[0]
        w1, w2 = self.mlp[0].weight.detach().cpu(), self.mlp[2].weight.detach().cpu()
        t["tc"] = torch.stack([umma_block(b) for b in (
            self.q_proj.weight, self.merge.weight, w1[:C, :C], w1[:C, C:], w1[C:, :C], w1[C:, C:],
            w2[:, :C], w2[:, C:])]).contiguous()
        t["kv_tc"] = torch.stack([umma_block(self.k_proj.weight), umma_block(self.v_proj.weight)]).contiguous()
        keep.extend(t.values())
        return _lib.CfpLoftrW(**{k: keep.ref(v) for k, v in t.items()})


class LocallyGroupedAttn(nn.Module):
    """LSA, transformer.py:75-87 (8 heads by default, :78)."""

    def __init__(self, dim, num_heads=8, ws=1):
        assert ws != 1
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim, self.num_heads, self.ws = dim, num_heads, ws
        self.encoder_layer = LoFTREncoderLayer(dim, num_heads)


class GlobalSubSampleAttn(nn.Module):
    """GSA, transformer.py:119-136."""

    def __init__(self, dim, num_heads=8, sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim, self.num_heads, self.sr_ratio = dim, num_heads, sr_ratio
        self.encoder_layer = LoFTREncoderLayer(dim, num_heads)
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.norm = nn.LayerNorm(dim)
        else:
            self.sr = None
            self.norm = None


class TwinsTransformer(nn.Module):
    """`image` layer, transformer.py:154-158.  ``num_heads`` is accepted and ignored
    exactly like the reference: LSA and GSA are built with their default 8 heads."""

    def __init__(self, dim, num_heads=8, ws=1):
        super().__init__()
        self.lga = LocallyGroupedAttn(dim=dim, ws=ws)
        self.gsa = GlobalSubSampleAttn(dim=dim, sr_ratio=ws)

    def pack(self, keep: Packer) -> _lib.CfpTwinsW:
        if self.gsa.sr is None:
            raise _lib.CfpError("libcfp serves GSA with sr_ratio > 1 only")
        C, ws = self.gsa.dim, self.gsa.sr_ratio
        w = _lib.CfpTwinsW()
        w.lsa = self.lga.encoder_layer.pack(keep)
        w.gsa = self.gsa.encoder_layer.pack(keep)
        # [Cout,Cin,ws,ws] -> [(dy,dx,cin)][Cout]
        srw = self.gsa.sr.weight.detach().cpu()
        sr_t = srw.float().permute(2, 3, 1, 0).reshape(ws * ws * C, C).contiguous()
        t = dict(sr_t=sr_t, sr_b=host(self.gsa.sr.bias).contiguous(),
                 srln_g=host(self.gsa.norm.weight).contiguous(),
                 srln_b=host(self.gsa.norm.bias).contiguous(),
                 sr_tc=torch.stack([umma_block(srw[:, :, dy, dx])
                                    for dy in range(ws) for dx in range(ws)]).contiguous())
        keep.extend(t.values())
        for k, v in t.items():
            setattr(w, k, keep.ref(v))
        w.ws = ws
        return w


class LoFTREncoderLayer_newcross9(nn.Module):
    """DAPM, transformer.py:169-202.  merge / mlp / norm1 / norm2 exist only for
    state_dict compatibility (never used by the reference's forward)."""

    def __init__(self, d_model, nhead):
        super().__init__()
        self.dim = d_model // nhead
        self.nhead = nhead
        self.q_proj = nn.Linear(d_model, d_model, bias=False)
        self.k_proj = nn.Linear(d_model, d_model, bias=False)
        self.v_proj = nn.Linear(d_model, d_model, bias=False)
        self.attention = LinearAttention()
        self.merge = nn.Linear(d_model, d_model, bias=False)
        self.mlp = nn.Sequential(
            nn.Linear(d_model * 2, d_model * 2, bias=False),
            nn.ReLU(True),
            nn.Linear(d_model * 2, d_model, bias=False),
        )
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.conv1 = nn.Conv2d(d_model * 2, d_model, kernel_size=3, bias=False, padding=1)
        self.bn1 = nn.BatchNorm2d(d_model)
        self.conv2 = nn.Conv2d(d_model, d_model, kernel_size=3, bias=False, padding=1)
        self.bn2 = nn.BatchNorm2d(d_model)
        self.relu = nn.ReLU()

    def pack(self, keep: Packer) -> _lib.CfpDapmW:
        if self.nhead != 4:
            raise _lib.CfpError("libcfp serves DAPM with 4 heads (fusion.py:13)")
        w = _lib.CfpDapmW()
        t = dict(
            wq_t=linear_t(self.q_proj.weight),
            wkv_t=torch.cat([linear_t(self.k_proj.weight), linear_t(self.v_proj.weight)], dim=1).contiguous(),
            tc=umma_block(self.q_proj.weight),
            kv_tc=torch.stack([umma_block(self.k_proj.weight), umma_block(self.v_proj.weight)]).contiguous(),
        )
        keep.extend(t.values())
        w.attn = _lib.CfpLoftrW(**{k: keep.ref(v) for k, v in t.items()})
        for i, (conv, bn) in enumerate(((self.conv1, self.bn1), (self.conv2, self.bn2)), start=1):
            scale, shift = fold_bn(bn)
            ws = host(conv.weight) * scale[:, None, None, None]                     # [Cout,Cin,3,3]
            cout, cin = ws.shape[0], ws.shape[1]
            wt = ws.permute(2, 3, 1, 0).reshape(9 * cin, cout).contiguous()
            # tensor-core blocks: one [Cout x Cout-wide K] block per (source, tap)
            pk = torch.stack([umma_block(ws[:, s0:s0 + cout, ky, kx])
                              for s0 in range(0, cin, cout) for ky in range(3) for kx in range(3)]).contiguous()
            shift = shift.contiguous()
            keep.extend((wt, shift, pk))
            setattr(w, f"conv{i}_t", keep.ref(wt))
            setattr(w, f"shift{i}", keep.ref(shift))
            setattr(w, f"conv{i}_pk", keep.ref(pk))
        return w


class LayerNorm(nn.Module):
    """channels_last LayerNorm container of convnext.py:60-72."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format != "channels_last":
            raise NotImplementedError
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps


class Block14(nn.Module):
    """LKPM, convnext.py:16-40.  ``conv1`` is never used by the reference's forward;
    gamma is None because layer_scale_init_value = 0."""

    def __init__(self, dim, drop_path=0., layer_scale_init_value=0, large_kernel=7):
        super().__init__()
        if drop_path > 0 or layer_scale_init_value > 0:
            raise NotImplementedError("libcfp serves Block14 as the reference instantiates it "
                                      "(drop_path=0, no layer scale; transformer.py:259)")
        self.dwconv2 = nn.Conv2d(dim, dim, kernel_size=large_kernel, padding=((large_kernel - 1) // 2), groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = None
        self.drop_path = nn.Identity()
        self.conv1 = nn.Conv2d(dim * 2, dim, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(dim)
        self.relu = nn.ReLU()

    def pack(self, keep: Packer) -> _lib.CfpLkpmW:
        k = self.dwconv2.kernel_size[0]
        scale, shift = fold_bn(self.bn1)
        C = scale.numel()
        taps = host(self.dwconv2.weight)[:, 0] * scale[:, None, None]                   # [C,k,k]
        t = dict(
            dw_t=taps.permute(1, 2, 0).reshape(k * k, C).contiguous(),
            dw_shift=(host(self.dwconv2.bias) * scale + shift).contiguous(),
            ln_g=host(self.norm.weight).contiguous(),
            ln_b=host(self.norm.bias).contiguous(),
            pw1_t=linear_t(self.pwconv1.weight),
            pw1_b=host(self.pwconv1.bias).contiguous(),
            pw2_t=linear_t(self.pwconv2.weight),
            pw2_b=host(self.pwconv2.bias).contiguous(),
        )
        # tensor-core MLP (csrc/k_chain_tc.cu: MlpTC): per 128-wide hidden slice j two blocks, W1_j then W2_j, each padded
        # to the ring-slot size.  The LayerNorm affine is folded into W1 / b1, and both biases ride in one extra 16-column
        # K step (first column = bias) that meets a constant ones-column of the A operand.  W1_j: bf16 [128][C+16];
        # W2_j: fp16 [C][128+16] (the GELU output feeds the second GEMM as fp16).
        g_ln, b_ln = host(self.norm.weight), host(self.norm.bias)
        w1 = host(self.pwconv1.weight)
        w2 = host(self.pwconv2.weight)
        b1 = host(self.pwconv1.bias) + w1 @ b_ln
        w1 = w1 * g_ln[None, :]
        b2 = host(self.pwconv2.bias)
        blk = max(128 * (C + 16) * 2, C * 144 * 2)          # bytes (MlpTC::BLK)
        blocks = []
        for j in range(4 * C // 128):
            sl = slice(j * 128, (j + 1) * 128)
            w1j = torch.cat([w1[sl, :], b1[sl, None], w1.new_zeros(128, 15)], dim=1)
            w2j = torch.cat([w2[:, sl], (b2 if j == 0 else torch.zeros_like(b2))[:, None], w2.new_zeros(C, 15)], dim=1)
            for blkw, dt in ((w1j, torch.bfloat16), (w2j, torch.float16)):
                raw = umma_block(blkw, dt).reshape(-1).view(torch.uint8)
                blocks.append(torch.cat([raw, raw.new_zeros(blk - raw.numel())]))
        t["tc"] = torch.cat(blocks).contiguous()
        # banded-Toeplitz blocks of the depthwise taps for the tensor-core path (csrc/k_dwconv_tc.cu):
        # T_dy[n][kk] = w[dy][kk - n]; vertical taps grouped as dy = 4 a + b with the four b's side by side along
        # the MMA's N dimension -> [C][a][k-step][k-group (2)][b*32 + n (128)][8] bf16 (UMMA K-major B blocks)
        nb = 4                                              # kDwNB in csrc/k_dwconv_tc.cu
        ks = (32 + k - 1 + 15) // 16
        na = (k + nb - 1) // nb
        n = torch.arange(32, device=taps.device)[:, None]
        kk = torch.arange(16 * ks, device=taps.device)[None, :]
        dx = kk - n
        band = ((dx >= 0) & (dx < k)).to(taps.dtype)
        toep = taps[:, :, dx.clamp(0, k - 1)] * band                      # [C, k(dy), 32, 16*ks]
        toep = torch.cat([toep, toep.new_zeros(C, nb * na - k, 32, 16 * ks)], dim=1)
        t["dw_toep"] = (toep.to(torch.bfloat16).view(C, na, nb, 32, ks, 2, 8).permute(0, 1, 4, 5, 2, 3, 6).contiguous())
        keep.extend(t.values())
        w = _lib.CfpLkpmW(**{n: keep.ref(v) for n, v in t.items()})
        w.ksize = k
        return w


    def forward(self, x):
        """Block14.forward (convnext.py:42-58) on an NCHW map, as a stand-alone module: train mode (BatchNorm on batch
        statistics, differentiable) through the fp32 training kernels of cfpnet_b200/train.py.  Inside
        ``TransformerFusion`` the eval-mode block is driven through ``cfp_lkpm_fwd`` on token-major maps instead."""
        _lib.require_cuda(x, "x")
        if not self.training:
            raise NotImplementedError("stand-alone Block14.forward serves train mode; eval mode runs inside "
                                      "TransformerFusion.forward (cfp_lkpm_fwd)")
        from .train import LkpmTrainFn
        return LkpmTrainFn.apply(self, x, *self.parameters())


class Combine1(nn.Module):
    """DAPM -> LKPM, transformer.py:251-259."""

    def __init__(self, d_model, nhead, large_kernel):
        super().__init__()
        self.transformer_path = LoFTREncoderLayer_newcross9(d_model, nhead)
        self.large_kernel_path = Block14(d_model, large_kernel=large_kernel)
