"""Host-side weight packing for libcfp.

The kernels want nn.Linear / conv weights transposed to [in][out] fp32 with
eval-mode BatchNorm folded in.  Packing is a one-off per parameter version
(cached), done with torch ops on the parameters' device; it is plumbing, not
part of the timed hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def fold_bn(bn: nn.modules.batchnorm._BatchNorm):
    """Eval-mode BN as y = x*scale + shift (fp32)."""
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def linear_t(lin_weight: torch.Tensor) -> torch.Tensor:
    """nn.Linear weight [out,in] -> [in,out] fp32 contiguous."""
    return lin_weight.detach().float().t().contiguous()


class PackCache:
    """Re-pack only when a parameter/buffer of `module` changed (in-place update,
    load_state_dict, .to(device))."""

    def __init__(self, module: nn.Module):
        object.__setattr__(self, "_module_ref", [module])     # not registered as a submodule
        self._key = None
        self._value = None

    def _current_key(self):
        m = self._module_ref[0]
        return tuple((t.data_ptr(), t._version, t.device.index) for t in
                     list(m.parameters()) + list(m.buffers()))

    def get(self, build):
        key = self._current_key()
        if key != self._key:
            with torch.no_grad():
                self._value = build()
            self._key = key
        return self._value


def umma_block(w: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """[N,K] weight -> 16-bit block in the canonical K-major no-swizzle UMMA layout
    [K/8][N][8] (csrc/umma.cuh): element (n,k) at ((k//8)*N + n)*8 + k%8."""
    n, k = w.shape
    assert k % 8 == 0
    return w.detach().to(dtype).contiguous().view(n, k // 8, 8).permute(1, 0, 2).contiguous()
