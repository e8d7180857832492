"""Batch sharding of the fusion path across the GPUs of one node.

Frames are independent in eval mode, so rank r of N simply takes frames
[r*B/N, (r+1)*B/N); weights are replicated and there is NO data-path collective.
The one thing ranks must agree on is the positional-encoding crop, which the reference
draws once per ``TransformerFusion`` call from the global CPU generator
(``fusion.py:88-91``): every rank seeds it identically before a forward.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``batch % world`` ranks get one extra frame."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(inputs: Dict, rank: int, world: int) -> Dict:
    """Slice every per-frame tensor (and the collated ``patch_info``) of a batch dict."""
    batch = inputs["mask"].shape[0]
    lo, hi = shard_range(batch, rank, world)
    out = {}
    for k, v in inputs.items():
        if k == "patch_info":
            out[k] = {kk: ({n: t[lo:hi] for n, t in vv.items()} if isinstance(vv, dict) else vv[lo:hi])
                      for kk, vv in v.items()}
        elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == batch:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def seed_posenc(step: int, base_seed: int = 2) -> None:
    """Same crop offsets on every rank for step ``step`` (call right before the forward)."""
    torch.manual_seed(base_seed + step)
