"""Batch sharding of the fusion path across the GPUs of one node.

Frames are independent in eval mode, so rank r of N simply takes frames
[r*B/N, (r+1)*B/N); weights are replicated and there is NO data-path collective.
The one thing ranks must agree on is the positional-encoding crop, which the reference
draws once per ``TransformerFusion`` call from the global CPU generator
(``fusion.py:88-91``): every rank seeds it identically before a forward.

``GradientAllreduce`` (below) is the host side of the training step's gradient exchange - the
only collective of the path; its producers (backward kernels) are a later row.
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


# ---------------------------------------------------------------------------------------------
# Training collective (BASELINE config 5): the one place the path has a
# real exchange step.  One process per GPU computes the gradients of its frames (BatchNorm uses
# per-replica batch statistics, as under the reference's nn.DataParallel, train.py:45), then the
# gradients are summed over ranks and divided by N.  The backward kernels are not built yet; this
# is the host plumbing they will feed, exercised on CPU with gloo (tests/test_multi_rank.py).
# ---------------------------------------------------------------------------------------------
class GradientAllreduce:
    """Averages ``.grad`` of a fixed parameter list over the ranks of a process group.

    * Parameters the forward never uses (the reference registers DAPM's merge / mlp / norms,
      LKPM's conv1, ... and never touches them; 16 to 28 tensors per ``TransformerFusion``) have
      ``grad is None`` on every rank: they are left out of the exchange and stay ``None``, as the
      reference's optimizer sees them.  The set must be the same on all ranks - checked with one
      tiny collective per call.
    * Gradients travel in flat buckets per dtype.  NVSwitch gives every GPU full bandwidth to every
      peer and reduces in the switch, so buckets are sized for launch latency, not link count: the
      whole hot path (6.65 M parameters, 26.6 MB fp32) fits the default 64 MB bucket - one
      all-reduce per step.
    """

    def __init__(self, params, group=None, bucket_bytes: int = 64 << 20):
        import torch.distributed as dist
        self.dist = dist
        self.params = [p for p in params]
        self.group = group
        self.bucket_bytes = int(bucket_bytes)
        self.collectives = 0                 # bucket all-reduces of the last call (the set check is one more)

    def _signature(self):
        return tuple(i for i, p in enumerate(self.params) if p.grad is not None)

    def _check_same_set(self, sig, device):
        # (count, index sum, index square sum) of the used set and their negatives: one MAX all-reduce yields max and
        # -min, which must agree.  Run on EVERY call: a rank cannot know that another rank's set changed, and
        # bucket all-reduces of different sizes hang (NCCL) instead of failing.
        s = torch.tensor([len(sig), sum(sig), sum(i * i for i in sig)], dtype=torch.int64, device=device)
        both = torch.cat([s, -s])
        self.dist.all_reduce(both, op=self.dist.ReduceOp.MAX, group=self.group)
        if not torch.equal(both[:3], -both[3:]):
            raise RuntimeError("ranks disagree on which parameters received a gradient; the gradient exchange "
                               "would pair different tensors (mine: %d of %d)" % (len(sig), len(self.params)))

    @torch.no_grad()
    def __call__(self) -> int:
        """All-reduce (sum) and divide by the world size, in place.  Returns the number of elements exchanged."""
        world = self.dist.get_world_size(self.group)
        sig = self._signature()
        self.collectives = 0
        self._check_same_set(sig, self.params[0].device)     # before any early return: every rank takes part
        if not sig:
            return 0
        grads = [self.params[i].grad for i in sig]
        total = 0
        by_dtype = {}
        for g in grads:
            by_dtype.setdefault((g.dtype, g.device), []).append(g)
        for (_dt, _dev), gs in by_dtype.items():
            start = 0
            while start < len(gs):                   # greedy buckets of at most bucket_bytes (at least one tensor)
                end, size = start, 0
                while end < len(gs) and (end == start or size + gs[end].numel() * gs[end].element_size() <= self.bucket_bytes):
                    size += gs[end].numel() * gs[end].element_size()
                    end += 1
                flat = torch.cat([g.reshape(-1) for g in gs[start:end]])
                self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group)
                self.collectives += 1
                flat.div_(world)
                off = 0
                for g in gs[start:end]:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
                total += off
                start = end
        return total
