"""N>1 path on CPU: two gloo ranks shard a batch, run their share (through the CPU oracle — the
only thing that can compute without a GPU), and the gathered result must equal the single-process
run bit for bit.  Covers shard_range/shard_inputs, the shared positional-encoding seed, and the
max-over-ranks timing reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cfpnet_b200 import shard, synth
from helpers import ref_keys
from oracle import cfp_oracle as O

LEVEL, BATCH = 3, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_share(inp, step):
    C, _, max_res, _ = synth.LEVELS[LEVEL]
    sd = synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{LEVEL}"], seed=LEVEL)
    hsd = synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0)
    feats = O.hist_encoder(hsd, inp["hist_data"])
    shard.seed_posenc(step)
    return O.transformer_fusion(sd, synth.COMBINE1_LAYERS, max_res, inp[f"x{LEVEL}"], feats[2], inp["mask"],
                                inp["patch_info"])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    full = synth.make_inputs("G416", BATCH, seed=4, levels=(LEVEL,))
    mine = shard.shard_inputs(full, rank, world)
    out = _run_share(mine, step=3)
    # gather variable-sized shares on rank 0
    sizes = [shard.shard_range(BATCH, r, world) for r in range(world)]
    pad = torch.zeros(max(hi - lo for lo, hi in sizes), *out.shape[1:])
    pad[: out.shape[0]] = out
    bucket = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bucket, dst=0)
    # the timing reduction of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        got = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bucket, sizes)])
        q.put((got, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, tmax = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    full = synth.make_inputs("G416", BATCH, seed=4, levels=(LEVEL,))
    want = _run_share(full, step=3)
    assert tmax == float(world)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_shard_range_is_a_partition():
    for batch in (1, 5, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def test_shard_inputs_slices_patch_info():
    full = synth.make_inputs("G416", 6, levels=(3,))
    part = shard.shard_inputs(full, 1, 4)
    lo, hi = shard.shard_range(6, 1, 4)
    assert part["x3"].shape[0] == hi - lo and torch.equal(part["mask"], full["mask"][lo:hi])
    assert part["patch_info"][16]["pad_size"].shape[0] == hi - lo
    assert part["patch_info"]["zone_num"].shape[0] == hi - lo


# ------------------------------------------------------------------ training collective
def _train_grads(inp, sd, hsd):
    """One train-mode forward + backward of this share through the oracle (per-replica BatchNorm statistics, as
    under the reference's DataParallel); gradients land in the leaves' .grad"""
    C, _, max_res, _ = synth.LEVELS[LEVEL]
    feats = O.hist_encoder(hsd, inp["hist_data"].double(), bn_stats={})
    shard.seed_posenc(0)
    out = O.transformer_fusion(sd, synth.COMBINE1_LAYERS, max_res, inp[f"x{LEVEL}"].double(), feats[2], inp["mask"],
                               inp["patch_info"], bn_stats={})
    ct = torch.randn(out.shape, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    (out * ct).sum().backward()


def _train_leaves():
    def leaves(sd):
        return {k: (v.double().requires_grad_(True)
                    if v.is_floating_point() and not k.endswith(("running_mean", "running_var")) else v)
                for k, v in sd.items()}
    sd = leaves(synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{LEVEL}"], seed=LEVEL))
    hsd = leaves(synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0))
    params = [v for d in (sd, hsd) for v in d.values() if v.requires_grad]
    return sd, hsd, params


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    full = synth.make_inputs("G416z6", 4, seed=6, levels=(LEVEL,))
    sd, hsd, params = _train_leaves()
    _train_grads(shard.shard_inputs(full, rank, world), sd, hsd)
    exchange = shard.GradientAllreduce(params, bucket_bytes=4 << 20)       # several buckets: 31 MB of fp64 gradients
    n = exchange()
    unused, buckets = sum(p.grad is None for p in params), exchange.collectives
    # a rank whose used-parameter set differs must be caught, not silently paired with other tensors
    mismatch_caught = False
    if rank == 1:
        params[0].grad = None
    try:
        exchange()
    except RuntimeError as e:
        mismatch_caught = "disagree" in str(e)
    flags = torch.tensor([float(mismatch_caught)])
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(([None if p.grad is None else p.grad.numpy().copy() for p in params[1:]], n, unused, buckets,   # by value
               float(flags)))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks():
    """sum over ranks / N of per-replica gradients == the same two shares computed one after the other; parameters
    the forward never uses stay grad=None and out of the exchange; mismatching sets are refused on every rank."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, n, unused, collectives, caught = q.get()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    full = synth.make_inputs("G416z6", 4, seed=6, levels=(LEVEL,))
    sd, hsd, params = _train_leaves()
    for r in range(world):                       # .grad accumulates the two shares
        _train_grads(shard.shard_inputs(full, r, world), sd, hsd)
    want = [None if p.grad is None else p.grad / world for p in params]
    assert unused == sum(w is None for w in want) and unused >= 16
    assert n == sum(w.numel() for w in want if w is not None)
    assert collectives > 1 and caught == 1.0
    for g, w in zip(got, want[1:]):
        assert (g is None) == (w is None)
        if w is not None:
            # atol: biases in front of a train-mode BatchNorm have an exactly-zero gradient (rounding noise ~1e-11)
            assert torch.allclose(torch.from_numpy(g), w, rtol=1e-8, atol=1e-9)
