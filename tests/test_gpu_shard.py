"""N>1 path of the PRODUCT on a GPU: two ranks (one process each, gloo for the plumbing so that the test also runs on
a one-GPU box - both ranks then share cuda:0; with two or more devices rank r takes cuda:r) shard a batch with
cfpnet_b200.shard, run their frames through FusionPath -> libcfp, and the concatenated rank outputs must equal the
one-rank run of the whole batch (frames are independent, the positional-encoding crop is seeded
identically on every rank, there is no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cfpnet_b200 import shard, synth
from helpers import ref_keys, rel_l2

pytestmark = pytest.mark.gpu
BATCH = 5            # odd on purpose: ranks get 3 + 2 frames


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build_path(dev, dtype):
    from cfpnet_b200 import FusionPath
    path = FusionPath(synth.COMBINE1_LAYERS)
    path.hist_encoder.load_state_dict(synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0))
    for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
        getattr(path, name).load_state_dict(synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{lv}"], seed=lv))
    return path.to(dev).eval().set_dtype(dtype)


def _run(path, inp, dev, dtype, step):
    shard.seed_posenc(step)
    with torch.no_grad():
        outs = path(inp["x3"].to(dev, dtype), inp["x2"].to(dev, dtype), inp["x1"].to(dev, dtype), inp["hist_data"].to(dev),
                    inp["mask"].to(dev), inp["patch_info"])
    torch.cuda.synchronize(dev)
    return [o.float().cpu() for o in outs]


def _worker(rank, world, port, dtype_name, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dtype = getattr(torch, dtype_name)
    full = synth.make_inputs("G416", BATCH, seed=4)
    mine = shard.shard_inputs(full, rank, world)
    outs = _run(_build_path(dev, dtype), mine, dev, dtype, step=3)
    sizes = [shard.shard_range(BATCH, r, world) for r in range(world)]
    gathered = []
    for o in outs:                                   # variable-sized shares: pad to the largest, gather on rank 0
        pad = torch.zeros(max(hi - lo for lo, hi in sizes), *o.shape[1:])
        pad[: o.shape[0]] = o
        bucket = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bucket, dst=0)
        if rank == 0:
            gathered.append(torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bucket, sizes)]))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,tol", [("float32", 1e-5), ("bfloat16", 1e-2)])
def test_two_rank_outputs_equal_the_one_rank_run(dtype_name, tol):
    """fp32: only the order of the fp32 atomics of straddling attention groups differs between the batch sizes;
    bf16: the same plus one bf16 rounding of the affected rows (same bound as the frame-independence test)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dtype_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=600)          # raises queue.Empty instead of hanging when a rank died
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    dtype = getattr(torch, dtype_name)
    want = _run(_build_path(dev, dtype), synth.make_inputs("G416", BATCH, seed=4), dev, dtype, step=3)
    assert len(got) == len(want) == 3
    for g_, w_ in zip(got, want):
        assert g_.shape == w_.shape
        assert rel_l2(g_, w_) <= tol
