"""The step around the training kernels (cfpnet_b200.train.FlatTrainer: flat gradient bucket -> all-reduce ->
cfp_tr_sumsq + cfp_tr_adamw) against torch's own clip_grad_norm_ + AdamW on the same gradients (what the reference's loop
calls, train.py:124-129), and the two-rank gradient exchange against the serial average."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cfpnet_b200 import synth
from cfpnet_b200.layers import Block14
from cfpnet_b200.train import FlatTrainer
from helpers import rel_l2

pytestmark = pytest.mark.gpu
LEVEL = 3


def _block(dev):
    C, _, _, k = synth.LEVELS[LEVEL]
    blk = Block14(C, large_kernel=k)
    blk.load_state_dict(synth.synthetic_state_dict({kk: v.shape for kk, v in blk.state_dict().items()}, 3))
    return blk.to(dev).train()


def _inputs(seed, batch, dev):
    C = synth.LEVELS[LEVEL][0]
    H, W = synth.level_hw("G416", LEVEL)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, C, H, W, generator=g).to(dev)
    ct = (torch.randn(batch, C, H, W, generator=g) * 3e-3).to(dev)
    return x, ct


def _backward(blk, x, ct):
    out = blk(x.clone().requires_grad_(True))
    torch.autograd.backward([out], [ct])


def test_flat_trainer_matches_torch_clip_and_adamw():
    dev = torch.device("cuda", 0)
    blk = _block(dev)
    lr_of = lambda p: 1e-3 if p.dim() > 1 else 1e-4          # noqa: E731  two learning-rate groups, interleaved
    tr = FlatTrainer([blk], weight_decay=0.1, max_norm=0.1, lr_of=lr_of)
    shadow, opt = None, None
    for it in range(3):
        x, ct = _inputs(10 + it, 2, dev)
        tr.zero_grad()
        _backward(blk, x, ct)
        if it == 0:
            used = [p for p in blk.parameters() if p.grad is not None]
            unused = [p for p in blk.parameters() if p.grad is None]
            assert unused, "Block14 registers conv1, which its forward never uses"
            before_unused = [p.detach().clone() for p in unused]
            shadow = [p.detach().clone().requires_grad_(True) for p in used]
            opt = torch.optim.AdamW([{"params": [s], "lr": lr_of(s)} for s in shadow], weight_decay=0.1)
            tr.adopt()
        for s, p in zip(shadow, used):
            s.grad = p.grad.detach().clone()
        gn = torch.nn.utils.clip_grad_norm_(shadow, 0.1)
        opt.step()
        tr.exchange()
        tr.update()
        torch.cuda.synchronize()
        assert abs(float(tr.sumsq[0].sqrt()) - float(gn)) <= 1e-4 * float(gn)
        for s, p in zip(shadow, used):
            assert rel_l2(p, s) <= 1e-5, it
    for p, b in zip(unused, before_unused):
        assert p.grad is None and torch.equal(p, b)          # never-used parameters: untouched, as the reference's optimizer leaves them
    # the parameters now live in one flat buffer and the gradients in another
    assert used[0].data_ptr() == tr.flat_p.data_ptr() and used[0].grad.data_ptr() == tr.flat_g.data_ptr()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"            # one-GPU box: both ranks share cuda:0, gloo carries the bucket
    dev = torch.device("cuda", rank % ngpu)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    blk = _block(dev)
    tr = FlatTrainer([blk], max_norm=0.1)
    for it in range(2):
        x, ct = _inputs(100 + 10 * it + rank, 2, dev)
        tr.zero_grad()
        _backward(blk, x, ct)
        if it == 0:
            tr.adopt()
        tr.exchange()
        tr.update()
    torch.cuda.synchronize()
    if backend == "nccl":
        bucket = [torch.zeros_like(tr.flat_p) for _ in range(world)]
        dist.all_gather(bucket, tr.flat_p.detach())
        gathered = [b.cpu() for b in bucket]
    else:
        flat = tr.flat_p.detach().cpu()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((backend, [g.clone() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_equals_serial_average():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    backend, got = q.get(timeout=600)
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert torch.equal(got[0], got[1]), "replicas diverged after the averaged update"
    # serial emulation on one device: per-rank gradients from replicas with per-replica BatchNorm statistics, averaged
    dev = torch.device("cuda", 0)
    reps = [_block(dev) for _ in range(world)]
    trs = [FlatTrainer([b], max_norm=0.1) for b in reps]
    for it in range(2):
        for r in range(world):
            x, ct = _inputs(100 + 10 * it + r, 2, dev)
            trs[r].zero_grad()
            _backward(reps[r], x, ct)
            if it == 0:
                trs[r].adopt()
        avg = sum(t.flat_g for t in trs) / world
        for t in trs:
            t.flat_g.copy_(avg)
            t.update()
    torch.cuda.synchronize()
    assert rel_l2(got[0], trs[0].flat_p.cpu()) <= 1e-5, backend      # fp32 atomics order in the weight-gradient reductions, through Adam
