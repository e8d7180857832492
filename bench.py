#!/usr/bin/env python
"""Benchmark of the CFP fusion + cross-zone propagation path (BASELINE.json metric:
"CFP fusion frames/s @416x544, 8x8 zones").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cfp|reference]

One *step* = one pass of the hot path (histogram encoder + the three TransformerFusion
calls) over one batch of synthetic frames.  Workload = BASELINE.json configs[2]:
combine1 layer list, bf16, 64 frames per GPU, 416x544 input with 8x8 zones of 48 px.
Launched under torchrun for N>1: one process per GPU, frames sharded by batch, no
data-path collective (weak scaling); the time is the max over ranks.

``--impl reference`` times the reference's CPU algorithm (the oracle port of the PyTorch
modules, oracle/cfp_oracle.py) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cfpnet_b200 import synth  # noqa: E402

METRIC = "CFP fusion frames/s @416x544, 8x8 zones"
GEOMETRY = "G416"
# SURVEY.md §8d / BASELINE.md §3 (combine1, 416x544): algorithmic work per frame
FLOP_PER_FRAME = 13.22e9
DENSE_FLOP_PER_FRAME = 11.26e9
ELEMS_PER_FRAME = 15.43e6


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md).  The poller is started well before
    the timed region (its NVML start-up perturbs the GPU for a few hundred ms); only the samples
    that arrive between mark_begin() and mark_end() are summarised."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None
        self.t0 = self.t1 = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.perf_counter(), [v.strip() for v in line.split(",")]))
        except Exception:
            pass

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def finish(self):
        if self.proc is not None:
            self.proc.terminate()
        inside = [s for t, s in self.samples if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        if not inside:      # region shorter than one polling period: take the nearest samples
            inside = [s for _, s in self.samples[-2:]]
        sm = [float(s[0]) for s in inside if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in inside if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in inside:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- reference arm
def oracle_modules():
    from oracle import cfp_oracle as O
    from cfpnet_b200 import FusionPath
    path = FusionPath(synth.COMBINE1_LAYERS)
    sds = {"hist_encoder": synth.synthetic_state_dict({k: v.shape for k, v in path.hist_encoder.state_dict().items()}, 0)}
    for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
        sds[name] = synth.synthetic_state_dict({k: v.shape for k, v in getattr(path, name).state_dict().items()}, lv)
    return O, sds


def time_cpu_frames(n_frames, steps, warmup, seed=1):
    """Time the oracle port (fp32, all host threads) on `n_frames`-frame batches."""
    O, sds = oracle_modules()
    torch.set_num_threads(os.cpu_count() or 1)
    inp = synth.make_inputs(GEOMETRY, n_frames, seed=seed)
    xs = [inp["x3"], inp["x2"], inp["x1"]]
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            torch.manual_seed(2)
            t0 = time.perf_counter()
            O.fusion_path(sds, synth.COMBINE1_LAYERS, xs, inp["hist_data"], inp["mask"], inp["patch_info"])
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    probe = time_cpu_frames(1, 1, 1)[0]
    budget_s = 90.0
    n = max(1, min(64, int(budget_s / ((a.steps + a.warmup) * max(probe, 1e-3)))))
    times = time_cpu_frames(n, a.steps, a.warmup)
    total = sum(times)
    fps = n * a.steps / total
    sample = f"{n} frames/step x {a.steps} steps (+{a.warmup} warm-up), fp32, torch CPU ops, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, per_gpu_batch=n),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(a, per_gpu_batch):
    return {"workload": "CFPNet combine1 fusion path (hist encoder + cross_atten3/2/1), 416x544, 8x8 zones of 48 px, "
                        f"{per_gpu_batch} frames per GPU, batch-sharded (BASELINE.json configs[2])",
            "layers": list(synth.COMBINE1_LAYERS), "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * a.gpus,
            "l2_policy": "inputs rotate over 3 distinct batches + workspace > 126 MB L2 between reuse"}


# ---------------------------------------------------------------------------------- product arm
def run_cfp(a):
    import torch.distributed as dist
    from cfpnet_b200 import FusionPath, _lib, shard
    from cfpnet_b200.build import build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl cfp needs a CUDA device: the fusion path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build()
    dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[a.dtype]
    B = a.batch

    path = FusionPath(synth.COMBINE1_LAYERS)
    path.hist_encoder.load_state_dict(synth.synthetic_state_dict(
        {k: v.shape for k, v in path.hist_encoder.state_dict().items()}, 0))
    for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
        m = getattr(path, name)
        m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, lv))
    path = path.to(dev).eval().set_dtype(dtype)

    # three distinct input batches per rank (rotated so consecutive steps never reuse L2 contents)
    NSETS = 3
    host_sets, dev_sets = [], []
    for s in range(NSETS):
        inp = synth.make_inputs(GEOMETRY, B, seed=100 + rank * NSETS + s)
        h = {"x3": inp["x3"].to(dtype).pin_memory(), "x2": inp["x2"].to(dtype).pin_memory(),
             "x1": inp["x1"].to(dtype).pin_memory(), "hist_data": inp["hist_data"].pin_memory(),
             "mask": inp["mask"].pin_memory()}
        host_sets.append(h)
        dev_sets.append({k: v.to(dev) for k, v in h.items()})
        patch_info = inp["patch_info"]
    h2d = sum(v.numel() * v.element_size() for v in host_sets[0].values())

    def step(i):
        d = dev_sets[i % NSETS]
        shard.seed_posenc(i)                # same positional-encoding crop on every rank
        return path(d["x3"], d["x2"], d["x1"], d["hist_data"], d["mask"], patch_info)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None
    sampler_t = time.perf_counter()
    if sampler:
        sampler.start()
    with torch.no_grad():
        for i in range(a.warmup):
            outs = step(i)
        d2h = sum(o.numel() * o.element_size() for o in outs)
        barrier()
        if sampler:                         # let nvidia-smi finish its start-up before timing (rank 0 only)
            while not sampler.samples and sampler.is_alive() and time.perf_counter() - sampler_t < 5.0:
                time.sleep(0.05)
        for i in range(2):                  # every rank: same sequence of steps and collectives
            step(i)
        barrier()
        # ---- device-resident throughput ("value")
        if sampler:
            sampler.mark_begin()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            step(i)
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        if sampler:
            sampler.mark_end()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.finish() if sampler else None

        # ---- end-to-end through the public host-buffer API ("e2e"): pinned host inputs in, pinned host
        # outputs back, every step; FusionPath.stream_host overlaps the copies of neighbouring steps
        # with compute on separate streams (each step still pays its own H2D + D2H).
        def host_batches(n):
            for i in range(n):
                yield host_sets[i % NSETS]

        for _ in path.stream_host(host_batches(max(a.warmup, 3)), patch_info, dev, seeds=range(2, 2 + max(a.warmup, 3))):
            pass
        barrier()
        t_wall = time.perf_counter()
        e0.record()
        n_out = 0
        for _idx, _outs in path.stream_host(host_batches(a.steps), patch_info, dev, seeds=range(2, 2 + a.steps)):
            n_out += 1
        e1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        assert n_out == a.steps
        # the last D2H completes on a side stream: take the larger of the device-event and wall-clock spans
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), t_wall * 1e3))

        # ---- per-kernel breakdown with CUDA events on the launch stream (roofline)
        prof = None
        if rank == 0:
            torch.cuda.synchronize()
            path.concurrent_levels = False      # one stream: the gap between events is one kernel's time
            _lib.profile_start()
            for i in range(a.steps):
                step(i)
            prof = _lib.profile_stop()
            path.concurrent_levels = True
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    frames = B * world * a.steps
    value = frames / (ms_total * 1e-3)
    es = 2 if dtype == torch.bfloat16 else 4
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.dtype, "data": "synthetic", "config": workload_config(a, B),
        "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    # whole-path roofline view (algorithmic bytes / dense flops per frame x measured frames/s)
    per_gpu_fps = value / world
    line["path"] = {"algorithmic_GBps": ELEMS_PER_FRAME * es * per_gpu_fps / 1e9,
                    "dense_TFLOPs": DENSE_FLOP_PER_FRAME * per_gpu_fps / 1e12,
                    "hbm_peak_GBps": peaks["hbm_gbs"], "bf16_peak_TFLOPs": peaks["bf16_tflops_sustained"]}
    if prof:
        line["roofline"], line["kernels"] = roofline_from_profile(prof, a.steps, B, es, peaks)
    if not a.no_cpu:
        cores = os.cpu_count() or 1
        t = time_cpu_frames(1, 10, 3)
        med = statistics.median(t)
        line["cpu_baseline"] = {"value": 1.0 / med, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "1 frame/step, 3 warm-up + 10 timed, median, fp32 torch CPU ops on all host threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# Algorithmic work per LAUNCH of each kernel at G416 (derivations: DESIGN.md §4).
#   flops  = useful multiply-adds x 2 of the op as the reference defines it (no padding / Toeplitz zeros)
#   bytes  = compulsory HBM traffic: every activation the op reads + writes once, weights amortised
LEVEL = {  # C: (N tokens, k, zone patch side p, window ws, inside Ni)
    32: dict(N=14144, k=31, p=12, ws=12, Ni=9216, H=104, W=136),
    64: dict(N=3536, k=15, p=6, ws=9, Ni=2304, H=52, W=68),
    128: dict(N=884, k=7, p=3, ws=6, Ni=576, H=26, W=34),
}


def kernel_work(name, B, es):
    import re
    m = re.search(r"(\d+)>$", name)
    if not m:
        return None
    v = int(m.group(1))
    if name.startswith("dwconv"):                       # dwconv<k> / dwconv_tc<k>: v is the kernel size
        C = {31: 32, 15: 64, 7: 128}[v]
        g = LEVEL[C]
        return dict(flops=2.0 * g["N"] * C * v * v * B, bytes=2.0 * g["N"] * C * es * B, bound="tensor" if "_tc" in name else "fma")
    C = v
    g = LEVEL[C]
    N, Ni, p, ws = g["N"], g["Ni"], g["p"], g["ws"]
    No = N - Ni
    nwin = -(-g["H"] // ws) * -(-g["W"] // ws)
    Ns = (g["H"] // ws) * (g["W"] // ws)
    chain = 16.0 * C * C                                # q, merge, mlp0 (2Cx2C), mlp2 (2C->C): flops per query row
    rows = {"hist2image": 64 * p * p, "lsa": nwin * ws * ws, "gsa": N, "dapm": No}
    src = {"hist2image": 64 * 16, "lsa": nwin * ws * ws, "gsa": Ns, "dapm": Ni}
    heads = {"hist2image": 4, "lsa": 8, "gsa": 8, "dapm": 4}
    for kind in rows:
        if f"<{kind}," in name:
            dh = C // heads[kind]
            if name.startswith("loftr_query"):
                return dict(flops=(chain + 2.0 * C * dh) * rows[kind] * B, bytes=2.0 * rows[kind] * C * es * B, bound="tensor")
            if name.startswith("attn_query"):
                return dict(flops=(2.0 * C * C + 2.0 * C * dh) * rows[kind] * B, bytes=2.0 * rows[kind] * C * es * B, bound="tensor")
            if name.startswith("kv_state"):
                return dict(flops=(4.0 * C * C + 2.0 * C * dh) * src[kind] * B, bytes=src[kind] * C * es * B, bound="tensor")
    if name.startswith("lkpm_mlp"):
        return dict(flops=16.0 * N * C * C * B, bytes=3.0 * N * C * es * B, bound="tensor")
    if name.startswith("conv3x3"):
        cin = 2 * C if "2C->C" in name else C
        return dict(flops=2.0 * N * cin * 9 * C * B, bytes=(cin + C + (C if cin == C else 0)) * N * es * B, bound="tensor")
    if name.startswith("sr_conv"):
        return dict(flops=2.0 * Ns * ws * ws * C * C * B, bytes=N * C * es * B, bound="tensor")
    return None


# dram bytes per launch of the top kernels from the committed `ncu --set full` captures (profiles/*.txt)
NCU_TRAFFIC = {}


def roofline_from_profile(prof, steps, B, es, peaks):
    total_ms = sum(v[1] for v in prof.values())
    kernels = {}
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        rec = {"launches_per_step": v[0] / steps, "ms_per_step": v[1] / steps, "share": v[1] / total_ms if total_ms else None}
        w = kernel_work(k, B, es)
        if w:
            avg_s = v[1] / v[0] * 1e-3
            rec["TFLOPs"] = w["flops"] / avg_s / 1e12
            rec["GBps"] = w["bytes"] / avg_s / 1e9
        kernels[k] = rec
    name, (count, ms) = max(prof.items(), key=lambda kv: kv[1][1])
    avg_s = ms / count * 1e-3
    w = kernel_work(name, B, es)
    roof = {"kernel": name, "avg_launch_ms": avg_s * 1e3, "peak_source": peaks["source"], "traffic": NCU_TRAFFIC.get(name)}
    if w and w["bound"] == "tensor":
        ach = w["flops"] / avg_s / 1e12
        roof.update(bound="tensor", achieved=ach, peak=peaks["bf16_tflops_sustained"], unit="TFLOP/s",
                    frac=ach / peaks["bf16_tflops_sustained"],
                    note="algorithmic (useful) flops of the op / CUDA-event time; peak = sustained cuBLAS bf16 (kernel timed inside a long step)",
                    hbm={"achieved_GBps": w["bytes"] / avg_s / 1e9, "peak_GBps": peaks["hbm_gbs"],
                         "frac": w["bytes"] / avg_s / 1e9 / peaks["hbm_gbs"]})
    elif w:
        ach = w["bytes"] / avg_s / 1e9
        roof.update(bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"])
    else:
        roof.update(bound="hbm", achieved=None, peak=peaks["hbm_gbs"], unit="GB/s", frac=None)
    return roof, kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cfp", choices=["cfp", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "cfp" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_cfp(a)


if __name__ == "__main__":
    main()
