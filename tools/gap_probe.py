"""Where does the non-kernel time of a step go?  (diagnostic, GPU box)"""
import sys, time
sys.path.insert(0, ".")
import torch
from cfpnet_b200 import FusionPath, synth, _lib

dev = torch.device("cuda:0")
path = FusionPath(synth.COMBINE1_LAYERS)
path.hist_encoder.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in path.hist_encoder.state_dict().items()}, 0))
for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
    m = getattr(path, name)
    m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, lv))
path = path.to(dev).eval().set_dtype(torch.bfloat16)
B = 64
inp = synth.make_inputs("G416", B, seed=3)
d = {k: inp[k].to(dev) for k in ("hist_data", "mask")}
for k in ("x3", "x2", "x1"):
    d[k] = inp[k].to(dev, torch.bfloat16)
pi = inp["patch_info"]

def step():
    return path(d["x3"], d["x2"], d["x1"], d["hist_data"], d["mask"], pi)

with torch.no_grad():
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    for trial in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(10):
            step()
        t_enq = time.perf_counter() - t0
        e1.record(); torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print(f"trial {trial}: enqueue {t_enq*100:.2f} ms/step, gpu {e0.elapsed_time(e1)/10:.2f} ms/step, wall {t_all*100:.2f} ms/step")
    # same under the event profiler
    _lib.profile_start()
    for _ in range(10):
        step()
    prof = _lib.profile_stop()
    print("profile sum ms/step", sum(v[1] for v in prof.values()) / 10)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    pass
    import subprocess
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap"
    for lms in (100, 500, 2000):
        proc = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", str(lms)],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        time.sleep(0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            torch.manual_seed(2 + i)
            step()
        e1.record(); torch.cuda.synchronize()
        proc.terminate()
        out = proc.stdout.read().strip().splitlines()
        print(f"nvidia-smi -lms {lms}: gpu {e0.elapsed_time(e1)/20:.2f} ms/step, {len(out)} samples, last {out[-1] if out else None}")
    e0.record()
    for i in range(20):
        torch.manual_seed(2 + i)
        step()
    e1.record(); torch.cuda.synchronize()
    print(f"no sampler + manual_seed: gpu {e0.elapsed_time(e1)/20:.2f} ms/step")
