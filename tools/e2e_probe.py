import sys, time
sys.path.insert(0, ".")
import torch
from cfpnet_b200 import FusionPath, synth
dev = torch.device("cuda:0")
path = FusionPath(synth.COMBINE1_LAYERS)
path.hist_encoder.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in path.hist_encoder.state_dict().items()}, 0))
for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
    m = getattr(path, name)
    m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, lv))
path = path.to(dev).eval().set_dtype(torch.bfloat16)
B = 64
sets = []
for s in range(3):
    inp = synth.make_inputs("G416", B, seed=3 + s)
    sets.append({"x3": inp["x3"].bfloat16().pin_memory(), "x2": inp["x2"].bfloat16().pin_memory(), "x1": inp["x1"].bfloat16().pin_memory(),
                 "hist_data": inp["hist_data"].pin_memory(), "mask": inp["mask"].pin_memory()})
pi = inp["patch_info"]
def batches(n):
    for i in range(n):
        yield sets[i % 3]
with torch.no_grad():
    for _ in path.stream_host(batches(4), pi, dev): pass
    torch.cuda.synchronize()
    for trial in range(6):
        ms0 = torch.cuda.memory_stats(dev)
        t0 = time.perf_counter(); stamps = []
        for idx, outs in path.stream_host(batches(40), pi, dev):
            stamps.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
        ms1 = torch.cuda.memory_stats(dev)
        d = [b - a for a, b in zip(stamps, stamps[1:])]
        big = [(i, round(x * 1e3, 1)) for i, x in enumerate(d) if x > 0.012]
        print(f"trial {trial}: {tot*1e3/40:.2f} ms/step; first yield {stamps[0]*1e3:.1f} ms; gaps>12ms: {big}; dev allocs {ms1['num_device_alloc']-ms0['num_device_alloc']}")
