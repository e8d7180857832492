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
    for depth in (2, 3, 4):
        t0 = time.perf_counter(); stamps = []
        for idx, outs in path.stream_host(batches(30), pi, dev, depth=depth):
            stamps.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
        d = [b - a for a, b in zip(stamps, stamps[1:])]
        print(f"depth {depth}: total {tot*1e3/30:.2f} ms/step; yield intervals (ms): " + " ".join(f"{x*1e3:.1f}" for x in d[:12]))
    # enqueue cost alone
    t0 = time.perf_counter()
    for i in range(10):
        d_ = sets[0]
        outs = path(*(d_[k].to(dev, non_blocking=True) for k in ("x3", "x2", "x1", "hist_data", "mask")), pi)
    t1 = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"plain enqueue incl H2D: {t1*100:.2f} ms/step host time")
