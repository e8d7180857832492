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
inp = synth.make_inputs("G416", B, seed=3)
d = {k: inp[k].to(dev) for k in ("hist_data", "mask")}
for k in ("x3", "x2", "x1"): d[k] = inp[k].to(dev, torch.bfloat16)
def sl(pi, lo, hi):
    return {kk: ({n: t[lo:hi] for n, t in vv.items()} if isinstance(vv, dict) else vv[lo:hi]) for kk, vv in pi.items()}
pi = inp["patch_info"]
streams = [torch.cuda.Stream() for _ in range(4)]
def run(nsplit):
    cur = torch.cuda.current_stream()
    if nsplit == 1:
        return path(d["x3"], d["x2"], d["x1"], d["hist_data"], d["mask"], pi)
    ev = torch.cuda.Event(); ev.record(cur)
    step = B // nsplit
    dones = []
    for i in range(nsplit):
        lo, hi = i * step, (i + 1) * step
        s = streams[i]; s.wait_event(ev)
        with torch.cuda.stream(s):
            path(d["x3"][lo:hi], d["x2"][lo:hi], d["x1"][lo:hi], d["hist_data"][lo:hi], d["mask"][lo:hi], sl(pi, lo, hi))
            e = torch.cuda.Event(); e.record(s); dones.append(e)
    for e in dones: cur.wait_event(e)
with torch.no_grad():
    for ns in (1, 2, 4, 1, 2):
        for _ in range(3): run(ns)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run(ns)
        e1.record(); torch.cuda.synchronize()
        print(f"micro-batches {ns}: {e0.elapsed_time(e1)/10:.2f} ms/step")
