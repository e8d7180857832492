"""Run the bench workload (combine1, bf16, 64 frames, 416x544) for a few steps with the three levels on ONE
stream - the target of ncu captures of a whole step (tooling).  usage: prof_step.py [steps] [B]

Prints the number of libcfp launches per step so that ``--launch-skip`` / ``--launch-count`` can select one step:
    ncu --set full --import-source on --clock-control none --launch-skip <n> --launch-count <n> ... prof_step.py 2
"""
import os
import sys

os.environ["CFP_SEQUENTIAL_LEVELS"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cfpnet_b200 import FusionPath, _lib, shard, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
path = FusionPath(synth.COMBINE1_LAYERS)
path.hist_encoder.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in path.hist_encoder.state_dict().items()}, 0))
for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
    m = getattr(path, name)
    m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, lv))
path = path.eval().set_dtype(torch.bfloat16).to(dev)       # cast on the host: no ATen cast kernels in front of the capture window
inp = synth.make_inputs("G416", B, seed=100)
d = {k: (inp[k].to(torch.bfloat16) if k.startswith("x") else inp[k]).to(dev) for k in ("x3", "x2", "x1", "hist_data", "mask")}
with torch.no_grad():
    for i in range(steps):
        n0 = _lib.launch_count()
        shard.seed_posenc(i)
        path(d["x3"], d["x2"], d["x1"], d["hist_data"], d["mask"], inp["patch_info"])
        torch.cuda.synchronize()
        print(f"step {i}: {_lib.launch_count() - n0} libcfp launches")
