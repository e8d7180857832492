"""Which stage turns the bf16 feature error into final-depth error?  Runs the final-depth harness of
tests/test_gpu_depth.py with every combination of per-level engines (fp32 / bf16 for cross_atten3 / 2 / 1 and the
histogram encoder) and prints abs-rel against the reference's depth map (tests/golden/depth_G416_B1.npz)."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cfpnet_b200  # noqa: E402
from cfpnet_b200 import synth  # noqa: E402
from cfpnet_b200.config import args  # noqa: E402
from oracle import cfp_oracle as O  # noqa: E402
from test_depth_golden import GOLDEN, MAX_VAL, MIN_VAL, tail_state  # noqa: E402

DEV = "cuda:0"
# the harness (decoder shell + head in torch ops) must not add its own error: no TF32 in cuDNN / cuBLAS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
LEVEL_OF = {"cross_atten3": 3, "cross_atten2": 2, "cross_atten1": 1}
args.attention_layer = list(synth.COMBINE1_LAYERS)
sd = tail_state()
sdd = {k: v.to(DEV) for k, v in sd.items()}
z = np.load(os.path.join(GOLDEN, "depth_G416_B1.npz"))
gt = torch.from_numpy(z["pred"])
inp = synth.make_inputs("G416", 1, seed=5, levels=())
feats = [t.to(DEV) for t in synth.encoder_features("G416", 1, seed=5)]


def build(name, dtype):
    lv = LEVEL_OF[name]
    C, _, max_res, lk = synth.LEVELS[lv]
    m = cfpnet_b200.TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1])
    m.load_state_dict({k[len(f"decoder.{name}."):]: v for k, v in sd.items() if k.startswith(f"decoder.{name}.")}, strict=True)
    return m.to(DEV).to(dtype).eval()


mods = {(n, dt): build(n, dt) for n in LEVEL_OF for dt in (torch.float32, torch.bfloat16)}
enc = cfpnet_b200.HistogramEncoder()
enc.load_state_dict({k[len("hist_encoder."):]: v for k, v in sd.items() if k.startswith("hist_encoder.")}, strict=True)
enc = enc.to(DEV).eval()
S = {torch.float32: "f32 ", torch.bfloat16: "bf16"}
print("ref bf16 abs-rel (reference modules cast to bf16):", float(z["ref_bf16_abs_rel"]))
print("hist  L3    L2    L1    abs-rel")
combos = [c for c in itertools.product((torch.float32, torch.bfloat16), repeat=4) if sum(d == torch.bfloat16 for d in c[1:]) <= 1 or all(d == torch.bfloat16 for d in c[1:])]
for he, d3, d2, d1 in combos:
    dts = {"cross_atten3": d3, "cross_atten2": d2, "cross_atten1": d1}
    with torch.no_grad():
        hist = {}
        for dt in (torch.float32, torch.bfloat16):
            enc.out_dtype = he
            hist[dt] = [h.to(dt) for h in enc(inp["hist_data"].to(DEV).unsqueeze(-1))]

        def fuse(name, x, feat1):
            dt = dts[name]
            lv = LEVEL_OF[name]
            f1 = hist[dt][{3: 2, 2: 1, 1: 0}[lv]]
            out = mods[(name, dt)](x.to(dt).contiguous(), f1, rect_data=inp["rect_data"], mask=inp["mask"].to(DEV),
                                   patch_info=inp["patch_info"], rgb=None)
            return out.float()

        torch.manual_seed(2)
        unet = O.decoder_shell(O.sub(sdd, "decoder."), feats, hist[torch.float32], fuse)
        _, pred = O.depth_tail(sdd, unet, MIN_VAL, MAX_VAL)
    print(S[he], S[d3], S[d2], S[d1], f"{O.abs_rel(pred.cpu(), gt):.3e}")
