"""Final-depth parity (north_star: "final depth within 0.5 % abs-rel"): the reference's decoder shell and adaptive-bins
head (restated in oracle/, torch ops on the GPU - they are the CALLER of the path, not part of it) wrapped around
the CUDA HistogramEncoder + the three CUDA TransformerFusion modules, against the depth map the reference's own
Decoder / DepthRegression produced for the same weights and inputs (tests/golden/depth_G416_B1.npz)."""
import os

import numpy as np
import pytest
import torch

import cfpnet_b200
from cfpnet_b200 import synth
from cfpnet_b200.config import args
from oracle import cfp_oracle as O
from test_depth_golden import GOLDEN, MAX_VAL, MIN_VAL, tail_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LEVEL_OF = {"cross_atten3": 3, "cross_atten2": 2, "cross_atten1": 1}


# the harness around the path (decoder shell + head as torch ops) must not add an error of its own: cuDNN's default
# TF32 convolutions alone cost 2.5e-3 abs-rel here
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def cuda_modules(sd, dtype, hist_dtype=None):
    args.attention_layer = list(synth.COMBINE1_LAYERS)
    mods = {}
    for name, lv in LEVEL_OF.items():
        C, _, max_res, lk = synth.LEVELS[lv]
        m = cfpnet_b200.TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1])
        m.load_state_dict({k[len(f"decoder.{name}."):]: v for k, v in sd.items() if k.startswith(f"decoder.{name}.")}, strict=True)
        mods[name] = m.to(DEV).to(dtype).eval()
    enc = cfpnet_b200.HistogramEncoder()
    enc.load_state_dict({k[len("hist_encoder."):]: v for k, v in sd.items() if k.startswith("hist_encoder.")}, strict=True)
    enc = enc.to(DEV).eval()
    enc.out_dtype = hist_dtype or dtype
    return enc, mods


# Per-stage attribution (tools/depth_attrib.py on a B200, profiles/r2n_depth_attrib.txt): fp32 engine everywhere 8.6e-6;
# bf16 histogram encoder alone 1.5e-3; ANY one fusion level in bf16 2.4e-2 ... 2.9e-2; all bf16 5.2e-2 (the reference with
# its fusion modules cast to bf16: 4.6e-2).  The 256-bin softmax head at random init turns a feature error e into ~3 e of
# depth error, so the 0.5 % bound needs features good to ~1.5e-3 at EVERY level - out of reach of any engine whose GEMM
# operands are bf16 (unit round-off 4e-3).  The cheapest setting that meets the north-star bound is therefore the named
# mode "hist_bf16": bf16 histogram encoder (tcgen05) + the exact fp32 engine for the three fusion levels; the pure bf16
# engine is held to 1.5 x the reference's own bf16 deviation.
@pytest.mark.parametrize("mode,dtype,hist_dtype,tol", [("fp32", torch.float32, torch.float32, 1e-4),
                                                       ("hist_bf16", torch.float32, torch.bfloat16, 5e-3),
                                                       ("bf16", torch.bfloat16, torch.bfloat16, None)])
def test_final_depth_abs_rel(mode, dtype, hist_dtype, tol):
    saved = list(args.attention_layer)
    try:
        sd = tail_state()
        enc, mods = cuda_modules(sd, dtype, hist_dtype)
        inp = synth.make_inputs("G416", 1, seed=5, levels=())
        feats = [t.to(DEV) for t in synth.encoder_features("G416", 1, seed=5)]
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        with torch.no_grad():
            hist = enc(inp["hist_data"].to(DEV).unsqueeze(-1))

            def fuse(name, x, feat1):
                out = mods[name](x.to(dtype).contiguous(), feat1.to(dtype), rect_data=inp["rect_data"], mask=inp["mask"].to(DEV),
                                 patch_info=inp["patch_info"], rgb=None)
                return out.float()

            torch.manual_seed(2)
            unet = O.decoder_shell(O.sub(sdd, "decoder."), feats, hist, fuse)
            _, pred = O.depth_tail(sdd, unet, MIN_VAL, MAX_VAL)
        z = np.load(os.path.join(GOLDEN, "depth_G416_B1.npz"))
        gt = torch.from_numpy(z["pred"])
        if tol is None:
            tol = 1.5 * float(z["ref_bf16_abs_rel"])
        err = O.abs_rel(pred.cpu(), gt)
        print(f"final depth abs-rel ({mode}): {err:.3e}")
        assert err <= tol, f"final depth abs-rel {err:.3e} > {tol} ({mode})"
    finally:
        args.attention_layer = saved
