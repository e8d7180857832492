"""Run the decoder + fusion + head workload (bench.py --workload tail_b16) for a few steps - the target of ncu captures
of the f1 / f2 kernels (tooling).  usage: prof_tail.py [steps] [B]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cfpnet_b200  # noqa: E402
from cfpnet_b200 import _lib, decoder as D, shard, synth  # noqa: E402
from cfpnet_b200.config import args as cargs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
cargs.attention_layer = list(synth.COMBINE1_LAYERS)
sd = synth.synthetic_state_dict(json.load(open(os.path.join(ROOT, "tests", "golden", "depth_tail_keys.json"))), seed=11)
dec = D.Decoder(num_classes=128)
dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
dec = dec.to(dev).eval()
for m in (dec.cross_atten1, dec.cross_atten2, dec.cross_atten3):
    m.to(torch.bfloat16)
head = D.DepthHead(n_bins=256, min_val=1e-3, max_val=10.0)
head.load_state_dict({k: v for k, v in sd.items() if k.startswith(("depth_head.", "conv_out."))}, strict=True)
head = head.to(dev).eval()
enc = cfpnet_b200.HistogramEncoder()
enc.load_state_dict({k[len("hist_encoder."):]: v for k, v in sd.items() if k.startswith("hist_encoder.")}, strict=True)
enc = enc.to(dev).eval()
enc.out_dtype = torch.bfloat16
inp = synth.make_inputs("G416", B, seed=200, levels=())
feats = [t.to(dev) for t in synth.encoder_features("G416", B, seed=200)]
hist_in, mask = inp["hist_data"].to(dev), inp["mask"].to(dev)
with torch.no_grad():
    for i in range(steps):
        n0 = _lib.launch_count()
        shard.seed_posenc(i)
        hist = enc(hist_in.unsqueeze(-1))
        unet, H, W = dec.forward_nhwc(feats, hist, rect_data=None, mask=mask, patch_info=inp["patch_info"], rgb=None)
        head.forward_nhwc(unet, H, W)
        torch.cuda.synchronize()
        print(f"step {i}: {_lib.launch_count() - n0} libcfp launches")
