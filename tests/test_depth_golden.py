"""End-to-end pin of the oracle against the reference's decoder + depth head (tests/golden/depth_G416_B1.npz, made
by tools/make_golden_depth.py from the reference's own Decoder / DepthRegression / HistogramEncoder classes)."""
import json
import os

import numpy as np
import torch

from cfpnet_b200 import synth
from oracle import cfp_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MIN_VAL, MAX_VAL = 1e-3, 10.0          # --min_depth / --max_depth of the reference's combine1 config


def tail_state():
    shapes = json.load(open(os.path.join(GOLDEN, "depth_tail_keys.json")))
    return synth.synthetic_state_dict(shapes, seed=11)


def run_tail(sd, fuse, dtype, device="cpu"):
    inp = synth.make_inputs("G416", 1, seed=5, levels=())
    feats = [t.to(device, dtype) for t in synth.encoder_features("G416", 1, seed=5)]
    sd = {k: (v.to(device) if v.dtype == torch.long else v.to(device, dtype)) for k, v in sd.items()}
    hist = O.hist_encoder(O.sub(sd, "hist_encoder."), inp["hist_data"].to(device, dtype))
    torch.manual_seed(2)                 # positional-encoding crops, drawn in call order L3, L2, L1
    unet = O.decoder_shell(O.sub(sd, "decoder."), feats, hist, lambda name, x, f: fuse(sd, name, x, f, inp))
    return O.depth_tail(sd, unet, MIN_VAL, MAX_VAL)


def oracle_fuse(sd, name, x, feat1, inp):
    level = {n: (mr, ) for n, _, mr, _ in O.LEVELS}[name][0]
    return O.transformer_fusion(O.sub(sd, f"decoder.{name}."), synth.COMBINE1_LAYERS, level, x, feat1, inp["mask"], inp["patch_info"])


def test_oracle_decoder_and_depth_head_match_the_reference():
    z = np.load(os.path.join(GOLDEN, "depth_G416_B1.npz"))
    with torch.no_grad():
        edges, pred = run_tail(tail_state(), oracle_fuse, torch.float64)
    gt = torch.from_numpy(z["pred"])
    assert pred.shape == gt.shape == (1, 1, 208, 272)
    assert float((edges.float() - torch.from_numpy(z["bin_edges"])).abs().max()) <= 1e-5
    assert O.abs_rel(pred, gt) <= 1e-6           # fp64 restatement vs fp64 reference (stored as fp32)
