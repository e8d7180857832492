"""Generate ``tests/golden/depth_G416_B1.npz`` from the REFERENCE decoder + depth head (container only).

    python tools/make_golden_depth.py

End-to-end pin for the "final depth within 0.5 % abs-rel" clause: the reference's own ``Decoder`` (with its three
``TransformerFusion`` modules), ``HistogramEncoder``, ``DepthRegression`` and ``conv_out`` softmax, wired exactly as
``Deltar.forward`` wires them (``src/models/deltar.py:39-61``), run in float64 on synthetic decoder-level inputs:
the five image-encoder feature maps are drawn from a seed (the timm EfficientNetV2 backbone is third-party, absent
and out of scope - SURVEY.md §8c - and both sides of the comparison get the same maps), weights are the deterministic
``cfpnet_b200.synth`` ones.  Stored: the predicted depth map, the bin edges, and the key -> shape list of the
decoder-shell / head parameters so that the tests can regenerate the identical weights from the seed.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ref_import import import_reference  # noqa: E402
from cfpnet_b200 import synth  # noqa: E402

ref = import_reference()
args = ref["args"]
from src.models import decoder as ref_decoder  # noqa: E402  (needs the shims installed by import_reference)

OUT = os.path.join(ROOT, "tests", "golden")
N_BINS, MIN_VAL, MAX_VAL = args.n_bins, args.min_depth, args.max_depth     # 256, 1e-3, 10 (combine1 config)


class Tail(nn.Module):
    """The part of Deltar below the image encoder (deltar.py:15-19)."""

    def __init__(self):
        super().__init__()
        self.hist_encoder = ref["encoder"].HistogramEncoder()
        self.depth_head = ref_decoder.DepthRegression(128, dim_out=N_BINS, norm=args.norm)
        self.decoder = ref_decoder.Decoder(num_classes=128)
        self.conv_out = nn.Sequential(nn.Conv2d(128, N_BINS, kernel_size=1, stride=1, padding=0), nn.Softmax(dim=1))

    def forward(self, img_features, add):                    # deltar.py:40-61
        hist_features = self.hist_encoder(add["hist_data"].unsqueeze(-1))
        unet_out = self.decoder(img_features, hist_features, rect_data=add["rect_data"], mask=add["mask"],
                                patch_info=add["patch_info"], rgb=None)
        bin_widths_normed, range_attention_maps = self.depth_head(unet_out)
        out = self.conv_out(range_attention_maps)
        bin_widths = (MAX_VAL - MIN_VAL) * bin_widths_normed
        bin_widths = nn.functional.pad(bin_widths, (1, 0), mode="constant", value=MIN_VAL)
        bin_edges = torch.cumsum(bin_widths, dim=1)
        centers = 0.5 * (bin_edges[:, :-1] + bin_edges[:, 1:])
        pred = torch.sum(out * centers.view(*centers.shape, 1, 1), dim=1, keepdim=True)
        return bin_edges, pred


def main():
    B = 1
    torch.manual_seed(0)
    m = Tail().eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(synth.synthetic_state_dict(shapes, seed=11), strict=True)
    m = m.double()
    inp = synth.make_inputs("G416", B, seed=5, levels=())
    feats = [t.double() for t in synth.encoder_features("G416", B, seed=5)]
    add = dict(hist_data=inp["hist_data"].double(), rect_data=inp["rect_data"].double(), mask=inp["mask"],
               patch_info=inp["patch_info"])
    with torch.no_grad():
        torch.manual_seed(2)                                  # positional-encoding crops (fusion.py:88-91)
        edges, pred = m(feats, add)
        m32 = m.float()
        torch.manual_seed(2)
        _, pred32 = m32([f.float() for f in feats], dict(add, hist_data=inp["hist_data"], rect_data=inp["rect_data"]))
    ref_drift = float(((pred32.double() - pred).abs() / pred).mean())

    # the reference itself with its three TransformerFusion modules cast to bf16 (shell and head stay fp32): the yardstick
    # for the bf16 path - at random init the 256-bin softmax turns ~1e-2 feature error into several percent of depth
    class Bf16(nn.Module):
        def __init__(self, mod):
            super().__init__()
            self.mod = mod.bfloat16()

        def forward(self, x, f, **kw):
            return self.mod(x.bfloat16(), f.bfloat16(), **kw).float()

    for n in ("cross_atten1", "cross_atten2", "cross_atten3"):
        setattr(m32.decoder, n, Bf16(getattr(m32.decoder, n)))
    with torch.no_grad():
        torch.manual_seed(2)
        _, pred16 = m32([f.float() for f in feats], dict(add, hist_data=inp["hist_data"], rect_data=inp["rect_data"]))
    ref_bf16 = float(((pred16.double() - pred).abs() / pred).mean())
    print("pred", tuple(pred.shape), "range", float(pred.min()), float(pred.max()),
          "reference fp32-vs-fp64 abs-rel", ref_drift, "reference bf16-fusion abs-rel", ref_bf16)
    np.savez_compressed(os.path.join(OUT, "depth_G416_B1.npz"), pred=pred.float().numpy(), bin_edges=edges.float().numpy(),
                        ref_fp32_abs_rel=np.float64(ref_drift), ref_bf16_abs_rel=np.float64(ref_bf16))
    with open(os.path.join(OUT, "depth_tail_keys.json"), "w") as fh:
        json.dump(shapes, fh)


if __name__ == "__main__":
    main()
