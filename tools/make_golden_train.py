"""Generate ``tests/golden/train_*.npz`` from the REFERENCE modules in TRAIN mode (container only).

    python tools/make_golden_train.py

First step of the training row of the scope table: before any backward kernel
exists, pin what one training step of the path computes.  The reference's ``HistogramEncoder`` and
``TransformerFusion`` are imported from /root/reference (tools/ref_import.py), put in ``.train()`` mode
(BatchNorm batch statistics, train.py:75) in float64, and run forward + backward on the synthetic inputs with the
scalar ``L = sum(out * cotangent)`` (cotangent = seeded N(0,1), so every output element gets its own weight).  Stored:

  * the forward output (float32 copy of the float64 run; maps above 100k elements as a seeded sample + channel sums);
  * d L / d x and d L / d hist_data - the gradients that leave the path on the input side;
  * for every parameter: whether it received a gradient at all (the reference registers parameters it never uses;
    SURVEY.md section 8e: they must stay out of the gradient allreduce), its gradient's L2 norm, sum, and the
    gradient values at 48 seeded positions;
  * every BatchNorm buffer after the step (running_mean / running_var / num_batches_tracked).

tests/test_oracle_train_golden.py pins ``oracle/cfp_oracle.py`` (train-mode BN + autograd over the restatement) on
these files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ref_import import import_reference  # noqa: E402
from cfpnet_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_PROBE = 48
SAMPLE_N = 32768
FULL_LIMIT = 100_000        # elements; larger maps are stored sampled

ref = import_reference()
args = ref["args"]
TransformerFusion = ref["fusion"].TransformerFusion
HistogramEncoder = ref["encoder"].HistogramEncoder


def load_double(mod, seed):
    shapes = {k: v.shape for k, v in mod.state_dict().items()}
    sd = synth.synthetic_state_dict(shapes, seed=seed)
    mod.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}, strict=True)
    return mod


def probe_index(name, numel):
    g = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    return torch.randint(0, numel, (N_PROBE,), generator=g)


def hash_name(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000000007
    return h


def train_case(tag, geometry, level, batch, layers):
    C, _, max_res, lk = synth.LEVELS[level]
    args.attention_layer = list(layers)
    args.change_embedding, args.no_skip_inside = True, False
    mod = load_double(TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1]).double(), level)
    henc = load_double(HistogramEncoder().double(), 0)
    mod.train()
    henc.train()

    inp = synth.make_inputs(geometry, batch, seed=1, levels=(level,))
    pi = synth.collate_patch_info([ref["dataloader"].patch_info_from_rect_data(r) for r in inp["rect_data"]])
    x = inp[f"x{level}"].double().requires_grad_(True)
    hist = inp["hist_data"].double().requires_grad_(True)
    feats = henc(hist.unsqueeze(-1))
    feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[C]
    torch.manual_seed(2)
    out = mod(x, feat1, rect_data=inp["rect_data"], mask=inp["mask"], patch_info=pi, rgb=None)
    ct = torch.randn(out.shape, generator=torch.Generator().manual_seed(77), dtype=torch.float64)
    (out * ct).sum().backward()

    rec = {"grad_hist": hist.grad.to(torch.float32).numpy(),
           "meta": np.array([geometry, str(level), str(batch), ",".join(layers)])}
    for key, t in (("out", out.detach()), ("grad_x", x.grad)):
        rec[key + "_shape"] = np.array(t.shape)
        if t.numel() <= FULL_LIMIT:
            rec[key] = t.to(torch.float32).numpy()
        else:                   # large maps: a seeded sample + the per-(frame, channel) sums
            idx = torch.randperm(t.numel(), generator=torch.Generator().manual_seed(1234))[:SAMPLE_N]
            rec[key + "_idx"] = idx.numpy().astype(np.int64)
            rec[key + "_sample"] = t.reshape(-1)[idx].to(torch.float32).numpy()
            rec[key + "_perchan"] = t.sum(dim=(2, 3)).to(torch.float32).numpy()
    names, has_grad, norms, sums, probes = [], [], [], [], []
    for prefix, m in (("fusion.", mod), ("hist.", henc)):
        for n, p in m.named_parameters():
            names.append(prefix + n)
            has_grad.append(p.grad is not None)
            gr = p.grad if p.grad is not None else torch.zeros_like(p)
            norms.append(float(gr.norm()))
            sums.append(float(gr.sum()))
            probes.append(gr.reshape(-1)[probe_index(prefix + n, p.numel())].numpy())
    rec.update(param_names=np.array(names), param_has_grad=np.array(has_grad), param_grad_norm=np.array(norms),
               param_grad_sum=np.array(sums), param_grad_probe=np.stack(probes))
    bnames, bvals = [], {}
    for prefix, m in (("fusion.", mod), ("hist.", henc)):
        for n, b in m.named_buffers():
            bnames.append(prefix + n)
            bvals["buf:" + prefix + n] = b.detach().numpy()
    rec["buffer_names"] = np.array(bnames)
    rec.update(bvals)
    np.savez_compressed(os.path.join(OUT, f"train_{tag}.npz"), **rec)
    used = sum(has_grad)
    print(f"{tag:24s} out {tuple(out.shape)}  params {len(names)} ({used} with grad, {len(names) - used} never used)  "
          f"|grad_x| {float(x.grad.norm()):.4g}  |grad_hist| {float(hist.grad.norm()):.4g}")


if __name__ == "__main__":
    C1 = synth.COMBINE1_LAYERS
    train_case("G416z6_L3_B2", "G416z6", 3, 2, C1)      # the reference's training layout: 6x6 zones of 64 px
    train_case("G416_L2_B2", "G416", 2, 2, C1)
    train_case("G416z6_L1_B1", "G416z6", 1, 1, C1)
    train_case("G480_L3_B1", "G480", 3, 1, C1)           # bilinear-resize branch of hist2image (480x640, 8x8 zones of 56 px)      # the heavy level: 31x31 depthwise, 12x12 windows
