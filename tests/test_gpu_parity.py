"""Parity of the CUDA path (through the drop-in modules -> C ABI -> sm_100a kernels) against
the CPU oracle and the committed reference outputs.  Tolerances:
masks / indices bit-exact; fp32 features rel-L2 <= 1e-3; bf16 rel-L2 <= 2e-2 (vs fp64 truth)."""
import ctypes
import os

import numpy as np
import pytest
import torch

import cfpnet_b200
from cfpnet_b200 import _lib, geometry, synth
from cfpnet_b200.config import args
from helpers import FUSION_CASES as _BASE_CASES, FUSION_CASES_Z6, FusionCase, GOLDEN, ref_keys, rel_l2

# 6x6 zones (the reference's training layout).  They failed on their first GPU run because cfp_twins_fwd /
# cfp_lkpm_fwd refused the workspace (their layout assumes 64 zones, the module had sized it for 36): reproduced and
# fixed on the CPU (cfp_workspace_bytes now covers both layouts, tests/test_host.py::
# test_workspace_covers_every_entry_point) and confirmed on a B200 with the round's last GPU seconds
# (tools/z6_quick.py -> profiles/r1t_z6_gpu_check.log: fp32 rel-L2 1.2e-6 / 8.3e-7, bf16 8.8e-3 / 8.6e-3).
FUSION_CASES = _BASE_CASES + FUSION_CASES_Z6
from oracle import cfp_oracle as O

pytestmark = pytest.mark.gpu
FP32_TOL, BF16_TOL = 1e-3, 2e-2
DEV = "cuda:0"


def build_fusion(case_or_level, layers=synth.COMBINE1_LAYERS, change_embedding=True, no_skip_inside=False,
                 dtype=torch.float32):
    if isinstance(case_or_level, FusionCase):
        c = case_or_level
        level, layers, change_embedding, no_skip_inside = c.level, c.layers, c.change_embedding, c.no_skip_inside
    else:
        level = case_or_level
    C, _, max_res, lk = synth.LEVELS[level]
    args.attention_layer = list(layers)
    args.change_embedding, args.no_skip_inside = change_embedding, no_skip_inside
    m = cfpnet_b200.TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1])
    kind = "baseline" if tuple(layers) == synth.BASELINE_LAYERS else "combine1"
    sd = synth.synthetic_state_dict(ref_keys()[f"fusion_{kind}_L{level}"], seed=level)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).to(dtype).eval(), sd


def build_hist(dtype=torch.float32):
    enc = cfpnet_b200.HistogramEncoder()
    sd = synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).eval()
    enc.out_dtype = dtype
    return enc, sd


@pytest.fixture(autouse=True)
def _restore_flags():
    saved = (list(args.attention_layer), args.change_embedding, args.no_skip_inside)
    yield
    args.attention_layer, args.change_embedding, args.no_skip_inside = saved


# ------------------------------------------------------------------ a1
# bf16: the tensor-core kernel keeps bf16 activations between the nine stages (as the reference cast to bf16
# does); measured 3e-3..6e-3, stated tolerance 1e-2 (the fused features downstream are held to 2e-2)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_hist_encoder(dtype, tol):
    enc, sd = build_hist(dtype)
    inp = synth.make_inputs("G416", 2, seed=1, levels=())
    outs = enc(inp["hist_data"].to(DEV).unsqueeze(-1))
    z = np.load(os.path.join(GOLDEN, "hist_encoder_B2.npz"))
    truth = O.hist_encoder(sd, inp["hist_data"].double())
    for c, o, t in zip((32, 64, 128), outs, truth):
        assert o.shape == (2, 64, 16, c) and o.dtype == dtype
        assert rel_l2(o, torch.from_numpy(z[f"out{c}"])) <= tol      # the reference's own output
        assert rel_l2(o, t) <= tol                                     # fp64 oracle


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_hist_encoder_ragged_rows(dtype, tol):
    """Row counts that are not a multiple of the 64 / 128-row tiles, a single zone, and more tiles than CTAs."""
    enc, sd = build_hist(dtype)
    for B, Z in ((1, 1), (3, 5), (1, 64), (41, 64)):
        h = torch.rand(B, Z, 16, generator=torch.Generator().manual_seed(B * 7 + Z)) * 4
        outs = enc(h.to(DEV).unsqueeze(-1))
        for o, t in zip(outs, O.hist_encoder(sd, h.double())):
            assert rel_l2(o, t) <= tol


# ------------------------------------------------------------------ a2/a3
@pytest.mark.parametrize("tag", ["G416_L3_B2", "G480_L3_B1", "G480pad_L3_B1", "G480pad_L2_B1", "G416_L1_B1", "G480_L2_B1", "G480_L1_B1",
                                 "G480pad_L1_B1", "G416_L3_B16_baseline"] + FUSION_CASES_Z6)
def test_masks_bit_exact(tag):
    case = FusionCase(tag)
    inp = case.inputs()
    H, W = synth.level_hw(case.geometry, case.level)
    g = geometry.zone_geometry(inp["patch_info"], case.max_res[1], H, W)
    cg = _lib.CfpGeom.from_geometry(g)
    B, P = case.batch, g.p1 * g.p2
    mask = inp["mask"].to(DEV, torch.uint8).contiguous()
    zm = torch.empty(B, H * W, dtype=torch.uint8, device=DEV)
    hm = torch.empty(B * g.zone_num ** 2, P, dtype=torch.uint8, device=DEV)
    pm = torch.empty(B, g.tzh, g.tzw, dtype=torch.uint8, device=DEV)
    _lib.call("cfp_zone_masks", mask.data_ptr(), zm.data_ptr(), hm.data_ptr(), pm.data_ptr(), B, H, W,
              ctypes.byref(cg), _lib.stream_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(zm.cpu().numpy().astype(bool), case.mask_bits("zone_mask"))
    assert np.array_equal(hm.cpu().numpy().astype(bool), case.mask_bits("hist_mask"))
    assert np.array_equal(pm.cpu().numpy().astype(bool), case.mask_bits("pad_mask"))
    # and against the oracle's restatement, with the channel repeat
    ozm, ohm, opm = O.zone_masks(g.asdict(), inp["mask"], B, H, W, case.C)
    assert torch.equal(zm.cpu().bool(), ozm[:, :, 0]) and torch.equal(hm.cpu().bool(), ohm[:, :, 0])
    assert torch.equal(pm.cpu().bool().reshape(-1), opm.view(-1, case.C)[:, 0])


# ------------------------------------------------------------------ a4-a9 through TransformerFusion
def run_case(case, dtype):
    m, sd = build_fusion(case, dtype=dtype)
    enc, hsd = build_hist(dtype)
    inp = case.inputs()
    feats = enc(inp["hist_data"].to(DEV).unsqueeze(-1))
    feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[case.C]
    torch.manual_seed(2)                       # same positional-encoding crop draws as the reference run
    out = m(inp[f"x{case.level}"].to(DEV, dtype), feat1, rect_data=inp["rect_data"], mask=inp["mask"].to(DEV),
            patch_info=inp["patch_info"], rgb=None)
    torch.cuda.synchronize()
    return out, inp, sd, hsd


@pytest.mark.parametrize("tag", FUSION_CASES)
def test_fusion_fp32_vs_reference(tag):
    case = FusionCase(tag)
    out, *_ = run_case(case, torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous()
    case.check_output(out, FP32_TOL, f"cuda fp32 {tag}")


@pytest.mark.parametrize("tag", FUSION_CASES)
def test_fusion_bf16_vs_reference(tag):
    case = FusionCase(tag)
    out, *_ = run_case(case, torch.bfloat16)
    assert out.dtype == torch.bfloat16
    case.check_output(out, BF16_TOL, f"cuda bf16 {tag}")


# the 6x6 training layout at every level against the oracle (L1 = 256 cells per zone)
@pytest.mark.parametrize("geom", ["G416", "G416z6"])
@pytest.mark.parametrize("level", [3, 2, 1])
def test_fusion_vs_oracle_batch3(level, geom):
    """Batch > 1 with per-frame masks, against the fp64 oracle on the same seeded inputs."""
    m, sd = build_fusion(level)
    enc, hsd = build_hist()
    C, _, max_res, _ = synth.LEVELS[level]
    inp = synth.make_inputs(geom, 3, seed=5, levels=(level,))
    feats = enc(inp["hist_data"].to(DEV).unsqueeze(-1))
    feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[C]
    torch.manual_seed(11)
    out = m(inp[f"x{level}"].to(DEV), feat1, rect_data=inp["rect_data"], mask=inp["mask"].to(DEV),
            patch_info=inp["patch_info"], rgb=None)
    torch.manual_seed(11)
    of = O.hist_encoder(hsd, inp["hist_data"].double())
    truth = O.transformer_fusion(sd, synth.COMBINE1_LAYERS, max_res, inp[f"x{level}"].double(),
                                 {32: of[0], 64: of[1], 128: of[2]}[C], inp["mask"], inp["patch_info"])
    assert rel_l2(out, truth) <= FP32_TOL


def test_all_zones_invalid_leaves_zone_cells_untouched_by_hist2image():
    """mask all False: hist2image contributes nothing (fusion.py:144), the rest still runs."""
    level = 3
    m, sd = build_fusion(level)
    enc, hsd = build_hist()
    inp = synth.make_inputs("G416", 1, seed=3, levels=(level,))
    mask = torch.zeros_like(inp["mask"])
    feats = enc(inp["hist_data"].to(DEV).unsqueeze(-1))
    torch.manual_seed(4)
    out = m(inp["x3"].to(DEV), feats[2], rect_data=inp["rect_data"], mask=mask.to(DEV),
            patch_info=inp["patch_info"], rgb=None)
    torch.manual_seed(4)
    of = O.hist_encoder(hsd, inp["hist_data"].double())
    truth = O.transformer_fusion(sd, synth.COMBINE1_LAYERS, (30, 40), inp["x3"].double(), of[2], mask,
                                 inp["patch_info"])
    assert rel_l2(out, truth) <= FP32_TOL


def test_same_seed_same_output_and_rng_consumption():
    """The drop-in draws from the global CPU generator exactly like the reference:
    two draws at 416x544, none at 480x640 (fusion.py:88-91)."""
    m, _ = build_fusion(3)
    enc, _ = build_hist()
    inp = synth.make_inputs("G416", 1, levels=(3,))
    feats = enc(inp["hist_data"].to(DEV).unsqueeze(-1))
    kw = dict(rect_data=inp["rect_data"], mask=inp["mask"].to(DEV), patch_info=inp["patch_info"], rgb=None)
    torch.manual_seed(2)
    a = m(inp["x3"].to(DEV), feats[2], **kw)
    after = torch.randint(0, 1000, [1])
    torch.manual_seed(2)
    b = m(inp["x3"].to(DEV), feats[2], **kw)
    assert rel_l2(a, b) <= 1e-5       # fp32 atomics in the attention state: last bits may differ
    torch.manual_seed(2)
    torch.randint(0, 5, [1]); torch.randint(0, 7, [1])
    assert torch.equal(after, torch.randint(0, 1000, [1]))
    inp = synth.make_inputs("G480", 1, levels=(3,))
    state = torch.get_rng_state()
    m(inp["x3"].to(DEV), feats[2], rect_data=inp["rect_data"], mask=inp["mask"].to(DEV),
      patch_info=inp["patch_info"], rgb=None)
    assert torch.equal(state, torch.get_rng_state())


def test_full_batch_properties():
    """BASELINE.json config 3 size (B=64, bf16, G416): size-independent properties instead of a
    64-frame oracle run — (1) frames are independent: frame i of the batch equals the same frame run
    alone (identical kernels, identical per-frame arithmetic -> bit-exact), (2) outputs finite."""
    dtype = torch.bfloat16
    enc, _ = build_hist(dtype)
    B = 64
    inp = synth.make_inputs("G416", B, seed=9)
    feats = enc(inp["hist_data"].to(DEV).unsqueeze(-1))
    for level in (3, 2, 1):
        m, _ = build_fusion(level, dtype=dtype)
        C = synth.LEVELS[level][0]
        feat1 = {32: feats[0], 64: feats[1], 128: feats[2]}[C]
        x = inp[f"x{level}"].to(DEV, dtype)
        kw = dict(rect_data=inp["rect_data"], rgb=None)
        torch.manual_seed(2)
        full = m(x, feat1, mask=inp["mask"].to(DEV), patch_info=inp["patch_info"], **kw)
        assert torch.isfinite(full).all()
        for i in (0, 37, 63):
            pi = {k: ({kk: vv[i:i + 1] for kk, vv in v.items()} if isinstance(v, dict) else v[i:i + 1])
                  for k, v in inp["patch_info"].items()}
            torch.manual_seed(2)
            one = m(x[i:i + 1], feat1[i:i + 1], mask=inp["mask"][i:i + 1].to(DEV), patch_info=pi, **kw)
            # fp32 atomics in the attention state make the last bits order-dependent
            assert rel_l2(one, full[i:i + 1]) <= 1e-2


@pytest.mark.parametrize("graph,B", [(False, 2), (True, 2), (True, 16)])
def test_stream_host_equals_direct_forward(graph, B):
    """The pipelined host-buffer API (separate H2D / compute / D2H streams, per-input "landed" and per-level "done" events;
    with ``graph`` every slot replays a captured forward in which those events are external wait / record nodes) returns,
    for every batch, what a plain forward on device-resident copies of the same batch returns."""
    from cfpnet_b200 import FusionPath
    path = FusionPath(synth.COMBINE1_LAYERS)
    path.hist_encoder.load_state_dict(synth.synthetic_state_dict(ref_keys()["hist_encoder"], seed=0))
    for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
        getattr(path, name).load_state_dict(synth.synthetic_state_dict(ref_keys()[f"fusion_combine1_L{lv}"], seed=lv))
    path = path.to(DEV).eval().set_dtype(torch.bfloat16)
    batches, pi = [], None
    for s in range(5):
        inp = synth.make_inputs("G416", B, seed=20 + s)
        pi = inp["patch_info"]
        batches.append({"x3": inp["x3"].bfloat16().pin_memory(), "x2": inp["x2"].bfloat16().pin_memory(),
                        "x1": inp["x1"].bfloat16().pin_memory(), "hist_data": inp["hist_data"].pin_memory(),
                        "mask": inp["mask"].pin_memory()})
    got = {}
    for idx, outs in path.stream_host(batches, pi, DEV, depth=2, seeds=range(100, 105), graph=graph):
        got[idx] = [o.clone() for o in outs]            # buffers are recycled after `depth` more batches
    assert sorted(got) == list(range(5))
    for i, hb in enumerate(batches):
        torch.manual_seed(100 + i)
        want = path(hb["x3"].to(DEV), hb["x2"].to(DEV), hb["x1"].to(DEV), hb["hist_data"].to(DEV), hb["mask"].to(DEV), pi)
        torch.cuda.synchronize()
        for g_, w_ in zip(got[i], want):
            assert rel_l2(g_, w_) <= 1e-2


def test_graph_replay_equals_eager_forward():
    """FusionPath.make_graphed at the 480x640 geometry (no positional-encoding crop is drawn there): a replay on new
    inputs equals the eager forward on them; at 416x544 the crop offsets are drawn per replay and read from device memory."""
    path = cfpnet_b200.FusionPath(synth.COMBINE1_LAYERS)
    path.hist_encoder.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in path.hist_encoder.state_dict().items()}, 0))
    for lv, name in ((3, "cross_atten3"), (2, "cross_atten2"), (1, "cross_atten1")):
        m = getattr(path, name)
        m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, lv))
    path = path.to(DEV).eval().set_dtype(torch.bfloat16)

    def dev_inputs(geom, seed):
        inp = synth.make_inputs(geom, 1, seed=seed)
        return [inp[k].to(torch.bfloat16).to(DEV) for k in ("x3", "x2", "x1")] + [inp["hist_data"].to(DEV), inp["mask"].to(DEV)], inp["patch_info"]

    a, pi = dev_inputs("G480", 3)
    b, _ = dev_inputs("G480", 4)
    with torch.no_grad():
        run = path.make_graphed(*a, pi)
        got = [o.clone() for o in run(*b)]
        torch.cuda.synchronize()
        want = path(*b, pi)
    for g_, w_ in zip(got, want):      # not bit-equal: the split-K / straddling-group fp32 atomics are order-dependent
        assert rel_l2(g_, w_) <= 5e-3
    # 416x544: the maps are smaller than the positional-encoding tables, so every forward draws crop offsets; a replay
    # takes them from device memory (same generator, same order) - nothing random is frozen into the graph
    c, pi416 = dev_inputs("G416", 3)
    d, _ = dev_inputs("G416", 4)
    with torch.no_grad():
        run416 = path.make_graphed(*c, pi416)
        for seed, inputs in ((11, d), (12, c), (13, d)):
            torch.manual_seed(seed)
            got = [o.clone() for o in run416(*inputs)]
            torch.cuda.synchronize()
            torch.manual_seed(seed)
            want = path(*inputs, pi416)
            for g_, w_ in zip(got, want):
                assert rel_l2(g_, w_) <= 5e-3, seed
        torch.manual_seed(11)
        first = [o.clone() for o in run416(*d)]
        torch.manual_seed(12)
        other = [o.clone() for o in run416(*d)]
        torch.cuda.synchronize()
        assert rel_l2(first[0], other[0]) > 1e-3, "different seeds must give different crops (hence outputs) on replay"
