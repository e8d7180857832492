"""Host-side weight packing checked on the CPU against plain PyTorch ops: the packed buffers are decoded with the
index formulas the kernels use (csrc/umma.cuh, k_dwconv_tc.cu, k_chain_tc.cu) and the kernels' arithmetic is emulated
in float64, so a layout or fold mistake in ``cfpnet_b200/layers.py`` / ``encoder.py`` fails here without a GPU."""
import pytest
import torch
import torch.nn.functional as F

import cfpnet_b200
from cfpnet_b200 import _lib, synth
from cfpnet_b200.config import args
from cfpnet_b200.packing import Packer, fold_bn, umma_block


def from_umma(block, n, k):
    """Inverse of umma_block: [K/8][N][8] -> [N, K]."""
    return block.view(k // 8, n, 8).permute(1, 0, 2).reshape(n, k)


def test_umma_block_layout_formula():
    w = torch.arange(48 * 32, dtype=torch.float32).view(48, 32)            # [N=48, K=32]
    blk = umma_block(w, torch.float32).reshape(-1)
    for n_, k_ in ((0, 0), (5, 7), (47, 31), (13, 8), (30, 17)):
        assert blk[((k_ // 8) * 48 + n_) * 8 + k_ % 8] == w[n_, k_]       # element (n,k) at ((k//8)*N + n)*8 + k%8
    assert torch.equal(from_umma(blk, 48, 32), w)


def make_block(C, k):
    blk = cfpnet_b200.layers.Block14(C, large_kernel=k)
    sd = synth.synthetic_state_dict({n: v.shape for n, v in blk.state_dict().items()}, seed=3)
    blk.load_state_dict(sd)
    return blk.eval()


@pytest.mark.parametrize("C,k,H,W,B", [(32, 31, 21, 40, 1), (64, 15, 13, 20, 3)])
def test_toeplitz_pack_reproduces_the_depthwise_conv(C, k, H, W, B):
    """Emulates dwconv_tc: vertical taps dy = 4a + b stacked along N, out[r] = sum_b E_b[r + b], on zero-padded planes."""
    blk = make_block(C, k)
    keep = Packer()
    w = blk.pack(keep)
    toep = next(t for t in keep if t.dtype == torch.bfloat16 and t.dim() == 7)      # [C][NA][KS][2][NB][32][8]
    nb, pad = 4, (k - 1) // 2
    na, ks = (k + nb - 1) // nb, (32 + k - 1 + 15) // 16
    assert tuple(toep.shape) == (C, na, ks, 2, nb, 32, 8) and w.ksize == k
    T = toep.double().permute(0, 1, 4, 5, 2, 3, 6).reshape(C, na, nb, 32, 16 * ks)   # [c][a][b][n][kk]
    scale, shift = fold_bn(blk.bn1)
    taps = (blk.dwconv2.weight.detach()[:, 0] * scale[:, None, None]).to(torch.bfloat16).double()
    x = torch.randn(B, C, H, W, dtype=torch.float64)
    ref = F.conv2d(x, taps[:, None], padding=pad, groups=C)
    nx = (W + 31) // 32
    planes = torch.zeros(B, C, H + 2 * pad + nb * na, 32 * nx + 16 * ks, dtype=torch.float64)
    planes[:, :, pad:pad + H, pad:pad + W] = x
    out = torch.zeros_like(ref)
    for xt in range(nx):
        E = torch.zeros(B, C, H + nb, nb, 32, dtype=torch.float64)
        for a in range(na):
            A = planes[:, :, nb * a:nb * a + H + nb, 32 * xt:32 * xt + 16 * ks]     # rows m + NB a
            E += torch.einsum("bcmk,cjnk->bcmjn", A, T[:, a])
        acc = sum(E[:, :, b:b + H, b] for b in range(nb))                             # out[r] = sum_b E_b[r + b]
        wv = min(32, W - 32 * xt)
        out[..., 32 * xt:32 * xt + wv] = acc[..., :wv]
    assert float((out - ref).abs().max()) <= 1e-9 * float(ref.abs().max() + 1)


@pytest.mark.parametrize("C", [32, 64, 128])
def test_mlp_fold_reproduces_layernorm_mlp(C):
    """W1' = W1 diag(g), b1' = b1 + W1 b ride in [LNhat(y) | 1]; b2 rides in [GELU(h) | 1] of slice 0 (MlpTC)."""
    blk = make_block(C, 7)
    keep = Packer()
    blk.pack(keep)
    tc = next(t for t in keep if t.dtype == torch.uint8)
    bsz = max(128 * (C + 16) * 2, C * 144 * 2)
    nsl = 4 * C // 128
    assert tc.numel() == 2 * nsl * bsz
    y = torch.randn(50, C, dtype=torch.float64)
    yhat = (y - y.mean(1, keepdim=True)) / torch.sqrt(y.var(1, unbiased=False, keepdim=True) + 1e-6)
    a0 = torch.cat([yhat, torch.ones(50, 1, dtype=torch.float64), torch.zeros(50, 15, dtype=torch.float64)], 1)
    out = torch.zeros(50, C, dtype=torch.float64)
    for j in range(nsl):
        w1 = from_umma(tc[(2 * j) * bsz:(2 * j) * bsz + 128 * (C + 16) * 2].view(torch.bfloat16), 128, C + 16).double()
        w2 = from_umma(tc[(2 * j + 1) * bsz:(2 * j + 1) * bsz + C * 144 * 2].view(torch.float16), C, 144).double()
        h = F.gelu(a0 @ w1.t())
        a1 = torch.cat([h, torch.ones(50, 1, dtype=torch.float64), torch.zeros(50, 15, dtype=torch.float64)], 1)
        out += a1 @ w2.t()
    with torch.no_grad():
        ln = F.layer_norm(y, (C,), blk.norm.weight.double(), blk.norm.bias.double(), eps=1e-6)
        ref = F.linear(F.gelu(F.linear(ln, blk.pwconv1.weight.double(), blk.pwconv1.bias.double())),
                       blk.pwconv2.weight.double(), blk.pwconv2.bias.double())
    assert float((out - ref).norm() / ref.norm()) <= 1e-2          # bf16 / fp16 weight rounding only


def test_hist_encoder_tc_blocks_decode_to_the_folded_stages():
    enc = cfpnet_b200.HistogramEncoder()
    enc.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in enc.state_dict().items()}, 0))
    w, (stages, tc, _buf) = enc.eval()._pack()
    off = 0
    for wt, _b in stages[1:]:                                   # stage i: wt [Cin, Cout] -> block [Cin/8][Cout][8]
        cin, cout = wt.shape
        blk = tc[off:off + cin * cout]
        assert torch.equal(from_umma(blk, cout, cin), wt.t().to(torch.bfloat16))
        off += cin * cout
    assert off == tc.numel() == 106496 // 2 and ctypes_ptr_ok(w)


def ctypes_ptr_ok(w):
    return bool(w.tc) and all(bool(w.w_t[i]) and bool(w.b[i]) for i in range(9))


# ------------------------------------------------------------------ DAPM 3x3 convs, LoFTR chain, GSA sub-sampling conv
def _loaded(mod, seed):
    mod.load_state_dict(synth.synthetic_state_dict({n: v.shape for n, v in mod.state_dict().items()}, seed=seed))
    return mod.eval()


@pytest.mark.parametrize("C", [32, 64])
def test_dapm_conv_blocks_reproduce_conv_bn(C):
    """conv{1,2}_pk: one [Cout x Cout] block per (source, tap), eval-mode BN scale folded into the weights, shift
    added by the epilogue (k_conv_tc.cu walks source-major, tap = ky*3+kx, over the zero-padded raster) - and the
    fp32 engine's [(tap, cin)][Cout] layout - both against conv2d + batch_norm."""
    m = _loaded(cfpnet_b200.layers.LoFTREncoderLayer_newcross9(C, 4), seed=5).double()
    keep = Packer()
    m.pack(keep)
    attn, convs = keep[:4], keep[4:]
    assert len(convs) == 6
    H, W, B = 7, 9, 2
    g = torch.Generator().manual_seed(0)
    for i, (conv, bn, cin) in enumerate(((m.conv1, m.bn1, 2 * C), (m.conv2, m.bn2, C))):
        wt, shift, pk = convs[3 * i:3 * i + 3]
        x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64)
        want = F.batch_norm(F.conv2d(x, conv.weight, padding=1), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                            False, 0.0, bn.eps)
        xp = F.pad(x, (1, 1, 1, 1))
        # tensor-core blocks (bf16 storage: decode, then compare at bf16 weight precision)
        assert pk.shape[0] == (cin // C) * 9 and pk.dtype == torch.bfloat16
        got = torch.zeros(B, C, H, W, dtype=torch.float64)
        exact = torch.zeros_like(got)
        scale, _ = fold_bn(bn)
        for s in range(cin // C):
            for ky in range(3):
                for kx in range(3):
                    blk = from_umma(pk[s * 9 + ky * 3 + kx].reshape(-1), C, C).double()        # [Cout, K = C inputs of source s]
                    tap = xp[:, s * C:(s + 1) * C, ky:ky + H, kx:kx + W]
                    got += torch.einsum("ok,bkhw->bohw", blk, tap)
                    exact += torch.einsum("ok,bkhw->bohw", conv.weight[:, s * C:(s + 1) * C, ky, kx] * scale[:, None].double(), tap)
        got += shift.double().view(1, C, 1, 1)
        exact += shift.double().view(1, C, 1, 1)
        assert (exact - want).abs().max() <= 1e-5 * want.abs().max()               # fold is exact (fp32 scale / shift)
        assert (got - want).norm() / want.norm() <= 4e-3                           # + bf16 rounding of the weights
        # fp32 engine layout: row (tap*cin + c) of wt, column = output channel
        got32 = torch.zeros(B, C, H, W, dtype=torch.float64)
        for ky in range(3):
            for kx in range(3):
                rows = wt[(ky * 3 + kx) * cin:(ky * 3 + kx + 1) * cin].double()                # [cin, Cout]
                got32 += torch.einsum("ko,bkhw->bohw", rows, xp[:, :, ky:ky + H, kx:kx + W])
        got32 += shift.double().view(1, C, 1, 1)
        assert (got32 - want).abs().max() <= 1e-5 * want.abs().max()
    # the attention half: q block and k | v blocks are the plain projection weights
    assert torch.equal(from_umma(attn[2].reshape(-1), C, C), m.q_proj.weight.to(torch.bfloat16))
    assert torch.equal(from_umma(attn[3][1].reshape(-1), C, C), m.v_proj.weight.to(torch.bfloat16))


@pytest.mark.parametrize("C,nhead", [(32, 4), (64, 8), (128, 8)])
def test_loftr_chain_blocks_follow_the_consumption_order(C, nhead):
    """cfp_loftr_w.tc: eight [C x C] blocks in the order the chain consumes them (k_chain_tc.cu): Wq | Wm |
    W1 quadrants accumulated as  hidden[:, :C] = x B2^T + msg B3^T,  hidden[:, C:] = x B4^T + msg B5^T  |
    out = hidden[:, :C] B6^T + hidden[:, C:] B7^T.  Emulated in float64 against the nn.Linear layers."""
    m = _loaded(cfpnet_b200.layers.LoFTREncoderLayer(C, nhead), seed=9)
    keep = Packer()
    m.pack(keep)
    names = ["wq_t", "wkv_t", "wm_t", "w1_t", "w2_t", "ln1_g", "ln1_b", "ln2_g", "ln2_b", "tc", "kv_tc"]
    t = dict(zip(names, keep))
    blocks = [from_umma(t["tc"][i].reshape(-1), C, C).double() for i in range(8)]
    bf = lambda w: w.detach().to(torch.bfloat16).double()      # noqa: E731  (the blocks store bf16)
    g = torch.Generator().manual_seed(1)
    x, msg = torch.randn(11, C, generator=g, dtype=torch.float64), torch.randn(11, C, generator=g, dtype=torch.float64)
    assert torch.allclose(x @ blocks[0].t(), F.linear(x, bf(m.q_proj.weight)))
    assert torch.allclose(msg @ blocks[1].t(), F.linear(msg, bf(m.merge.weight)))
    hidden = torch.cat([x @ blocks[2].t() + msg @ blocks[3].t(), x @ blocks[4].t() + msg @ blocks[5].t()], dim=1)
    assert torch.allclose(hidden, F.linear(torch.cat([x, msg], dim=1), bf(m.mlp[0].weight)))
    hidden = torch.relu(hidden)
    out = hidden[:, :C] @ blocks[6].t() + hidden[:, C:] @ blocks[7].t()
    assert torch.allclose(out, F.linear(hidden, bf(m.mlp[2].weight)))
    kvb = [from_umma(t["kv_tc"][i].reshape(-1), C, C).double() for i in range(2)]
    assert torch.allclose(x @ kvb[0].t(), F.linear(x, bf(m.k_proj.weight))) and \
        torch.allclose(x @ kvb[1].t(), F.linear(x, bf(m.v_proj.weight)))
    # fp32 engine: [in][out] transposes, k | v side by side
    assert torch.equal(t["wkv_t"], torch.cat([m.k_proj.weight.t(), m.v_proj.weight.t()], dim=1))
    assert torch.equal(t["w1_t"], m.mlp[0].weight.t()) and torch.equal(t["w2_t"], m.mlp[2].weight.t())


@pytest.mark.parametrize("C,ws", [(32, 12), (128, 6)])
def test_gsa_subsampling_blocks_reproduce_the_strided_conv(C, ws):
    """sr_tc: one [C x C] block per tap (dy, dx) of the stride-ws, ws x ws conv (transformer.py:144-147): the kernel
    accumulates  out[token] = sum_taps x[token*ws + (dy,dx)] B_tap^T  over a split of the taps; sr_t is the fp32
    engine's [(dy, dx, cin)][Cout] layout."""
    m = _loaded(cfpnet_b200.layers.TwinsTransformer(C, ws=ws), seed=2)
    keep = Packer()
    m.pack(keep)
    sr_t, sr_b, _, _, sr_tc = keep[-5:]
    H, W = 2 * ws, 3 * ws
    x = torch.randn(1, C, H, W, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    w64 = m.gsa.sr.weight.detach().double()
    want = F.conv2d(x, w64, m.gsa.sr.bias.detach().double(), stride=ws)
    want_bf = F.conv2d(x, w64.to(torch.bfloat16).double(), m.gsa.sr.bias.detach().double(), stride=ws)
    got = torch.zeros_like(want)
    got32 = torch.zeros_like(want)
    for dy in range(ws):
        for dx in range(ws):
            tap = x[:, :, dy::ws, dx::ws]                                   # [1, C, H/ws, W/ws]
            got += torch.einsum("ok,bkhw->bohw", from_umma(sr_tc[dy * ws + dx].reshape(-1), C, C).double(), tap)
            got32 += torch.einsum("ko,bkhw->bohw", sr_t[(dy * ws + dx) * C:(dy * ws + dx + 1) * C].double(), tap)
    bias = sr_b.double().view(1, C, 1, 1)
    assert torch.allclose(got + bias, want_bf, rtol=1e-9, atol=1e-9)
    assert torch.allclose(got32 + bias, want, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ pack cache / flat buffer plumbing
def test_packer_lays_tensors_out_in_one_flat_buffer_and_relocate_turns_offsets_into_pointers():
    from cfpnet_b200.packing import relocate
    keep = Packer()
    a, b = torch.arange(5, dtype=torch.float32), torch.arange(7, dtype=torch.int16).to(torch.bfloat16)
    oa, ob = keep.ref(a), keep.ref(b)
    assert oa == Packer.ALIGN and ob == 2 * Packer.ALIGN and keep.ref(a) == oa         # 0 stays the null pointer
    flat = keep.upload("cpu")
    assert torch.equal(flat[oa:oa + 20].view(torch.float32), a) and torch.equal(flat[ob:ob + 14].view(torch.bfloat16), b)
    w = _lib.CfpTwinsW()
    w.lsa.wq_t, w.sr_b, w.ws = oa, ob, 6
    relocate(w, 1 << 20)
    assert w.lsa.wq_t == (1 << 20) + oa and w.sr_b == (1 << 20) + ob and w.ws == 6 and not w.gsa.wq_t and not w.sr_t
    h = _lib.CfpHistW()
    h.w_t[3] = oa
    relocate(h, 4096)
    assert h.w_t[3] == 4096 + oa and not h.w_t[0] and not h.tc


def test_data_parallel_replicas_own_their_pack_cache_and_scratch():
    """nn.DataParallel replicas share the original's __dict__ entries: a shared cache would publish device-0
    weight pointers to every replica.  Replicas get fresh caches, are never cached (their parameters are re-broadcast
    on every forward), and the original's cache is keyed by the caller's own tensors."""
    import copy
    args.attention_layer = list(synth.COMBINE1_LAYERS)
    m = cfpnet_b200.TransformerFusion(32, [120, 160], large_kernel=31, patch_size=16).eval()
    r = m._replicate_for_data_parallel()
    assert r._cache is not m._cache and r._scratch is not m._scratch and r._is_replica
    enc = cfpnet_b200.HistogramEncoder().eval()
    assert enc._replicate_for_data_parallel()._cache is not enc._cache
    calls = []
    build = lambda: calls.append(1) or len(calls)                                      # noqa: E731
    assert m._cache.get(m, build) == 1 and m._cache.get(m, build) == 1                 # cached for the original
    with torch.no_grad():
        m.positional_encodings.add_(1.0)                                               # version bump -> re-pack
    assert m._cache.get(m, build) == 2
    assert r._cache.get(r, build) == 3 and r._cache.get(r, build) == 4                 # replicas: never cached
    m2 = copy.deepcopy(m)                                                              # locks inside do not break deepcopy
    assert m2._cache is not m._cache and m2._cache.get(m2, build) == 5
