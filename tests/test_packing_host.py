"""Host-side weight packing checked on the CPU against plain PyTorch ops: the packed buffers are decoded with the
index formulas the kernels use (csrc/umma.cuh, k_dwconv_tc.cu, k_chain_tc.cu) and the kernels' arithmetic is emulated
in float64, so a layout or fold mistake in ``cfpnet_b200/layers.py`` / ``encoder.py`` fails here without a GPU."""
import pytest
import torch
import torch.nn.functional as F

import cfpnet_b200
from cfpnet_b200 import _lib, synth
from cfpnet_b200.config import args
from cfpnet_b200.packing import fold_bn, umma_block


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
    keep = []
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


@pytest.mark.parametrize("C", [32, 64])
def test_mlp_fold_reproduces_layernorm_mlp(C):
    """W1' = W1 diag(g), b1' = b1 + W1 b ride in [LNhat(y) | 1]; b2 rides in [GELU(h) | 1] of slice 0 (MlpTC)."""
    blk = make_block(C, 7)
    keep = []
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
    w, (stages, tc) = enc.eval()._pack()
    off = 0
    for wt, _b in stages[1:]:                                   # stage i: wt [Cin, Cout] -> block [Cin/8][Cout][8]
        cin, cout = wt.shape
        blk = tc[off:off + cin * cout]
        assert torch.equal(from_umma(blk, cout, cin), wt.t().to(torch.bfloat16))
        off += cin * cout
    assert off == tc.numel() == 106496 // 2 and ctypes_ptr_ok(w)


def ctypes_ptr_ok(w):
    return bool(w.tc) and all(bool(w.w_t[i]) and bool(w.b[i]) for i in range(9))
