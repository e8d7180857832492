// Dense spatial convolutions of the fusion path, as implicit GEMMs on token-major maps:
//   a6  DAPM  conv3x3(2C->C)+BN, conv3x3(C->C)+BN (+ residual)   (transformer.py:239-247)
//   a9  GSA   stride-ws "sr" conv + LayerNorm                    (transformer.py:144-147)
// Eval-mode BN is folded by the host wrapper: weights carry the BN scale, `shift` the rest.
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

// ------------------------------------------------------------------ conv3x3
// CTA = one TH x TW pixel tile of one frame x all C output channels.  The (TH+2)x(TW+2) halo
// of all input channels is staged once in shared memory; the 9 taps are 9 GEMM calls whose A
// rows are shifted views of the halo (no im2col copy).  The second input of DAPM's conv1 is
// the message map, which is zero inside the zone rectangle by construction
// (transformer.py:233-234): those cells are never read from memory.
template <int C> struct ConvTile {
    static constexpr int TH = C >= 128 ? 4 : 8, TW = 8, BM = TH * TW;
};

template <int C, typename T, bool kTwoSrc>
__global__ void __launch_bounds__(kThreads) conv3x3_kernel(const T* __restrict__ in0, const T* __restrict__ in1,
                                                           const float* __restrict__ w_t,
                                                           const float* __restrict__ shift,
                                                           const T* __restrict__ residual, T* __restrict__ out,
                                                           int H, int W, int zy0, int zy1, int zx0, int zx1) {
    constexpr int TH = ConvTile<C>::TH, TW = ConvTile<C>::TW, BM = ConvTile<C>::BM;
    constexpr int CIN = kTwoSrc ? 2 * C : C, LDH = CIN + 4, HW2 = TW + 2;
    extern __shared__ __align__(16) float smem[];
    float* halo = smem;                                  // [(TH+2)*(TW+2)][LDH]
    float* wbuf = halo + (TH + 2) * HW2 * LDH;
    const int b = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const size_t frame = (size_t)b * H * W;

    constexpr int V = CIN / 4;
    for (int i = threadIdx.x; i < (TH + 2) * HW2 * V; i += kThreads) {
        const int cell = i / V, c = (i % V) * 4;
        const int y = y0 - 1 + cell / HW2, x = x0 - 1 + cell % HW2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const size_t n = frame + (size_t)y * W + x;
            if (c < C) v = IO<T>::ld4(in0 + n * C + c);
            else if (!(y >= zy0 && y < zy1 && x >= zx0 && x < zx1)) v = IO<T>::ld4(in1 + n * C + (c - C));
        }
        *reinterpret_cast<float4*>(halo + cell * LDH + c) = v;
    }
    __syncthreads();

    using G = RowsGemm<BM, C>;
    float acc[G::TM][G::TN];
    G::zero(acc);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
        auto arow = [&](int r) { return halo + ((r / TW + dy) * HW2 + (r % TW + dx)) * LDH; };
        G::run(acc, arow, w_t + (size_t)tap * CIN * C, C, CIN, wbuf);
    }
    G::foreach(acc, [&](int r, int c, float v) {
        const int y = y0 + r / TW, x = x0 + r % TW;
        if (y < H && x < W) {
            const size_t o = (frame + (size_t)y * W + x) * C + c;
            v += shift[c];
            if (residual) v += IO<T>::ld(residual + o);
            IO<T>::st(out + o, v);
        }
    });
}

template <int C, typename T>
static int conv3x3_impl(const void* in0, const void* in1, const float* w_t, const float* shift, const void* residual,
                        void* out, int B, int H, int W, int zy0, int zy1, int zx0, int zx1, cudaStream_t st) {
    constexpr int TH = ConvTile<C>::TH, TW = ConvTile<C>::TW;
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
    const int cin = in1 ? 2 * C : C;
    const size_t smem = (size_t)((TH + 2) * (TW + 2) * (cin + 4) + 2 * 32 * C) * sizeof(float);
    if (in1) {
        auto k = conv3x3_kernel<C, T, true>;
        if (int e = set_smem(k, smem)) return e;
        k<<<grid, kThreads, smem, st>>>((const T*)in0, (const T*)in1, w_t, shift, (const T*)residual, (T*)out, H, W,
                                        zy0, zy1, zx0, zx1);
    } else {
        auto k = conv3x3_kernel<C, T, false>;
        if (int e = set_smem(k, smem)) return e;
        k<<<grid, kThreads, smem, st>>>((const T*)in0, nullptr, w_t, shift, (const T*)residual, (T*)out, H, W, 0, 0,
                                        0, 0);
    }
    return check_launch(in1 ? "conv3x3<2C->C>" : "conv3x3<C->C>");
}

// ------------------------------------------------------------------ GSA sr conv + LN
// out[b][(i,j)][:] = LN( sum_{dy,dx} feat[b][i*ws+dy][j*ws+dx][:] W[dy][dx] + bias ),  fp32 out.
// Non-overlapping ws x ws patches, no padding: Ns = (H/ws)*(W/ws) rows per frame.
template <int C, typename T>
__global__ void __launch_bounds__(kThreads) sr_conv_ln_kernel(const T* __restrict__ feat, float* __restrict__ sr_tok,
                                                              int64_t rows, int H, int W, int ws, int nsx, int Ns,
                                                              const float* __restrict__ sr_t,
                                                              const float* __restrict__ sr_b,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta) {
    constexpr int BM = C >= 128 ? 32 : 64, LD = C + 4;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                 // [BM][LD]
    float* wbuf = xs + BM * LD;
    __shared__ int64_t base[BM];      // element offset of the patch's top-left cell, -1 if row is padding
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    if (threadIdx.x < BM) {
        int64_t r = row0 + threadIdx.x;
        if (r < rows) {
            int b = (int)(r / Ns), s = (int)(r % Ns);
            base[threadIdx.x] = ((int64_t)b * H * W + (int64_t)(s / nsx) * ws * W + (s % nsx) * ws) * C;
        } else base[threadIdx.x] = -1;
    }
    using G = RowsGemm<BM, C>;
    float acc[G::TM][G::TN];
    G::zero(acc);
    __syncthreads();
#pragma unroll 1
    for (int tap = 0; tap < ws * ws; ++tap) {
        const int64_t toff = ((int64_t)(tap / ws) * W + tap % ws) * C;
        for (int i = threadIdx.x; i < BM * (C / 4); i += kThreads) {
            int r = i / (C / 4), c = (i % (C / 4)) * 4;
            float4 v = base[r] >= 0 ? IO<T>::ld4(feat + base[r] + toff + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(xs + r * LD + c) = v;
        }
        __syncthreads();
        G::run(acc, SmemRows{xs, LD}, sr_t + (size_t)tap * C * C, C, C, wbuf);   // ends with __syncthreads
    }
    G::foreach(acc, [&](int r, int c, float v) { xs[r * LD + c] = v + sr_b[c]; });
    __syncthreads();
    layernorm_rows<BM, C>(xs, LD, gamma, beta, kLnEps);
    __syncthreads();
    for (int i = threadIdx.x; i < BM * (C / 4); i += kThreads) {
        int r = i / (C / 4), c = (i % (C / 4)) * 4;
        if (row0 + r < rows)
            *reinterpret_cast<float4*>(sr_tok + (row0 + r) * C + c) = *reinterpret_cast<const float4*>(xs + r * LD + c);
    }
}

template <int C, typename T>
static int sr_impl(const void* feat0, float* sr_tok, int B, int H, int W, int ws, const float* sr_t, const float* sr_b,
                   const float* g, const float* b, cudaStream_t st) {
    constexpr int BM = C >= 128 ? 32 : 64;
    const int nsx = W / ws, Ns = (H / ws) * nsx;
    const int64_t rows = (int64_t)B * Ns;
    if (rows == 0) return 0;
    const size_t smem = (size_t)(BM * (C + 4) + 2 * 32 * C) * sizeof(float);
    auto k = sr_conv_ln_kernel<C, T>;
    if (int e = set_smem(k, smem)) return e;
    k<<<(unsigned)((rows + BM - 1) / BM), kThreads, smem, st>>>((const T*)feat0, sr_tok, rows, H, W, ws, nsx, Ns, sr_t,
                                                                sr_b, g, b);
    return check_launch("sr_conv_ln_kernel");
}

#define CFP_DISPATCH_C_T(FN, ...)                                                        \
    if (dtype == CFP_F32) {                                                              \
        if (C == 32) return FN<32, float>(__VA_ARGS__);                                  \
        if (C == 64) return FN<64, float>(__VA_ARGS__);                                  \
        if (C == 128) return FN<128, float>(__VA_ARGS__);                                \
    } else if (dtype == CFP_BF16) {                                                      \
        if (C == 32) return FN<32, bf16>(__VA_ARGS__);                                   \
        if (C == 64) return FN<64, bf16>(__VA_ARGS__);                                   \
        if (C == 128) return FN<128, bf16>(__VA_ARGS__);                                 \
    }                                                                                    \
    return fail("unsupported (C=%d, dtype=%d): libcfp serves C in {32,64,128}, fp32/bf16", C, dtype);

int conv3x3(const void* in0, const void* in1, const float* w_t, const float* shift, const void* residual, void* out,
            int B, int H, int W, int C, int zy0, int zy1, int zx0, int zx1, int dtype, cudaStream_t st) {
    CFP_DISPATCH_C_T(conv3x3_impl, in0, in1, w_t, shift, residual, out, B, H, W, zy0, zy1, zx0, zx1, st)
}
int sr_conv_ln(const void* feat0, float* sr_tok, int B, int H, int W, int C, int ws, const float* sr_t,
               const float* sr_b, const float* g, const float* b, int dtype, cudaStream_t st) {
    CFP_DISPATCH_C_T(sr_impl, feat0, sr_tok, B, H, W, ws, sr_t, sr_b, g, b, st)
}

}  // namespace cfp
