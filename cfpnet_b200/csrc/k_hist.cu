// a1. Per-zone ToF histogram encoder (src/models/encoder.py:6-50).
//
// Three chained PointNet blocks, each 3 x [Conv1d(k=1) + BatchNorm1d + ReLU], on every
// depth sample independently: 1 -> 32 -> 32 -> 32 | -> 64 -> 64 -> 64 | -> 128 -> 128 -> 128.
// Eval-mode BN is folded into the pointwise weights by the host wrapper, so a stage is
// relu(x W^T + b).  One CTA keeps a 64-sample tile resident in shared memory through all
// nine stages (55 k weights stream through cp.async, L2-resident) and writes the three
// token tensors the decoder levels consume; the only HBM traffic is 4 B in and
// (32+64+128) elements out per sample.
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

constexpr int kHistBM = 64;
constexpr int kHistLd = 128 + 4;

template <typename T, int C>
__device__ __forceinline__ void store_tile(const float* buf, T* __restrict__ out, int64_t row0, int64_t rows) {
    constexpr int V = C / 4;
    for (int i = threadIdx.x; i < kHistBM * V; i += kThreads) {
        int r = i / V, c4 = i % V;
        if (row0 + r < rows) {
            float4 v = *reinterpret_cast<const float4*>(buf + r * kHistLd + c4 * 4);
            IO<T>::st4(out + (row0 + r) * C + c4 * 4, v);
        }
    }
}

template <int N, int K>
__device__ __forceinline__ void hist_stage(const float* in, float* out, const float* __restrict__ wt,
                                           const float* __restrict__ bias, float* wbuf) {
    gemm_to_smem<kHistBM, N>(SmemRows{in, kHistLd}, wt, K, wbuf, out, kHistLd,
                             [&](int c, float v) { return fmaxf(v + bias[c], 0.f); });
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kThreads) hist_encoder_kernel(const float* __restrict__ hist, T* __restrict__ o32,
                                                                T* __restrict__ o64, T* __restrict__ o128,
                                                                int64_t rows, cfp_hist_w w) {
    extern __shared__ __align__(16) float smem[];
    float* bufA = smem;
    float* bufB = bufA + kHistBM * kHistLd;
    float* wbuf = bufB + kHistBM * kHistLd;
    const int64_t row0 = (int64_t)blockIdx.x * kHistBM;

    // stage 0: K = 1
    for (int i = threadIdx.x; i < kHistBM * 32; i += kThreads) {
        int r = i >> 5, c = i & 31;
        float x = row0 + r < rows ? hist[row0 + r] : 0.f;
        bufA[r * kHistLd + c] = fmaxf(fmaf(x, w.w_t[0][c], w.b[0][c]), 0.f);
    }
    __syncthreads();
    hist_stage<32, 32>(bufA, bufB, w.w_t[1], w.b[1], wbuf);
    hist_stage<32, 32>(bufB, bufA, w.w_t[2], w.b[2], wbuf);
    store_tile<T, 32>(bufA, o32, row0, rows);
    hist_stage<64, 32>(bufA, bufB, w.w_t[3], w.b[3], wbuf);
    hist_stage<64, 64>(bufB, bufA, w.w_t[4], w.b[4], wbuf);
    hist_stage<64, 64>(bufA, bufB, w.w_t[5], w.b[5], wbuf);
    store_tile<T, 64>(bufB, o64, row0, rows);
    hist_stage<128, 64>(bufB, bufA, w.w_t[6], w.b[6], wbuf);
    hist_stage<128, 128>(bufA, bufB, w.w_t[7], w.b[7], wbuf);
    hist_stage<128, 128>(bufB, bufA, w.w_t[8], w.b[8], wbuf);
    store_tile<T, 128>(bufA, o128, row0, rows);
}

template <typename T>
static int launch(const float* hist, void* o32, void* o64, void* o128, int64_t rows, const cfp_hist_w& w,
                  cudaStream_t st) {
    const size_t smem = (2 * kHistBM * kHistLd + 2 * 32 * 128) * sizeof(float);
    if (int e = set_smem(hist_encoder_kernel<T>, smem)) return e;
    const unsigned grid = (unsigned)((rows + kHistBM - 1) / kHistBM);
    hist_encoder_kernel<T><<<grid, kThreads, smem, st>>>(hist, (T*)o32, (T*)o64, (T*)o128, rows, w);
    return check_launch("hist_encoder_kernel");
}

int hist_encoder(const float* hist, void* o32, void* o64, void* o128, int64_t rows, const cfp_hist_w& w,
                 int dtype, cudaStream_t st) {
    return dtype == CFP_F32 ? launch<float>(hist, o32, o64, o128, rows, w, st)
                            : launch<bf16>(hist, o32, o64, o128, rows, w, st);
}

}  // namespace cfp
