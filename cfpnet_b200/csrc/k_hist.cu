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
#include "umma.cuh"

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

// ---------------------------------------------------------------------------------------
// bf16 path: the eight K >= 32 stages run on the tensor cores (tcgen05.mma, M = 128 samples, N = Cout,
// fp32 accumulators in TMEM).  All eight weight blocks (104 KB bf16, canonical K-major UMMA layout,
// umma.cuh) are loaded ONCE per CTA by the bulk-copy engine and stay resident in shared memory; a CTA
// runs two independent 128-sample pipelines (warps 0-3 and 4-7: thread r <-> sample r <-> TMEM lane r)
// so that the MMA of one overlaps the bias + ReLU epilogue of the other.  A stage's epilogue writes the
// bf16 activations back into the pipeline's single A buffer in place (its MMA has completed) and, after
// stages 2 / 5 / 8, to the three token tensors.
struct HistTC {
    static constexpr uint32_t LBO = 129 * 16;                 // A buffer: [16 k-groups][128 rows + pad][16 B]
    static constexpr int ABUF = 16 * (int)LBO;
    static constexpr int WBYTES = 2 * (32 * 32 * 2 + 64 * 32 + 64 * 64 * 2 + 128 * 64 + 128 * 128 * 2);   // 106496
    static constexpr size_t SMEM = (size_t)WBYTES + 2 * (size_t)ABUF;
};
struct HistBars {
    uint64_t w_full, acc_ready[2];
    uint32_t tmem_slot;
};

// one stage of one pipeline: D = A[128 x K] * W^T (N columns), then relu(D + b) -> A (in place) [+ global]
template <int K, int N, int WOFF, bool kStore>
__device__ __forceinline__ void hist_tc_stage(uint8_t* a, const uint8_t* wsm, const float* __restrict__ bias, uint32_t tmem,
                                              uint64_t* acc_ready, uint32_t& ph, int grp, int wq, int tid_g, bf16* __restrict__ out,
                                              int64_t row, int64_t rows) {
    umma::fence_async_smem();
    umma::fence_before_sync();
    asm volatile("bar.sync %0, 128;\n" ::"r"(1 + grp) : "memory");
    if (wq == 0) {
        umma::fence_after_sync();
        uint64_t ad = umma::smem_desc(umma::smem_u32(a), HistTC::LBO);
        uint64_t wd = umma::smem_desc(umma::smem_u32(wsm + WOFF), N * 16);
        constexpr uint32_t idesc = umma::idesc_bf16(128, N);
#pragma unroll
        for (int ks = 0; ks < K / 16; ++ks) {
            umma::mma_bf16(tmem, ad, wd, idesc, ks > 0);
            ad = umma::desc_advance(ad, 2 * HistTC::LBO);
            wd = umma::desc_advance(wd, 2 * N * 16);
        }
        umma::commit(acc_ready);
    }
    umma::mbar_wait(acc_ready, ph); ph ^= 1;
    umma::fence_after_sync();
    umma::tmem_for_each16<N>(umma::tmem_addr(tmem, wq * 32, 0), [&](int c0, const float (&t)[16]) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + c0 + j), b1 = *reinterpret_cast<const float4*>(bias + c0 + j + 4);
            uint4 u;
            u.x = umma::pack_bf16(fmaxf(t[j + 0] + b0.x, 0.f), fmaxf(t[j + 1] + b0.y, 0.f));
            u.y = umma::pack_bf16(fmaxf(t[j + 2] + b0.z, 0.f), fmaxf(t[j + 3] + b0.w, 0.f));
            u.z = umma::pack_bf16(fmaxf(t[j + 4] + b1.x, 0.f), fmaxf(t[j + 5] + b1.y, 0.f));
            u.w = umma::pack_bf16(fmaxf(t[j + 6] + b1.z, 0.f), fmaxf(t[j + 7] + b1.w, 0.f));
            *reinterpret_cast<uint4*>(a + (size_t)((c0 + j) / 8) * HistTC::LBO + tid_g * 16) = u;
            if (kStore && row < rows) *reinterpret_cast<uint4*>(out + row * N + c0 + j) = u;
        }
    });
}

__global__ void __launch_bounds__(256, 1) hist_encoder_tc_kernel(const float* __restrict__ hist, bf16* __restrict__ o32,
                                                                 bf16* __restrict__ o64, bf16* __restrict__ o128, int64_t rows,
                                                                 cfp_hist_w w, int ntiles) {
    extern __shared__ __align__(128) uint8_t smem_tc[];
    __shared__ HistBars bars;
    uint8_t* wsm = smem_tc;
    const int tid = threadIdx.x, warp = umma::warp_idx_sync();
    const int grp = warp >> 2, wq = warp & 3, tid_g = tid & 127;
    uint8_t* a = smem_tc + HistTC::WBYTES + (size_t)grp * HistTC::ABUF;

    if (tid == 0) {
        umma::mbar_init(&bars.w_full, 1);
        umma::mbar_init(&bars.acc_ready[0], 1);
        umma::mbar_init(&bars.acc_ready[1], 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&bars.tmem_slot, 256);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot + grp * 128;
    if (warp == 0 && umma::elect_one()) {                      // all weight blocks, once
        umma::mbar_expect_tx(&bars.w_full, HistTC::WBYTES);
        for (int off = 0; off < HistTC::WBYTES; off += 16384)
            umma::bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(w.tc) + off,
                           min(16384, HistTC::WBYTES - off), &bars.w_full);
    }
    pdl_wait();                            // the token tensors may still be read by the previous forward's kernels
    uint32_t ph = 0;
    bool w_ready = false;
    for (int tile = blockIdx.x * 2 + grp; tile < ntiles; tile += gridDim.x * 2) {
        const int64_t row = (int64_t)tile * 128 + tid_g;
        const float x = row < rows ? hist[row] : 0.f;
        // stage 0 (K = 1): relu(x * w0 + b0) -> A[:, 0:32)
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            float o8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o8[i] = fmaxf(fmaf(x, __ldg(w.w_t[0] + j + i), __ldg(w.b[0] + j + i)), 0.f);
            umma::store_chunk(a, HistTC::LBO, tid_g, j / 8, o8);
        }
        if (!w_ready) { umma::mbar_wait(&bars.w_full, 0); w_ready = true; }
        constexpr int O1 = 0, O2 = O1 + 2048, O3 = O2 + 2048, O4 = O3 + 4096, O5 = O4 + 8192, O6 = O5 + 8192, O7 = O6 + 16384,
                      O8 = O7 + 32768;
        uint64_t* bar = &bars.acc_ready[grp];
        hist_tc_stage<32, 32, O1, false>(a, wsm, w.b[1], tmem, bar, ph, grp, wq, tid_g, nullptr, row, rows);
        hist_tc_stage<32, 32, O2, true>(a, wsm, w.b[2], tmem, bar, ph, grp, wq, tid_g, o32, row, rows);
        hist_tc_stage<32, 64, O3, false>(a, wsm, w.b[3], tmem, bar, ph, grp, wq, tid_g, nullptr, row, rows);
        hist_tc_stage<64, 64, O4, false>(a, wsm, w.b[4], tmem, bar, ph, grp, wq, tid_g, nullptr, row, rows);
        hist_tc_stage<64, 64, O5, true>(a, wsm, w.b[5], tmem, bar, ph, grp, wq, tid_g, o64, row, rows);
        hist_tc_stage<64, 128, O6, false>(a, wsm, w.b[6], tmem, bar, ph, grp, wq, tid_g, nullptr, row, rows);
        hist_tc_stage<128, 128, O7, false>(a, wsm, w.b[7], tmem, bar, ph, grp, wq, tid_g, nullptr, row, rows);
        hist_tc_stage<128, 128, O8, true>(a, wsm, w.b[8], tmem, bar, ph, grp, wq, tid_g, o128, row, rows);
    }
    pdl_trigger();
    if (!w_ready) umma::mbar_wait(&bars.w_full, 0);           // never leave with a bulk copy in flight
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        umma::fence_after_sync();
        umma::tmem_dealloc(bars.tmem_slot, 256);
    }
}

static int launch_tc(const float* hist, void* o32, void* o64, void* o128, int64_t rows, const cfp_hist_w& w, cudaStream_t st) {
    CFP_REQUIRE(w.tc != nullptr, "hist_encoder: bf16 path needs the packed tensor-core weights (cfp_hist_w.tc)");
    if (int e = set_smem(hist_encoder_tc_kernel, HistTC::SMEM)) return e;
    const int64_t ntiles = (rows + 127) / 128;
    CFP_REQUIRE(ntiles < ((int64_t)1 << 30), "hist_encoder: too many rows");
    const int grid = (int)((ntiles + 1) / 2 < sm_count() ? (ntiles + 1) / 2 : sm_count());
    launch_pdl(hist_encoder_tc_kernel, grid, 256, HistTC::SMEM, st, hist, (bf16*)o32, (bf16*)o64, (bf16*)o128, rows, w, (int)ntiles);
    return check_launch("hist_encoder_tc");
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
    return dtype == CFP_F32 ? launch<float>(hist, o32, o64, o128, rows, w, st) : launch_tc(hist, o32, o64, o128, rows, w, st);
}

}  // namespace cfp
