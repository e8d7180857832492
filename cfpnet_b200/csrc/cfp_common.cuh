// Shared device building blocks of libcfp (sm_100a).
//
// Everything here is hand-written CUDA for the CFP fusion path; no libraries.
// Layout conventions:
//   * activations are token-major ("NHWC"): [B, N=H*W, C], C innermost;
//   * weights are pre-packed by the host wrapper as W^T, i.e. [K][N] fp32 with N
//     innermost, so a K-chunk of a weight matrix is a set of contiguous rows that
//     cp.async can stream into shared memory without a transpose;
//   * the exact-fp32 engine (RowsGemm) keeps a BM-row activation tile resident in
//     shared memory across a whole chain of GEMMs; accumulation is fp32 FFMA.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace cfp {

typedef __nv_bfloat16 bf16;

constexpr int kThreads = 256;           // every row-chain kernel uses 8 warps
constexpr float kAttnEps = 1e-6f;       // attention.py:11
constexpr float kLnEps = 1e-5f;         // nn.LayerNorm default
constexpr float kLkpmLnEps = 1e-6f;     // convnext.py:31

// ---------------------------------------------------------------- dtype I/O
template <typename T> struct IO;
template <> struct IO<float> {
    static __device__ __forceinline__ float4 ld4(const float* p) {
        return *reinterpret_cast<const float4*>(p);
    }
    static __device__ __forceinline__ void st4(float* p, float4 v) {
        *reinterpret_cast<float4*>(p) = v;
    }
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct IO<bf16> {
    static __device__ __forceinline__ float4 ld4(const bf16* p) {
        uint2 u = *reinterpret_cast<const uint2*>(p);
        float4 v;
        v.x = __uint_as_float(u.x << 16);
        v.y = __uint_as_float(u.x & 0xffff0000u);
        v.z = __uint_as_float(u.y << 16);
        v.w = __uint_as_float(u.y & 0xffff0000u);
        return v;
    }
    static __device__ __forceinline__ void st4(bf16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
        __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&a);
        u.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = u;
    }
    static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------- small math
__device__ __forceinline__ float elu1(float x) {          // elu(x)+1, attention.py:10-11
    // branch-free: written as `x > 0 ? x + 1 : __expf(x)` ptxas put a divergent branch (BSSY / BSYNC pair) around every
    // exp - 116 of them in a C = 128 kv_state epilogue, which made that epilogue 6 us where the same walk with a relu
    // takes 1 us.  The exponent's argument is clamped so that the unused lane of the select cannot overflow.
    // ex2.approx.ftz directly: __expf without -ftz carries a denormal-range fix-up (compare with -126, halve, square) that
    // is three more instructions per value; results below 2^-126 may flush to zero here.
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 0.f) * 1.4426950408889634f));
    return x > 0.f ? x + 1.f : e;
}
__device__ __forceinline__ float gelu_erf(float x) {      // nn.GELU() default (erf form)
    return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
// GELU(erf) with the Abramowitz-Stegun 7.1.25 erf (|error| <= 2.5e-5, far below bf16 resolution):
// ~12 instructions instead of erff's ~30; used by the bf16 tensor-core MLP whose epilogue is
// ALU-bound at small C.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.f, fmaf(0.47047f, z, 1.f));
    const float poly = t * fmaf(t, fmaf(t, 0.7478556f, -0.0958798f), 0.3480242f);
    const float erf_abs = 1.f - poly * __expf(-z * z);
    return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------- cp.async
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------- RowsGemm
// acc[BM rows x NT cols] += A[BM x K] * Wt[K x NT]
//   A   : shared memory, fp32, row r at arow(r) (16-byte aligned, K contiguous)
//   Wt  : global fp32, row k at Wt + k*ldw (NT contiguous columns used)
//   wbuf: shared staging, 2 * KT * NT floats, 16-byte aligned
// 8 warps; warp w owns rows [w*TM, (w+1)*TM); lane owns TN columns (col(i)).
// The caller must __syncthreads() after producing A; run() ends with all of its
// own reads of wbuf complete (trailing __syncthreads()).
template <int BM, int NT, int KT = 32>
struct RowsGemm {
    static_assert(BM % 8 == 0 && NT % 32 == 0 && KT % 4 == 0, "tile shape");
    static constexpr int TM = BM / 8;
    static constexpr int TN = NT / 32;
    static constexpr int kWbufFloats = 2 * KT * NT;

    static __device__ __forceinline__ int col(int lane, int i) {
        if (TN >= 4) return (i >> 2) * 128 + lane * 4 + (i & 3);
        return lane * TN + i;
    }
    static __device__ __forceinline__ void zero(float (&acc)[TM][TN]) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    }
    static __device__ __forceinline__ void stage(float* dst, const float* __restrict__ Wt, int ldw, int k0) {
        constexpr int kVecPerRow = NT / 4;
        constexpr int kVecs = KT * kVecPerRow;
        for (int v = threadIdx.x; v < kVecs; v += kThreads) {
            int kk = v / kVecPerRow, c4 = v % kVecPerRow;
            cp_async16(dst + kk * NT + c4 * 4, Wt + (size_t)(k0 + kk) * ldw + c4 * 4);
        }
        cp_async_commit();
    }
    template <class ARow>
    static __device__ __forceinline__ void run(float (&acc)[TM][TN], ARow arow, const float* __restrict__ Wt,
                                               int ldw, int K, float* wbuf) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const float* arows[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) arows[i] = arow(warp * TM + i);
        const int nchunk = K / KT;
        stage(wbuf, Wt, ldw, 0);
        for (int ch = 0; ch < nchunk; ++ch) {
            float* cur = wbuf + (ch & 1) * (KT * NT);
            if (ch + 1 < nchunk) {
                stage(wbuf + ((ch + 1) & 1) * (KT * NT), Wt, ldw, (ch + 1) * KT);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const int k0 = ch * KT;
#pragma unroll
            for (int kk = 0; kk < KT; kk += 4) {
                float4 a[TM];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(arows[i] + k0 + kk);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float w[TN];
                    const float* wr = cur + (kk + j4) * NT;
                    if (TN >= 4) {
#pragma unroll
                        for (int g = 0; g < TN / 4; ++g) {
                            float4 t = *reinterpret_cast<const float4*>(wr + g * 128 + lane * 4);
                            w[g * 4 + 0] = t.x; w[g * 4 + 1] = t.y; w[g * 4 + 2] = t.z; w[g * 4 + 3] = t.w;
                        }
                    } else if (TN == 2) {
                        float2 t = *reinterpret_cast<const float2*>(wr + lane * 2);
                        w[0] = t.x; w[1] = t.y;
                    } else {
                        w[0] = wr[lane];
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        float av = j4 == 0 ? a[i].x : j4 == 1 ? a[i].y : j4 == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, w[j], acc[i][j]);
                    }
                }
            }
            __syncthreads();
        }
    }
    // visit every accumulator: f(row, col, value)
    template <class F>
    static __device__ __forceinline__ void foreach(float (&acc)[TM][TN], F f) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) f(warp * TM + i, col(lane, j), acc[i][j]);
    }
};

// Plain strided rows in shared memory.
struct SmemRows {
    const float* base;
    int ld;
    __device__ __forceinline__ const float* operator()(int r) const { return base + r * ld; }
};

// out[BM][N] (smem, row stride ldo) = epi(A[BM][K] * Wt[K][N]); N may exceed 256 (passes of NT).
template <int BM, int N, class ARow, class Epi>
__device__ __forceinline__ void gemm_to_smem(ARow arow, const float* __restrict__ Wt, int K, float* wbuf,
                                             float* out, int ldo, Epi epi) {
    constexpr int NT = N > 256 ? 256 : N;
    static_assert(N % NT == 0, "N must be a multiple of the pass width");
    using G = RowsGemm<BM, NT>;
#pragma unroll 1
    for (int n0 = 0; n0 < N; n0 += NT) {
        float acc[G::TM][G::TN];
        G::zero(acc);
        G::run(acc, arow, Wt + n0, N, K, wbuf);
        G::foreach(acc, [&](int r, int c, float v) { out[r * ldo + n0 + c] = epi(n0 + c, v); });
    }
}

// LayerNorm over C contiguous floats of each of BM smem rows, in place (warp per row).
template <int BM, int C>
__device__ __forceinline__ void layernorm_rows(float* buf, int ld, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PER = C / 32;
    for (int r = warp; r < BM; r += kThreads / 32) {
        float* row = buf + r * ld;
        float v[PER], s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = row[lane + 32 * i]; s += v[i]; }
        const float mean = warp_sum(s) * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] -= mean; q += v[i] * v[i]; }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = lane + 32 * i;
            row[c] = v[i] * rstd * gamma[c] + beta[c];
        }
    }
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the bf16 path is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may
// become resident while the previous kernel of the stream is still draining, run their prologue (barrier init, TMEM
// allocation, constant shared-memory columns, first weight copies - nothing the previous kernel produced) and block
// in pdl_wait() until the previous grid has completed and its writes are visible.  Rules followed everywhere:
//   * pdl_trigger() right after the prologue, so the next kernel can be scheduled as soon as the last wave of this
//     grid is resident;
//   * every thread that reads or writes an activation / workspace buffer calls pdl_wait() first (weights and packed
//     constants may be fetched before it);
//   * at least the row / epilogue threads of every kernel wait, so grid N cannot finish before grid N-1 and the
//     ordering stays transitive along the stream.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// ---------------------------------------------------------------- host side
struct ErrorState {
    char msg[512];
};
ErrorState& tls_error();
int fail(const char* fmt, ...);
int check_launch(const char* what);

#define CFP_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) return ::cfp::fail(__VA_ARGS__); \
    } while (0)

bool pdl_enabled();      // CFP_NO_PDL=1 launches everything fully serialised (A/B measurements)
int sm_count();          // multiprocessors of the current device (cached per device)

// k<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute
template <class... P, class... A>
inline void launch_pdl(void (*k)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, k, P(args)...);      // a failure is picked up by check_launch (cudaGetLastError)
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e));
    }
    return 0;
}

}  // namespace cfp
