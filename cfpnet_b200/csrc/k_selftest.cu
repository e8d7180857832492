// Known-answer self-test of the tcgen05 engine (umma.cuh): one CTA computes
//   D[m][n] = sum_k A[m + row_shift][k] * B[n][k],   m < 128, n < N, k < K
// from operands staged in the canonical K-major no-swizzle layout.  `row_shift` exercises the
// shifted-view property the 3x3 implicit-GEMM convolution relies on.  Used by tests only.
#include <cuda_bf16.h>
#include "cfp_common.cuh"
#include "cfp_internal.h"
#include "umma.cuh"

namespace cfp {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B,
                                                            float* __restrict__ D, int rows_a, int N, int K,
                                                            int row_shift, int tmem_cols) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const uint32_t lbo_a = rows_a * 16, lbo_b = N * 16;
    uint8_t* a_tile = smem_raw;
    uint8_t* b_tile = a_tile + (K / 8) * lbo_a;
    const int tid = threadIdx.x, warp = umma::warp_idx_sync();

    for (int i = tid; i < rows_a * (K / 8); i += 128) {
        int r = i / (K / 8), kg = i % (K / 8);
        *reinterpret_cast<uint4*>(a_tile + (size_t)kg * lbo_a + r * 16) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kg * 8);
    }
    for (int i = tid; i < N * (K / 8); i += 128) {
        int r = i / (K / 8), kg = i % (K / 8);
        *reinterpret_cast<uint4*>(b_tile + (size_t)kg * lbo_b + r * 16) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + kg * 8);
    }
    if (tid == 0) {
        umma::mbar_init(&mbar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, tmem_cols);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        const uint32_t idesc = umma::idesc_bf16(128, N);
        const uint32_t a0 = umma::smem_u32(a_tile) + row_shift * 16, b0 = umma::smem_u32(b_tile);
        for (int ks = 0; ks < K / 16; ++ks)
            umma::mma_bf16(tmem, umma::smem_desc(a0 + ks * 2 * lbo_a, lbo_a), umma::smem_desc(b0 + ks * 2 * lbo_b, lbo_b),
                           idesc, ks > 0);
        umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, 0);
    umma::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, c0), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, tmem_cols);
}

int umma_selftest(const void* A, const void* B, float* D, int rows_a, int N, int K, int row_shift, cudaStream_t st) {
    CFP_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K > 0, "selftest: bad N=%d K=%d", N, K);
    CFP_REQUIRE(row_shift >= 0 && rows_a >= 128 + row_shift, "selftest: rows_a=%d too small", rows_a);
    int cols = 32;
    while (cols < N) cols *= 2;
    const size_t smem = (size_t)(K / 8) * (rows_a + N) * 16;
    CFP_REQUIRE(smem <= 200 * 1024, "selftest: %zu B smem", smem);
    if (int e = set_smem(umma_selftest_kernel, smem)) return e;
    umma_selftest_kernel<<<1, 128, smem, st>>>((const bf16*)A, (const bf16*)B, D, rows_a, N, K, row_shift, cols);
    return check_launch("umma_selftest_kernel");
}

}  // namespace cfp
