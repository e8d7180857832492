// Tooling (not part of libcfp): how long does one tcgen05.mma (M = 128, K = 16, bf16, no-swizzle K-major operands as
// libcfp stages them) take as a function of N, of the A view's start alignment and of the operand stride?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -shared -Xcompiler -fPIC \
//        -I cfpnet_b200/csrc -I include tools/probe/mma_probe.cu -o tools/probe/libmma_probe.so
// One CTA per SM; the elected lane of warp 0 issues `reps` MMAs back to back (operands: zeros in shared memory), one commit,
// clock64 around the issue + completion wait.  out[cta] = clocks per MMA.
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
namespace cfp { typedef __nv_bfloat16 bf16; }
#include "umma.cuh"
using namespace cfp;

__global__ void __launch_bounds__(128) mma_probe_kernel(float* out, int N, int a_off, int lbo_a, int ksteps, int reps, int nacc,
                                                        int b_same, int sbo_a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = umma::warp_idx_sync();
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = umma::idesc_bf16(128, N);
        const uint32_t a0 = umma::smem_u32(smem) + a_off, b0 = umma::smem_u32(smem) + 96 * 1024;
        const uint32_t lbo_b = N * 16;
        // descriptors of the four unrolled issues are loop-invariant: the timed loop is four predicated UTCHMMA + a counter
        uint64_t ad[4], bd[4];
        uint32_t dc[4];
        for (int u = 0; u < 4; ++u) {
            const int ks = u % ksteps;
            ad[u] = umma::smem_desc(a0 + ks * 2 * lbo_a, lbo_a, sbo_a);
            bd[u] = umma::smem_desc(b0 + (b_same ? 0 : ks * 2 * lbo_b), lbo_b);
            dc[u] = tmem + (u % nacc) * 256;
        }
        long long t0 = clock64();
        if (umma::elect_one()) {
#pragma unroll 1
            for (int r = 0; r < reps; r += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dc[u]), "l"(ad[u]), "l"(bd[u]),
                                 "r"(idesc), "r"(1u)
                                 : "memory");
            }
        }
        __syncwarp();
        umma::commit(&mbar);
        umma::mbar_wait(&mbar, 0);
        long long t1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x] = (float)(t1 - t0) / reps;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) { umma::fence_after_sync(); umma::tmem_dealloc(tmem, 512); }
}

extern "C" int mma_probe(float* out, int ctas, int N, int a_off, int lbo_a, int ksteps, int reps, int nacc, int b_same, int sbo_a) {
    cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    mma_probe_kernel<<<ctas, 128, 160 * 1024>>>(out, N, a_off, lbo_a, ksteps, reps, nacc, b_same, sbo_a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "mma_probe: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
