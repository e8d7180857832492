// Fused LoFTR query chain on the 5th-gen tensor cores (bf16 path):
//
//   x --Wq--> q --elu+1--> Q --(Q KV)/(Q.Ksum+eps)--> msg --Wm--> LN1 --[x|msg] W1--> relu --W2--> LN2 --(+x)--> out
//
// (transformer.py:54-71, attention.py:31-49).  One CTA keeps a 128-token tile on chip through the
// whole chain: the five GEMMs run as tcgen05.mma (M=128, N=C, K=16) with fp32 accumulators in
// TMEM; between GEMMs the 128 "row" threads (thread r <-> token r <-> TMEM lane r) read their
// accumulator row with tcgen05.ld, apply the row-wise math in registers and write the next A
// operand straight into shared memory in the canonical K-major UMMA layout (umma.cuh).  The tile is
// read from HBM once and written once.
//
// Weights: every GEMM of the chain is cut into [C x C] bf16 blocks (8 per tile: Wq, Wm, four
// quadrants of W1, two K-halves of W2), pre-packed by the host in consumption order; a producer
// lane streams them through a shared-memory ring with the bulk-copy (TMA) engine while MMAs and
// row math of earlier stages run.
//
//   warps 0-3 : row threads (stage x, epilogues)      warp 4 : weight producer      warp 5 : MMA issuer
//
// The same kernel serves hist2image, LSA, GSA (full chain) and the DAPM attention (q -> msg only)
// through the row providers of providers.cuh.
#include "cfp_common.cuh"
#include "cfp_internal.h"
#include "providers.cuh"
#include "umma.cuh"

namespace cfp {

template <int C> struct ChainTC {
    static constexpr int KG = C / 8;                       // 16-byte k-groups per C columns
    static constexpr uint32_t LBO = 129 * 16;              // 128 rows + one pad chunk: conflict-free staging
    static constexpr int SLOT = 2 * C * C;                 // bytes of one [C x C] bf16 weight block
    static constexpr int NSLOT = C >= 128 ? 2 : 4;
    static constexpr int ABUF = 2 * KG * (int)LBO;         // a tile of up to 2C columns
    static constexpr int TMEM_COLS = 2 * C < 32 ? 32 : 2 * C;
    static constexpr size_t SMEM = 2 * (size_t)ABUF + (size_t)NSLOT * SLOT;
};

struct ChainBars {
    uint64_t full[4], empty[4], a_ready, acc_ready;
    uint32_t tmem_slot;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rows_sync() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

__device__ __forceinline__ void unpack8(uint4 u, float (&v)[8]) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}

// Row r's accumulator columns [0, C) -> registers.
template <int C>
__device__ __forceinline__ void load_row(uint32_t tmem, int warp, float (&v)[C]) {
#pragma unroll
    for (int c0 = 0; c0 < C; c0 += 16) {
        float t[16];
        umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, c0), t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[c0 + j] = t[j];
    }
}
template <int C>
__device__ __forceinline__ void layernorm_reg(float (&v)[C], const float* __restrict__ g, const float* __restrict__ b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) s += v[i];
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) { v[i] -= mean; q += v[i] * v[i]; }
    const float rstd = rsqrtf(q * (1.f / C) + kLnEps);
#pragma unroll
    for (int i = 0; i < C; ++i) v[i] = v[i] * rstd * g[i] + b[i];
}

template <int C, int NH, bool kAttnOnly, class Q>
__global__ void __launch_bounds__(192) loftr_query_tc_kernel(Q q, cfp_loftr_w w, const float* __restrict__ kv,
                                                             const float* __restrict__ ksum, int ntiles) {
    using P = ChainTC<C>;
    constexpr int DH = C / NH, KG = P::KG, G = DH < 16 ? 16 : DH;
    constexpr int NCHUNK = kAttnOnly ? 1 : 8;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ ChainBars bars;
    uint8_t* a0 = smem;                  // [x | LN1(merge(msg))]  2C columns
    uint8_t* a1 = a0 + P::ABUF;          // msg (C columns), later the MLP hidden (2C columns)
    uint8_t* ring = a1 + P::ABUF;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < P::NSLOT; ++i) { umma::mbar_init(&bars.full[i], 1); umma::mbar_init(&bars.empty[i], 1); }
        umma::mbar_init(&bars.a_ready, 128);
        umma::mbar_init(&bars.acc_ready, 1);
        umma::fence_mbar_init();
    }
    if (warp == 4) umma::tmem_alloc(&bars.tmem_slot, P::TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot;

    if (warp < 4) {
        // =============================================================== row threads
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t row0 = (int64_t)tile * 128;
            // ---- stage x (coalesced: consecutive threads take consecutive 16-byte channel groups)
            for (int i = tid; i < 128 * KG; i += 128) {
                const int r = i / KG, kg = i % KG;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (row0 + r < q.rows) v = load8_bf16(q, row0 + r, kg * 8);
                *reinterpret_cast<uint4*>(a0 + (size_t)kg * P::LBO + r * 16) = v;
            }
            const int64_t row = row0 + tid;
            const int g = row < q.rows ? q.group(row) : -1;
            umma::fence_async_smem();
            mbar_arrive(&bars.a_ready);

            // ---- epilogue 1: Q = elu(q)+1, msg = (Q KV) / (Q.Ksum + eps)
            umma::mbar_wait(&bars.acc_ready, ph); ph ^= 1;
            umma::fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += G) {
                float qv[G], out[G];
#pragma unroll
                for (int j = 0; j < G; j += 16) {
                    float t[16];
                    umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, c0 + j), t);
#pragma unroll
                    for (int i = 0; i < 16; ++i) qv[j + i] = elu1(t[i]);
                }
#pragma unroll
                for (int hh = 0; hh < G / DH; ++hh) {
                    const int h0 = c0 + hh * DH;             // first channel of this head
                    float num[DH], den = kAttnEps;
#pragma unroll
                    for (int v = 0; v < DH; ++v) num[v] = 0.f;
                    if (g >= 0) {
                        const float* kvh = kv + (size_t)g * (C * DH) + (size_t)h0 * DH;
                        const float* ksh = ksum + (size_t)g * C + h0;
#pragma unroll
                        for (int d = 0; d < DH; ++d) {
                            const float qd = qv[hh * DH + d];
                            den = fmaf(qd, ksh[d], den);
#pragma unroll
                            for (int v = 0; v < DH; v += 4) {
                                const float4 k4 = *reinterpret_cast<const float4*>(kvh + d * DH + v);
                                num[v] = fmaf(qd, k4.x, num[v]); num[v + 1] = fmaf(qd, k4.y, num[v + 1]);
                                num[v + 2] = fmaf(qd, k4.z, num[v + 2]); num[v + 3] = fmaf(qd, k4.w, num[v + 3]);
                            }
                        }
                    }
                    const float inv = 1.f / den;
#pragma unroll
                    for (int v = 0; v < DH; ++v) out[hh * DH + v] = num[v] * inv;
                }
#pragma unroll
                for (int j = 0; j < G; j += 8) {
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o8[i] = out[j + i];
                    if (kAttnOnly) {
                        if (g >= 0) store8(q, row, c0 + j, o8);
                    } else {
                        umma::store_chunk(a1, P::LBO, tid, (c0 + j) / 8, o8);
                    }
                }
            }
            if (kAttnOnly) {
                umma::fence_before_sync();
                rows_sync();
                continue;
            }
            umma::fence_async_smem();
            umma::fence_before_sync();
            mbar_arrive(&bars.a_ready);

            // ---- epilogue 2: LN1(merge) -> second half of the cat tile
            umma::mbar_wait(&bars.acc_ready, ph); ph ^= 1;
            umma::fence_after_sync();
            {
                float v[C];
                load_row<C>(tmem, warp, v);
                layernorm_reg<C>(v, w.ln1_g, w.ln1_b);
#pragma unroll
                for (int j = 0; j < C; j += 8) {
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o8[i] = v[j + i];
                    umma::store_chunk(a0, P::LBO, tid, KG + j / 8, o8);
                }
            }
            umma::fence_async_smem();
            umma::fence_before_sync();
            mbar_arrive(&bars.a_ready);

            // ---- epilogue 3: relu(W1 [x|msg]) -> hidden tile (2C columns)
            umma::mbar_wait(&bars.acc_ready, ph); ph ^= 1;
            umma::fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < 2 * C; c0 += 16) {
                float t[16];
                umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, c0), t);
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o8[i] = fmaxf(t[j + i], 0.f);
                    umma::store_chunk(a1, P::LBO, tid, (c0 + j) / 8, o8);
                }
            }
            umma::fence_async_smem();
            umma::fence_before_sync();
            mbar_arrive(&bars.a_ready);

            // ---- epilogue 4: x + LN2(W2 hidden) -> scatter
            umma::mbar_wait(&bars.acc_ready, ph); ph ^= 1;
            umma::fence_after_sync();
            {
                float v[C];
                load_row<C>(tmem, warp, v);
                layernorm_reg<C>(v, w.ln2_g, w.ln2_b);
                if (g >= 0) {
#pragma unroll
                    for (int j = 0; j < C; j += 8) {
                        float x8[8], o8[8];
                        unpack8(*reinterpret_cast<const uint4*>(a0 + (size_t)(j / 8) * P::LBO + tid * 16), x8);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o8[i] = x8[i] + v[j + i];
                        store8(q, row, j, o8);
                    }
                }
            }
            umma::fence_before_sync();
            rows_sync();          // every row has read its x from a0 before the next tile is staged
        }
    } else if (warp == 4) {
        // =============================================================== weight producer
        if (lane == 0) {
            const bf16* wsrc = reinterpret_cast<const bf16*>(w.tc);
            int cc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int c = 0; c < NCHUNK; ++c, ++cc) {
                    const int slot = cc % P::NSLOT, round = cc / P::NSLOT;
                    if (round > 0) umma::mbar_wait(&bars.empty[slot], (round - 1) & 1);
                    umma::mbar_expect_tx(&bars.full[slot], P::SLOT);
                    umma::bulk_g2s(ring + (size_t)slot * P::SLOT, wsrc + (size_t)c * C * C, P::SLOT, &bars.full[slot]);
                }
        }
    } else {
        // =============================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma::idesc_bf16(128, C);
            const uint32_t a0s = umma::smem_u32(a0), a1s = umma::smem_u32(a1), rs = umma::smem_u32(ring);
            constexpr uint32_t LBO_B = C * 16;
            uint32_t ph = 0;
            int cc = 0;
            // one [C x C] block: D[:, dcol:dcol+C] (+)= A[:, kg0*8 : kg0*8+C] * Wblock^T
            auto block = [&](uint32_t abase, int kg0, int dcol, bool acc_first) {
                const int slot = cc % P::NSLOT, round = cc / P::NSLOT;
                umma::mbar_wait(&bars.full[slot], round & 1);
                umma::fence_after_sync();
                const uint32_t wb = rs + slot * P::SLOT;
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks)
                    umma::mma_bf16(tmem + dcol, umma::smem_desc(abase + (kg0 + 2 * ks) * P::LBO, P::LBO),
                                   umma::smem_desc(wb + ks * 2 * LBO_B, LBO_B), idesc, acc_first || ks > 0);
                umma::commit(&bars.empty[slot]);
                ++cc;
            };
            auto wait_a = [&]() { umma::mbar_wait(&bars.a_ready, ph); ph ^= 1; umma::fence_after_sync(); };
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                wait_a();
                block(a0s, 0, 0, false);                       // q
                umma::commit(&bars.acc_ready);
                if (kAttnOnly) continue;
                wait_a();
                block(a1s, 0, 0, false);                       // merge
                umma::commit(&bars.acc_ready);
                wait_a();
                block(a0s, 0, 0, false);                       // W1 quadrants: (n0,k0) (n0,k1) (n1,k0) (n1,k1)
                block(a0s, KG, 0, true);
                block(a0s, 0, C, false);
                block(a0s, KG, C, true);
                umma::commit(&bars.acc_ready);
                wait_a();
                block(a1s, 0, 0, false);                       // W2 K-halves
                block(a1s, KG, 0, true);
                umma::commit(&bars.acc_ready);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        umma::fence_after_sync();
        umma::tmem_dealloc(tmem, P::TMEM_COLS);
    }
}

template <int C, int NH, bool kAttnOnly, class Q>
static int run_query_tc(const char* name, const Q& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                        cudaStream_t st) {
    using P = ChainTC<C>;
    CFP_REQUIRE(w.tc != nullptr, "%s: bf16 path needs the packed tensor-core weights (cfp_loftr_w.tc)", name);
    auto k = loftr_query_tc_kernel<C, NH, kAttnOnly, Q>;
    if (int e = set_smem(k, P::SMEM)) return e;
    const int64_t ntiles = (q.rows + 127) / 128;
    const int per_sm = C >= 128 ? 1 : (C == 64 ? 2 : 4);
    const int grid = (int)(ntiles < 148 * per_sm ? ntiles : 148 * per_sm);
    k<<<grid, 192, P::SMEM, st>>>(q, w, kv, ksum, (int)ntiles);
    return check_launch(name);
}

// =====================================================================================
// LKPM pointwise half on tensor cores (convnext.py:49-58):
//   out = x + W2 GELU(W1 LN(y) + b1) + b2,   W1: C -> 4C, W2: 4C -> C
// The 4C hidden dimension is walked in four C-wide slices j: h_j = LN(y) W1_j^T (TMEM cols
// [0,C)), GELU in registers -> bf16 A tile, out += h_j W2_j^T (TMEM cols [C,2C)); the out
// accumulator stays in TMEM across the four slices.  MMA2_j and MMA1_{j+1} are issued back to
// back, so the tensor pipe works on slice j+1 while the row threads apply GELU to slice j.
// Weight blocks in consumption order: W1_0, W2_0, W1_1, W2_1, W1_2, W2_2, W1_3, W2_3.
template <int C>
__global__ void __launch_bounds__(192) lkpm_mlp_tc_kernel(bf16* __restrict__ feat0, const bf16* __restrict__ y,
                                                          int64_t rows, cfp_lkpm_w w, int ntiles) {
    using P = ChainTC<C>;
    constexpr int KG = P::KG;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ ChainBars bars;
    uint8_t* a0 = smem;                          // LN(y)        [KG][129][16 B]
    uint8_t* a1 = a0 + KG * P::LBO;              // GELU(h_j)    [KG][129][16 B]
    uint8_t* ring = a1 + KG * P::LBO;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < P::NSLOT; ++i) { umma::mbar_init(&bars.full[i], 1); umma::mbar_init(&bars.empty[i], 1); }
        umma::mbar_init(&bars.a_ready, 128);
        umma::mbar_init(&bars.acc_ready, 1);
        umma::fence_mbar_init();
    }
    if (warp == 4) umma::tmem_alloc(&bars.tmem_slot, P::TMEM_COLS);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot;

    if (warp < 4) {
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t row0 = (int64_t)tile * 128;
            for (int i = tid; i < 128 * KG; i += 128) {          // coalesced copy of the y tile
                const int r = i / KG, kg = i % KG;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (row0 + r < rows) v = *reinterpret_cast<const uint4*>(y + (row0 + r) * C + kg * 8);
                *reinterpret_cast<uint4*>(a0 + (size_t)kg * P::LBO + r * 16) = v;
            }
            rows_sync();
            {   // channels-last LayerNorm (eps 1e-6) of row `tid`, in place
                float v[C];
#pragma unroll
                for (int j = 0; j < C; j += 8) {
                    float t[8];
                    unpack8(*reinterpret_cast<const uint4*>(a0 + (size_t)(j / 8) * P::LBO + tid * 16), t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[j + i] = t[i];
                }
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < C; ++i) s += v[i];
                const float mean = s * (1.f / C);
                float q = 0.f;
#pragma unroll
                for (int i = 0; i < C; ++i) { v[i] -= mean; q += v[i] * v[i]; }
                const float rstd = rsqrtf(q * (1.f / C) + kLkpmLnEps);
#pragma unroll
                for (int j = 0; j < C; j += 8) {
                    float o8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o8[i] = v[j + i] * rstd * w.ln_g[j + i] + w.ln_b[j + i];
                    umma::store_chunk(a0, P::LBO, tid, j / 8, o8);
                }
            }
            umma::fence_async_smem();
            mbar_arrive(&bars.a_ready);
#pragma unroll 1
            for (int js = 0; js < 4; ++js) {                     // hidden slice js: GELU(h + b1) -> a1
                umma::mbar_wait(&bars.acc_ready, ph); ph ^= 1;
                umma::fence_after_sync();
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += 16) {
                    float t[16];
                    umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, c0), t);
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                        float o8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) o8[i] = gelu_erf(t[j + i] + w.pw1_b[js * C + c0 + j + i]);
                        umma::store_chunk(a1, P::LBO, tid, (c0 + j) / 8, o8);
                    }
                }
                umma::fence_async_smem();
                umma::fence_before_sync();
                mbar_arrive(&bars.a_ready);
            }
            umma::mbar_wait(&bars.acc_ready, ph); ph ^= 1;       // out accumulator complete
            umma::fence_after_sync();
            const int64_t row = row0 + tid;
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 16) {
                float t[16];
                umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, C + c0), t);
                if (row < rows) {
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                        bf16* p = feat0 + row * C + c0 + j;
                        float x8[8];
                        unpack8(*reinterpret_cast<const uint4*>(p), x8);
                        uint4 u;
                        u.x = umma::pack_bf16(x8[0] + t[j + 0] + w.pw2_b[c0 + j + 0], x8[1] + t[j + 1] + w.pw2_b[c0 + j + 1]);
                        u.y = umma::pack_bf16(x8[2] + t[j + 2] + w.pw2_b[c0 + j + 2], x8[3] + t[j + 3] + w.pw2_b[c0 + j + 3]);
                        u.z = umma::pack_bf16(x8[4] + t[j + 4] + w.pw2_b[c0 + j + 4], x8[5] + t[j + 5] + w.pw2_b[c0 + j + 5]);
                        u.w = umma::pack_bf16(x8[6] + t[j + 6] + w.pw2_b[c0 + j + 6], x8[7] + t[j + 7] + w.pw2_b[c0 + j + 7]);
                        *reinterpret_cast<uint4*>(p) = u;
                    }
                }
            }
            umma::fence_before_sync();
            rows_sync();
        }
    } else if (warp == 4) {
        if (lane == 0) {
            const bf16* wsrc = reinterpret_cast<const bf16*>(w.tc);
            int cc = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int c = 0; c < 8; ++c, ++cc) {
                    const int slot = cc % P::NSLOT, round = cc / P::NSLOT;
                    if (round > 0) umma::mbar_wait(&bars.empty[slot], (round - 1) & 1);
                    umma::mbar_expect_tx(&bars.full[slot], P::SLOT);
                    umma::bulk_g2s(ring + (size_t)slot * P::SLOT, wsrc + (size_t)c * C * C, P::SLOT, &bars.full[slot]);
                }
        }
    } else {
        if (lane == 0) {
            const uint32_t idesc = umma::idesc_bf16(128, C);
            const uint32_t a0s = umma::smem_u32(a0), a1s = umma::smem_u32(a1), rs = umma::smem_u32(ring);
            constexpr uint32_t LBO_B = C * 16;
            uint32_t ph = 0;
            int cc = 0;
            auto block = [&](uint32_t abase, int dcol, bool acc_first) {
                const int slot = cc % P::NSLOT, round = cc / P::NSLOT;
                umma::mbar_wait(&bars.full[slot], round & 1);
                umma::fence_after_sync();
                const uint32_t wb = rs + slot * P::SLOT;
#pragma unroll
                for (int ks = 0; ks < C / 16; ++ks)
                    umma::mma_bf16(tmem + dcol, umma::smem_desc(abase + 2 * ks * P::LBO, P::LBO),
                                   umma::smem_desc(wb + ks * 2 * LBO_B, LBO_B), idesc, acc_first || ks > 0);
                umma::commit(&bars.empty[slot]);
                ++cc;
            };
            auto wait_a = [&]() { umma::mbar_wait(&bars.a_ready, ph); ph ^= 1; umma::fence_after_sync(); };
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                wait_a();
                block(a0s, 0, false);                          // h_0
                umma::commit(&bars.acc_ready);
                for (int js = 0; js < 4; ++js) {
                    wait_a();                                  // GELU(h_js) staged in a1, h columns free
                    block(a1s, C, js > 0);                     // out += GELU(h_js) W2_js^T
                    if (js < 3) block(a0s, 0, false);          // h_{js+1}
                    umma::commit(&bars.acc_ready);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        umma::fence_after_sync();
        umma::tmem_dealloc(tmem, P::TMEM_COLS);
    }
}

template <int C>
static int run_lkpm_mlp_tc(void* feat0, const void* y, int64_t rows, const cfp_lkpm_w& w, cudaStream_t st) {
    using P = ChainTC<C>;
    CFP_REQUIRE(w.tc != nullptr, "lkpm_mlp: bf16 path needs the packed tensor-core weights (cfp_lkpm_w.tc)");
    constexpr size_t smem = 2 * (size_t)P::KG * P::LBO + (size_t)P::NSLOT * P::SLOT;
    auto k = lkpm_mlp_tc_kernel<C>;
    if (int e = set_smem(k, smem)) return e;
    const int64_t ntiles = (rows + 127) / 128;
    const int per_sm = C >= 128 ? 1 : (C == 64 ? 2 : 4);
    const int grid = (int)(ntiles < 148 * per_sm ? ntiles : 148 * per_sm);
    k<<<grid, 192, smem, st>>>((bf16*)feat0, (const bf16*)y, rows, w, (int)ntiles);
    return check_launch("lkpm_mlp_tc");
}

int lkpm_mlp_tc(void* feat0, const void* y, int64_t rows, int C, const cfp_lkpm_w& w, cudaStream_t st) {
    if (C == 32) return run_lkpm_mlp_tc<32>(feat0, y, rows, w, st);
    if (C == 64) return run_lkpm_mlp_tc<64>(feat0, y, rows, w, st);
    if (C == 128) return run_lkpm_mlp_tc<128>(feat0, y, rows, w, st);
    return fail("unsupported C=%d", C);
}

// ---- entry points used by k_loftr.cu's layer implementations (bf16 only)
#define CFP_TC_DISPATCH(NH, ATTN, NAME)                                                  \
    if (C == 32) return run_query_tc<32, NH, ATTN>(NAME, q, w, kv, ksum, st);            \
    if (C == 64) return run_query_tc<64, NH, ATTN>(NAME, q, w, kv, ksum, st);            \
    if (C == 128) return run_query_tc<128, NH, ATTN>(NAME, q, w, kv, ksum, st);          \
    return fail("unsupported C=%d", C);

int query_tc_h2i(int C, const ZonePatchRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st) {
    CFP_TC_DISPATCH(4, false, "loftr_query_tc<hist2image>")
}
int query_tc_lsa(int C, const WindowRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st) {
    CFP_TC_DISPATCH(8, false, "loftr_query_tc<lsa>")
}
int query_tc_gsa(int C, const FrameRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st) {
    CFP_TC_DISPATCH(8, false, "loftr_query_tc<gsa>")
}
int query_tc_dapm(int C, const OutsideRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                  cudaStream_t st) {
    CFP_TC_DISPATCH(4, true, "attn_query_tc<dapm>")
}

}  // namespace cfp
