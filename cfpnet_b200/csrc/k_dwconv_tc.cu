// LKPM large-kernel depthwise convolution on the tensor cores (bf16 path).
//
//   y[b][r][x][c] = ReLU( sum_{dy,dx} w[c][dy][dx] * in[b][r+dy-p][x+dx-p][c] + shift[c] ),  p = (k-1)/2
//                                                                   (convnext.py:45-47, BN folded)
//
// As a direct stencil the 31x31 case is bound by the fp32 FMA pipe (961 FMA per output).
// Here each channel plane is a GEMM instead: for one channel and one tile of 32 output columns
//
//   D[r][n] = sum_dy sum_kk  A_dy[r][kk] * T_dy[n][kk],     A_dy[r][kk] = plane[r + dy][x0 + kk]
//                                                            T_dy[n][kk] = w[dy][kk - n]  (0 outside the band)
//
// i.e. one M=128 (rows) x N=32 (columns) x K=16 tcgen05.mma per (dy, 16-column slice): the banded
// Toeplitz matrix T_dy carries the horizontal taps, the vertical tap dy is just a row shift of the A
// view (start address + 16 B * dy in the canonical K-major layout, umma.cuh).  About half of every
// T_dy is structural zeros; the tensor pipe is still ~10x faster than the FMA pipe on this op.
//
// Data flow: dw_plane_pack_kernel rewrites the token-major map into zero-padded channel planes stored
// directly in the UMMA layout ([b][c][x-group][padded row][8 columns]), so that the A operand of a
// (frame, channel, column tile) is ONE contiguous block brought in by a single bulk (TMA) copy.
// dwconv_tc_kernel: CTA = (column tile, channel); the channel's K Toeplitz blocks (host-packed) stay
// resident in shared memory while the CTA walks over the frames with a 3-stage A ring and two TMEM
// accumulators (MMA of frame i+1 overlaps the epilogue of frame i).  The epilogue writes planar
// rows; dw_plane_unpack_kernel transposes back to token-major for the pointwise MLP.
#include "cfp_common.cuh"
#include "cfp_internal.h"
#include "umma.cuh"

namespace cfp {

struct DwGeom {
    int H, W, C, K, PAD;
    int F;         // frames stacked vertically in one plane (PAD zero rows between them): fills the 128-row
                   // M tile when the map is short (H=52 at 1/8 scale -> 2 frames per MMA)
    int RS;        // row stride between stacked frames = H + PAD
    int NB;        // plane stacks = ceil(B / F)
    int nM;        // 128-row blocks per stack
    int HP;        // padded plane rows      = nM*128 + K - 1
    int KS;        // 16-column K-steps per dy = ceil((32 + K - 1) / 16)
    int nX;        // 32-column output tiles
    int WG;        // 8-column groups per plane = 4*nX + 2*KS - 4
};
static DwGeom dw_geom(int B, int H, int W, int C, int K) {
    DwGeom g;
    g.H = H; g.W = W; g.C = C; g.K = K; g.PAD = (K - 1) / 2;
    g.RS = H + g.PAD;
    g.F = (128 + g.PAD) / g.RS;
    if (g.F < 1) g.F = 1;
    if (g.F > B) g.F = B;
    g.NB = (B + g.F - 1) / g.F;
    const int rows = g.F * H + (g.F - 1) * g.PAD;      // output rows of one stack
    g.nM = (rows + 127) / 128;
    g.HP = g.nM * 128 + K - 1;
    g.KS = (32 + K - 1 + 15) / 16;
    g.nX = (W + 31) / 32;
    g.WG = 4 * g.nX + 2 * g.KS - 4;
    return g;
}
size_t dwconv_tc_plane_bytes(int B, int H, int W, int C, int K) {
    DwGeom g = dw_geom(B, H, W, C, K);
    const size_t in_bytes = (size_t)g.NB * C * g.WG * g.HP * 16;
    const size_t out_bytes = (size_t)B * C * H * W * 2;
    return ((in_bytes + 255) & ~(size_t)255) + ((out_bytes + 255) & ~(size_t)255);
}

// ---------------------------------------------------------------- token-major -> padded planes
// CTA = (frame, 8 consecutive image rows): coalesced read of the [8][W][C] slab, transpose through
// shared memory, then each thread emits 16-byte chunks (8 columns of one channel, one padded row);
// 8 consecutive rows of one (channel, x-group) are 128 contiguous bytes of the plane.
__global__ void __launch_bounds__(256) dw_plane_pack_kernel(const bf16* __restrict__ in, bf16* __restrict__ planes, DwGeom g) {
    extern __shared__ __align__(16) uint32_t slab[];          // [8 rows][W][C/2 + 1] channel pairs (odd stride: no conflicts)
    const int b = blockIdx.y, y0 = blockIdx.x * 8;
    const int C = g.C, W = g.W, C2 = C / 2, LD = C2 + 1;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + ((size_t)b * g.H + y0) * W * C2;
    const int rows = min(8, g.H - y0);
    for (int i = threadIdx.x; i < rows * W * C2; i += 256) {
        const int pix = i / C2, c2 = i - pix * C2;
        slab[pix * LD + c2] = src[i];
    }
    __syncthreads();
    // one thread = (channel pair, x-group, row): 8 four-byte reads -> two 16-byte chunks (channels 2*c2, 2*c2+1);
    // r fastest so that 8 consecutive threads write 128 contiguous bytes of a plane
    const int xg_lo = g.PAD / 8, xg_hi = (g.PAD + W - 1) / 8;          // x-groups that contain image columns
    const int nxg = xg_hi - xg_lo + 1;
    const int stack = b / g.F, yoff = (b % g.F) * g.RS + g.PAD;
    for (int i = threadIdx.x; i < C2 * nxg * 8; i += 256) {
        const int r = i & 7, xg = xg_lo + (i >> 3) % nxg, c2 = (i >> 3) / nxg;
        if (r >= rows) continue;
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = xg * 8 + j - g.PAD;
            v[j] = (x >= 0 && x < W) ? slab[(r * W + x) * LD + c2] : 0u;
        }
        uint4 lo, hi;
        lo.x = __byte_perm(v[0], v[1], 0x5410); hi.x = __byte_perm(v[0], v[1], 0x7632);
        lo.y = __byte_perm(v[2], v[3], 0x5410); hi.y = __byte_perm(v[2], v[3], 0x7632);
        lo.z = __byte_perm(v[4], v[5], 0x5410); hi.z = __byte_perm(v[4], v[5], 0x7632);
        lo.w = __byte_perm(v[6], v[7], 0x5410); hi.w = __byte_perm(v[6], v[7], 0x7632);
        const size_t off = ((((size_t)stack * C + 2 * c2) * g.WG + xg) * g.HP + (y0 + r + yoff)) * 8;
        *reinterpret_cast<uint4*>(planes + off) = lo;
        *reinterpret_cast<uint4*>(planes + off + (size_t)g.WG * g.HP * 8) = hi;
    }
}

// ---------------------------------------------------------------- planar rows -> token-major
// 64 tokens x 64 channels per CTA; 4-byte accesses on both sides (two tokens of a channel in, two channels of
// a token out).
__global__ void __launch_bounds__(256) dw_plane_unpack_kernel(const bf16* __restrict__ planar, bf16* __restrict__ out, int H,
                                                              int W, int C) {
    __shared__ uint16_t tile[64][66];                        // [channel][token]
    const int b = blockIdx.z, n0 = blockIdx.x * 64, c0 = blockIdx.y * 64, N = H * W;
    const uint16_t* src = reinterpret_cast<const uint16_t*>(planar);
    uint16_t* dst = reinterpret_cast<uint16_t*>(out);
    const bool even = (N & 1) == 0;
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {       // (channel, token pair)
        const int c = c0 + i / 32, n = n0 + (i % 32) * 2;
        if (c >= C || n >= N) continue;
        const size_t o = ((size_t)b * C + c) * N + n;
        if (even && n + 1 < N) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(src + o);
            tile[i / 32][(i % 32) * 2] = (uint16_t)(v & 0xffffu);
            tile[i / 32][(i % 32) * 2 + 1] = (uint16_t)(v >> 16);
        } else {
            tile[i / 32][(i % 32) * 2] = src[o];
            if (n + 1 < N) tile[i / 32][(i % 32) * 2 + 1] = src[o + 1];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {       // (token, channel pair)
        const int n = n0 + i / 32, c = c0 + (i % 32) * 2;
        if (n >= N || c >= C) continue;
        const uint32_t v = (uint32_t)tile[(i % 32) * 2][i / 32] | ((uint32_t)tile[(i % 32) * 2 + 1][i / 32] << 16);
        *reinterpret_cast<uint32_t*>(dst + ((size_t)b * N + n) * C + c) = v;
    }
}

// ---------------------------------------------------------------- the Toeplitz GEMM
struct DwBars {
    uint64_t t_full, t_empty, a_full[3], a_empty[3], acc_full[2], acc_empty[2];
    uint32_t tmem_slot;
};

// Work items are (channel, column tile, frame, row block), channel-major; every CTA takes an equal
// contiguous share, so the grid is one balanced wave and the Toeplitz blocks are reloaded only when
// a CTA's share crosses a channel boundary.
struct DwItem { int c, xt, b, mt; };                      // b = plane stack index
__device__ __forceinline__ DwItem dw_item(int i, const DwGeom& g) {
    DwItem it;
    it.mt = i % g.nM; i /= g.nM;
    it.b = i % g.NB; i /= g.NB;
    it.xt = i % g.nX;
    it.c = i / g.nX;
    return it;
}

__global__ void __launch_bounds__(192) dwconv_tc_kernel(const bf16* __restrict__ planes, const bf16* __restrict__ toep,
                                                        const float* __restrict__ shift, bf16* __restrict__ planar_out,
                                                        int B, DwGeom g, int items_per_cta, int total_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ DwBars bars;
    const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
    const uint32_t t_blk = 2 * g.KS * 32 * 16;                 // bytes of one T_dy block  [2*KS groups][32][8]
    const uint32_t t_bytes = g.K * t_blk;
    const uint32_t lbo_a = g.HP * 16;
    const uint32_t a_bytes = 2 * g.KS * lbo_a;
    uint8_t* t_sm = smem;
    uint8_t* a_sm = smem + ((t_bytes + 127) & ~127u);          // [3][a_bytes]
    const uint32_t a_stride = (a_bytes + 127) & ~127u;
    const int i0 = blockIdx.x * items_per_cta, i1 = min(i0 + items_per_cta, total_items);

    if (tid == 0) {
        umma::mbar_init(&bars.t_full, 1);
        umma::mbar_init(&bars.t_empty, 1);
        for (int i = 0; i < 3; ++i) { umma::mbar_init(&bars.a_full[i], 1); umma::mbar_init(&bars.a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars.acc_full[i], 1); umma::mbar_init(&bars.acc_empty[i], 128); }
        umma::fence_mbar_init();
    }
    if (warp == 4) umma::tmem_alloc(&bars.tmem_slot, 64);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot;

    if (warp < 4) {
        // ---------------- epilogue: thread = output row, 32 columns
        for (int i = i0; i < i1; ++i) {
            const int n = i - i0, ab = n & 1;
            const DwItem it = dw_item(i, g);
            const float sh = shift[it.c];
            umma::mbar_wait(&bars.acc_full[ab], (n >> 1) & 1);
            umma::fence_after_sync();
            float v0[16], v1[16], v[32];
            umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, ab * 32), v0);
            umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, ab * 32 + 16), v1);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = v0[j]; v[16 + j] = v1[j]; }
            umma::fence_before_sync();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(&bars.acc_empty[ab])) : "memory");
            const int m = it.mt * 128 + tid, x0 = it.xt * 32;
            const int f = m / g.RS, y = m - f * g.RS, frame = it.b * g.F + f;      // stacked frame and its row
            if (f < g.F && y < g.H && frame < B) {
                bf16* dst = planar_out + (((size_t)frame * g.C + it.c) * g.H + y) * g.W + x0;
                const bool aligned = (((size_t)(dst - planar_out)) & 7) == 0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (x0 + j + 8 <= g.W && aligned) {
                        uint4 u;
                        u.x = umma::pack_bf16(fmaxf(v[j + 0] + sh, 0.f), fmaxf(v[j + 1] + sh, 0.f));
                        u.y = umma::pack_bf16(fmaxf(v[j + 2] + sh, 0.f), fmaxf(v[j + 3] + sh, 0.f));
                        u.z = umma::pack_bf16(fmaxf(v[j + 4] + sh, 0.f), fmaxf(v[j + 5] + sh, 0.f));
                        u.w = umma::pack_bf16(fmaxf(v[j + 6] + sh, 0.f), fmaxf(v[j + 7] + sh, 0.f));
                        *reinterpret_cast<uint4*>(dst + j) = u;
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (x0 + j + q < g.W) dst[j + q] = __float2bfloat16_rn(fmaxf(v[j + q] + sh, 0.f));
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ---------------- producer: Toeplitz blocks per channel, one bulk copy per work item
        {
            int cur_c = -1, nt = 0;
            for (int i = i0; i < i1; ++i) {
                const int n = i - i0, s = n % 3;
                const DwItem it = dw_item(i, g);
                if (it.c != cur_c) {
                    if (nt > 0) umma::mbar_wait(&bars.t_empty, (nt - 1) & 1);   // MMAs on the old blocks are done
                    umma::bulk_load(t_sm, toep + (size_t)it.c * g.K * (t_blk / 2), t_bytes, &bars.t_full);
                    cur_c = it.c;
                    ++nt;
                }
                if (n >= 3) umma::mbar_wait(&bars.a_empty[s], ((n / 3) - 1) & 1);
                const bf16* src = planes + ((((size_t)it.b * g.C + it.c) * g.WG + 4 * it.xt) * g.HP) * 8;
                umma::bulk_load(a_sm + (size_t)s * a_stride, src, a_bytes, &bars.a_full[s]);
            }
        }
    } else {
        // ---------------- MMA issuer (warp-uniform control flow; one elected lane issues)
        {
            const uint32_t idesc = umma::idesc_bf16(128, 32);
            const uint64_t tdesc0 = umma::smem_desc(umma::smem_u32(t_sm), 512);
            const uint32_t as0 = umma::smem_u32(a_sm);
            int cur_c = -1, nt = 0;
            for (int i = i0; i < i1; ++i) {
                const int n = i - i0, s = n % 3, ab = n & 1;
                const DwItem it = dw_item(i, g);
                if (it.c != cur_c) {
                    if (nt > 0) umma::commit(&bars.t_empty);
                    umma::mbar_wait(&bars.t_full, nt & 1);
                    cur_c = it.c;
                    ++nt;
                }
                umma::mbar_wait(&bars.a_full[s], (n / 3) & 1);
                if (n >= 2) umma::mbar_wait(&bars.acc_empty[ab], ((n >> 1) - 1) & 1);
                umma::fence_after_sync();
                uint64_t ad = umma::smem_desc(as0 + s * a_stride + (uint32_t)(it.mt * 128) * 16, lbo_a);
                uint64_t td = tdesc0;
                const uint32_t dcol = tmem + ab * 32;
                for (int dy = 0; dy < g.K; ++dy) {
                    uint64_t adj = ad, tdj = td;
                    for (int j = 0; j < g.KS; ++j) {
                        umma::mma_bf16(dcol, adj, tdj, idesc, (dy | j) != 0);
                        adj = umma::desc_advance(adj, 2 * lbo_a);
                        tdj = umma::desc_advance(tdj, 2 * 512);
                    }
                    ad = umma::desc_advance(ad, 16);           // next vertical tap: one row down
                    td = umma::desc_advance(td, t_blk);
                }
                umma::commit(&bars.a_empty[s]);
                umma::commit(&bars.acc_full[ab]);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        umma::fence_after_sync();
        umma::tmem_dealloc(tmem, 64);
    }
}

int dwconv_tc(const void* in, void* out, int B, int H, int W, int C, int K, const void* toep, const float* shift,
              char* plane_ws, cudaStream_t st) {
    CFP_REQUIRE(toep != nullptr, "dwconv: bf16 path needs the packed Toeplitz blocks (cfp_lkpm_w.dw_toep)");
    CFP_REQUIRE(B <= 65535 && C <= 65535, "grid limits");
    DwGeom g = dw_geom(B, H, W, C, K);
    const size_t in_bytes = ((size_t)g.NB * C * g.WG * g.HP * 16 + 255) & ~(size_t)255;
    bf16* planes = reinterpret_cast<bf16*>(plane_ws);
    bf16* planar_out = reinterpret_cast<bf16*>(plane_ws + in_bytes);
    cudaError_t e = cudaMemsetAsync(planes, 0, in_bytes, st);           // zero padding of the planes
    if (e != cudaSuccess) return fail("cudaMemsetAsync(planes): %s", cudaGetErrorString(e));
    {
        const size_t smem = (size_t)8 * W * (C / 2 + 1) * 4;
        CFP_REQUIRE(smem <= 200 * 1024, "dw_plane_pack: %zu B shared memory", smem);
        if (int err = set_smem(dw_plane_pack_kernel, smem)) return err;
        dw_plane_pack_kernel<<<dim3((H + 7) / 8, B), 256, smem, st>>>((const bf16*)in, planes, g);
        if (int err = check_launch("dw_plane_pack")) return err;
    }
    {
        const uint32_t t_bytes = g.K * 2 * g.KS * 32 * 16, a_bytes = 2 * g.KS * g.HP * 16;
        const size_t smem = ((t_bytes + 127) & ~127u) + 3 * (size_t)((a_bytes + 127) & ~127u);
        CFP_REQUIRE(smem <= 225 * 1024, "dwconv (tensor-core path): %zu B shared memory (H=%d too tall)", smem, H);
        if (int err = set_smem(dwconv_tc_kernel, smem)) return err;
        const int total = C * g.nX * g.NB * g.nM;
        const int grid = total < 148 ? total : 148;
        const int per = (total + grid - 1) / grid;
        dwconv_tc_kernel<<<(total + per - 1) / per, 192, smem, st>>>(planes, (const bf16*)toep, shift, planar_out, B, g, per, total);
        if (int err = check_launch(K == 31 ? "dwconv_tc<31>" : K == 15 ? "dwconv_tc<15>" : "dwconv_tc<7>")) return err;
    }
    dw_plane_unpack_kernel<<<dim3((H * W + 63) / 64, (C + 63) / 64, B), 256, 0, st>>>(planar_out, (bf16*)out, H, W, C);
    return check_launch("dw_plane_unpack");
}

}  // namespace cfp
