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
// The banded Toeplitz matrix T_dy carries the horizontal taps; the vertical tap is a row shift.  A row-shifted
// A view (start address + 16 B * dy in the canonical K-major layout, umma.cuh) would work, but with N = 32 every
// MMA is bound by the shared-memory read of its 4 KB A tile (twice that when the shift is not a multiple of
// 8 rows).  So the vertical taps are split as dy = 8 a + b: the eight b's of one a use the SAME 8-row-aligned A
// view and their Toeplitz blocks sit side by side along N:
//
//   E[m][(b, n)] = sum_a sum_kk plane[m + 8 a][x0 + kk] * T_{8a+b}[n][kk]        one M=128 x N=256 x K=16 MMA per (a, kk/16)
//   out[r][n]    = sum_b E[r + b][(b, n)]                                        row shift b applied by the epilogue
//
// One A read now feeds 256 accumulator columns (the MMA runs at the tensor pipe's rate) and a 31 x 31 tile costs
// 16 MMAs instead of 124.  The epilogue thread of accumulator row m gets E[m + b] from lane m + b with warp
// shuffles; the rows that live in the next warp's TMEM quarter travel through a small shared-memory exchange.
// About half of every T_dy is structural zeros; the tensor pipe is still >10x faster than the FMA pipe on this op.
//
// Data flow: dw_plane_pack_kernel rewrites the token-major map into zero-padded channel planes stored
// directly in the UMMA layout ([b][c][x-group][padded row][8 columns]), so that the A operand of a
// (frame, channel, column tile) is ONE contiguous block brought in by a single bulk (TMA) copy.
// dwconv_tc_kernel: CTA = (column tile, channel); the channel's K Toeplitz blocks (host-packed) stay
// resident in shared memory while the CTA walks over the frames with a 3-stage A ring and two TMEM
// accumulators (MMA of frame i+1 overlaps the epilogue of frame i).  The epilogue writes planar
// rows ([B][C][H][W]); the pointwise MLP (lkpm_mlp_tc_kernel) reads that map in place - lanes are consecutive
// tokens, so a channel's 32 values are one coalesced 64-byte segment - and no transpose back is needed.
#include "cfp_common.cuh"
#include "cfp_internal.h"
#include "umma.cuh"

namespace cfp {

constexpr int kDwMS = 120;     // output rows per 128-row M block: the epilogue reads accumulator rows m .. m+7

struct DwGeom {
    int H, W, C, K, PAD;
    int F;         // frames stacked vertically in one plane (PAD zero rows between them): fills the M block
                   // when the map is short (H=52 at 1/8 scale -> 2 frames per MMA)
    int RS;        // row stride between stacked frames = H + PAD
    int NB;        // plane stacks = ceil(B / F)
    int nM;        // M blocks (kDwMS output rows each) per stack
    int HP;        // padded plane rows
    int KS;        // 16-column K-steps per tap group = ceil((32 + K - 1) / 16)
    int NA;        // groups of eight vertical taps = ceil(K / 8)
    int nX;        // 32-column output tiles
    int WG;        // 8-column groups per plane = 4*nX + 2*KS - 4
};
static DwGeom dw_geom(int B, int H, int W, int C, int K) {
    DwGeom g;
    g.H = H; g.W = W; g.C = C; g.K = K; g.PAD = (K - 1) / 2;
    g.RS = H + g.PAD;
    g.F = (kDwMS + g.PAD) / g.RS;                      // F*H + (F-1)*PAD <= kDwMS
    if (g.F < 1) g.F = 1;
    if (g.F > B) g.F = B;
    g.NB = (B + g.F - 1) / g.F;
    const int rows = g.F * H + (g.F - 1) * g.PAD;      // output rows of one stack
    g.nM = (rows + kDwMS - 1) / kDwMS;
    g.KS = (32 + K - 1 + 15) / 16;
    g.NA = (K + 7) / 8;
    // rows an A view can touch: last block start + 128 accumulator rows + 8 rows per further tap group;
    // and every row that feeds a stored output (PAD + F x (H + PAD)) must exist for the pack kernel
    g.HP = (g.nM - 1) * kDwMS + 128 + 8 * (g.NA - 1);
    if (g.HP < g.F * g.RS + g.PAD) g.HP = g.F * g.RS + g.PAD;
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
// CTA = (plane stack, 8 consecutive PLANE rows): coalesced read of the image rows among them ([W][C] slabs),
// transpose through shared memory, then each thread emits 16-byte chunks (8 columns of one channel, one
// plane row); 8 consecutive rows of one (channel, x-group) are 128 contiguous bytes of the plane.  The kernel
// writes EVERY chunk of the planes - image cells and zeros everywhere else (PAD rows above / between / below the
// stacked frames, PAD columns left and right, the slack rows and column groups the 128-row / 64-column operand
// views reach into) - so the planes need no memset and stale workspace bytes never enter an MMA.
__global__ void __launch_bounds__(256) dw_plane_pack_kernel(const bf16* __restrict__ in, bf16* __restrict__ planes, DwGeom g, int B) {
    extern __shared__ __align__(16) uint32_t slab[];          // [8 rows][W][C/2 + 1] channel pairs (odd stride: no conflicts)
    const int stack = blockIdx.y, pr0 = blockIdx.x * 8;
    const int C = g.C, W = g.W, C2 = C / 2, LD = C2 + 1;
    // every plane row an A view can touch (HP >= PAD + F x (H + PAD)): the rows below the last frame's halo are only
    // multiplied by all-zero Toeplitz blocks (taps dy >= K of the last group of eight), but a stale NaN would survive that
    const int rows = min(8, g.HP - pr0);
    // plane row pr -> (stacked frame f, image row y) or halo
    int fy[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int q = pr0 + r - g.PAD;                         // row relative to the first frame's first row
        const int f = q >= 0 ? q / g.RS : -1, y = q - f * g.RS;
        const int frame = stack * g.F + f;
        fy[r] = (r < rows && q >= 0 && f < g.F && y < g.H && frame < B) ? frame * g.H + y : -1;
    }
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in);
    for (int i = threadIdx.x; i < 8 * W * C2; i += 256) {
        const int r = i / (W * C2), j = i - r * (W * C2);
        if (fy[r] >= 0) slab[(r * W) * LD + (j / C2) * LD + (j % C2)] = src[(size_t)fy[r] * W * C2 + j];
    }
    __syncthreads();
    // one thread = (channel pair, x-group, row): 8 four-byte reads -> two 16-byte chunks (channels 2*c2, 2*c2+1);
    // r fastest so that 8 consecutive threads write 128 contiguous bytes of a plane
    const int nxg = g.WG;      // every x-group an A operand covers (a stale NaN times a structural zero of T would poison the row)
    for (int i = threadIdx.x; i < C2 * nxg * 8; i += 256) {
        const int r = i & 7, xg = (i >> 3) % nxg, c2 = (i >> 3) / nxg;
        if (r >= rows) continue;
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = xg * 8 + j - g.PAD;
            v[j] = (fy[r] >= 0 && x >= 0 && x < W) ? slab[(r * W + x) * LD + c2] : 0u;
        }
        uint4 lo, hi;
        lo.x = __byte_perm(v[0], v[1], 0x5410); hi.x = __byte_perm(v[0], v[1], 0x7632);
        lo.y = __byte_perm(v[2], v[3], 0x5410); hi.y = __byte_perm(v[2], v[3], 0x7632);
        lo.z = __byte_perm(v[4], v[5], 0x5410); hi.z = __byte_perm(v[4], v[5], 0x7632);
        lo.w = __byte_perm(v[6], v[7], 0x5410); hi.w = __byte_perm(v[6], v[7], 0x7632);
        const size_t off = ((((size_t)stack * C + 2 * c2) * g.WG + xg) * g.HP + (pr0 + r)) * 8;
        *reinterpret_cast<uint4*>(planes + off) = lo;
        *reinterpret_cast<uint4*>(planes + off + (size_t)g.WG * g.HP * 8) = hi;
    }
}

// ---------------------------------------------------------------- the Toeplitz GEMM
struct DwBars {
    uint64_t t_full, t_empty, a_full[3], a_empty[3], acc_full[2], acc_empty[2];
    uint32_t tmem_slot;
};

// Work items are (channel, column tile, frame stack, row block), channel-major; every CTA takes an equal
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

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int kXchLd = 36;                                 // floats per exchanged row (16-byte aligned, bank-staggered)
constexpr int kXchRows = 28;                               // rows a warp publishes per item: sum_{b=1..7} b
constexpr size_t kXchBytes = 2 * 3 * kXchRows * kXchLd * sizeof(float);

__global__ void __launch_bounds__(192) dwconv_tc_kernel(const bf16* __restrict__ planes, const bf16* __restrict__ toep,
                                                        const float* __restrict__ shift, bf16* __restrict__ planar_out,
                                                        int B, DwGeom g, int items_per_cta, int total_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ DwBars bars;
    const int tid = threadIdx.x, warp = umma::warp_idx_sync(), lane = tid & 31;
    constexpr uint32_t t_blk = 2 * 256 * 16;                   // bytes of one (a, ks) block  [2 k-groups][256 = (b, n)][8]
    const uint32_t t_bytes = g.NA * g.KS * t_blk;
    const uint32_t lbo_a = g.HP * 16;
    const uint32_t a_bytes = 2 * g.KS * lbo_a;
    uint8_t* t_sm = smem;
    uint8_t* a_sm = smem + t_bytes;                            // [3][a_stride]
    const uint32_t a_stride = (a_bytes + 127) & ~127u;
    float* xch = reinterpret_cast<float*>(a_sm + 3 * (size_t)a_stride);   // [2][3 warps][28 rows][kXchLd]
    const int i0 = blockIdx.x * items_per_cta, i1 = min(i0 + items_per_cta, total_items);

    if (tid == 0) {
        umma::mbar_init(&bars.t_full, 1);
        umma::mbar_init(&bars.t_empty, 1);
        for (int i = 0; i < 3; ++i) { umma::mbar_init(&bars.a_full[i], 1); umma::mbar_init(&bars.a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars.acc_full[i], 1); umma::mbar_init(&bars.acc_empty[i], 128); }
        umma::fence_mbar_init();
    }
    if (warp == 4) umma::tmem_alloc(&bars.tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot;

    if (warp < 4) {
        // ---------------- epilogue: thread = accumulator row m; out[m][n] = sum_b E_b[m + b][n]
        for (int i = i0; i < i1; ++i) {
            const int n = i - i0, ab = n & 1;
            const DwItem it = dw_item(i, g);
            const float sh = shift[it.c];
            float* xw = xch + (size_t)ab * (3 * kXchRows * kXchLd);
            umma::mbar_wait(&bars.acc_full[ab], (n >> 1) & 1);
            umma::fence_after_sync();
            float acc[32];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                float e[32];
                tmem_ld32(umma::tmem_addr(tmem, warp * 32, ab * 256 + b * 32), e);
                if (b == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = e[j];
                } else {
                    if (warp > 0 && lane < b) {                // rows the warp below needs: E_b[32 w + lane]
                        float* dst = xw + ((size_t)(warp - 1) * kXchRows + b * (b - 1) / 2 + lane) * kXchLd;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(e[j], e[j + 1], e[j + 2], e[j + 3]);
                    }
                    const bool in_warp = lane + b < 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float sv = __shfl_down_sync(0xffffffffu, e[j], b);
                        acc[j] += in_warp ? sv : 0.f;
                    }
                }
            }
            umma::fence_before_sync();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(&bars.acc_empty[ab])) : "memory");
            asm volatile("bar.sync 1, 128;\n" ::: "memory");  // boundary rows of this item are published
            if (warp < 3 && lane >= 25) {
                for (int b = 32 - lane; b < 8; ++b) {
                    const float* src = xw + ((size_t)warp * kXchRows + b * (b - 1) / 2 + (lane + b - 32)) * kXchLd;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(src + j);
                        acc[j] += q.x; acc[j + 1] += q.y; acc[j + 2] += q.z; acc[j + 3] += q.w;
                    }
                }
            }
            const int m = it.mt * kDwMS + tid, x0 = it.xt * 32;
            const int f = m / g.RS, y = m - f * g.RS, frame = it.b * g.F + f;      // stacked frame and its row
            if (tid < kDwMS && f < g.F && y < g.H && frame < B) {
                bf16* dst = planar_out + (((size_t)frame * g.C + it.c) * g.H + y) * g.W + x0;
                const bool aligned = (((size_t)(dst - planar_out)) & 7) == 0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    if (x0 + j + 8 <= g.W && aligned) {
                        uint4 u;
                        u.x = umma::pack_bf16(fmaxf(acc[j + 0] + sh, 0.f), fmaxf(acc[j + 1] + sh, 0.f));
                        u.y = umma::pack_bf16(fmaxf(acc[j + 2] + sh, 0.f), fmaxf(acc[j + 3] + sh, 0.f));
                        u.z = umma::pack_bf16(fmaxf(acc[j + 4] + sh, 0.f), fmaxf(acc[j + 5] + sh, 0.f));
                        u.w = umma::pack_bf16(fmaxf(acc[j + 6] + sh, 0.f), fmaxf(acc[j + 7] + sh, 0.f));
                        *reinterpret_cast<uint4*>(dst + j) = u;
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (x0 + j + q < g.W) dst[j + q] = __float2bfloat16_rn(fmaxf(acc[j + q] + sh, 0.f));
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
                    if (umma::elect_one()) {
                        umma::mbar_expect_tx(&bars.t_full, t_bytes);
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(toep) + (size_t)it.c * t_bytes;
                        for (uint32_t off = 0; off < t_bytes; off += 4 * t_blk)
                            umma::bulk_g2s(t_sm + off, src + off, min(4 * t_blk, t_bytes - off), &bars.t_full);
                    }
                    __syncwarp();
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
            const uint32_t idesc = umma::idesc_bf16(128, 256);
            const uint64_t tdesc0 = umma::smem_desc(umma::smem_u32(t_sm), 256 * 16);
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
                // vertical taps dy = 8 a + b: the eight b's of one a share ONE (8-row aligned) A view and sit side by
                // side along N (256 = 8 x 32 columns); their row shift b is applied by the epilogue
                uint64_t ad = umma::smem_desc(as0 + s * a_stride + (uint32_t)(it.mt * kDwMS) * 16, lbo_a);
                uint64_t td = tdesc0;
                const uint32_t dcol = tmem + ab * 256;
                for (int a = 0; a < g.NA; ++a) {
                    uint64_t adj = ad;
                    for (int j = 0; j < g.KS; ++j) {
                        umma::mma_bf16(dcol, adj, td, idesc, (a | j) != 0);
                        adj = umma::desc_advance(adj, 2 * lbo_a);
                        td = umma::desc_advance(td, t_blk);
                    }
                    ad = umma::desc_advance(ad, 8 * 16);       // next group of eight vertical taps: eight rows down
                }
                umma::commit(&bars.a_empty[s]);
                umma::commit(&bars.acc_full[ab]);
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        umma::fence_after_sync();
        umma::tmem_dealloc(tmem, 512);
    }
}

int dwconv_tc(const void* in, const void** planar_out_p, int B, int H, int W, int C, int K, const void* toep, const float* shift,
              char* plane_ws, cudaStream_t st) {
    CFP_REQUIRE(toep != nullptr, "dwconv: bf16 path needs the packed Toeplitz blocks (cfp_lkpm_w.dw_toep)");
    CFP_REQUIRE(B <= 65535 && C <= 65535, "grid limits");
    DwGeom g = dw_geom(B, H, W, C, K);
    const size_t in_bytes = ((size_t)g.NB * C * g.WG * g.HP * 16 + 255) & ~(size_t)255;
    bf16* planes = reinterpret_cast<bf16*>(plane_ws);
    bf16* planar_out = reinterpret_cast<bf16*>(plane_ws + in_bytes);
    *planar_out_p = planar_out;          // [B][C][H][W]; lkpm_mlp_tc reads it in place (no transpose back)
    {
        const size_t smem = (size_t)8 * W * (C / 2 + 1) * 4;
        CFP_REQUIRE(smem <= 200 * 1024, "dw_plane_pack: %zu B shared memory", smem);
        if (int err = set_smem(dw_plane_pack_kernel, smem)) return err;
        dw_plane_pack_kernel<<<dim3((g.HP + 7) / 8, g.NB), 256, smem, st>>>((const bf16*)in, planes, g, B);
        if (int err = check_launch("dw_plane_pack")) return err;
    }
    {
        const uint32_t t_bytes = g.NA * g.KS * 2 * 256 * 16, a_bytes = 2 * g.KS * g.HP * 16;
        const size_t smem = t_bytes + 3 * (size_t)((a_bytes + 127) & ~127u) + kXchBytes;
        CFP_REQUIRE(smem <= 225 * 1024, "dwconv (tensor-core path): %zu B shared memory (H=%d too tall)", smem, H);
        if (int err = set_smem(dwconv_tc_kernel, smem)) return err;
        const int total = C * g.nX * g.NB * g.nM;
        const int grid = total < 148 ? total : 148;
        const int per = (total + grid - 1) / grid;
        dwconv_tc_kernel<<<(total + per - 1) / per, 192, smem, st>>>(planes, (const bf16*)toep, shift, planar_out, B, g, per, total);
        if (int err = check_launch(K == 31 ? "dwconv_tc<31>" : K == 15 ? "dwconv_tc<15>" : "dwconv_tc<7>")) return err;
    }
    return 0;
}

}  // namespace cfp
