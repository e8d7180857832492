// Layout kernels of the fusion module boundary (memory-bound):
//   a4  positional-encoding add + NCHW -> token-major   (fusion.py:92-97)
//       token-major -> NCHW                              (fusion.py:186)
//   a3  bit-exact export of zone_mask / hist_mask / pad_mask (fusion.py:103-120)
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

// 32x32 (pixel x channel) transpose tile through shared memory so that both the
// NCHW side (pixels contiguous) and the token side (channels contiguous) are
// accessed with full 128-byte lines.
template <typename T, bool kToTokens>
__global__ void __launch_bounds__(256) layout_kernel(const T* __restrict__ src, const float* __restrict__ pos,
                                                     T* __restrict__ dst, int C, int H, int W, int pos_w,
                                                     int oy, int ox, const int* __restrict__ crop) {
    __shared__ float tile[32][33];
    if (crop) { oy = crop[0]; ox = crop[1]; }
    const int N = H * W;
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    if (kToTokens) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int c = c0 + ty + 8 * i, n = n0 + tx;
            if (c < C && n < N) tile[ty + 8 * i][tx] = IO<T>::ld(src + ((size_t)b * C + c) * N + n);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n = n0 + ty + 8 * i, c = c0 + tx;
            if (c < C && n < N) {
                int y = n / W, x = n - y * W;
                float p = pos[((size_t)(oy + y) * pos_w + (ox + x)) * C + c];
                IO<T>::st(dst + ((size_t)b * N + n) * C + c, tile[tx][ty + 8 * i] + p);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n = n0 + ty + 8 * i, c = c0 + tx;
            if (c < C && n < N) tile[ty + 8 * i][tx] = IO<T>::ld(src + ((size_t)b * N + n) * C + c);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int c = c0 + ty + 8 * i, n = n0 + tx;
            if (c < C && n < N) IO<T>::st(dst + ((size_t)b * C + c) * N + n, tile[tx][ty + 8 * i]);
        }
    }
}

// bf16 fast path: 64 tokens x 64 channels per CTA, 4-byte accesses on both sides (two tokens of one channel on
// the NCHW side, two channels of one token on the token side).  Requires H*W and C even.
template <bool kToTokens>
__global__ void __launch_bounds__(256) layout_bf16_kernel(const bf16* __restrict__ src_, const float* __restrict__ pos,
                                                          bf16* __restrict__ dst_, int C, int H, int W, int pos_w, int oy, int ox,
                                                          const int* __restrict__ crop) {
    __shared__ uint16_t tile[64][66];                        // [channel][token]
    if (crop) { oy = crop[0]; ox = crop[1]; }
    const int b = blockIdx.z, n0 = blockIdx.x * 64, c0 = blockIdx.y * 64, N = H * W;
    const uint16_t* src = reinterpret_cast<const uint16_t*>(src_);
    uint16_t* dst = reinterpret_cast<uint16_t*>(dst_);
    if (kToTokens) {
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {   // read (channel, token pair) from NCHW
            const int cl = i / 32, nl = (i % 32) * 2, c = c0 + cl, n = n0 + nl;
            if (c < C && n < N) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(src + ((size_t)b * C + c) * N + n);
                tile[cl][nl] = (uint16_t)(v & 0xffffu);
                tile[cl][nl + 1] = (uint16_t)(v >> 16);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {   // write (token, channel pair) + positional encoding
            const int nl = i / 32, cl = (i % 32) * 2, n = n0 + nl, c = c0 + cl;
            if (n < N && c < C) {
                const int y = n / W, x = n - y * W;
                const float2 p = *reinterpret_cast<const float2*>(pos + ((size_t)(oy + y) * pos_w + (ox + x)) * C + c);
                const float a = __uint_as_float((uint32_t)tile[cl][nl] << 16) + p.x;
                const float d = __uint_as_float((uint32_t)tile[cl + 1][nl] << 16) + p.y;
                __nv_bfloat162 o = __floats2bfloat162_rn(a, d);
                *reinterpret_cast<uint32_t*>(dst + ((size_t)b * N + n) * C + c) = *reinterpret_cast<uint32_t*>(&o);
            }
        }
    } else {
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {   // read (token, channel pair) from tokens
            const int nl = i / 32, cl = (i % 32) * 2, n = n0 + nl, c = c0 + cl;
            if (n < N && c < C) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(src + ((size_t)b * N + n) * C + c);
                tile[cl][nl] = (uint16_t)(v & 0xffffu);
                tile[cl + 1][nl] = (uint16_t)(v >> 16);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 32; i += 256) {   // write (channel, token pair) to NCHW
            const int cl = i / 32, nl = (i % 32) * 2, c = c0 + cl, n = n0 + nl;
            if (c < C && n < N) {
                const uint32_t v = (uint32_t)tile[cl][nl] | ((uint32_t)tile[cl][nl + 1] << 16);
                *reinterpret_cast<uint32_t*>(dst + ((size_t)b * C + c) * N + n) = v;
            }
        }
    }
}

// bf16 main path: TC channels x TN tokens per CTA (TC * TN = 4096) with 16-byte global accesses on BOTH sides
// (8 tokens of one channel on the NCHW side, 8 channels of one token on the token side); the transpose goes
// through a [channel][token] shared tile of 65-word rows: the 2-byte column accesses are conflict-free, the
// 4-byte row accesses 2-way.  Requires H*W % 8 == 0 (16-byte aligned channel rows) and C % TC == 0.
template <bool kToTokens, int TC>
__global__ void __launch_bounds__(256) layout_bf16_v8_kernel(const bf16* __restrict__ src_, const float* __restrict__ pos,
                                                             bf16* __restrict__ dst_, int C, int H, int W, int pos_w, int oy, int ox,
                                                             const int* __restrict__ crop) {
    constexpr int TN = 4096 / TC, LD = TN + 2;               // halfwords per tile row (odd word count)
    constexpr int NCH = TN / 8, CG = TC / 8;                 // 16-byte chunks per channel row / per token
    __shared__ __align__(4) uint16_t tile[TC * LD];
    const int b = blockIdx.z, n0 = blockIdx.x * TN, c0 = blockIdx.y * TC, N = H * W;
    const uint16_t* src = reinterpret_cast<const uint16_t*>(src_);
    uint16_t* dst = reinterpret_cast<uint16_t*>(dst_);
    pdl_trigger();                         // short, bandwidth-bound kernel: the next one may queue up behind it at once
    pdl_wait();
    if (crop) { oy = crop[0]; ox = crop[1]; }  // crop offsets from device memory (CUDA-graph replays: nothing random is frozen)
    if (kToTokens) {
#pragma unroll
        for (int i = threadIdx.x; i < TC * NCH; i += 256) {  // NCHW side: (channel, 8 tokens)
            const int cl = i / NCH, nl = (i % NCH) * 8, n = n0 + nl;
            if (n < N) {
                const uint4 v = *reinterpret_cast<const uint4*>(src + ((size_t)b * C + c0 + cl) * N + n);
                uint32_t* t = reinterpret_cast<uint32_t*>(tile + cl * LD + nl);
                t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = threadIdx.x; i < TN * CG; i += 256) {   // token side: (token, 8 channels) + positional encoding
            const int nl = i / CG, cl = (i % CG) * 8, n = n0 + nl;
            if (n < N) {
                const int y = n / W, x = n - y * W;
                const float* pp = pos + ((size_t)(oy + y) * pos_w + (ox + x)) * C + c0 + cl;
                const float4 p0 = *reinterpret_cast<const float4*>(pp), p1 = *reinterpret_cast<const float4*>(pp + 4);
                const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float a = __uint_as_float((uint32_t)tile[(cl + j) * LD + nl] << 16) + pv[j];
                    const float d = __uint_as_float((uint32_t)tile[(cl + j + 1) * LD + nl] << 16) + pv[j + 1];
                    __nv_bfloat162 t = __floats2bfloat162_rn(a, d);
                    o[j / 2] = *reinterpret_cast<uint32_t*>(&t);
                }
                *reinterpret_cast<uint4*>(dst + ((size_t)b * N + n) * C + c0 + cl) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    } else {
#pragma unroll
        for (int i = threadIdx.x; i < TN * CG; i += 256) {   // token side: (token, 8 channels)
            const int nl = i / CG, cl = (i % CG) * 8, n = n0 + nl;
            if (n < N) {
                const uint4 v = *reinterpret_cast<const uint4*>(src + ((size_t)b * N + n) * C + c0 + cl);
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    tile[(cl + j) * LD + nl] = (uint16_t)(w4[j / 2] & 0xffffu);
                    tile[(cl + j + 1) * LD + nl] = (uint16_t)(w4[j / 2] >> 16);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = threadIdx.x; i < TC * NCH; i += 256) {  // NCHW side: (channel, 8 tokens)
            const int cl = i / NCH, nl = (i % NCH) * 8, n = n0 + nl;
            if (n < N) {
                const uint32_t* t = reinterpret_cast<const uint32_t*>(tile + cl * LD + nl);
                *reinterpret_cast<uint4*>(dst + ((size_t)b * C + c0 + cl) * N + n) = make_uint4(t[0], t[1], t[2], t[3]);
            }
        }
    }
}

template <bool kToTokens, int TC>
static void launch_v8(const void* src, const float* pos, void* dst, int B, int C, int H, int W, int pos_w, int oy, int ox,
                      const int* crop, cudaStream_t st) {
    dim3 grid((H * W + 4096 / TC - 1) / (4096 / TC), C / TC, B);
    launch_pdl(layout_bf16_v8_kernel<kToTokens, TC>, grid, 256, 0, st, (const bf16*)src, pos, (bf16*)dst, C, H, W, pos_w, oy, ox, crop);
}

template <typename T>
static int launch_layout(bool to_tokens, const void* src, const float* pos, void* dst, int B, int C, int H,
                         int W, int pos_w, int oy, int ox, const int* crop, cudaStream_t st) {
    if (sizeof(T) == 2 && (H * W) % 8 == 0 && C % 32 == 0 && (((uintptr_t)src | (uintptr_t)dst | (uintptr_t)pos) & 15) == 0) {
        if (C % 64 == 0) {
            if (to_tokens) launch_v8<true, 64>(src, pos, dst, B, C, H, W, pos_w, oy, ox, crop, st);
            else launch_v8<false, 64>(src, pos, dst, B, C, H, W, pos_w, oy, ox, crop, st);
        } else {
            if (to_tokens) launch_v8<true, 32>(src, pos, dst, B, C, H, W, pos_w, oy, ox, crop, st);
            else launch_v8<false, 32>(src, pos, dst, B, C, H, W, pos_w, oy, ox, crop, st);
        }
        return check_launch("layout_kernel");
    }
    if (sizeof(T) == 2 && (H * W) % 2 == 0 && C % 2 == 0) {
        dim3 grid64((H * W + 63) / 64, (C + 63) / 64, B);
        if (to_tokens)
            layout_bf16_kernel<true><<<grid64, 256, 0, st>>>((const bf16*)src, pos, (bf16*)dst, C, H, W, pos_w, oy, ox, crop);
        else
            layout_bf16_kernel<false><<<grid64, 256, 0, st>>>((const bf16*)src, pos, (bf16*)dst, C, H, W, pos_w, oy, ox, crop);
        return check_launch("layout_kernel");
    }
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, B);
    if (to_tokens)
        layout_kernel<T, true><<<grid, 256, 0, st>>>((const T*)src, pos, (T*)dst, C, H, W, pos_w, oy, ox, crop);
    else
        layout_kernel<T, false><<<grid, 256, 0, st>>>((const T*)src, pos, (T*)dst, C, H, W, pos_w, oy, ox, crop);
    return check_launch("layout_kernel");
}

int posenc_tokens(const void* x, const float* pos, void* tokens, int B, int C, int H, int W, int pos_w,
                  int oy, int ox, int dtype, cudaStream_t st, const int* crop) {
    return dtype == CFP_F32 ? launch_layout<float>(true, x, pos, tokens, B, C, H, W, pos_w, oy, ox, crop, st)
                            : launch_layout<bf16>(true, x, pos, tokens, B, C, H, W, pos_w, oy, ox, crop, st);
}
int tokens_to_nchw(const void* tokens, void* out, int B, int C, int H, int W, int dtype, cudaStream_t st) {
    return dtype == CFP_F32 ? launch_layout<float>(false, tokens, nullptr, out, B, C, H, W, 0, 0, 0, nullptr, st)
                            : launch_layout<bf16>(false, tokens, nullptr, out, B, C, H, W, 0, 0, 0, nullptr, st);
}

// ---------------------------------------------------------------- masks
__global__ void zone_masks_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ zone_mask,
                                  uint8_t* __restrict__ hist_mask, uint8_t* __restrict__ pad_mask, int B,
                                  int H, int W, cfp_geom g) {
    const int Z = g.zone_num * g.zone_num, P = g.p1 * g.p2;
    const long n_zone = (long)B * H * W, n_hist = (long)B * Z * P, n_pad = (long)B * g.tzh * g.tzw;
    const long total = n_zone + n_hist + n_pad;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        if (i < n_zone) {
            int n = (int)(i % (H * W)), y = n / W, x = n - y * W;
            zone_mask[i] = (y >= g.ry0 && y < g.ry1 && x >= g.rx0 && x < g.rx1) ? 1 : 0;
        } else if (i < n_zone + n_hist) {
            long j = i - n_zone;
            hist_mask[j] = mask[j / P] ? 1 : 0;
        } else {
            long j = i - n_zone - n_hist;
            int cx = (int)(j % g.tzw), cy = (int)((j / g.tzw) % g.tzh);
            uint8_t v = 1;
            if (g.pad_h > 0 || g.pad_w > 0) {     // fusion.py:112-118
                int top = max(-g.sy_wo, 0), left = max(-g.sx_wo, 0);
                int bot = max(g.ey_wo - H, 0), right = max(g.ex_wo - W, 0);
                if (cy < top || cy >= g.tzh - bot || cx < left || cx >= g.tzw - right) v = 0;
            }
            pad_mask[j] = v;
        }
    }
}

int zone_masks(const uint8_t* mask, uint8_t* zm, uint8_t* hm, uint8_t* pm, int B, int H, int W,
               const cfp_geom& g, cudaStream_t st) {
    zone_masks_kernel<<<sm_count() * 4, 256, 0, st>>>(mask, zm, hm, pm, B, H, W, g);
    return check_launch("zone_masks_kernel");
}

}  // namespace cfp
