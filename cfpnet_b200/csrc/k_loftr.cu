// Linear-attention layers of the fusion path:
//   a5  hist2image   (fusion.py:132-157, transformer.py:41-71, attention.py:20-52)
//   a6  DAPM attention half (transformer.py:215-234)
//   a9  Twins LSA / GSA (transformer.py:89-116, 138-150)
//
// All four are the same two-phase computation and share two kernels:
//   phase 1  kv_state_kernel : source rows -> k|v projection -> K = elu(k)+1 ->
//            per-group state  KV[h] = sum_s K_s^T V_s  (dh x dh per head) and Ksum[h] = sum_s K_s
//   phase 2  loftr_query_kernel : query rows -> q projection -> Q = elu(q)+1 ->
//            msg = (Q KV) / (Q.Ksum + eps) -> merge -> LN -> MLP([x,msg]) -> LN -> x + msg
// (the reference divides V by S before the sum and multiplies the message by S afterwards,
// attention.py:42,49 — an fp16-overflow guard that cancels exactly in real arithmetic; the
// state here is accumulated in fp32 and skips it).
//
// What differs per layer is only WHICH rows form a group and where they live; that is a
// "row provider": it maps a dense row index to a gather address (zone patch cell, window
// cell, outside-zone cell, ...) and scatters the result back.  No mask tensor and no
// gathered copy of the tokens is ever materialised.
//
// A CTA keeps its BM-row tile in shared memory through the whole chain (RowsGemm), so a
// layer reads its tokens once and writes them once.
#include "cfp_common.cuh"
#include "cfp_internal.h"
#include "providers.cuh"
#include <type_traits>

namespace cfp {

// ------------------------------------------------------------------ phase 1
template <int C, int NH, class Src>
__global__ void __launch_bounds__(kThreads) kv_state_kernel(Src src, const float* __restrict__ wkv_t,
                                                            float* __restrict__ kv, float* __restrict__ ksum) {
    constexpr int BM = Tile<C>::BM, DH = C / NH, LDX = C + 4, LDK = 2 * C + 4;
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                       // [BM][LDX]
    float* kvs = xs + BM * LDX;             // [BM][LDK]  K | V
    float* wbuf = kvs + BM * LDK;
    const int64_t row0 = (int64_t)blockIdx.x * BM;

    __shared__ int gid[BM];
    for (int i = threadIdx.x; i < BM * (C / 4); i += kThreads) {
        int r = i / (C / 4), c4 = i % (C / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < src.rows) {
            const auto ref = src.locate(row0 + r);
            v = src.load4(ref, c4 * 4);
            if (c4 == 0) gid[r] = src.group(ref);
        }
        *reinterpret_cast<float4*>(xs + r * LDX + c4 * 4) = v;
    }
    __syncthreads();
    gemm_to_smem<BM, 2 * C>(SmemRows{xs, LDX}, wkv_t, C, wbuf, kvs, LDK,
                            [](int c, float v) { return c < C ? elu1(v) : v; });
    __syncthreads();

    const int nrows = (int)min((int64_t)BM, src.rows - row0);
    int r = 0;
    while (r < nrows) {                     // one segment per group present in the tile
        const int g = gid[r];
        int e = r + 1;
        while (e < nrows && gid[e] == g) ++e;
        for (int idx = threadIdx.x; idx < C * DH; idx += kThreads) {
            const int kc = idx / DH, vc = (kc / DH) * DH + idx % DH;
            float s = 0.f;
            for (int t = r; t < e; ++t) s = fmaf(kvs[t * LDK + kc], kvs[t * LDK + C + vc], s);
            atomicAdd(kv + (size_t)g * (C * DH) + idx, s);
        }
        for (int c = threadIdx.x; c < C; c += kThreads) {
            float s = 0.f;
            for (int t = r; t < e; ++t) s += kvs[t * LDK + c];
            atomicAdd(ksum + (size_t)g * C + c, s);
        }
        r = e;
    }
}

// ------------------------------------------------------------------ phase 2
template <int C, int NH, bool kAttnOnly, class Q>
__global__ void __launch_bounds__(kThreads) loftr_query_kernel(Q q, cfp_loftr_w w, const float* __restrict__ kv,
                                                               const float* __restrict__ ksum) {
    constexpr int BM = Tile<C>::BM, DH = C / NH, LD = 2 * C + 4;
    extern __shared__ __align__(16) float smem[];
    float* cat = smem;                      // [BM][LD]   x | msg
    float* hb = cat + BM * LD;              // [BM][LD]   q | raw msg, later MLP hidden
    float* wbuf = hb + BM * LD;
    __shared__ int gid[BM];
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < BM * (C / 4); i += kThreads) {
        int r = i / (C / 4), c4 = i % (C / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < q.rows) {
            const auto ref = q.locate(row0 + r);
            v = q.load4(ref, c4 * 4);
            if (c4 == 0) gid[r] = q.group(ref);
        } else if (c4 == 0) gid[r] = -1;
        *reinterpret_cast<float4*>(cat + r * LD + c4 * 4) = v;
    }
    __syncthreads();

    // Q = elu(x Wq^T) + 1
    gemm_to_smem<BM, C>(SmemRows{cat, LD}, w.wq_t, C, wbuf, hb, LD, [](int, float v) { return elu1(v); });
    __syncthreads();

    // msg[h] = (Q_h KV_h) / (Q_h . Ksum_h + eps)           attention.py:44-49
    for (int r = warp; r < BM; r += kThreads / 32) {
        const int g = gid[r];
        const float* qr = hb + r * LD;
#pragma unroll
        for (int i = 0; i < C / 32; ++i) {
            const int c = lane + 32 * i, h0 = (c / DH) * DH, v = c - h0;
            float m = 0.f;
            if (g >= 0) {
                const float* kvg = kv + (size_t)g * (C * DH) + (size_t)h0 * DH + v;
                const float* ksg = ksum + (size_t)g * C + h0;
                float num = 0.f, den = 0.f;
#pragma unroll 8
                for (int d = 0; d < DH; ++d) {
                    const float qv = qr[h0 + d];
                    num = fmaf(qv, kvg[d * DH], num);
                    den = fmaf(qv, ksg[d], den);
                }
                m = num / (den + kAttnEps);
            }
            hb[r * LD + C + c] = m;
        }
    }
    __syncthreads();

    if (kAttnOnly) {                        // DAPM: the message map is the output
        for (int i = threadIdx.x; i < BM * (C / 4); i += kThreads) {
            int r = i / (C / 4), c4 = i % (C / 4);
            if (row0 + r < q.rows)
                q.store4(q.locate(row0 + r), c4 * 4, *reinterpret_cast<const float4*>(hb + r * LD + C + c4 * 4));
        }
        return;
    }

    // message = LN1(merge(msg))                            transformer.py:64-65
    gemm_to_smem<BM, C>(SmemRows{hb + C, LD}, w.wm_t, C, wbuf, cat + C, LD, [](int, float v) { return v; });
    __syncthreads();
    layernorm_rows<BM, C>(cat + C, LD, w.ln1_g, w.ln1_b, kLnEps);
    __syncthreads();
    // message = LN2(W2 relu(W1 [x, message]))              transformer.py:68-69
    gemm_to_smem<BM, 2 * C>(SmemRows{cat, LD}, w.w1_t, 2 * C, wbuf, hb, LD,
                            [](int, float v) { return fmaxf(v, 0.f); });
    __syncthreads();
    gemm_to_smem<BM, C>(SmemRows{hb, LD}, w.w2_t, 2 * C, wbuf, cat + C, LD, [](int, float v) { return v; });
    __syncthreads();
    layernorm_rows<BM, C>(cat + C, LD, w.ln2_g, w.ln2_b, kLnEps);
    __syncthreads();
    // x + message                                          transformer.py:71
    for (int i = threadIdx.x; i < BM * (C / 4); i += kThreads) {
        int r = i / (C / 4), c4 = i % (C / 4);
        if (row0 + r < q.rows) {
            float4 x = *reinterpret_cast<const float4*>(cat + r * LD + c4 * 4);
            float4 m = *reinterpret_cast<const float4*>(cat + r * LD + C + c4 * 4);
            q.store4(q.locate(row0 + r), c4 * 4, make_float4(x.x + m.x, x.y + m.y, x.z + m.z, x.w + m.w));
        }
    }
}

// hist2image resize branch, second half (fusion.py:146-149,157): the layer output on the
// [zn*p1, zn*p2] canvas is resized back to [tzh, tzw] (bilinear, align_corners) and its
// in-image part is added onto (or assigned to) the zone rectangle of feat0.
template <typename T>
__global__ void canvas_resize_add_kernel(const T* __restrict__ canvas, T* __restrict__ feat0, int B, int H, int W,
                                         int C, cfp_geom g, int assign) {
    const int oh = g.zone_num * g.p1, ow = g.zone_num * g.p2;
    const int rh = g.ry1 - g.ry0, rw = g.rx1 - g.rx0, V = C / 4;
    const int64_t total = (int64_t)B * rh * rw * V;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % V) * 4;
        int64_t p = i / V;
        int x = g.rx0 + (int)(p % rw), y = g.ry0 + (int)((p / rw) % rh), b = (int)(p / ((int64_t)rw * rh));
        int ty = y - g.sy_wo, tx = x - g.sx_wo;            // cell of the [tzh,tzw] canvas
        float fy = g.tzh > 1 ? ty * ((float)(oh - 1) / (float)(g.tzh - 1)) : 0.f;
        float fx = g.tzw > 1 ? tx * ((float)(ow - 1) / (float)(g.tzw - 1)) : 0.f;
        int y0 = (int)fy, x0 = (int)fx, y1 = min(y0 + 1, oh - 1), x1 = min(x0 + 1, ow - 1);
        float ly = fy - y0, lx = fx - x0;
        const T* base = canvas + (int64_t)b * oh * ow * C + c;
        float4 a = IO<T>::ld4(base + ((int64_t)y0 * ow + x0) * C), bq = IO<T>::ld4(base + ((int64_t)y0 * ow + x1) * C);
        float4 cq = IO<T>::ld4(base + ((int64_t)y1 * ow + x0) * C), d = IO<T>::ld4(base + ((int64_t)y1 * ow + x1) * C);
        float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        float4 v = make_float4(w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x, w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y,
                               w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z, w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w);
        T* dst = feat0 + ((int64_t)b * H * W + (int64_t)y * W + x) * C + c;
        if (!assign) {
            float4 o = IO<T>::ld4(dst);
            v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
        }
        IO<T>::st4(dst, v);
    }
}

// ------------------------------------------------------------------ launchers
template <int C> constexpr size_t kv_smem() {
    constexpr int BM = Tile<C>::BM, NT = 2 * C > 256 ? 256 : 2 * C;
    return (size_t)(BM * (C + 4) + BM * (2 * C + 4) + 2 * 32 * NT) * sizeof(float);
}
template <int C> constexpr size_t query_smem() {
    constexpr int BM = Tile<C>::BM, NT = 2 * C > 256 ? 256 : 2 * C;
    return (size_t)(2 * BM * (2 * C + 4) + 2 * 32 * NT) * sizeof(float);
}

template <int C, int NH, class Src>
static int run_kv_state(const char* name, const Src& src, int groups, const float* wkv_t, float* kv, float* ksum,
                        cudaStream_t st) {
    constexpr int DH = C / NH;
    CFP_REQUIRE(src.rows < ((int64_t)1 << 31), "%s: %lld rows exceed the 32-bit row index", name, (long long)src.rows);
    cudaError_t e = cudaMemsetAsync(kv, 0, (size_t)groups * (C * DH + C) * sizeof(float), st);
    if (e != cudaSuccess) return fail("cudaMemsetAsync(kv state): %s", cudaGetErrorString(e));
    auto k = kv_state_kernel<C, NH, Src>;
    if (int err = set_smem(k, kv_smem<C>())) return err;
    const unsigned grid = (unsigned)((src.rows + Tile<C>::BM - 1) / Tile<C>::BM);
    k<<<grid, kThreads, kv_smem<C>(), st>>>(src, wkv_t, kv, ksum);
    return check_launch(name);
}
template <int C, int NH, bool kAttnOnly, class Q>
static int run_query(const char* name, const Q& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                     cudaStream_t st) {
    CFP_REQUIRE(q.rows < ((int64_t)1 << 31), "%s: %lld rows exceed the 32-bit row index", name, (long long)q.rows);
    auto k = loftr_query_kernel<C, NH, kAttnOnly, Q>;
    if (int err = set_smem(k, query_smem<C>())) return err;
    const unsigned grid = (unsigned)((q.rows + Tile<C>::BM - 1) / Tile<C>::BM);
    k<<<grid, kThreads, query_smem<C>(), st>>>(q, w, kv, ksum);
    return check_launch(name);
}

// KV state for `groups` groups with head dim DH lives at ws+L.kv: [groups][C*DH] then [groups][C].
template <int C, int NH>
static inline void kv_ptrs(char* ws, const WsLayout& L, int groups, float*& kv, float*& ksum) {
    kv = reinterpret_cast<float*>(ws + L.kv);
    ksum = kv + (size_t)groups * (C * (C / NH));
}

template <int C, typename T>
static int d2i_impl(void* feat0, const void* emb, const void* zone_tok, const float* pos2, const uint8_t* mask,
                    int B, int H, int W, int S, const cfp_geom& g, const cfp_loftr_w& w, int assign, char* ws,
                    const WsLayout& L, cudaStream_t st) {
    const int Z = g.zone_num * g.zone_num, groups = B * Z;
    float *kv, *ksum;
    kv_ptrs<C, 4>(ws, L, groups, kv, ksum);
    ZoneTokSrc<T> src((const T*)zone_tok, pos2, S, C, (int64_t)groups * S);
    if constexpr (std::is_same<T, bf16>::value) {
        if (int e = kv_tc_h2i(C, src, S, groups, w.kv_tc, kv, ksum, st)) return e;
    } else {
        if (int e = run_kv_state<C, 4>("kv_state<hist2image>", src, groups, w.wkv_t, kv, ksum, st)) return e;
    }
    ZonePatchRows<T> q((T*)feat0, (const T*)emb, (T*)(ws + L.canvas), mask, H, W, C, g.zone_num, g.p1, g.p2,
                       g.sy_wo, g.sx_wo, g.tzh, g.tzw, g.interpolate, assign, (int64_t)groups * g.p1 * g.p2);
    if constexpr (std::is_same<T, bf16>::value) {
        if (int e = query_tc_h2i(C, q, w, kv, ksum, st)) return e;
    } else {
        if (int e = run_query<C, 4, false>("loftr_query<hist2image>", q, w, kv, ksum, st)) return e;
    }
    if (g.interpolate) {
        const int64_t total = (int64_t)B * (g.ry1 - g.ry0) * (g.rx1 - g.rx0) * (C / 4);
        const int64_t want = (total + 255) / 256;
        const unsigned grid = (unsigned)(want < sm_count() * 8 ? want : sm_count() * 8);
        canvas_resize_add_kernel<T><<<grid, 256, 0, st>>>((const T*)(ws + L.canvas), (T*)feat0, B, H, W, C, g, assign);
        return check_launch("canvas_resize_add_kernel");
    }
    return 0;
}

template <int C, typename T>
static int dapm_impl(const void* feat0, void* msg_map, int B, int H, int W, const cfp_geom& g,
                     const cfp_loftr_w& w, char* ws, const WsLayout& L, cudaStream_t st) {
    const int rw = g.rx1 - g.rx0, Ni = (g.ry1 - g.ry0) * rw, No = H * W - Ni;
    float *kv, *ksum;
    kv_ptrs<C, 4>(ws, L, B, kv, ksum);
    if (Ni > 0) {
        InsideSrc<T> src((const T*)feat0, H, W, C, g.ry0, g.rx0, rw, Ni, (int64_t)B * Ni);
        if constexpr (std::is_same<T, bf16>::value) {
            if (int e = kv_tc_dapm(C, src, Ni, B, w.kv_tc, kv, ksum, st)) return e;
        } else {
            if (int e = run_kv_state<C, 4>("kv_state<dapm>", src, B, w.wkv_t, kv, ksum, st)) return e;
        }
    } else {
        cudaError_t e = cudaMemsetAsync(kv, 0, (size_t)B * (C * (C / 4) + C) * sizeof(float), st);
        if (e != cudaSuccess) return fail("cudaMemsetAsync(kv state): %s", cudaGetErrorString(e));
    }
    if (No == 0) return 0;
    OutsideRows<T> q((const T*)feat0, (T*)msg_map, H, W, C, g.ry0, g.ry1, g.rx0, g.rx1, No, (int64_t)B * No);
    if constexpr (std::is_same<T, bf16>::value) return query_tc_dapm(C, q, w, kv, ksum, st);
    else return run_query<C, 4, true>("attn_query<dapm>", q, w, kv, ksum, st);
}

template <int C, typename T>
static int twins_impl(void* feat0, void* out_nchw, int B, int H, int W, const cfp_twins_w& w, char* ws, const WsLayout& L,
                      cudaStream_t st) {
    const int wsz = w.ws;
    // LSA (transformer.py:94-116): pad to a multiple of ws, attention inside each window, 8 heads
    const int nwy = (H + wsz - 1) / wsz, nwx = (W + wsz - 1) / wsz, nwin = nwy * nwx, groups = B * nwin;
    float *kv, *ksum;
    kv_ptrs<C, 8>(ws, L, groups, kv, ksum);
    WindowRows<T> win((T*)feat0, H, W, C, wsz, nwx, nwin, (int64_t)groups * wsz * wsz);
    if constexpr (std::is_same<T, bf16>::value) {
        if (int e = kv_tc_lsa(C, win, wsz * wsz, groups, w.lsa.kv_tc, kv, ksum, st)) return e;
    } else {
        if (int e = run_kv_state<C, 8>("kv_state<lsa>", win, groups, w.lsa.wkv_t, kv, ksum, st)) return e;
    }
    if constexpr (std::is_same<T, bf16>::value) {
        if (int e = query_tc_lsa(C, win, w.lsa, kv, ksum, st)) return e;
    } else {
        if (int e = run_query<C, 8, false>("loftr_query<lsa>", win, w.lsa, kv, ksum, st)) return e;
    }
    // GSA (transformer.py:138-150): keys/values = LN(sr(x)), stride-ws conv without padding
    const int Ns = (H / wsz) * (W / wsz);
    float* sr_tok = reinterpret_cast<float*>(ws + L.sr);
    if constexpr (std::is_same<T, bf16>::value) {
        if (int e = sr_conv_ln_tc(feat0, sr_tok, B, H, W, C, wsz, w.sr_tc, w.sr_b, w.srln_g, w.srln_b, st)) return e;
    } else {
        if (int e = sr_conv_ln(feat0, sr_tok, B, H, W, C, wsz, w.sr_t, w.sr_b, w.srln_g, w.srln_b, CFP_F32, st)) return e;
    }
    kv_ptrs<C, 8>(ws, L, B, kv, ksum);
    SrTokSrc src(sr_tok, Ns, C, (int64_t)B * Ns);
    if constexpr (std::is_same<T, bf16>::value) {
        if (int e = kv_tc_gsa(C, src, Ns, B, w.gsa.kv_tc, kv, ksum, st)) return e;
    } else {
        if (int e = run_kv_state<C, 8>("kv_state<gsa>", src, B, w.gsa.wkv_t, kv, ksum, st)) return e;
    }
    FrameRows<T> fr((T*)feat0, H * W, C, (int64_t)B * H * W);
    if constexpr (std::is_same<T, bf16>::value) {
        if (out_nchw) {     // last layer of a call: the query chain's last epilogue writes the caller's NCHW map itself
            FrameRowsToNCHW<bf16> fo((const bf16*)feat0, (bf16*)out_nchw, H * W, C, (int64_t)B * H * W);
            return query_tc_gsa_nchw(C, fo, w.gsa, kv, ksum, st);
        }
        return query_tc_gsa(C, fr, w.gsa, kv, ksum, st);
    } else {
        if (int e = run_query<C, 8, false>("loftr_query<gsa>", fr, w.gsa, kv, ksum, st)) return e;
        return out_nchw ? tokens_to_nchw(feat0, out_nchw, B, C, H, W, CFP_F32, st) : 0;
    }
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

int d2i(void* feat0, const void* emb, const void* zone_tok, const float* pos2, const uint8_t* mask, int B, int H,
        int W, int C, int S, const cfp_geom& g, const cfp_loftr_w& w, int assign, char* ws, const WsLayout& L,
        int dtype, cudaStream_t st) {
    CFP_DISPATCH_C_T(d2i_impl, feat0, emb, zone_tok, pos2, mask, B, H, W, S, g, w, assign, ws, L, st)
}
int dapm_attention(const void* feat0, void* msg_map, int B, int H, int W, int C, const cfp_geom& g,
                   const cfp_loftr_w& w, char* ws, const WsLayout& L, int dtype, cudaStream_t st) {
    CFP_DISPATCH_C_T(dapm_impl, feat0, msg_map, B, H, W, g, w, ws, L, st)
}
int twins(void* feat0, void* out_nchw, int B, int H, int W, int C, const cfp_twins_w& w, char* ws, const WsLayout& L, int dtype,
          cudaStream_t st) {
    CFP_DISPATCH_C_T(twins_impl, feat0, out_nchw, B, H, W, w, ws, L, st)
}

}  // namespace cfp
