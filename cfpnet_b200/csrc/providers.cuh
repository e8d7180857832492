// Row providers of the linear-attention layers (shared by the exact-fp32 engine in k_loftr.cu
// and the tensor-core chain in k_chain_tc.cu).  A provider maps a dense row index of a layer's
// query / source set to where that token lives: zone-patch cell of the hist2image canvas, cell
// of a (zero-padded) LSA window, outside-/inside-zone cell of DAPM, plain frame token, ...
// No mask tensor and no gathered copy of the tokens is ever materialised.
#pragma once
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

template <int C> struct Tile { static constexpr int BM = C >= 128 ? 32 : 64; };

// ------------------------------------------------------------------ providers
// Common interface:
//   int64_t rows;                     total dense rows
//   int group(int64_t r)              attention group of row r
//   float4 load4(int64_t r, int c)    channels c..c+3 of row r (zeros for padding)
//   void store4(int64_t r, int c, float4 v)      (query providers only)

template <typename T>
struct ZoneTokSrc {            // hist2image keys/values: zone tokens + positional_encodings2
    const T* tok; const float* pos2; int S, C; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / S); }
    __device__ float4 load4(int64_t r, int c) const {
        float4 v = IO<T>::ld4(tok + r * C + c);
        float4 p = *reinterpret_cast<const float4*>(pos2 + (r % S) * C + c);
        return make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
    }
};

template <typename T>
struct WindowRows {            // LSA: ws x ws windows over the zero-padded map (queries and keys)
    T* feat; int H, W, C, ws, nwx, nwin; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / (ws * ws)); }
    __device__ bool locate(int64_t r, int64_t& off) const {
        int g = (int)(r / (ws * ws)), l = (int)(r % (ws * ws));
        int b = g / nwin, wi = g % nwin;
        int y = (wi / nwx) * ws + l / ws, x = (wi % nwx) * ws + l % ws;
        off = ((int64_t)b * H * W + (int64_t)y * W + x) * C;
        return y < H && x < W;
    }
    __device__ float4 load4(int64_t r, int c) const {
        int64_t off;
        return locate(r, off) ? IO<T>::ld4(feat + off + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ void store4(int64_t r, int c, float4 v) const {
        int64_t off;
        if (locate(r, off)) IO<T>::st4(feat + off + c, v);
    }
};

template <typename T>
struct FrameRows {             // GSA queries: every token of a frame, group = frame
    T* feat; int N, C; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / N); }
    __device__ float4 load4(int64_t r, int c) const { return IO<T>::ld4(feat + r * C + c); }
    __device__ void store4(int64_t r, int c, float4 v) const { IO<T>::st4(feat + r * C + c, v); }
};

struct SrTokSrc {              // GSA keys/values: fp32 sub-sampled tokens [B][Ns][C]
    const float* tok; int Ns, C; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / Ns); }
    __device__ float4 load4(int64_t r, int c) const { return *reinterpret_cast<const float4*>(tok + r * C + c); }
};

template <typename T>
struct InsideSrc {             // DAPM keys/values: tokens inside the zone rectangle, raster order
    const T* feat; int H, W, C, ry0, rx0, rw, Ni; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / Ni); }
    __device__ float4 load4(int64_t r, int c) const {
        int b = (int)(r / Ni), i = (int)(r % Ni);
        int y = ry0 + i / rw, x = rx0 + i % rw;
        return IO<T>::ld4(feat + ((int64_t)b * H * W + (int64_t)y * W + x) * C + c);
    }
};

template <typename T>
struct OutsideRows {           // DAPM queries: tokens outside the rectangle; message map out
    const T* feat; T* msg; int H, W, C, ry0, ry1, rx0, rx1, No; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / No); }
    __device__ int64_t locate(int64_t r) const {
        int b = (int)(r / No), o = (int)(r % No);
        const int rw = rx1 - rx0, top = ry0 * W, mid = (ry1 - ry0) * (W - rw);
        int n;
        if (o < top) n = o;
        else if (o < top + mid) {
            int q = o - top, row = q / (W - rw), j = q % (W - rw);
            n = (ry0 + row) * W + (j < rx0 ? j : j + rw);
        } else n = ry1 * W + (o - top - mid);
        return ((int64_t)b * H * W + n) * C;
    }
    __device__ float4 load4(int64_t r, int c) const { return IO<T>::ld4(feat + locate(r) + c); }
    __device__ void store4(int64_t r, int c, float4 v) const { IO<T>::st4(msg + locate(r) + c, v); }
};

template <typename T>
struct ZonePatchRows {         // hist2image queries: cells of the zone canvas, grouped per zone
    T* feat0; const T* emb; T* canvas; const uint8_t* mask;
    int H, W, C, zn, p1, p2, sy_wo, sx_wo, tzh, tzw, interpolate, assign; int64_t rows;
    __device__ int group(int64_t r) const { return (int)(r / (p1 * p2)); }
    __device__ void cell(int64_t r, int& b, int& cy, int& cx) const {
        int g = (int)(r / (p1 * p2)), l = (int)(r % (p1 * p2));
        b = g / (zn * zn);
        int z = g % (zn * zn);
        cy = (z / zn) * p1 + l / p2;
        cx = (z % zn) * p2 + l % p2;
    }
    // value of the zero-padded map at canvas cell (ty,tx) of the un-resized canvas
    __device__ float4 canvas_at(int b, int ty, int tx, int c) const {
        int y = sy_wo + ty, x = sx_wo + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return IO<T>::ld4(emb + ((int64_t)b * H * W + (int64_t)y * W + x) * C + c);
    }
    __device__ float4 load4(int64_t r, int c) const {
        int b, cy, cx;
        cell(r, b, cy, cx);
        if (!interpolate) return canvas_at(b, cy, cx, c);
        // F.interpolate(bilinear, align_corners=True) from [tzh,tzw] to [zn*p1, zn*p2]  (fusion.py:141)
        const int oh = zn * p1, ow = zn * p2;
        float fy = oh > 1 ? cy * ((float)(tzh - 1) / (float)(oh - 1)) : 0.f;
        float fx = ow > 1 ? cx * ((float)(tzw - 1) / (float)(ow - 1)) : 0.f;
        int y0 = (int)fy, x0 = (int)fx;
        int y1 = min(y0 + 1, tzh - 1), x1 = min(x0 + 1, tzw - 1);
        float ly = fy - y0, lx = fx - x0;
        float4 a = canvas_at(b, y0, x0, c), bq = canvas_at(b, y0, x1, c);
        float4 cq = canvas_at(b, y1, x0, c), d = canvas_at(b, y1, x1, c);
        float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
        return make_float4(w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x,
                           w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y,
                           w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z,
                           w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w);
    }
    __device__ void store4(int64_t r, int c, float4 v) const {
        int b, cy, cx;
        cell(r, b, cy, cx);
        const bool valid = mask[group(r)] != 0;           // zone_feature[~hist_mask] = 0  (fusion.py:144)
        if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (interpolate) {                                // resized back by canvas_resize_add_kernel
            IO<T>::st4(canvas + (((int64_t)b * zn * p1 + cy) * (zn * p2) + cx) * C + c, v);
            return;
        }
        int y = sy_wo + cy, x = sx_wo + cx;
        if (y < 0 || y >= H || x < 0 || x >= W) return;    // pad_mask (fusion.py:112-118)
        T* dst = feat0 + ((int64_t)b * H * W + (int64_t)y * W + x) * C + c;
        if (assign) { IO<T>::st4(dst, v); return; }       // --no_skip_inside (fusion.py:154-155)
        if (!valid) return;
        float4 o = IO<T>::ld4(dst);                       // feat0[zone_mask] += ...  (fusion.py:157)
        IO<T>::st4(dst, make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w));
    }
};


// 8 consecutive channels of a row as packed bf16 (tensor-core path staging)
template <class P>
__device__ __forceinline__ uint4 load8_bf16(const P& p, int64_t r, int c) {
    float4 a = p.load4(r, c), b = p.load4(r, c + 4);
    __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x, a.y), t1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(b.x, b.y), t3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
    u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
    return u;
}
template <class P>
__device__ __forceinline__ void store8(const P& p, int64_t r, int c, const float (&v)[8]) {
    p.store4(r, c, make_float4(v[0], v[1], v[2], v[3]));
    p.store4(r, c + 4, make_float4(v[4], v[5], v[6], v[7]));
}

// k_chain_tc.cu: the query chain on tcgen05 (bf16 activations only)
int query_tc_h2i(int C, const ZonePatchRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st);
int query_tc_lsa(int C, const WindowRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st);
int query_tc_gsa(int C, const FrameRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st);
int query_tc_dapm(int C, const OutsideRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                  cudaStream_t st);
// attention state on tcgen05; S = rows per group
int kv_tc_h2i(int C, const ZoneTokSrc<bf16>& src, int S, int groups, const void* wkv_tc, float* kv, float* ksum,
              cudaStream_t st);
int kv_tc_lsa(int C, const WindowRows<bf16>& src, int S, int groups, const void* wkv_tc, float* kv, float* ksum,
              cudaStream_t st);
int kv_tc_gsa(int C, const SrTokSrc& src, int S, int groups, const void* wkv_tc, float* kv, float* ksum,
              cudaStream_t st);
int kv_tc_dapm(int C, const InsideSrc<bf16>& src, int S, int groups, const void* wkv_tc, float* kv, float* ksum,
               cudaStream_t st);

}  // namespace cfp
