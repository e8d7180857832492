// Row providers of the linear-attention layers (shared by the exact-fp32 engine in k_loftr.cu
// and the tensor-core chain in k_chain_tc.cu).  A provider maps a dense row index of a layer's
// query / source set to where that token lives: zone-patch cell of the hist2image canvas, cell
// of a (zero-padded) LSA window, outside-/inside-zone cell of DAPM, plain frame token, ...
// No mask tensor and no gathered copy of the tokens is ever materialised.
//
// Interface (R = provider-specific reference to a located row):
//   int64_t rows                          total dense rows
//   R     locate(int64_t r)               all integer geometry of row r, done ONCE per row
//   int   group(const R&)                 attention group of the row
//   float4 load4(const R&, int c)         channels c..c+3 (zeros for padding cells)
//   void  store4(const R&, int c, float4) query providers only
// The index arithmetic uses FastDiv (multiply-high by a host-computed reciprocal): the divisors
// (window size, zones per side, rows per group, ...) are runtime values and a hardware-emulated
// 64-bit division per chunk used to dominate the staging cost of the small-C layers.
#pragma once
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

template <int C> struct Tile { static constexpr int BM = C >= 128 ? 32 : 64; };

__device__ __forceinline__ uint4 pack8_bf16_fwd(const float (&v)[8]) {
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
    u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
    return u;
}

struct FastDiv {               // exact n / d for 0 <= n < 2^32, 1 <= d < 2^32
    uint64_t m;
    uint32_t d;
    FastDiv() : m(0), d(1) {}
    explicit FastDiv(uint32_t div) : m(div > 1 ? (~0ull / div) + 1 : 0), d(div ? div : 1) {}
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (uint32_t)__umul64hi((uint64_t)n, m); }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

struct RowRef {                // a located token: element offset into the map, group, validity
    int64_t off;
    int g;
    bool ok;
};

template <typename T>
struct ZoneTokSrc {            // hist2image keys/values: zone tokens + positional_encodings2
    const T* tok; const float* pos2; int S, C; int64_t rows; FastDiv dS;
    ZoneTokSrc(const T* t, const float* p, int S_, int C_, int64_t n) : tok(t), pos2(p), S(S_), C(C_), rows(n), dS(S_) {}
    struct R { int64_t off; int g, s; };
    __device__ R locate(int64_t r) const {
        uint32_t g, s;
        dS.divmod((uint32_t)r, g, s);
        return R{r * C, (int)g, (int)s};
    }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const {
        float4 v = IO<T>::ld4(tok + x.off + c);
        float4 p = *reinterpret_cast<const float4*>(pos2 + x.s * C + c);
        return make_float4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
    }
};

template <typename T>
struct WindowRows {            // LSA: ws x ws windows over the zero-padded map (queries and keys)
    T* feat; int H, W, C, ws, nwx, nwin; int64_t rows; FastDiv dL, dWin, dNwx, dWs;
    WindowRows(T* f, int H_, int W_, int C_, int ws_, int nwx_, int nwin_, int64_t n)
        : feat(f), H(H_), W(W_), C(C_), ws(ws_), nwx(nwx_), nwin(nwin_), rows(n), dL(ws_ * ws_), dWin(nwin_), dNwx(nwx_), dWs(ws_) {}
    typedef RowRef R;
    __host__ __device__ uint32_t rows_per_group() const { return dL.d; }
    __device__ int group_of_row(int64_t r) const { return (int)dL.div((uint32_t)r); }
    __device__ R locate(int64_t r) const {
        uint32_t g, l, b, wi, wy, wx, iy, ix;
        dL.divmod((uint32_t)r, g, l);
        dWin.divmod(g, b, wi);
        dNwx.divmod(wi, wy, wx);
        dWs.divmod(l, iy, ix);
        const int y = wy * ws + iy, x = wx * ws + ix;
        return R{((int64_t)b * H * W + (int64_t)y * W + x) * C, (int)g, y < H && x < W};
    }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const {
        return x.ok ? IO<T>::ld4(feat + x.off + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ void store4(const R& x, int c, float4 v) const {
        if (x.ok) IO<T>::st4(feat + x.off + c, v);
    }
    // bf16 storage: one 16-byte access per 8 channels
    __device__ uint4 raw8(const R& x, int c) const {
        static_assert(sizeof(T) == 2, "raw8 is the bf16 fast path");
        return x.ok ? *reinterpret_cast<const uint4*>(feat + x.off + c) : make_uint4(0u, 0u, 0u, 0u);
    }
    __device__ void put8(const R& x, int c, uint4 v) const {
        static_assert(sizeof(T) == 2, "put8 is the bf16 fast path");
        if (x.ok) *reinterpret_cast<uint4*>(feat + x.off + c) = v;
    }
};

template <typename T>
struct FrameRows {             // GSA queries: every token of a frame, group = frame
    T* feat; int N, C; int64_t rows; FastDiv dN;
    FrameRows(T* f, int N_, int C_, int64_t n) : feat(f), N(N_), C(C_), rows(n), dN(N_) {}
    typedef RowRef R;
    __host__ __device__ uint32_t rows_per_group() const { return dN.d; }
    __device__ int group_of_row(int64_t r) const { return (int)dN.div((uint32_t)r); }
    __device__ R locate(int64_t r) const { return R{r * C, (int)dN.div((uint32_t)r), true}; }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const { return IO<T>::ld4(feat + x.off + c); }
    __device__ void store4(const R& x, int c, float4 v) const { IO<T>::st4(feat + x.off + c, v); }
    __device__ uint4 raw8(const R& x, int c) const {
        static_assert(sizeof(T) == 2, "raw8 is the bf16 fast path");
        return *reinterpret_cast<const uint4*>(feat + x.off + c);
    }
    __device__ void put8(const R& x, int c, uint4 v) const {
        static_assert(sizeof(T) == 2, "put8 is the bf16 fast path");
        *reinterpret_cast<uint4*>(feat + x.off + c) = v;
    }
};

// GSA queries of the LAST layer of a call: rows are read token-major like FrameRows, the result goes straight to the
// caller's NCHW map (fusion.py:186, the rearrange back) - lanes are consecutive tokens, so a channel's 32 two-byte stores
// of a warp are one contiguous 64-byte segment.  Saves the tokens -> NCHW pass (a read + a write of the whole map).
template <typename T>
struct FrameRowsToNCHW {
    const T* feat; T* out; int N, C; int64_t rows; FastDiv dN;
    FrameRowsToNCHW(const T* f, T* o, int N_, int C_, int64_t n) : feat(f), out(o), N(N_), C(C_), rows(n), dN(N_) {}
    struct R { int64_t off, ooff; int g; };
    __host__ __device__ uint32_t rows_per_group() const { return dN.d; }
    __device__ int group_of_row(int64_t r) const { return (int)dN.div((uint32_t)r); }
    __device__ R locate(int64_t r) const {
        const uint32_t b = dN.div((uint32_t)r), n = (uint32_t)r - b * (uint32_t)N;
        return R{r * C, (int64_t)b * C * N + n, (int)b};
    }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const { return IO<T>::ld4(feat + x.off + c); }
    __device__ uint4 raw8(const R& x, int c) const {
        static_assert(sizeof(T) == 2, "raw8 is the bf16 fast path");
        return *reinterpret_cast<const uint4*>(feat + x.off + c);
    }
    __device__ void store4(const R& x, int c, float4 v) const {
        T* o = out + x.ooff + (int64_t)c * N;
        IO<T>::st(o, v.x); IO<T>::st(o + N, v.y); IO<T>::st(o + 2 * (int64_t)N, v.z); IO<T>::st(o + 3 * (int64_t)N, v.w);
    }
};

struct SrTokSrc {              // GSA keys/values: fp32 sub-sampled tokens [B][Ns][C]
    const float* tok; int Ns, C; int64_t rows; FastDiv dN;
    SrTokSrc(const float* t, int Ns_, int C_, int64_t n) : tok(t), Ns(Ns_), C(C_), rows(n), dN(Ns_) {}
    typedef RowRef R;
    __device__ R locate(int64_t r) const { return R{r * C, (int)dN.div((uint32_t)r), true}; }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const { return *reinterpret_cast<const float4*>(tok + x.off + c); }
};

template <typename T>
struct InsideSrc {             // DAPM keys/values: tokens inside the zone rectangle, raster order
    const T* feat; int H, W, C, ry0, rx0, rw, Ni; int64_t rows; FastDiv dNi, dRw;
    InsideSrc(const T* f, int H_, int W_, int C_, int ry0_, int rx0_, int rw_, int Ni_, int64_t n)
        : feat(f), H(H_), W(W_), C(C_), ry0(ry0_), rx0(rx0_), rw(rw_), Ni(Ni_), rows(n), dNi(Ni_), dRw(rw_) {}
    typedef RowRef R;
    __device__ R locate(int64_t r) const {
        uint32_t b, i, iy, ix;
        dNi.divmod((uint32_t)r, b, i);
        dRw.divmod(i, iy, ix);
        return R{((int64_t)b * H * W + (int64_t)(ry0 + iy) * W + rx0 + ix) * C, (int)b, true};
    }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const { return IO<T>::ld4(feat + x.off + c); }
    __device__ uint4 raw8(const R& x, int c) const {
        static_assert(sizeof(T) == 2, "raw8 is the bf16 fast path");
        return *reinterpret_cast<const uint4*>(feat + x.off + c);
    }
};

template <typename T>
struct OutsideRows {           // DAPM queries: tokens outside the rectangle; message map out
    const T* feat; T* msg; int H, W, C, ry0, ry1, rx0, rx1, No; int64_t rows; FastDiv dNo, dOut;
    OutsideRows(const T* f, T* m, int H_, int W_, int C_, int ry0_, int ry1_, int rx0_, int rx1_, int No_, int64_t n)
        : feat(f), msg(m), H(H_), W(W_), C(C_), ry0(ry0_), ry1(ry1_), rx0(rx0_), rx1(rx1_), No(No_), rows(n), dNo(No_),
          dOut(W_ - (rx1_ - rx0_) > 0 ? W_ - (rx1_ - rx0_) : 1) {}
    typedef RowRef R;
    __host__ __device__ uint32_t rows_per_group() const { return dNo.d; }
    __device__ int group_of_row(int64_t r) const { return (int)dNo.div((uint32_t)r); }
    __device__ R locate(int64_t r) const {
        uint32_t b, o;
        dNo.divmod((uint32_t)r, b, o);
        const int rw = rx1 - rx0, top = ry0 * W, mid = (ry1 - ry0) * (W - rw);
        int n;
        if ((int)o < top) n = o;
        else if ((int)o < top + mid) {
            uint32_t row, j;
            dOut.divmod(o - top, row, j);
            n = (ry0 + row) * W + ((int)j < rx0 ? j : j + rw);
        } else n = ry1 * W + (o - top - mid);
        return R{((int64_t)b * H * W + n) * C, (int)b, true};
    }
    __device__ int group(const R& x) const { return x.g; }
    __device__ float4 load4(const R& x, int c) const { return IO<T>::ld4(feat + x.off + c); }
    __device__ void store4(const R& x, int c, float4 v) const { IO<T>::st4(msg + x.off + c, v); }
    __device__ uint4 raw8(const R& x, int c) const {
        static_assert(sizeof(T) == 2, "raw8 is the bf16 fast path");
        return *reinterpret_cast<const uint4*>(feat + x.off + c);
    }
    __device__ void put8(const R& x, int c, uint4 v) const {
        static_assert(sizeof(T) == 2, "put8 is the bf16 fast path");
        *reinterpret_cast<uint4*>(msg + x.off + c) = v;
    }
};

// kFast (a compile-time copy of "no resize, no --no_skip_inside, canvas cut from feat0 itself", the geometry every
// benchmarked 416x544 call has): the bilinear-resize gather and the general scatter vanish from the instantiation.  Inlined
// into every 16-byte chunk of the tensor-core chain's fully unrolled stage / epilogue code those branches made
// loftr_query_tc<hist2image,128> 24 k SASS lines (390 KB), 26 % of its stall samples "no instruction" (instruction-cache
// misses) on paths that geometry never takes.
template <typename T, bool kFast = false>
struct ZonePatchRows {         // hist2image queries: cells of the zone canvas, grouped per zone
    T* feat0; const T* emb; T* canvas; const uint8_t* mask;
    int H, W, C, zn, p1, p2, sy_wo, sx_wo, tzh, tzw, interpolate, assign; int64_t rows;
    FastDiv dP, dZ, dZn, dP2;
    bool fast_eligible() const { return !interpolate && !assign && emb == feat0; }
    explicit ZonePatchRows(const ZonePatchRows<T, !kFast>& o)
        : feat0(o.feat0), emb(o.emb), canvas(o.canvas), mask(o.mask), H(o.H), W(o.W), C(o.C), zn(o.zn), p1(o.p1), p2(o.p2),
          sy_wo(o.sy_wo), sx_wo(o.sx_wo), tzh(o.tzh), tzw(o.tzw), interpolate(o.interpolate), assign(o.assign), rows(o.rows),
          dP(o.dP), dZ(o.dZ), dZn(o.dZn), dP2(o.dP2) {}
    ZonePatchRows(T* f, const T* e, T* cv, const uint8_t* m, int H_, int W_, int C_, int zn_, int p1_, int p2_, int sy, int sx,
                  int tzh_, int tzw_, int interp, int assign_, int64_t n)
        : feat0(f), emb(e), canvas(cv), mask(m), H(H_), W(W_), C(C_), zn(zn_), p1(p1_), p2(p2_), sy_wo(sy), sx_wo(sx),
          tzh(tzh_), tzw(tzw_), interpolate(interp), assign(assign_), rows(n), dP(p1_ * p2_), dZ(zn_ * zn_), dZn(zn_), dP2(p2_) {}
    // o / w: resize branch only - element offsets into emb of the four canvas cells a resized cell blends (-1: outside the
    // map, reads as zero) and their bilinear weights, computed ONCE per row in locate().  (Recomputed inside every
    // 4-channel load they made the unrolled stage / last-epilogue code of the chain 14.5 k SASS lines; at batch 1 - the
    // 480x640 latency workload, one wave, cold instruction cache - fetching them was most of the launch.  As an
    // out-of-line call the gather was slower still: 0.40 vs 0.26 ms.)
    struct R { int b, cy, cx, g; bool valid; int64_t o[4]; float w[4]; };
    __host__ __device__ uint32_t rows_per_group() const { return dP.d; }
    __device__ int group_of_row(int64_t r) const { return (int)dP.div((uint32_t)r); }
    __device__ int64_t canvas_off(int b, int ty, int tx) const {
        const int y = sy_wo + ty, x = sx_wo + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) return -1;
        return ((int64_t)b * H * W + (int64_t)y * W + x) * C;
    }
    __device__ R locate(int64_t r) const {
        uint32_t g, l, b, z, zy, zx, py, px;
        dP.divmod((uint32_t)r, g, l);
        dZ.divmod(g, b, z);
        dZn.divmod(z, zy, zx);
        dP2.divmod(l, py, px);
        R q{(int)b, (int)(zy * p1 + py), (int)(zx * p2 + px), (int)g, mask[g] != 0, {0, 0, 0, 0}, {0.f, 0.f, 0.f, 0.f}};
        if (!kFast && interpolate) {
            // F.interpolate(bilinear, align_corners=True) from [tzh,tzw] to [zn*p1, zn*p2]  (fusion.py:141)
            const int oh = zn * p1, ow = zn * p2;
            const float fy = oh > 1 ? q.cy * ((float)(tzh - 1) / (float)(oh - 1)) : 0.f;
            const float fx = ow > 1 ? q.cx * ((float)(tzw - 1) / (float)(ow - 1)) : 0.f;
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = min(y0 + 1, tzh - 1), x1 = min(x0 + 1, tzw - 1);
            const float ly = fy - y0, lx = fx - x0;
            q.o[0] = canvas_off(q.b, y0, x0); q.o[1] = canvas_off(q.b, y0, x1);
            q.o[2] = canvas_off(q.b, y1, x0); q.o[3] = canvas_off(q.b, y1, x1);
            q.w[0] = (1.f - ly) * (1.f - lx); q.w[1] = (1.f - ly) * lx; q.w[2] = ly * (1.f - lx); q.w[3] = ly * lx;
        }
        return q;
    }
    __device__ int group(const R& x) const { return x.g; }
    // value of the zero-padded map at canvas cell (ty,tx) of the un-resized canvas
    __device__ float4 canvas_at(int b, int ty, int tx, int c) const {
        int y = sy_wo + ty, x = sx_wo + tx;
        if (y < 0 || y >= H || x < 0 || x >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return IO<T>::ld4(emb + ((int64_t)b * H * W + (int64_t)y * W + x) * C + c);
    }
    // bf16 storage, no resize: the canvas cell is one 16-byte load (zeros outside the map); the resize branch goes through
    // the fp32 bilinear blend
    __device__ uint4 raw8(const R& q, int c) const {
        static_assert(sizeof(T) == 2, "raw8 is the bf16 fast path");
        if (kFast || !interpolate) {
            const int y = sy_wo + q.cy, x = sx_wo + q.cx;
            if (y < 0 || y >= H || x < 0 || x >= W) return make_uint4(0u, 0u, 0u, 0u);
            return *reinterpret_cast<const uint4*>(emb + ((int64_t)q.b * H * W + (int64_t)y * W + x) * C + c);
        }
        const float4 a = load4(q, c), b = load4(q, c + 4);
        const float v8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        return pack8_bf16_fwd(v8);
    }
    __device__ float4 load4(const R& q, int c) const {
        if (kFast || !interpolate) return canvas_at(q.b, q.cy, q.cx, c);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a = q.o[0] >= 0 ? IO<T>::ld4(emb + q.o[0] + c) : z4, bq = q.o[1] >= 0 ? IO<T>::ld4(emb + q.o[1] + c) : z4;
        const float4 cq = q.o[2] >= 0 ? IO<T>::ld4(emb + q.o[2] + c) : z4, d = q.o[3] >= 0 ? IO<T>::ld4(emb + q.o[3] + c) : z4;
        const float w00 = q.w[0], w01 = q.w[1], w10 = q.w[2], w11 = q.w[3];
        return make_float4(w00 * a.x + w01 * bq.x + w10 * cq.x + w11 * d.x,
                           w00 * a.y + w01 * bq.y + w10 * cq.y + w11 * d.y,
                           w00 * a.z + w01 * bq.z + w10 * cq.z + w11 * d.z,
                           w00 * a.w + w01 * bq.w + w10 * cq.w + w11 * d.w);
    }
    // feat0[zone] += out with out = x + msg: when the canvas is cut from feat0 itself (change_embedding, fusion.py:134),
    // no resize and no --no_skip_inside, the cell's current value IS the x the row was staged from, so the sum is
    // 2x + msg from registers: one 16-byte store instead of a read-modify-write whose load latency ends every tile.
    __device__ bool sums_in_place() const { return kFast || (!interpolate && !assign && emb == feat0); }
    __device__ void put8_sum(const R& q, int c, const float (&o8)[8], const float (&x8)[8]) const {
        const int y = sy_wo + q.cy, x = sx_wo + q.cx;
        if (!q.valid || y < 0 || y >= H || x < 0 || x >= W) return;      // hist_mask / pad_mask (fusion.py:144,112-118)
        float s8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s8[i] = o8[i] + x8[i];
        *reinterpret_cast<uint4*>(feat0 + ((int64_t)q.b * H * W + (int64_t)y * W + x) * C + c) = pack8_bf16_fwd(s8);
    }
    __device__ void store4(const R& q, int c, float4 v) const {
        if (!q.valid) v = make_float4(0.f, 0.f, 0.f, 0.f);   // zone_feature[~hist_mask] = 0  (fusion.py:144)
        if (interpolate) {                                // resized back by canvas_resize_add_kernel
            IO<T>::st4(canvas + (((int64_t)q.b * zn * p1 + q.cy) * (zn * p2) + q.cx) * C + c, v);
            return;
        }
        int y = sy_wo + q.cy, x = sx_wo + q.cx;
        if (y < 0 || y >= H || x < 0 || x >= W) return;    // pad_mask (fusion.py:112-118)
        T* dst = feat0 + ((int64_t)q.b * H * W + (int64_t)y * W + x) * C + c;
        if (assign) { IO<T>::st4(dst, v); return; }       // --no_skip_inside (fusion.py:154-155)
        if (!q.valid) return;
        float4 o = IO<T>::ld4(dst);                       // feat0[zone_mask] += ...  (fusion.py:157)
        IO<T>::st4(dst, make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w));
    }
};

// 8 consecutive channels of a located row as packed bf16 (tensor-core path staging).  Providers whose
// storage already is bf16 expose raw8()/put8() (one 16-byte access, no conversion round trip).
template <class P, class = void> struct HasRaw8 { static constexpr bool value = false; };
template <class P> struct HasRaw8<P, decltype((void)&P::raw8)> { static constexpr bool value = true; };
template <class P, class = void> struct HasPutSum { static constexpr bool value = false; };
template <class P> struct HasPutSum<P, decltype((void)&P::put8_sum)> { static constexpr bool value = true; };
template <class P, class = void> struct HasPut8 { static constexpr bool value = false; };
template <class P> struct HasPut8<P, decltype((void)&P::put8)> { static constexpr bool value = true; };

template <class P>
__device__ __forceinline__ uint4 load8_bf16(const P& p, const typename P::R& ref, int c) {
    if constexpr (HasRaw8<P>::value) {
        return p.raw8(ref, c);
    } else {
        float4 a = p.load4(ref, c), b = p.load4(ref, c + 4);
        __nv_bfloat162 t0 = __floats2bfloat162_rn(a.x, a.y), t1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(b.x, b.y), t3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
        u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
        return u;
    }
}
__device__ __forceinline__ uint4 pack8_bf16(const float (&v)[8]) {
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
    u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
    return u;
}
template <class P>
__device__ __forceinline__ void store8(const P& p, const typename P::R& ref, int c, const float (&v)[8]) {
    if constexpr (HasPut8<P>::value) {
        p.put8(ref, c, pack8_bf16(v));
    } else {
        p.store4(ref, c, make_float4(v[0], v[1], v[2], v[3]));
        p.store4(ref, c + 4, make_float4(v[4], v[5], v[6], v[7]));
    }
}

// k_chain_tc.cu: the query chain on tcgen05 (bf16 activations only)
int query_tc_h2i(int C, const ZonePatchRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st);
int query_tc_lsa(int C, const WindowRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st);
int query_tc_gsa(int C, const FrameRows<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
                 cudaStream_t st);
int query_tc_gsa_nchw(int C, const FrameRowsToNCHW<bf16>& q, const cfp_loftr_w& w, const float* kv, const float* ksum,
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
