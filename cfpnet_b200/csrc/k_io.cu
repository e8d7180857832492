// Either side of the path: what the reference does on the host CPU around the model.
//
//   f3  input side, src/utils/dataloader.py:
//         zone_hist       get_hist_parallel (:84-134): depth map -> per-zone 4 cm histogram (torch.histc), strongest
//                         contiguous cluster, (mu, sigma) + validity per zone.  The reference runs a Python loop over the
//                         zones of ONE frame inside the dataloader; here one CTA per (frame, zone), whole batch per launch.
//         zone_samples    sample_point_from_hist_parallel (:65-81): (mu, sigma) -> the 16 depth samples the histogram
//                         encoder consumes (uniform grid over mu +- 3 sigma, or normal quantiles).
//   f4  loss / metrics, src/loss.py:9-19 and src/utils/metrics.py:4-24:
//         silog_fwd/bwd   SILogLoss: bilinear (align_corners) resize of the prediction to the target, masked log
//                         difference, 10 sqrt(var + 0.15 mean^2) and its gradient w.r.t. the prediction.
//         depth_metrics   compute_errors: a1 a2 a3 abs_rel rmse log_10 rmse_log silog sq_rel over the valid pixels.
//
// All of it is memory-bound integer / elementwise / reduction work: coalesced loads, shared-memory histograms,
// warp-shuffle reductions, double-precision accumulators where the reference accumulates in float64.
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ f3: zone histograms
// CTA = (zone, frame).  bin = (int)((x - 0) * nbins / max_distance) in fp32 - the order of operations of torch.histc's CPU
// kernel (checked against torch.histc on edge values for max_distance that are not powers of two); x == max lands in the
// last bin, x outside [0, max] is ignored.  Then dataloader.py:108-118: bin 0 cleared, counts - 20 clipped at 0, only the
// contiguous run of non-zero bins with the largest sum survives (first one on ties, np.argmax), and :128-131 in float64:
// n = sum (float32), mu = sum(centre * count) / (n + 1e-9), sigma = sqrt(sum(count (centre - mu)^2) / (n + 1e-9)) + 1e-9.
__global__ void __launch_bounds__(256) zone_hist_kernel(const float* __restrict__ dep, int H, int W, int sy, int sx, int ph, int pw,
                                                        int zn, int nbins, float max_distance, const double* __restrict__ centres,
                                                        float* __restrict__ fh, uint8_t* __restrict__ mask, int* __restrict__ hist_out) {
    extern __shared__ int sh_hist[];                     // [nbins]
    __shared__ double red[3][8];
    __shared__ int run_lo, run_hi;
    const int z = blockIdx.x, b = blockIdx.y, zy = z / zn, zx = z % zn;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const float* src = dep + ((size_t)b * H + sy + zy * ph) * W + sx + zx * pw;
    const float fb = (float)nbins;
    for (int i = threadIdx.x; i < ph * pw; i += blockDim.x) {
        const float x = src[(size_t)(i / pw) * W + (i % pw)];
        if (x >= 0.f && x <= max_distance) {
            int pos = (int)__fdiv_rn(__fmul_rn(x, fb), max_distance);
            if (pos == nbins) pos = nbins - 1;
            atomicAdd(&sh_hist[pos], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        int c = i == 0 ? 0 : sh_hist[i] - 20;
        sh_hist[i] = c > 0 ? c : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // <= a few hundred bins: one thread scans the runs
        long best = -1;
        int lo = 0, hi = 0, i = 0;
        while (i < nbins) {
            if (sh_hist[i] == 0) { ++i; continue; }
            int j = i;
            long s = 0;
            while (j < nbins && sh_hist[j] != 0) s += sh_hist[j++];
            if (s > best) { best = s; lo = i; hi = j; }
            i = j;
        }
        run_lo = lo; run_hi = hi;
    }
    __syncthreads();
    const int lo = run_lo, hi = run_hi;
    double n = 0.0, m1 = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) { n += sh_hist[i]; m1 += centres[i] * sh_hist[i]; }
    n = warp_sum_d(n); m1 = warp_sum_d(m1);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = n; red[1][threadIdx.x >> 5] = m1; }
    __syncthreads();
    n = 0.0; m1 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { n += red[0][i]; m1 += red[1][i]; }
    // the reference's n is a float32 sum and "n + 1e-9" stays float32 (dataloader.py:126-131): the 1e-9 only matters for n = 0
    const double den = (double)__fadd_rn((float)n, 1e-9f);
    const double mu = m1 / den;
    double m2 = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) { const double d = centres[i] - mu; m2 += sh_hist[i] * d * d; }
    m2 = warp_sum_d(m2);
    if ((threadIdx.x & 31) == 0) red[2][threadIdx.x >> 5] = m2;
    __syncthreads();
    if (threadIdx.x == 0) {
        m2 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m2 += red[2][i];
        const size_t o = (size_t)b * zn * zn + z;
        fh[2 * o] = (float)mu;
        fh[2 * o + 1] = (float)(sqrt(m2 / den) + 1e-9);
        mask[o] = n > 0.0 ? 1 : 0;
    }
    if (hist_out)
        for (int i = threadIdx.x; i < nbins; i += blockDim.x)
            hist_out[((size_t)b * zn * zn + z) * nbins + i] = (i >= lo && i < hi) ? sh_hist[i] : 0;
}

int zone_hist(const float* dep, int B, int H, int W, int sy, int sx, int ph, int pw, int zn, int nbins, float max_distance,
              const double* centres, float* fh, uint8_t* mask, int* hist_out, cudaStream_t st) {
    CFP_REQUIRE(zn >= 1 && ph >= 1 && pw >= 1 && nbins >= 1 && nbins <= 8192, "zone_hist: bad zone grid / bin count");
    CFP_REQUIRE(sy >= 0 && sx >= 0 && sy + zn * ph <= H && sx + zn * pw <= W, "zone_hist: the %dx%d zones of %dx%d px at (%d,%d) leave the %dx%d map",
                zn, zn, ph, pw, sy, sx, H, W);
    CFP_REQUIRE(max_distance > 0.f, "zone_hist: max_distance must be positive");
    zone_hist_kernel<<<dim3(zn * zn, B), 256, nbins * sizeof(int), st>>>(dep, H, W, sy, sx, ph, pw, zn, nbins, max_distance, centres, fh,
                                                                        mask, hist_out);
    return check_launch("zone_hist");
}

// ------------------------------------------------------------------------------------------------ f3: zone samples
// mode 0 (sample_uniform): out[s] = w0[s] * (mu - 3 sigma) + w1[s] * (mu + 3 sigma), w0 / w1 = torch.linspace(1, 0, S) /
// (0, 1, S) tables - the reference blends two ramps (tensor_linspace, :43-58), each product and the sum rounded to fp32 (no
// FMA contraction here either: the result is bit-identical to the reference's).
// mode 1 (normal quantiles): out[s] = mu + (sigma * w0[s]) * sqrt(2), w0[s] = erfinv(2 ppf_s - 1) (torch's Normal.icdf order).
// Invalid zones give zeros.
__global__ void zone_samples_kernel(const float* __restrict__ fh, const uint8_t* __restrict__ mask, float* __restrict__ out,
                                    int64_t zones, int S, const float* __restrict__ w0, const float* __restrict__ w1, int mode) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= zones * S) return;
    const int64_t z = i / S;
    const int s = (int)(i - z * S);
    float r = 0.f;
    if (mask[z]) {
        const float mu = fh[2 * z], sg = fh[2 * z + 1];
        if (mode == 0) {
            const float three_s = __fmul_rn(3.0f, sg);
            const float start = __fsub_rn(mu, three_s), end = __fadd_rn(mu, three_s);
            r = __fadd_rn(__fmul_rn(w0[s], start), __fmul_rn(w1[s], end));
        } else {
            r = __fadd_rn(mu, __fmul_rn(__fmul_rn(sg, w0[s]), 1.41421356237309504880f));
        }
    }
    out[i] = r;
}
int zone_samples(const float* fh, const uint8_t* mask, float* out, int64_t zones, int S, const float* w0, const float* w1, int mode,
                 cudaStream_t st) {
    CFP_REQUIRE(zones >= 0 && S >= 1, "zone_samples: bad shape");
    CFP_REQUIRE(mode == 0 || mode == 1, "zone_samples: mode %d (0 = uniform grid, 1 = normal quantiles)", mode);
    CFP_REQUIRE(w0 && (mode == 1 || w1), "zone_samples: missing weight table");
    if (zones == 0) return 0;
    const int64_t total = zones * S;
    zone_samples_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(fh, mask, out, zones, S, w0, w1, mode);
    return check_launch("zone_samples");
}

// ------------------------------------------------------------------------------------------------ f4: SILog loss
// F.interpolate(pred, (H, W), bilinear, align_corners=True) at target pixel (oy, ox): torch's source index
// src = dst * (in - 1) / (out - 1) in fp32 (area_pixel_compute_scale / source_index), lambda = src - floor.
struct Bilin { int i0, i1; float l0, l1; };
__device__ __forceinline__ Bilin bilin_of(int o, int in, int out) {
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float f = scale * o;
    int i0 = (int)f;
    if (i0 > in - 1) i0 = in - 1;
    const int i1 = i0 + (i0 < in - 1 ? 1 : 0);
    const float l1 = f - (float)i0;
    return Bilin{i0, i1, 1.f - l1, l1};
}
__device__ __forceinline__ float upsample_at(const float* __restrict__ p, int w, const Bilin& by, const Bilin& bx) {
    return by.l0 * (bx.l0 * p[by.i0 * w + bx.i0] + bx.l1 * p[by.i0 * w + bx.i1]) +
           by.l1 * (bx.l0 * p[by.i1 * w + bx.i0] + bx.l1 * p[by.i1 * w + bx.i1]);
}
// scratch (doubles): [0] n, [1] mean, [2] D, [3] loss, [4] ticket (as unsigned), [8 + 3 k ..] per-CTA (count, sum g, sum g^2)
constexpr int kSilogParts = 512;
__global__ void __launch_bounds__(256) silog_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const uint8_t* __restrict__ mask, int B, int h, int w, int H, int W,
                                                        int interpolate, double* __restrict__ scratch, float* __restrict__ loss) {
    const int64_t total = (int64_t)B * H * W;
    double cnt = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        const int ox = (int)(i % W), oy = (int)((i / W) % H), b = (int)(i / ((int64_t)W * H));
        float up;
        if (interpolate) up = upsample_at(pred + (size_t)b * h * w, w, bilin_of(oy, h, H), bilin_of(ox, w, W));
        else up = pred[i];
        const double g = (double)(logf(up) - logf(target[i]));
        cnt += 1.0; s1 += g; s2 += g * g;
    }
    __shared__ double part[3][8];
    __shared__ bool last;
    cnt = warp_sum_d(cnt); s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
    if ((threadIdx.x & 31) == 0) { part[0][threadIdx.x >> 5] = cnt; part[1][threadIdx.x >> 5] = s1; part[2][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b1 = 0.0, c = 0.0;
        for (int i = 0; i < 8; ++i) { a += part[0][i]; b1 += part[1][i]; c += part[2][i]; }
        double* dst = scratch + 8 + 3 * blockIdx.x;
        dst[0] = a; dst[1] = b1; dst[2] = c;
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned*>(scratch + 4), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    double n = 0.0, a1 = 0.0, a2 = 0.0;
    for (int i = 0; i < (int)gridDim.x; ++i) {           // fixed order: the loss is reproducible bit for bit
        const volatile double* src = scratch + 8 + 3 * i;
        n += src[0]; a1 += src[1]; a2 += src[2];
    }
    const double mean = n > 0.0 ? a1 / n : 0.0;
    const double var = n > 1.0 ? (a2 - n * mean * mean) / (n - 1.0) : nan("");      // torch.var of one element is nan
    const double D = var + 0.15 * mean * mean;
    scratch[0] = n; scratch[1] = mean; scratch[2] = D; scratch[3] = 10.0 * sqrt(D);
    *reinterpret_cast<unsigned*>(scratch + 4) = 0u;
    *loss = (float)(10.0 * sqrt(D));
}
int silog_fwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W, int interpolate,
              double* scratch, float* loss, cudaStream_t st) {
    CFP_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "silog: bad shape");
    CFP_REQUIRE(interpolate || (h == H && w == W), "silog: without interpolation the prediction must have the target's size");
    const int64_t total = (int64_t)B * H * W;
    int64_t grid = (total + 255) / 256;
    if (grid > kSilogParts) grid = kSilogParts;
    silog_fwd_kernel<<<(unsigned)grid, 256, 0, st>>>(pred, target, mask, B, h, w, H, W, interpolate, scratch, loss);
    return check_launch("silog_fwd");
}
// dL/dpred: dL/dg_i = gout * (5 / sqrt(D)) * (2 (g_i - mean) / (n - 1) + 0.3 mean / n), dL/dup_i = dL/dg_i / up_i, pulled
// back through the bilinear resize (each target pixel adds to its four source pixels).  grad_pred must be zeroed.
__global__ void __launch_bounds__(256) silog_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const uint8_t* __restrict__ mask, int B, int h, int w, int H, int W,
                                                        int interpolate, const double* __restrict__ scratch, float gout,
                                                        float* __restrict__ grad_pred) {
    const double n = scratch[0], mean = scratch[1], D = scratch[2];
    const double k0 = (double)gout * 5.0 / sqrt(D), k1 = 2.0 / (n - 1.0), k2 = 0.3 * mean / n;
    const int64_t total = (int64_t)B * H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        const int ox = (int)(i % W), oy = (int)((i / W) % H), b = (int)(i / ((int64_t)W * H));
        if (!interpolate) {
            const float up = pred[i];
            const double g = (double)(logf(up) - logf(target[i]));
            grad_pred[i] = (float)(k0 * (k1 * (g - mean) + k2) / (double)up);
            continue;
        }
        const Bilin by = bilin_of(oy, h, H), bx = bilin_of(ox, w, W);
        const float* p = pred + (size_t)b * h * w;
        const float up = upsample_at(p, w, by, bx);
        const double g = (double)(logf(up) - logf(target[i]));
        const float d = (float)(k0 * (k1 * (g - mean) + k2) / (double)up);
        float* gp = grad_pred + (size_t)b * h * w;
        atomicAdd(gp + by.i0 * w + bx.i0, d * by.l0 * bx.l0);
        atomicAdd(gp + by.i0 * w + bx.i1, d * by.l0 * bx.l1);
        atomicAdd(gp + by.i1 * w + bx.i0, d * by.l1 * bx.l0);
        atomicAdd(gp + by.i1 * w + bx.i1, d * by.l1 * bx.l1);
    }
}
int silog_bwd(const float* pred, const float* target, const uint8_t* mask, int B, int h, int w, int H, int W, int interpolate,
              const double* scratch, float gout, float* grad_pred, cudaStream_t st) {
    CFP_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "silog: bad shape");
    cudaError_t e = cudaMemsetAsync(grad_pred, 0, (size_t)B * h * w * sizeof(float), st);
    if (e != cudaSuccess) return fail("silog_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int64_t total = (int64_t)B * H * W;
    const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
    silog_bwd_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(pred, target, mask, B, h, w, H, W, interpolate, scratch, gout,
                                                                          grad_pred);
    return check_launch("silog_bwd");
}

// ------------------------------------------------------------------------------------------------ f4: metrics
// compute_errors (metrics.py:4-24) over the pixels with valid != 0: out[0..8] = a1 a2 a3 abs_rel rmse log_10 rmse_log silog
// sq_rel (the order of the reference's dict), out[9] = number of valid pixels.  scratch: 12 doubles per CTA + 2.
constexpr int kMetricParts = 512, kMetricSums = 11;
__global__ void __launch_bounds__(256) depth_metrics_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                            const uint8_t* __restrict__ valid, int64_t n, double* __restrict__ scratch,
                                                            double* __restrict__ out) {
    double s[kMetricSums];
#pragma unroll
    for (int k = 0; k < kMetricSums; ++k) s[k] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        const double g = gt[i], p = pred[i];
        const double th = fmax(g / p, p / g), d = g - p, lg = log(g), lp = log(p), e = lp - lg;
        s[0] += 1.0;
        s[1] += th < 1.25 ? 1.0 : 0.0;
        s[2] += th < 1.25 * 1.25 ? 1.0 : 0.0;
        s[3] += th < 1.25 * 1.25 * 1.25 ? 1.0 : 0.0;
        s[4] += fabs(d) / g;
        s[5] += d * d / g;
        s[6] += d * d;
        s[7] += (lg - lp) * (lg - lp);
        s[8] += e;
        s[9] += e * e;
        s[10] += fabs(log10(g) - log10(p));
    }
    __shared__ double part[kMetricSums][8];
    __shared__ bool last;
#pragma unroll
    for (int k = 0; k < kMetricSums; ++k) {
        const double v = warp_sum_d(s[k]);
        if ((threadIdx.x & 31) == 0) part[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < kMetricSums) {
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += part[threadIdx.x][i];
        scratch[2 + (size_t)blockIdx.x * kMetricSums + threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned*>(scratch), 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    double t[kMetricSums];
    for (int k = 0; k < kMetricSums; ++k) t[k] = 0.0;
    for (int i = 0; i < (int)gridDim.x; ++i)
        for (int k = 0; k < kMetricSums; ++k) t[k] += *(const volatile double*)(scratch + 2 + (size_t)i * kMetricSums + k);
    const double c = t[0];
    out[0] = t[1] / c; out[1] = t[2] / c; out[2] = t[3] / c;
    out[3] = t[4] / c;
    out[4] = sqrt(t[6] / c);
    out[5] = t[10] / c;
    out[6] = sqrt(t[7] / c);
    out[7] = sqrt(t[9] / c - (t[8] / c) * (t[8] / c)) * 100.0;
    out[8] = t[5] / c;
    out[9] = c;
    *reinterpret_cast<unsigned*>(scratch) = 0u;
}
int depth_metrics(const float* gt, const float* pred, const uint8_t* valid, int64_t n, double* scratch, double* out, cudaStream_t st) {
    CFP_REQUIRE(n > 0, "depth_metrics: empty input");
    int64_t grid = (n + 255) / 256;
    if (grid > kMetricParts) grid = kMetricParts;
    depth_metrics_kernel<<<(unsigned)grid, 256, 0, st>>>(gt, pred, valid, n, scratch, out);
    return check_launch("depth_metrics");
}

}  // namespace cfp
