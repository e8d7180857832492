// Training step (BASELINE config 5), second slice: the fp32 building blocks the linear-attention layers and the
// DAPM convolutions need on top of k_train.cu - row regrouping (gather / scatter-add over host-built index vectors: zone
// canvas cells, LSA windows, DAPM inside / outside sets, conv taps, the strided sr conv), the per-group attention state
// and its application, and the per-head / per-group row arithmetic of attention.py:31-49 and of its closed-form backward
// (DESIGN.md section 8).  cfpnet_b200/train_seq.py orders them; cfpnet_b200/train.py
// (CudaOps) binds them.  Everything is a token-major fp32 [rows][C] matrix, C % 4 == 0.
//
// These are memory-bound one-pass kernels (the reference trains in fp32 with autograd; parity with ITS gradients is the
// bar here, 1e-3 per tensor) - the tensor-core engine of the inference path is not involved.
#include "cfp_common.cuh"
#include "cfp_internal.h"

namespace cfp {

static inline unsigned rows_grid(int64_t work, int per_block = 256) {
    int64_t g = (work + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------ gather / scatter-add
// out[i][:] = idx[i] >= 0 ? src[idx[i]][:] : 0
__global__ void __launch_bounds__(256) tr_gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                             float* __restrict__ out, int64_t n, int C4) {
    const int64_t total = n * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C4;
        const int c = (int)(i - r * C4);
        const int s = idx[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s >= 0) v = reinterpret_cast<const float4*>(src)[(int64_t)s * C4 + c];
        reinterpret_cast<float4*>(out)[i] = v;
    }
}
int tr_gather_rows(const float* src, const int* idx, float* out, int64_t n, int C, cudaStream_t st) {
    CFP_REQUIRE(C > 0 && C % 4 == 0, "gather_rows: C=%d not a multiple of 4", C);
    if (n == 0) return 0;
    tr_gather_rows_kernel<<<rows_grid(n * (C / 4)), 256, 0, st>>>(src, idx, out, n, C / 4);
    return check_launch("tr_gather_rows");
}

// out[idx[i]][:] += src[i][:] for idx[i] >= 0 (out already holds the base map).  Zone / window / inside-outside / sr-conv
// index vectors name a row at most once per call; the adjoint of the 3x3 im2col gather names every pixel nine times (a
// pixel collects from its nine neighbours), so the sums are fp32 atomics and their rounding depends on the arrival order.
__global__ void __launch_bounds__(256) tr_scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                                  float* __restrict__ out, int64_t n, int C) {
    const int64_t total = n * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const int d = idx[r];
        if (d >= 0) atomicAdd(out + (int64_t)d * C + c, src[i]);
    }
}
int tr_scatter_add_rows(const float* src, const int* idx, const float* base, float* out, int64_t n, int64_t base_rows, int C,
                        cudaStream_t st) {
    if (base != out) {
        cudaError_t e = cudaMemcpyAsync(out, base, (size_t)base_rows * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return fail("cudaMemcpyAsync(scatter base): %s", cudaGetErrorString(e));
    }
    if (n == 0) return 0;
    tr_scatter_add_rows_kernel<<<rows_grid(n * C), 256, 0, st>>>(src, idx, out, n, C);
    return check_launch("tr_scatter_add_rows");
}

// ------------------------------------------------------------------------------------------------ attention state
// KV[g][h][i][j] = sum_r A[g][r][h d + i] * B[g][r][h d + j],   As[g][c] = sum_r (w ? w[g][r][h(c)] : 1) * A[g][r][c]
// (forward: A = K = elu(k)+1, B = V: the K^T V state and the K sum of attention.py:39-44; backward: A = Q, B = dnum,
// w = dden: the state's gradient).  CTA = (row chunk, head, group): the chunk's [rows x d] slices of A and B are staged in
// shared memory, thread = (i, j) pairs; partial sums of the chunks meet in KV / As with fp32 atomics (outputs zeroed first).
constexpr int kRedRows = 128;
__global__ void __launch_bounds__(256) tr_attn_reduce_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                             const float* __restrict__ w, float* __restrict__ KV,
                                                             float* __restrict__ As, int R, int C, int nh) {
    extern __shared__ float sm[];
    const int d = C / nh, h = blockIdx.y, g = blockIdx.z;
    const int r0 = blockIdx.x * kRedRows, nr = min(kRedRows, R - r0);
    float* sa = sm;                       // [kRedRows][d]
    float* sb = sa + kRedRows * d;        // [kRedRows][d]
    float* sw = sb + kRedRows * d;        // [kRedRows]
    const int64_t row_base = (int64_t)g * R + r0;
    for (int i = threadIdx.x; i < nr * d; i += 256) {
        const int r = i / d, c = i - r * d;
        const int64_t off = (row_base + r) * C + h * d + c;
        sa[i] = A[off];
        sb[i] = Bm[off];
    }
    for (int r = threadIdx.x; r < nr; r += 256) sw[r] = w ? w[(row_base + r) * nh + h] : 1.f;
    __syncthreads();
    for (int p = threadIdx.x; p < d * d; p += 256) {
        const int i = p / d, j = p - i * d;
        float acc = 0.f;
        for (int r = 0; r < nr; ++r) acc = fmaf(sa[r * d + i], sb[r * d + j], acc);
        atomicAdd(KV + (((int64_t)g * nh + h) * d + i) * d + j, acc);
    }
    for (int i = threadIdx.x; i < d; i += 256) {
        float acc = 0.f;
        for (int r = 0; r < nr; ++r) acc = fmaf(sw[r], sa[r * d + i], acc);
        atomicAdd(As + (int64_t)g * C + h * d + i, acc);
    }
}
int tr_attn_reduce(const float* A, const float* Bm, const float* w, float* KV, float* As, int G, int R, int C, int nh,
                   cudaStream_t st) {
    CFP_REQUIRE(G > 0 && R > 0 && nh > 0 && C % nh == 0 && G <= 65535 && nh <= 65535, "attn_reduce: bad shape G=%d R=%d C=%d nh=%d", G, R, C, nh);
    const int d = C / nh;
    cudaError_t e = cudaMemsetAsync(KV, 0, (size_t)G * C * d * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(As, 0, (size_t)G * C * sizeof(float), st);
    if (e != cudaSuccess) return fail("cudaMemsetAsync(attention state): %s", cudaGetErrorString(e));
    const size_t smem = (size_t)(2 * kRedRows * d + kRedRows) * sizeof(float);
    dim3 grid((R + kRedRows - 1) / kRedRows, nh, G);
    tr_attn_reduce_kernel<<<grid, 256, smem, st>>>(A, Bm, w, KV, As, R, C, nh);
    return check_launch("tr_attn_reduce");
}

// out[g][r][h d + o] = sum_k X[g][r][h d + k] * (transpose ? KV[g][h][o][k] : KV[g][h][k][o])
// (forward: Q x KV; backward: dnum x KV^T, V x dKV^T, K x dKV).  Thread = one output element.
__global__ void __launch_bounds__(256) tr_attn_apply_kernel(const float* __restrict__ X, const float* __restrict__ KV,
                                                            float* __restrict__ out, int64_t rows, int R, int C, int nh, int transpose) {
    const int d = C / nh;
    const int64_t total = rows * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / C;
        const int c = (int)(i - row * C), h = c / d, o = c - h * d;
        const int64_t g = row / R;
        const float* x = X + row * C + h * d;
        const float* m = KV + ((g * nh + h) * d) * d;
        float acc = 0.f;
        if (transpose) {
            for (int k = 0; k < d; ++k) acc = fmaf(x[k], m[o * d + k], acc);
        } else {
            for (int k = 0; k < d; ++k) acc = fmaf(x[k], m[k * d + o], acc);
        }
        out[i] = acc;
    }
}
int tr_attn_apply(const float* X, const float* KV, float* out, int G, int R, int C, int nh, int transpose, cudaStream_t st) {
    CFP_REQUIRE(G > 0 && R > 0 && nh > 0 && C % nh == 0, "attn_apply: bad shape");
    const int64_t rows = (int64_t)G * R;
    tr_attn_apply_kernel<<<rows_grid(rows * C), 256, 0, st>>>(X, KV, out, rows, R, C, nh, transpose);
    return check_launch("tr_attn_apply");
}

// out[row][h] = sum_k a[row][h d + k] * b[brow][h d + k] + eps,   brow = rpg ? row / rpg : row
// (forward: the normaliser Q . Ksum + eps, attention.py:42; backward: dmsg . msg)
__global__ void __launch_bounds__(256) tr_head_dot_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          float* __restrict__ out, int64_t rows, int C, int nh, int rpg, float eps) {
    const int d = C / nh;
    const int64_t total = rows * nh;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / nh;
        const int h = (int)(i - row * nh);
        const int64_t brow = rpg ? row / rpg : row;
        const float* pa = a + row * C + h * d;
        const float* pb = b + brow * C + h * d;
        float acc = 0.f;
        for (int k = 0; k < d; ++k) acc = fmaf(pa[k], pb[k], acc);
        out[i] = acc + eps;
    }
}
int tr_head_dot(const float* a, const float* b, float* out, int64_t rows, int C, int nh, int rpg, float eps, cudaStream_t st) {
    CFP_REQUIRE(rows > 0 && nh > 0 && C % nh == 0 && rpg >= 0, "head_dot: bad shape");
    tr_head_dot_kernel<<<rows_grid(rows * nh), 256, 0, st>>>(a, b, out, rows, C, nh, rpg, eps);
    return check_launch("tr_head_dot");
}

// Row arithmetic with a per-(row, head) or per-group operand; s [rows][nh], b / m indexed by row / rpg:
//   op 0: out = a * s[row][h]          op 1: out = a / s[row][h]                      (head_scale)
//   op 2: out = a + s[row][h] * b[row / rpg][c]                                       (head_axpy: dQ += dden * Ksum)
//   op 3: out = a + b[row / rpg][c]                                                    (group_add: dK += dKsum)
//   op 4: out = a * m[row / rpg]                                                       (group_scale: the zone mask, fusion.py:144)
__global__ void __launch_bounds__(256) tr_rowop_kernel(const float* __restrict__ a, const float* __restrict__ s,
                                                       const float* __restrict__ b, float* __restrict__ out, int64_t rows, int C,
                                                       int nh, int rpg, int op) {
    const int d = C / nh;
    const int64_t total = rows * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / C;
        const int c = (int)(i - row * C);
        const float x = a[i];
        float r;
        switch (op) {
            case 0: r = x * s[row * nh + c / d]; break;
            case 1: r = x / s[row * nh + c / d]; break;
            case 2: r = fmaf(s[row * nh + c / d], b[(row / rpg) * C + c], x); break;
            case 3: r = x + b[(row / rpg) * C + c]; break;
            default: r = x * b[row / rpg]; break;
        }
        out[i] = r;
    }
}
int tr_rowop(const float* a, const float* s, const float* b, float* out, int64_t rows, int C, int nh, int rpg, int op,
             cudaStream_t st) {
    CFP_REQUIRE(op >= 0 && op <= 4, "rowop: unknown op %d", op);
    CFP_REQUIRE(rows > 0 && C > 0 && nh > 0 && C % nh == 0, "rowop: bad shape");
    CFP_REQUIRE((op > 2 || s != nullptr) && (op < 2 || (b != nullptr && rpg > 0)), "rowop: op %d is missing an operand", op);
    tr_rowop_kernel<<<rows_grid(rows * C), 256, 0, st>>>(a, s, b, out, rows, C, nh, rpg > 0 ? rpg : 1, op);
    return check_launch("tr_rowop");
}

}  // namespace cfp
