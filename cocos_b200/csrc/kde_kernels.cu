// Gaussian kernel density estimate: the step after the path in the reference's own
// workflow (notebooks/Getting Started.ipynb cells 65-83: KDE of the simulated terminal prices).
//   gaussian_kernel_estimate_vectorized_whitened      cocos/scientific/kde.py:152-197
//   whitening / whitened_points / normalization       kde.py:1147-1170
// The reference materialises an (n, m, d) residual array per batch and reduces it with array passes; here every
// (data point, evaluation point) pair is one FADD + FMUL + MUFU.EX2 + FFMA in registers:
//   estimate_j = norm * sum_i w_i * 2^(-|p_i - q_j|^2),   p, q = whitened coordinates pre-scaled by sqrt(log2(e)/2).
// MUFU-bound: one MUFU.EX2 per pair, no HBM traffic beyond reading the points once per block.
#include "kernels.h"

namespace cb {

constexpr int kKdeThreads = 128;
constexpr int kKdeTile = 256;      // data points staged per shared-memory tile

// raw weighted moments S1 = sum w x, S2 = sum w x x^T in double (d <= 8), one atomic per block and entry
__global__ void __launch_bounds__(256)
kde_moments_kernel(const float *__restrict__ pts, const float *__restrict__ wts, uint64_t n, int d,
                   double *__restrict__ out /* [1 + d + d*d]: sum w, S1, S2 */) {
    double acc[1 + 8 + 64];
    const int na = 1 + d + d * d;
    for (int a = 0; a < na; ++a) acc[a] = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double w = wts ? (double)wts[i] : 1.0;
        double x[8];
        for (int a = 0; a < d; ++a) x[a] = (double)pts[i * d + a];
        acc[0] += w;
        for (int a = 0; a < d; ++a) {
            acc[1 + a] += w * x[a];
            for (int b = 0; b < d; ++b) acc[1 + d + a * d + b] += w * x[a] * x[b];
        }
    }
    __shared__ double s_acc[8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int a = 0; a < na; ++a) {
        const double v = warp_sum(acc[a]);
        __syncthreads();
        if (lane == 0) s_acc[warp] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += s_acc[w];
            atomicAdd(out + a, t);
        }
    }
}

// out[i, :] = in[i, :] @ W  (W: d x d row-major, already carrying the sqrt(log2(e)/2) scale)
__global__ void __launch_bounds__(256)
kde_whiten_kernel(const float *__restrict__ in, uint64_t n, int d, const float *__restrict__ W,
                  float *__restrict__ out) {
    __shared__ float s_w[64];
    if (threadIdx.x < d * d) s_w[threadIdx.x] = W[threadIdx.x];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x[8], y[8];
        for (int a = 0; a < d; ++a) x[a] = in[i * d + a];
        for (int b = 0; b < d; ++b) {
            float t = 0.f;
            for (int a = 0; a < d; ++a) t = fmaf(x[a], s_w[a * d + b], t);
            y[b] = t;
        }
        for (int b = 0; b < d; ++b) out[i * d + b] = y[b];
    }
}

// grid = (tiles of evaluation points, chunks of data points); each thread owns PJ evaluation points
template <int D, int PJ, bool WEIGHTED>
__global__ void __launch_bounds__(kKdeThreads)
kde_eval_kernel(const float *__restrict__ p, const float *__restrict__ w, uint64_t n, const float *__restrict__ q,
                uint64_t m, uint64_t chunk, float *__restrict__ partial /* [gridDim.y][m] */) {
    __shared__ float s_p[kKdeTile * D];
    __shared__ float s_w[kKdeTile];
    const uint64_t j0 = ((uint64_t)blockIdx.x * kKdeThreads + threadIdx.x) * PJ;
    float qj[PJ][D], acc[PJ];
#pragma unroll
    for (int t = 0; t < PJ; ++t) {
        acc[t] = 0.f;
#pragma unroll
        for (int a = 0; a < D; ++a) qj[t][a] = (j0 + t < m) ? q[(j0 + t) * D + a] : 0.f;
    }
    const uint64_t i_begin = (uint64_t)blockIdx.y * chunk;
    const uint64_t i_end = i_begin + chunk < n ? i_begin + chunk : n;
    for (uint64_t base = i_begin; base < i_end; base += kKdeTile) {
        const int count = (int)(i_end - base < (uint64_t)kKdeTile ? i_end - base : kKdeTile);
        __syncthreads();
        for (int t = threadIdx.x; t < count * D; t += kKdeThreads) s_p[t] = p[base * D + t];
        if (WEIGHTED)
            for (int t = threadIdx.x; t < count; t += kKdeThreads) s_w[t] = w[base + t];
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < count; ++i) {
            float pi[D];
#pragma unroll
            for (int a = 0; a < D; ++a) pi[a] = s_p[i * D + a];          // broadcast reads
            const float wi = WEIGHTED ? s_w[i] : 1.0f;
#pragma unroll
            for (int t = 0; t < PJ; ++t) {
                float r2 = 0.f;
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float diff = pi[a] - qj[t][a];
                    r2 = fmaf(diff, diff, r2);
                }
                const float e = mufu_ex2(-r2);
                acc[t] = WEIGHTED ? fmaf(wi, e, acc[t]) : acc[t] + e;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < PJ; ++t)
        if (j0 + t < m) partial[(uint64_t)blockIdx.y * m + j0 + t] = acc[t];
}

__global__ void __launch_bounds__(256)
kde_finish_kernel(const float *__restrict__ partial, uint64_t m, int chunks, double scale, float *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        double t = 0.0;
        for (int c = 0; c < chunks; ++c) t += (double)partial[(uint64_t)c * m + j];      // fixed order
        out[j] = (float)(t * scale);
    }
}

cudaError_t launch_kde_moments(const float *d_pts, const float *d_w, uint64_t n, int d, double *d_out,
                               cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(double) * (1 + d + d * d), s);
    if (e != cudaSuccess) return e;
    uint64_t b = (n + 255) / 256;
    if (b > 148ull * 4) b = 148ull * 4;
    if (b < 1) b = 1;
    kde_moments_kernel<<<(int)b, 256, 0, s>>>(d_pts, d_w, n, d, d_out);
    return cudaGetLastError();
}

cudaError_t launch_kde_whiten(const float *d_in, uint64_t n, int d, const float *d_W, float *d_out, cudaStream_t s) {
    uint64_t b = (n + 255) / 256;
    if (b > 148ull * 8) b = 148ull * 8;
    if (b < 1) b = 1;
    kde_whiten_kernel<<<(int)b, 256, 0, s>>>(d_in, n, d, d_W, d_out);
    return cudaGetLastError();
}

template <int D, int PJ>
static cudaError_t launch_eval(const float *p, const float *w, uint64_t n, const float *q, uint64_t m, int chunks,
                               uint64_t chunk, float *partial, cudaStream_t s) {
    const uint64_t per_block = (uint64_t)kKdeThreads * PJ;
    dim3 grid((unsigned)((m + per_block - 1) / per_block), (unsigned)chunks);
    if (w) kde_eval_kernel<D, PJ, true><<<grid, kKdeThreads, 0, s>>>(p, w, n, q, m, chunk, partial);
    else kde_eval_kernel<D, PJ, false><<<grid, kKdeThreads, 0, s>>>(p, w, n, q, m, chunk, partial);
    return cudaGetLastError();
}

int kde_plan_chunks(uint64_t n, uint64_t m, int sm_count, int *pj_out) {
    const int pj = m >= (uint64_t)kKdeThreads * 4 * sm_count ? 4 : (m >= (uint64_t)kKdeThreads * 2 ? 2 : 1);
    const uint64_t tiles = (m + (uint64_t)kKdeThreads * pj - 1) / ((uint64_t)kKdeThreads * pj);
    uint64_t want = ((uint64_t)sm_count * 8 + tiles - 1) / tiles;       // ~8 blocks per SM in total
    const uint64_t max_by_data = (n + kKdeTile - 1) / kKdeTile;
    if (want > max_by_data) want = max_by_data;
    if (want > 4096) want = 4096;
    if (want < 1) want = 1;
    *pj_out = pj;
    return (int)want;
}

cudaError_t launch_kde_evaluate(const float *d_p, const float *d_w, uint64_t n, int d, const float *d_q, uint64_t m,
                                double scale, float *d_partial, int chunks, int pj, float *d_out, cudaStream_t s) {
    if (m == 0) return cudaSuccess;
    const uint64_t chunk = ((n + chunks - 1) / chunks + kKdeTile - 1) / kKdeTile * kKdeTile;
    cudaError_t e;
#define CB_KDE_CASE(D)                                                                                  \
    case D:                                                                                              \
        e = pj == 4 ? launch_eval<D, 4>(d_p, d_w, n, d_q, m, chunks, chunk, d_partial, s)                \
          : pj == 2 ? launch_eval<D, 2>(d_p, d_w, n, d_q, m, chunks, chunk, d_partial, s)                \
                    : launch_eval<D, 1>(d_p, d_w, n, d_q, m, chunks, chunk, d_partial, s);               \
        break;
    switch (d) {
        CB_KDE_CASE(1) CB_KDE_CASE(2) CB_KDE_CASE(3) CB_KDE_CASE(4)
        default: return cudaErrorInvalidValue;
    }
#undef CB_KDE_CASE
    if (e != cudaSuccess) return e;
    uint64_t b = (m + 255) / 256;
    if (b > 148ull * 8) b = 148ull * 8;
    kde_finish_kernel<<<(int)b, 256, 0, s>>>(d_partial, m, chunks, scale, d_out);
    return cudaGetLastError();
}

}  // namespace cb
