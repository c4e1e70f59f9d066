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

// integrate_box_1d (cocos/scientific/kde.py:869-902): sum_i w_i * (Phi((high - x_i)/s) - Phi((low - x_i)/s)) in double;
// Phi(b) - Phi(a) = (erfc(a/sqrt2) - erfc(b/sqrt2)) / 2 keeps the tails accurate.  One atomic per block.
__global__ void __launch_bounds__(256)
kde_box_1d_kernel(const float *__restrict__ pts, const float *__restrict__ wts, uint64_t n, double low, double high,
                  double inv_sd_sqrt2, double *__restrict__ out) {
    double acc = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = (double)pts[i];
        const double w = wts ? (double)wts[i] : 1.0;
        acc += w * 0.5 * (erfc((low - x) * inv_sd_sqrt2) - erfc((high - x) * inv_sd_sqrt2));
    }
    __shared__ double s_acc[8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double v = warp_sum(acc);
    if (lane == 0) s_acc[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += s_acc[w];
        atomicAdd(out, t);
    }
}

// resample (kde.py:987-1025): sample s = dataset[:, index_s] + L z_s with index_s ~ weights and z_s ~ N(0, I_d),
// L = lower Cholesky factor of the kernel covariance.  One Philox4x32-10 block per sample (index = sample number):
// words 2 and 3 pick the data point -- by binary search in the cumulative weights, or floor(u*n) for uniform
// weights -- words 0 and 1 give the Box-Muller pair for dimensions 0 and 1; a second block (block = 1) of the same
// counter feeds dimensions 2 and 3.  Output is (d, size) row-major like the reference's.
__global__ void __launch_bounds__(256)
kde_resample_kernel(const float *__restrict__ pts /* (n, d) */, const double *__restrict__ cumw /* (n) or NULL */,
                    uint64_t n, int d, const float *__restrict__ chol /* d x d lower, row-major */, uint64_t size,
                    const __grid_constant__ PhiloxKeys k, uint32_t subseq, float *__restrict__ out /* (d, size) */) {
    __shared__ float s_l[16];
    if (threadIdx.x < d * d) s_l[threadIdx.x] = chol[threadIdx.x];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < size; s += stride) {
        const uint4 o = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), 0u, subseq, k);
        // data point: 32 + 23 bits of uniform (double), so that large n and small weights are resolved
        const double u01 = ((double)o.w + (double)(o.z >> 9) * (1.0 / 8388608.0)) * (1.0 / 4294967296.0);   // [0,1)
        uint64_t idx;
        if (cumw == nullptr) {
            idx = (uint64_t)(u01 * (double)n);
            if (idx >= n) idx = n - 1;
        } else {
            const double u = u01 * cumw[n - 1];
            uint64_t lo = 0, hi = n - 1;                       // first i with cumw[i] > u
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (cumw[mid] > u) hi = mid; else lo = mid + 1;
            }
            idx = lo;
        }
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        box_muller(o.x, o.y, -1.3862943611198906f, kTwoPi, z[0], z[1]);
        if (d > 2) {
            const uint4 o2 = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), 1u, subseq, k);
            box_muller(o2.x, o2.y, -1.3862943611198906f, kTwoPi, z[2], z[3]);
        }
        for (int a = 0; a < d; ++a) {
            float t = pts[idx * d + a];
            for (int b = 0; b <= a; ++b) t = fmaf(s_l[a * d + b], z[b], t);
            out[(uint64_t)a * size + s] = t;
        }
    }
}

cudaError_t launch_kde_box_1d(const float *d_pts, const float *d_w, uint64_t n, double low, double high, double sd,
                              double *d_out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(double), s);
    if (e != cudaSuccess) return e;
    uint64_t b = (n + 255) / 256;
    if (b > 148ull * 8) b = 148ull * 8;
    if (b < 1) b = 1;
    kde_box_1d_kernel<<<(int)b, 256, 0, s>>>(d_pts, d_w, n, low, high, 1.0 / (sd * 1.4142135623730951), d_out);
    return cudaGetLastError();
}

cudaError_t launch_kde_resample(const float *d_pts, const double *d_cumw, uint64_t n, int d, const float *d_chol,
                                uint64_t size, const PhiloxKeys &k, uint32_t subseq, float *d_out, cudaStream_t s) {
    if (size == 0) return cudaSuccess;
    uint64_t b = (size + 255) / 256;
    if (b > 148ull * 8) b = 148ull * 8;
    kde_resample_kernel<<<(int)b, 256, 0, s>>>(d_pts, d_cumw, n, d, d_chol, size, k, subseq, d_out);
    return cudaGetLastError();
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
