// Derived distributions of cocos.numerics.random on the Philox core:
//   uniform / exponential / normal / lognormal / logistic   cocos/numerics/random.py:330-365, 488-520   (4 values per block)
//   wald / gamma / beta (chisquare = gamma)                 random.py:372-485  (one value per block, rejection in registers)
//   randint / choice                                        random.py:252-303
//   multivariate_normal                                     random.py:523-531 (z @ cholesky(cov).T)
// The reference builds each of these from several ArrayFire array passes (rand/randn, log, exp, masks, boolean
// gathers for the Marsaglia-Tsang rejection loop).  Here every value is produced in registers from its own
// Philox counter: element e of a launch uses index = e / 4 (simple kinds) or index = e (rejection kinds) and
// block = attempt, so results do not depend on the launch geometry.  CPU model: oracle/distributions_numpy.py.
#include "kernels.h"
#include "fill_loop.cuh"

namespace cb {

enum DistKind : int { kUniform = 0, kExponential = 1, kNormal = 2, kLognormal = 3, kLogistic = 4, kWald = 5,
                      kGamma = 6, kBeta = 7 };

__device__ __forceinline__ float u_open(uint32_t w) {            // (0,1): (k + 0.5) * 2^-23, never 0 or 1
    return __fadd_rn(u01_f32(w), 5.9604644775390625e-08f);
}

// Marsaglia-Tsang (2000) for shape >= 1 (the reference's "Algorithm 4.33", random.py:383-411); attempts are
// Philox blocks (w0,w1 -> candidate normal, w2 -> acceptance uniform).  Returns d*v (unit scale).
__device__ __forceinline__ float gamma_mt(float alpha, uint32_t ilo, uint32_t ihi, uint32_t stream, uint32_t subseq,
                                          const PhiloxKeys &k) {
    const float d = alpha - 1.0f / 3.0f;
    const float c = rsqrtf(9.0f * d);
    float v = 1.0f;
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {
        const uint4 o = philox4x32_10(ilo, ihi, attempt | (stream << 16), subseq, k);
        float z, z_unused;
        box_muller(o.x, o.y, -1.3862943611198906f, kTwoPi, z, z_unused);
        const float y = fmaf(c, z, 1.0f);
        if (y <= 0.0f) continue;
        v = y * y * y;
        const float u = u_open(o.z);
        if (__logf(u) < 0.5f * z * z + d - d * v + d * __logf(v)) break;
    }
    return d * v;
}

__device__ __forceinline__ float gamma_any(float alpha, uint32_t ilo, uint32_t ihi, uint32_t stream, uint32_t subseq,
                                           const PhiloxKeys &k) {
    if (alpha >= 1.0f) return gamma_mt(alpha, ilo, ihi, stream, subseq, k);
    // shape < 1: Gamma(a) = Gamma(a+1) * U^(1/a)  (random.py:413-415); U from the spare word of attempt block 0
    const uint4 o = philox4x32_10(ilo, ihi, stream << 16, subseq, k);
    const float boost = __powf(u_open(o.w), 1.0f / alpha);
    return gamma_mt(alpha + 1.0f, ilo, ihi, stream, subseq, k) * boost;
}

// the four values of one Philox block for the simple kinds
template <int KIND>
__device__ __forceinline__ void dist_values(const uint4 &o, float a, float b, float (&v)[4]) {
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
    if constexpr (KIND == kNormal || KIND == kLognormal) {
        box_muller(o.x, o.y, -1.3862943611198906f, kTwoPi, v[0], v[1]);
        box_muller(o.z, o.w, -1.3862943611198906f, kTwoPi, v[2], v[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = fmaf(b, v[e], a);                                   // loc + scale*z
            if constexpr (KIND == kLognormal) v[e] = __expf(v[e]);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if constexpr (KIND == kUniform) {
                v[e] = __fadd_rn(__fmul_rn(u01_f32(w[e]), b), a);      // u*(high-low)+low, reference roundings
            } else if constexpr (KIND == kExponential) {
                v[e] = -a * __logf(__fsub_rn(1.0f, u01_f32(w[e])));    // log(1-u)*(-scale), 1-u in (0,1]
            } else {                                                   // logistic: loc - scale*log(1/u - 1)
                const float u = u_open(w[e]);
                v[e] = a + b * __logf(__fdividef(u, 1.0f - u));
            }
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
dist_fill_kernel(float *__restrict__ out, uint64_t n, float a, float b, const __grid_constant__ PhiloxKeys k,
                 uint32_t subseq) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if constexpr (KIND <= kLogistic) {
        if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {            // the shared fast loop (fill_loop.cuh)
            philox_fill_aligned<float, 4>(out, n, k, subseq,
                                          [a, b](const uint4 &o, float (&v)[4]) { dist_values<KIND>(o, a, b, v); });
            return;
        }
        const uint64_t n_blocks = (n + 3) / 4;
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_blocks; j += stride) {
            float v[4];
            dist_values<KIND>(philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, subseq, k), a, b, v);
            const uint64_t base = j * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) if (base + e < n) out[base + e] = v[e];
        }
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint32_t ilo = (uint32_t)i, ihi = (uint32_t)(i >> 32);
            float x;
            if constexpr (KIND == kWald) {                                    // random.py:442-452, mu = a, lambda = b
                const uint4 o = philox4x32_10(ilo, ihi, 0u, subseq, k);
                float z, z_unused;
                box_muller(o.x, o.y, -1.3862943611198906f, kTwoPi, z, z_unused);
                const float y = z * z;
                const float mu = a, lam = b;
                x = mu + mu * mu / (2.0f * lam) * y - mu / (2.0f * lam) * sqrtf(4.0f * mu * lam * y + mu * mu * y * y);
                if (u01_f32(o.z) > mu / (mu + x)) x = mu * mu / x;
            } else if constexpr (KIND == kGamma) {                            // shape = a, scale = b
                x = gamma_any(a, ilo, ihi, 0u, subseq, k) * b;
            } else {                                                          // beta(a, b) = X/(X+Y), random.py:431-435
                const float gx = gamma_any(a, ilo, ihi, 0u, subseq, k);
                const float gy = gamma_any(b, ilo, ihi, 1u, subseq, k);
                x = gx / (gx + gy);
            }
            out[i] = x;
        }
    }
}

// randint: trunc(u / divisor) + low with divisor = f32(1/(high-low)) -- the reference's float32 arithmetic
// (random.py:277-285); 4 values per Philox block.
template <typename Int>
__global__ void __launch_bounds__(256)
randint_kernel(Int *__restrict__ out, uint64_t n, float divisor, long long low, const __grid_constant__ PhiloxKeys k,
               uint32_t subseq) {
    const uint64_t n_blocks = (n + 3) / 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_blocks; j += stride) {
        const uint4 o = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, subseq, k);
        const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint64_t i = j * 4 + e;
            if (i < n) out[i] = (Int)((long long)__fdiv_rn(u01_f32(w[e]), divisor) + low);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_kernel(T *__restrict__ out, const T *__restrict__ src, const int *__restrict__ idx, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = src[idx[i]];
}

// rows of standard normals -> mean + L z, in place (L lower-triangular, row-major, in shared memory)
template <typename real>
__global__ void __launch_bounds__(128)
mvn_transform_kernel(real *__restrict__ zx, uint64_t rows, int d, const real *__restrict__ chol,
                     const real *__restrict__ mean) {
    extern __shared__ unsigned char smem_raw[];
    real *L = reinterpret_cast<real *>(smem_raw);
    real *m = L + d * d;
    for (int t = threadIdx.x; t < d * d; t += blockDim.x) L[t] = chol[t];
    for (int t = threadIdx.x; t < d; t += blockDim.x) m[t] = mean[t];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        real *row = zx + r * d;
        for (int j = d - 1; j >= 0; --j) {          // y_j needs z_0..z_j only: descending j may overwrite in place
            real acc = m[j];
            for (int c = 0; c <= j; ++c) acc = fma(L[j * d + c], row[c], acc);
            row[j] = acc;
        }
    }
}

static int dist_grid(uint64_t work) {
    uint64_t b = (work + 255) / 256;
    if (b < 1) b = 1;
    const uint64_t cap = 148ull * 16;
    return (int)(b > cap ? cap : b);
}

cudaError_t launch_dist_fill(int kind, float *d_out, uint64_t n, float a, float b, const PhiloxKeys &k,
                             uint32_t subseq, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t g4_blocks = ((n + 3) / 4 + 511) / 512;      // two Philox blocks per thread and trip (fill_loop.cuh)
    const int g1 = dist_grid(n);
    switch (kind) {
        case kUniform: dist_fill_kernel<kUniform><<<resident_grid((const void *)dist_fill_kernel<kUniform>, g4_blocks), 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kExponential: dist_fill_kernel<kExponential><<<resident_grid((const void *)dist_fill_kernel<kExponential>, g4_blocks), 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kNormal: dist_fill_kernel<kNormal><<<resident_grid((const void *)dist_fill_kernel<kNormal>, g4_blocks), 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kLognormal: dist_fill_kernel<kLognormal><<<resident_grid((const void *)dist_fill_kernel<kLognormal>, g4_blocks), 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kLogistic: dist_fill_kernel<kLogistic><<<resident_grid((const void *)dist_fill_kernel<kLogistic>, g4_blocks), 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kWald: dist_fill_kernel<kWald><<<g1, 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kGamma: dist_fill_kernel<kGamma><<<g1, 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        case kBeta: dist_fill_kernel<kBeta><<<g1, 256, 0, s>>>(d_out, n, a, b, k, subseq); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_randint(void *d_out, int out_is_64, uint64_t n, float divisor, long long low, const PhiloxKeys &k,
                           uint32_t subseq, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int g = dist_grid((n + 3) / 4);
    if (out_is_64) randint_kernel<long long><<<g, 256, 0, s>>>((long long *)d_out, n, divisor, low, k, subseq);
    else randint_kernel<int><<<g, 256, 0, s>>>((int *)d_out, n, divisor, low, k, subseq);
    return cudaGetLastError();
}

cudaError_t launch_gather(void *d_out, const void *d_src, const int *d_idx, uint64_t n, int elem_bytes,
                          cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int g = dist_grid(n);
    if (elem_bytes == 4) gather_kernel<uint32_t><<<g, 256, 0, s>>>((uint32_t *)d_out, (const uint32_t *)d_src, d_idx, n);
    else if (elem_bytes == 8) gather_kernel<uint64_t><<<g, 256, 0, s>>>((uint64_t *)d_out, (const uint64_t *)d_src, d_idx, n);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

template <typename real>
cudaError_t launch_mvn_transform(real *d_zx, uint64_t rows, int d, const real *d_chol, const real *d_mean,
                                 cudaStream_t s) {
    if (rows == 0) return cudaSuccess;
    const size_t smem = sizeof(real) * ((size_t)d * d + d);
    uint64_t b = (rows + 127) / 128;
    if (b > 148ull * 16) b = 148ull * 16;
    mvn_transform_kernel<real><<<(int)b, 128, smem, s>>>(d_zx, rows, d, d_chol, d_mean);
    return cudaGetLastError();
}
template cudaError_t launch_mvn_transform<float>(float *, uint64_t, int, const float *, const float *, cudaStream_t);
template cudaError_t launch_mvn_transform<double>(double *, uint64_t, int, const double *, const double *, cudaStream_t);

}  // namespace cb
