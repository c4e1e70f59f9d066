// Random-array kernels behind cocos.numerics.random's array API (HBM-write bound).
//
//   rand / randn                   cocos/numerics/random.py:127-153 (af.data.randu / af.data.randn)
//   randn_antithetic               cocos/numerics/random.py:190-216
//   rand_antithetic                cocos/numerics/random.py:219-245
//
// Element order (CPU model: oracle/philox_numpy.py rand_f32 / randn_f32_model / ...):
//   f32 uniform: elements 4j..4j+3 = u01(o0..o3) of Philox(index=j, block=0, subsequence)
//   f32 normal (randn): elements 6j..6j+5 = the three Box-Muller pairs of block j, split like the fused Heston
//        kernel (radius w_d >> 9, angle w_d[8:0]:w3 bits); the antithetic kernels keep two pairs per block:
//        Box-Muller(o0,o1) -> (4j, 4j+1), Box-Muller(o2,o3) -> (4j+2, 4j+3)
//   f64: elements 2j, 2j+1 from Philox(index=j): uniform u52(o0,o1), u52(o2,o3);
//        normal: one Box-Muller pair from U = 1-u52(o0,o1), t = u52(o2,o3)
// Each thread produces 16 bytes per Philox call and stores them as one vector.
#include "kernels.h"
#include "fill_loop.cuh"

namespace cb {

template <typename real, bool NORMAL>
struct Draw;   // fills v[0..W) from one Philox block

template <>
struct Draw<float, false> {
    static constexpr int W = 4;
    static __device__ __forceinline__ void make(const uint4 &o, float (&v)[4]) {
        v[0] = u01_f32(o.x); v[1] = u01_f32(o.y); v[2] = u01_f32(o.z); v[3] = u01_f32(o.w);
    }
    static __device__ __forceinline__ float reflect(float u) { return __fsub_rn(1.0f, u); }
};
template <>
struct Draw<float, true> {
    static constexpr int W = 4;
    static __device__ __forceinline__ void make(const uint4 &o, float (&v)[4]) {
        box_muller(o.x, o.y, -1.3862943611198906f /* -2 ln 2 */, kTwoPi, v[0], v[1]);
        box_muller(o.z, o.w, -1.3862943611198906f, kTwoPi, v[2], v[3]);
    }
    static __device__ __forceinline__ float reflect(float z) { return -z; }
};
template <>
struct Draw<double, false> {
    static constexpr int W = 2;
    static __device__ __forceinline__ void make(const uint4 &o, double (&v)[2]) {
        v[0] = u01_f64(o.x, o.y); v[1] = u01_f64(o.z, o.w);
    }
    static __device__ __forceinline__ double reflect(double u) { return 1.0 - u; }
};
template <>
struct Draw<double, true> {
    static constexpr int W = 2;
    static __device__ __forceinline__ void make(const uint4 &o, double (&v)[2]) {
        const double U = 1.0 - u01_f64(o.x, o.y);              // (0,1]
        const double t = u01_f64(o.z, o.w);                    // [0,1)
        const double r = sqrt(-2.0 * log(U));
        double s, c;
        sincospi(2.0 * t - 1.0, &s, &c);                       // theta = 2 pi t - pi
        v[0] = r * c;
        v[1] = r * s;
    }
    static __device__ __forceinline__ double reflect(double z) { return -z; }
};

__global__ void __launch_bounds__(256)
philox_words_kernel(uint4 *__restrict__ out, uint64_t n_blocks, uint64_t first, uint32_t block,
                    const __grid_constant__ PhiloxKeys k, uint32_t subseq) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_blocks; i += stride) {
        const uint64_t idx = first + i;
        out[i] = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), block, subseq, k);
    }
}

template <typename real, bool NORMAL>
__global__ void __launch_bounds__(256)
random_fill_kernel(real *__restrict__ out, uint64_t n, const __grid_constant__ PhiloxKeys k, uint32_t subseq) {
    using D = Draw<real, NORMAL>;
    constexpr int W = D::W;
    const uint64_t n_blocks = (n + W - 1) / W;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_blocks; j += stride) {
        real v[W];
        D::make(philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, subseq, k), v);
        const uint64_t base = j * W;
        if (aligned && base + W <= n) {
            if constexpr (W == 4) __stcs(reinterpret_cast<float4 *>(out + base), make_float4(v[0], v[1], v[2], v[3]));
            else __stcs(reinterpret_cast<double2 *>(out + base), make_double2(v[0], v[1]));
        } else {
#pragma unroll
            for (int e = 0; e < W; ++e) if (base + e < n) out[base + e] = v[e];
        }
    }
}

// The fast path of random_fill_kernel for a 16-byte aligned output (philox_fill_aligned, fill_loop.cuh): same elements.
template <typename real, bool NORMAL>
__global__ void __launch_bounds__(256)
random_fill_fast_kernel(real *__restrict__ out, uint64_t n, const __grid_constant__ PhiloxKeys kc, uint32_t subseq) {
    using D = Draw<real, NORMAL>;
    philox_fill_aligned<real, D::W>(out, n, kc, subseq, [](const uint4 &o, real (&v)[D::W]) { D::make(o, v); });
}

// Output shape (outer, d, inner) row-major, antithetic along the middle axis:
// draws fill [.., 0:ceil(d/2), ..] in draw order (flattened (outer, ceil(d/2), inner)),
// reflections fill [.., ceil(d/2):d, ..].
template <typename real, bool NORMAL>
__global__ void __launch_bounds__(256)
random_antithetic_kernel(real *__restrict__ out, uint64_t outer, uint64_t d, uint64_t inner,
                         const __grid_constant__ PhiloxKeys k, uint32_t subseq) {
    using D = Draw<real, NORMAL>;
    constexpr int W = D::W;
    const uint64_t da = (d + 1) / 2, df = d / 2;
    const uint64_t slab = da * inner;                 // draws per outer index
    const uint64_t n_draw = outer * slab;
    const uint64_t n_blocks = (n_draw + W - 1) / W;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // fast path (outer == 1, e.g. the Heston (R,2) draw): draw g lands at g, its reflection at slab+g
    const bool vec = (outer == 1) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                     ((slab * sizeof(real)) % 16 == 0);
    const uint64_t n_mirrored = df * inner;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_blocks; j += stride) {
        real v[W];
        D::make(philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, subseq, k), v);
        const uint64_t base = j * W;
        if (vec && base + W <= n_draw) {
            if constexpr (W == 4) __stcs(reinterpret_cast<float4 *>(out + base), make_float4(v[0], v[1], v[2], v[3]));
            else __stcs(reinterpret_cast<double2 *>(out + base), make_double2(v[0], v[1]));
            if (base + W <= n_mirrored) {
                if constexpr (W == 4)
                    __stcs(reinterpret_cast<float4 *>(out + slab + base),
                           make_float4(D::reflect(v[0]), D::reflect(v[1]), D::reflect(v[2]), D::reflect(v[3])));
                else
                    __stcs(reinterpret_cast<double2 *>(out + slab + base),
                           make_double2(D::reflect(v[0]), D::reflect(v[1])));
            } else {
#pragma unroll
                for (int e = 0; e < W; ++e) if (base + e < n_mirrored) out[slab + base + e] = D::reflect(v[e]);
            }
            continue;
        }
#pragma unroll
        for (int e = 0; e < W; ++e) {
            const uint64_t g = base + e;
            if (g >= n_draw) break;
            const uint64_t o = g / slab, rem = g % slab;
            const uint64_t a = rem / inner, in = rem % inner;
            const uint64_t dst = (o * d + a) * inner + in;
            out[dst] = v[e];
            if (a < df) out[dst + da * inner] = D::reflect(v[e]);
        }
    }
}

// float32 normals, six per Philox block: the three (23-bit radius, 19-bit angle) draws of the fused Heston kernel
// (the Philox call is what bounds these fills, so the block is used to the last bit).
// A warp produces 192 consecutive floats per tile (32 blocks); they are transposed through shared memory so that the
// global stores are 16-byte vectors covering 768 contiguous bytes.  Like the other fills: UniformTail rounds, keys
// in registers, 32-bit tile arithmetic inside segments with a uniform high counter word, no bounds test per tile --
// the last, ragged tile of the array is written by one warp with scalar stores.
__device__ __forceinline__ void randn6_draws(const uint4 &o, float (&v)[6]) {
    const uint32_t w[3] = {o.x, o.y, o.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const uint32_t field = (__funnelshift_l(o.w << (10 * d), w[d], 14) & 0x007FFFF0u) | 0x3F800000u;
        const float r = mufu_sqrt(mufu_lg2(2.0f - mantissa_float(w[d])) * -1.3862943611198906f);
        const float theta = fmaf(__uint_as_float(field), kTwoPi, -kThreePi);
        v[2 * d] = r * mufu_cos(theta);
        v[2 * d + 1] = r * mufu_sin(theta);
    }
}

__global__ void __launch_bounds__(256)
randn6_fill_kernel(float *__restrict__ out, uint64_t n, const __grid_constant__ PhiloxKeys kc, uint32_t subseq) {
    __shared__ __align__(16) float s_tile[8][192];
    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) { k.k0[r] = pin(kc.k0[r]); k.k1[r] = pin(kc.k1[r]); }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const uint64_t full_tiles = aligned ? n / 192 : 0;        // tiles written with vector stores
    float2 *mine = reinterpret_cast<float2 *>(&s_tile[warp][lane * 6]);
    const float4 *src = reinterpret_cast<const float4 *>(s_tile[warp]);
    // segments of 2^27 tiles share the high word of the block index (block = 32*tile + lane)
    for (uint64_t seg = 0; seg < full_tiles;) {
        const uint32_t hi = (uint32_t)(seg >> 27);
        const uint64_t seg_end = ((uint64_t)hi + 1) << 27 < full_tiles ? ((uint64_t)hi + 1) << 27 : full_tiles;
        const uint64_t span = seg_end - seg;
        UniformTail tail;
        tail.init(hi, subseq, k);
        const uint32_t trips = span > gwarp ? (uint32_t)((span - gwarp + nwarps - 1) / nwarps) : 0;
        uint32_t c0 = (uint32_t)((seg + gwarp) << 5) + lane;
        float4 *dst = reinterpret_cast<float4 *>(out + (seg + gwarp) * 192) + lane;
        for (uint32_t t = trips; t > 0; --t) {
            float v[6];
            randn6_draws(tail.call(c0, k), v);
            mine[0] = make_float2(v[0], v[1]);
            mine[1] = make_float2(v[2], v[3]);
            mine[2] = make_float2(v[4], v[5]);
            __syncwarp();
            __stcs(dst, src[lane]);
            if (lane < 16) __stcs(dst + 32, src[32 + lane]);
            __syncwarp();
            c0 += nwarps << 5;
            dst += (size_t)nwarps * 48;
        }
        seg = seg_end;
    }
    // what the vector path did not cover: the ragged end (or everything, for an unaligned output)
    const uint64_t n_blocks = (n + 5) / 6;
    const uint64_t first_block = full_tiles * 32;
    for (uint64_t j = first_block + (uint64_t)gwarp * 32 + lane; j < n_blocks; j += (uint64_t)nwarps * 32) {
        float v[6];
        randn6_draws(philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), 0u, subseq, k), v);
#pragma unroll
        for (int e = 0; e < 6; ++e) if (j * 6 + e < n) out[j * 6 + e] = v[e];
    }
}

static int fill_grid(uint64_t n_blocks) {
    uint64_t b = (n_blocks + 255) / 256;
    if (b < 1) b = 1;
    const uint64_t cap = 148ull * 8 * 4;
    return (int)(b > cap ? cap : b);
}

cudaError_t launch_philox_words(uint32_t *d_out, uint64_t n_blocks, uint64_t first, uint32_t block,
                                const PhiloxKeys &k, uint32_t subseq, cudaStream_t s) {
    if (n_blocks == 0) return cudaSuccess;
    philox_words_kernel<<<fill_grid(n_blocks), 256, 0, s>>>(reinterpret_cast<uint4 *>(d_out), n_blocks, first, block,
                                                           k, subseq);
    return cudaGetLastError();
}

template <typename real, bool NORMAL>
cudaError_t launch_random_fill(real *d_out, uint64_t n, const PhiloxKeys &k, uint32_t subseq, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    if constexpr (NORMAL && sizeof(real) == 4) {
        const uint64_t tiles = ((n + 5) / 6 + 31) / 32;
        randn6_fill_kernel<<<resident_grid((const void *)randn6_fill_kernel, (tiles + 7) / 8), 256, 0, s>>>(d_out, n, k,
                                                                                                          subseq);
        return cudaGetLastError();
    }
    constexpr int W = Draw<real, NORMAL>::W;
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
        const void *fn = (const void *)random_fill_fast_kernel<real, NORMAL>;
        // two blocks per thread and trip
        random_fill_fast_kernel<real, NORMAL><<<resident_grid(fn, ((n + W - 1) / W + 511) / 512), 256, 0, s>>>(
            d_out, n, k, subseq);
        return cudaGetLastError();
    }
    random_fill_kernel<real, NORMAL><<<fill_grid((n + W - 1) / W), 256, 0, s>>>(d_out, n, k, subseq);
    return cudaGetLastError();
}

template <typename real, bool NORMAL>
cudaError_t launch_random_antithetic(real *d_out, uint64_t outer, uint64_t d_axis, uint64_t inner,
                                     const PhiloxKeys &k, uint32_t subseq, cudaStream_t s) {
    const uint64_t n_draw = outer * ((d_axis + 1) / 2) * inner;
    if (n_draw == 0) return cudaSuccess;
    constexpr int W = Draw<real, NORMAL>::W;
    random_antithetic_kernel<real, NORMAL><<<fill_grid((n_draw + W - 1) / W), 256, 0, s>>>(d_out, outer, d_axis,
                                                                                           inner, k, subseq);
    return cudaGetLastError();
}

#define CB_INSTANTIATE(real, normal)                                                                              \
    template cudaError_t launch_random_fill<real, normal>(real *, uint64_t, const PhiloxKeys &, uint32_t,          \
                                                          cudaStream_t);                                          \
    template cudaError_t launch_random_antithetic<real, normal>(real *, uint64_t, uint64_t, uint64_t,              \
                                                                const PhiloxKeys &, uint32_t, cudaStream_t);
CB_INSTANTIATE(float, false)
CB_INSTANTIATE(float, true)
CB_INSTANTIATE(double, false)
CB_INSTANTIATE(double, true)

}  // namespace cb
