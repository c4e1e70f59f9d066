// Fused Monte-Carlo pi: uniform draws, quarter-circle test and count in one kernel.
//
//   estimate_pi                examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py:18-33
//   rand_antithetic([n], 0)    cocos/numerics/random.py:219-245 (point i+ceil(n/2) = 1 - point i)
//
// Draw j (a point (ux,uy)) uses words (2*(j%2), 2*(j%2)+1) of Philox(index=j/2, block=0,
// subsequence).  The test is rn(rn(x*x) + rn(y*y)) <= 1 with explicit roundings so the
// count is bit-identical to NumPy's float32 evaluation (oracle_pi_inside).
// No memory traffic: issue-bound (Philox: 10 instr per word).
#include "kernels.h"

namespace cb {

__device__ __forceinline__ unsigned inside(float x, float y) {
    return __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)) <= 1.0f ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
pi_count_kernel(uint64_t n, int antithetic, uint64_t draw_first, uint64_t draw_count,
                const __grid_constant__ PhiloxKeys k, uint32_t subseq, unsigned long long *__restrict__ total) {
    // draws [draw_first, draw_first+draw_count) of ndraw = antithetic ? ceil(n/2) : n;
    // reflections exist for draws j < floor(n/2)
    const uint64_t n_reflected = antithetic ? n / 2 : 0;
    const uint64_t draw_end = draw_first + draw_count;
    const uint64_t blk_first = draw_first / 2, blk_end = (draw_end + 1) / 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long count = 0;
    for (uint64_t b = blk_first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < blk_end; b += stride) {
        const uint4 o = philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), 0u, subseq, k);
        const float ux[2] = {u01_f32(o.x), u01_f32(o.z)};
        const float uy[2] = {u01_f32(o.y), u01_f32(o.w)};
        unsigned c = 0;
        const uint64_t j0 = 2 * b;
        if (j0 >= draw_first && j0 + 1 < draw_end && (j0 + 1 < n_reflected || n_reflected == 0)) {
            // interior block (almost all of them): both draws are owned and, if antithetic, both are reflected
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                c += inside(ux[h], uy[h]);
                if (n_reflected) c += inside(__fsub_rn(1.0f, ux[h]), __fsub_rn(1.0f, uy[h]));
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t j = j0 + h;
                const bool own = j >= draw_first && j < draw_end;
                const unsigned a = inside(ux[h], uy[h]);
                const unsigned r = inside(__fsub_rn(1.0f, ux[h]), __fsub_rn(1.0f, uy[h]));
                c += own ? a : 0u;
                c += (own && j < n_reflected) ? r : 0u;
            }
        }
        count += c;
    }
    // warp -> block -> one atomic per block
    unsigned lo = (unsigned)count, hi = (unsigned)(count >> 32);
    unsigned long long wsum = (unsigned long long)__reduce_add_sync(0xffffffffu, lo);   // counts per thread < 2^32 in practice
    wsum += (unsigned long long)__reduce_add_sync(0xffffffffu, hi) << 32;
    __shared__ unsigned long long s_w[8];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += s_w[w];
        atomicAdd(total, t);
    }
}

cudaError_t launch_pi_count(uint64_t n, int antithetic, uint64_t draw_first, uint64_t draw_count,
                            const PhiloxKeys &k, uint32_t subseq, unsigned long long *d_inside, int sm_count,
                            cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_inside, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (draw_count == 0) return cudaSuccess;
    const uint64_t blocks_needed = ((draw_count + 1) / 2 + 1 + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 8;
    const int grid = (int)(blocks_needed < cap ? (blocks_needed ? blocks_needed : 1) : cap);
    pi_count_kernel<<<grid, 256, 0, s>>>(n, antithetic, draw_first, draw_count, k, subseq, d_inside);
    return cudaGetLastError();
}

}  // namespace cb
