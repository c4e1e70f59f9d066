// Fused Monte-Carlo pi: uniform draws, quarter-circle test and count in one kernel.
//
//   estimate_pi                examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py:18-33
//   rand_antithetic([n], 0)    cocos/numerics/random.py:219-245 (point i+ceil(n/2) = 1 - point i)
//
// Draw j (a point (ux,uy)) uses words (2*(j%2), 2*(j%2)+1) of Philox(index=j/2, block=0,
// subsequence).  The test is rn(rn(x*x) + rn(y*y)) <= 1 with explicit roundings so the
// count is bit-identical to NumPy's float32 evaluation (oracle_pi_inside).
//
// No memory traffic; bound by the Philox multiplies (IMAD.WIDE on the FMA pipe).  What the kernel does about it:
//   * the counter is (index_lo, index_hi, 0, subsequence): three of its four words are launch- or segment-uniform,
//     so round 0 needs ONE 32x32->64 multiply and round 1 ONE (the others are uniform and precomputed once per
//     thread): 18 multiplies per call instead of 20 (UniformTail below);
//   * round keys live in registers; the loop index is 32 bits (the 64-bit block index is split into segments with
//     a uniform high word); two Philox calls per trip give the scheduler independent chains;
//   * the interior of the range (both draws of a block owned and reflected) runs branch-free; the at most two
//     ragged blocks at the ends of a shard are counted by one thread with the plain code;
//   * the inside test is ONE saturating FMA that yields 1.0 or 0.0 (inside_indicator) and the count is kept in a
//     float (exact below 2^24, flushed to the 64-bit counter in bursts): full-rate FMA-pipe work instead of a
//     compare and a predicated add on the half-rate ALU pipe (profiles/r02_pipe_model.md).
// Grid = every resident block the device can hold (occupancy query), not a fixed blocks-per-SM guess.
// The last block to finish (atomic ticket) owns the total; with a communicator that has NVLink peer mappings it
// also runs the cross-GPU sum there (peer_exchange_fold), so no collective follows the kernel.
#include "kernels.h"
#include "payoff_reduce.cuh"

namespace cb {

__device__ __forceinline__ unsigned inside(float x, float y) {
    return __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)) <= 1.0f ? 1u : 0u;
}

// 1.0f when rn(rn(x*x) + rn(y*y)) <= 1, else 0.0f -- on the FMA pipe: s takes the values ..., 1-2^-24, 1, 1+2^-23, ...
// around the threshold, and 2^23*(1-s) + 1 (exact in an FMA) is >= 1 for s <= 1 and <= 0 for s >= 1+2^-23, so the
// saturating FMA is the indicator.  A compare + predicated add would be two ALU-pipe instructions at half rate.
__device__ __forceinline__ float inside_indicator(float x, float y) {
    const float s = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    float r;
    asm("fma.rn.sat.f32 %0, %1, 0fCB000000, 0f4B000001;" : "=f"(r) : "f"(s));   // sat(s * -2^23 + (2^23 + 1))
    return r;
}

// points inside among the two draws of one block and (ANTI) their reflections (exact small integers in float)
template <bool ANTI>
__device__ __forceinline__ float block_count(const uint4 &o) {
    const float mx0 = mantissa_float(o.x), my0 = mantissa_float(o.y);
    const float mx1 = mantissa_float(o.z), my1 = mantissa_float(o.w);
    float c = inside_indicator(__fadd_rn(mx0, -1.0f), __fadd_rn(my0, -1.0f)) +
              inside_indicator(__fadd_rn(mx1, -1.0f), __fadd_rn(my1, -1.0f));
    if constexpr (ANTI) {      // 1 - u = 2 - mantissa_float, exact like u itself
        c += inside_indicator(__fsub_rn(2.0f, mx0), __fsub_rn(2.0f, my0)) +
             inside_indicator(__fsub_rn(2.0f, mx1), __fsub_rn(2.0f, my1));
    }
    return c;
}

template <bool ANTI>
__global__ void __launch_bounds__(256)
pi_count_kernel(uint64_t n, uint64_t draw_first, uint64_t draw_count, const __grid_constant__ PhiloxKeys kc,
                uint32_t subseq, unsigned long long *__restrict__ total, ReduceWorkspace ws) {
    // draws [draw_first, draw_first+draw_count) of ndraw = ANTI ? ceil(n/2) : n; reflections exist for draws
    // j < floor(n/2), i.e. for all but the last draw of an odd n
    const uint64_t n_reflected = ANTI ? n / 2 : 0;
    const uint64_t draw_end = draw_first + draw_count;
    const uint64_t blk_first = draw_first / 2, blk_end = (draw_end + 1) / 2;
    // interior blocks: both draws owned (and reflected)
    uint64_t in_first = (draw_first + 1) / 2;
    uint64_t in_end = draw_end / 2;
    if (ANTI && n_reflected / 2 < in_end) in_end = n_reflected / 2;
    if (in_end < in_first) in_end = in_first;

    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) { k.k0[r] = pin(kc.k0[r]); k.k1[r] = pin(kc.k1[r]); }

    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long count = 0;
    // segments of the block index with a uniform high word (one segment unless n exceeds 2^34 points)
    for (uint64_t seg = in_first; seg < in_end;) {
        const uint32_t hi = (uint32_t)(seg >> 32);
        const uint64_t seg_end = ((uint64_t)hi + 1) << 32 < in_end ? ((uint64_t)hi + 1) << 32 : in_end;
        const uint32_t lo_first = (uint32_t)seg;
        const uint64_t span = seg_end - seg;                    // <= 2^32
        UniformTail tail;
        tail.init(hi, subseq, k);
        // every thread takes blocks lo_first + gtid + i*stride, two per trip; 32-bit arithmetic inside the loop
        const uint64_t mine = span > gtid ? (span - gtid + stride - 1) / stride : 0;
        uint32_t c0 = lo_first + gtid;
        // float accumulators hold exact counts below 2^24: 8 points per trip, flushed every 2^20 trips
        uint32_t left = (uint32_t)(mine >> 1);
        while (left > 0) {
            const uint32_t burst = left < (1u << 20) ? left : (1u << 20);
            float acc = 0.0f;
            for (uint32_t t = burst; t > 0; --t) {
                const uint4 a = tail.call(c0, k);
                const uint4 b = tail.call(c0 + stride, k);
                c0 += 2u * stride;
                acc += block_count<ANTI>(a) + block_count<ANTI>(b);
            }
            count += (unsigned long long)acc;
            left -= burst;
        }
        if (mine & 1) count += (unsigned long long)block_count<ANTI>(tail.call(c0, k));
        seg = seg_end;
    }
    // ragged blocks at the ends of the shard: a draw not owned, or the unreflected last draw of an odd n
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int e = 0; e < 2; ++e) {
            // blocks of [blk_first, blk_end) outside the interior: below in_first, and from in_end on
            const uint64_t from = e == 0 ? blk_first : (in_end > in_first ? in_end : in_first);
            const uint64_t to = e == 0 ? (in_first < blk_end ? in_first : blk_end) : blk_end;
            for (uint64_t b = from; b < to; ++b) {
                const uint4 o = philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), 0u, subseq, k);
                const float ux[2] = {u01_f32(o.x), u01_f32(o.z)};
                const float uy[2] = {u01_f32(o.y), u01_f32(o.w)};
                for (int h = 0; h < 2; ++h) {
                    const uint64_t j = 2 * b + h;
                    const bool own = j >= draw_first && j < draw_end;
                    if (own) count += inside(ux[h], uy[h]);
                    if (own && j < n_reflected) count += inside(__fsub_rn(1.0f, ux[h]), __fsub_rn(1.0f, uy[h]));
                }
            }
        }
    }
    // warp -> block -> one atomic per block; the last block owns the total
    unsigned lo = (unsigned)count, hi32 = (unsigned)(count >> 32);
    unsigned long long wsum = (unsigned long long)__reduce_add_sync(0xffffffffu, lo);
    wsum += (unsigned long long)__reduce_add_sync(0xffffffffu, hi32) << 32;
    __shared__ unsigned long long s_w[8];
    __shared__ bool s_last;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_w[warp] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) t += s_w[w];
        atomicAdd(total, t);
        __threadfence();
        s_last = atomicAdd(ws.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *ws.ticket = 0u;
    if (ws.px != nullptr) {
        __threadfence();
        __shared__ unsigned long long s_total;
        if (threadIdx.x == 0) s_total = __ldcg(total);
        __syncthreads();
        peer_exchange_fold<unsigned long long>(1, &s_total, ws.px, ws.epoch, ws.sig);
        __syncthreads();
        if (threadIdx.x == 0) *total = s_total;
    }
}

cudaError_t pi_occupancy(int antithetic, int *blocks_per_sm) {
    const void *fn = antithetic ? (const void *)pi_count_kernel<true> : (const void *)pi_count_kernel<false>;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, 256, 0);
}

cudaError_t launch_pi_count(uint64_t n, int antithetic, uint64_t draw_first, uint64_t draw_count,
                            const PhiloxKeys &k, uint32_t subseq, unsigned long long *d_inside, int resident_blocks,
                            ReduceWorkspace ws, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_inside, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (draw_count == 0 && ws.px == nullptr) return cudaSuccess;      // with the exchange an empty shard still publishes 0
    // two Philox blocks per thread and trip
    const uint64_t blocks_needed = ((draw_count + 1) / 2 + 1 + 511) / 512;
    const uint64_t cap = (uint64_t)resident_blocks;
    const int grid = (int)(blocks_needed < cap ? (blocks_needed ? blocks_needed : 1) : cap);
    if (antithetic) pi_count_kernel<true><<<grid, 256, 0, s>>>(n, draw_first, draw_count, k, subseq, d_inside, ws);
    else pi_count_kernel<false><<<grid, 256, 0, s>>>(n, draw_first, draw_count, k, subseq, d_inside, ws);
    return cudaGetLastError();
}

}  // namespace cb
