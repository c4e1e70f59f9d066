// The loop every array fill shares (sm_100a): one Philox4x32-10 block -> W elements -> one 16-byte vector store.
//
//   rand / randn / uniform / exponential / normal / lognormal / logistic     cocos/numerics/random.py:127-153, 330-520
//
// For a 16-byte aligned output:
//   * rounds 0 and 1 of Philox cost one multiply each (UniformTail, philox.cuh: three of the four counter words are
//     uniform), round keys live in registers;
//   * the 64-bit block index is cut into segments with a uniform high word, inside which a thread owns the blocks
//     first + gtid + t*stride: 32-bit arithmetic, a trip count known up front, no bounds test per block and one
//     64-bit pointer bump per trip; two Philox calls per trip are independent chains for the scheduler;
//   * the ragged last block (n not a multiple of W) is written by one thread with scalar stores.
// Element e of the array is value e % W of block e / W whatever the launch geometry.
#pragma once
#include <map>
#include <mutex>
#include <type_traits>

#include "kernels.h"

namespace cb {

template <typename real, int W, typename Make>
__device__ __forceinline__ void philox_fill_aligned(real *__restrict__ out, uint64_t n, const PhiloxKeys &kc,
                                                    uint32_t subseq, Make make) {
    static_assert(W * sizeof(real) == 16, "one 16-byte vector per Philox block");
    using vec_t = typename std::conditional<W == 4, float4, double2>::type;
    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) { k.k0[r] = pin(kc.k0[r]); k.k1[r] = pin(kc.k1[r]); }
    const uint64_t full = n / W;                              // whole blocks: one 16-byte vector each
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    auto store = [](vec_t *dst, const real (&v)[W]) {
        if constexpr (W == 4) __stcs(dst, make_float4(v[0], v[1], v[2], v[3]));
        else __stcs(dst, make_double2(v[0], v[1]));
    };
    for (uint64_t seg = 0; seg < full;) {
        const uint32_t hi = (uint32_t)(seg >> 32);
        const uint64_t seg_end = ((uint64_t)hi + 1) << 32 < full ? ((uint64_t)hi + 1) << 32 : full;
        const uint64_t span = seg_end - seg;
        UniformTail tail;
        tail.init(hi, subseq, k);
        const uint64_t mine = span > gtid ? (span - gtid + stride - 1) / stride : 0;
        uint32_t c0 = (uint32_t)seg + gtid;
        vec_t *dst = reinterpret_cast<vec_t *>(out) + seg + gtid;
        for (uint32_t t = (uint32_t)(mine >> 1); t > 0; --t) {
            real va[W], vb[W];
            const uint4 a = tail.call(c0, k);
            const uint4 b = tail.call(c0 + stride, k);
            make(a, va);
            make(b, vb);
            store(dst, va);
            store(dst + stride, vb);
            c0 += 2u * stride;
            dst += 2ull * stride;
        }
        if (mine & 1) {
            real va[W];
            make(tail.call(c0, k), va);
            store(dst, va);
        }
        seg = seg_end;
    }
    if (gtid == 0 && full * W < n) {                          // the ragged last block
        real v[W];
        make(philox4x32_10((uint32_t)full, (uint32_t)(full >> 32), 0u, subseq, k), v);
        for (int e = 0; e < W; ++e) if (full * W + e < n) out[full * W + e] = v[e];
    }
}

// grid of the fast fills: every block the device can hold at once (occupancy query, cached per kernel), each thread
// then walks its strided share of the Philox blocks
static inline int resident_grid(const void *fn, uint64_t blocks_wanted) {
    static std::map<const void *, int> cache;
    static std::mutex lock;
    int resident = 0;
    {
        std::lock_guard<std::mutex> g(lock);
        auto it = cache.find(fn);
        if (it == cache.end()) {
            int dev = 0, sms = 148, bps = 8;
            if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, 256, 0) != cudaSuccess || bps < 1) bps = 1;
            it = cache.emplace(fn, sms * bps).first;
        }
        resident = it->second;
    }
    if (blocks_wanted < 1) blocks_wanted = 1;
    return (int)(blocks_wanted < (uint64_t)resident ? blocks_wanted : (uint64_t)resident);
}

}  // namespace cb
