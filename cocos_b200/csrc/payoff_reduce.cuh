// Call-payoff accumulation and the grid-wide reduction shared by the fused Heston kernels.
//
//   compute_option_price_from_simulated_paths     examples/stochastic_volatility/heston_utilities.py:99-119
//   host left fold of ComputeDevicePool           cocos/multi_processing/device_pool.py:246-251
//
// Per thread: f64 accumulators of sum(payoff) and sum(antithetic-unit average ^2) -- one strike, or (MULTI) the
// lane-transposed sweep in which lane l keeps strikes l and l+32 for its whole warp and terminal prices are
// broadcast by shuffle (2 shuffles + 2 x ~10 instructions per source lane).  Per block: warp shuffles / shared memory -> partials[block].  The last block to arrive
// (atomic ticket) folds the partials in block order into ws.result, so a launch is bit-reproducible, and resets
// the ticket for the next launch on the stream.  ws.result is also the NCCL send/receive buffer.
#pragma once
#include "kernels.h"

namespace cb {

constexpr int kFusedMaxThreads = 256;
constexpr int kFusedMaxWarps = kFusedMaxThreads / 32;

// The fused cross-GPU step (see PeerExchange in kernels.h).  Runs in the LAST block of a launch only, after
// result[0..n) holds this rank's totals (doubles, or 64-bit counts for the pi kernel).  On return every rank holds
// the rank-ordered sum in result[] -- identical bits everywhere -- or a poison value when a peer never arrived
// (watchdog on the wall clock, %globaltimer: 60 s) or published another launch's data (tag mismatch).
// A real call, not inlined: the time loop of the calling kernel keeps the schedule it has without this code.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long as_word(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ unsigned long long as_word(unsigned long long v) { return v; }
template <typename T> __device__ __forceinline__ T from_word(unsigned long long w);
template <> __device__ __forceinline__ double from_word<double>(unsigned long long w) { return __longlong_as_double((long long)w); }
template <> __device__ __forceinline__ unsigned long long from_word<unsigned long long>(unsigned long long w) { return w; }

constexpr unsigned long long kExchangeWatchdogNs = 60ull * 1000000000ull;

template <typename T>
static __device__ __noinline__ void peer_exchange_fold(int n, T *result, const PeerExchange *table, unsigned int epoch,
                                                       unsigned int sig) {
    const PeerExchange &px = *table;
    const size_t parity_base = (size_t)(epoch & 1u) * kMaxPeers * kSlotStride;
    const unsigned long long tag = ((unsigned long long)(sig & 0xFFFFFFu) << 40) | ((unsigned long long)(n & 0xFF) << 32) | epoch;
    __syncthreads();                                        // result[] was written by threads of this block
    for (int a = threadIdx.x; a <= n; a += blockDim.x) {    // element n of a slot is its tag
        const unsigned long long w = a < n ? as_word(result[a]) : tag;
        const size_t at = parity_base + (size_t)px.rank * kSlotStride + (a < n ? a : kMaxAcc);
        for (int p = 0; p < px.nranks; ++p) px.slots[p][at] = w;
    }
    __threadfence_system();                                 // the slots are out before any flag is
    __syncthreads();
    __shared__ int s_fault;                                 // 0 fine, 1 timed out, 2 tag mismatch
    if (threadIdx.x == 0) s_fault = 0;
    __syncthreads();
    if ((int)threadIdx.x < px.nranks) {
        unsigned int *remote = px.flags[threadIdx.x] + px.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
        const unsigned int *mine = px.flags[px.rank] + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        for (;;) {
            unsigned int seen;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
            if ((int)(seen - epoch) >= 0) break;
            if (global_timer_ns() - t0 > kExchangeWatchdogNs) { atomicMax(&s_fault, 1); break; }   // a rank never launched
            __nanosleep(64);
        }
        // the slot of rank threadIdx.x must hold THIS launch (same number, same element count)
        const unsigned long long got =
            __ldcg(px.slots[px.rank] + parity_base + (size_t)threadIdx.x * kSlotStride + kMaxAcc);
        if (got != tag) atomicMax(&s_fault, 2);
    }
    __syncthreads();
    const int fault = s_fault;
    for (int a = threadIdx.x; a < n; a += blockDim.x) {
        T t = (T)0;
        for (int r = 0; r < px.nranks; ++r)
            t += from_word<T>(__ldcg(px.slots[px.rank] + parity_base + (size_t)r * kSlotStride + a));
        if (fault == 1) t = from_word<T>(kExchangeTimeoutBits);
        else if (fault == 2) t = from_word<T>(kExchangeMismatchBits);
        result[a] = t;
    }
}

template <typename real, bool MULTI>
struct PayoffAccumulator {
    double sum[MULTI ? 2 : 1] = {};
    double sq[MULTI ? 2 : 1] = {};
    // strike sweep (MULTI): the scaled positive part each type accumulates, and what publish() takes out again
    static __device__ __forceinline__ float positive_part(float d) { return fmaxf(d, 0.0f); }       // max(d, 0)
    static __device__ __forceinline__ double positive_part(double d) { return d + fabs(d); }        // 2*max(d, 0)
    static constexpr double kSumScale = sizeof(real) == 8 ? 0.5 : 1.0;
    static constexpr double kSqScale = sizeof(real) == 8 ? 0.0625 : 0.25;

    // S1, S2: terminal prices of a pair; active: the pair exists; paired: its partner exists.
    // Every lane of the warp must call this together (MULTI shuffles).
    // sign = +1 prices calls max(S-K,0), sign = -1 puts max(K-S,0).
    __device__ __forceinline__ void add(real S1, real S2, bool active, bool paired, real K0, const real *s_strikes,
                                        unsigned lane, real sign = (real)1) {
        if constexpr (!MULTI) {
            const real p1 = active ? fmax(sign * (S1 - K0), (real)0) : (real)0;
            const real p2 = paired ? fmax(sign * (S2 - K0), (real)0) : (real)0;
            const double tot = (double)(p1 + p2);
            const double unit = paired ? 0.5 * tot : tot;
            sum[0] += tot;
            sq[0] = fma(unit, unit, sq[0]);
        } else {
            // double: the positive part is taken as d + |d| = 2*max(d, 0) -- ONE add with the |.| operand modifier
            // instead of fmax (DSETP + selects + NaN quieting, ~7 instructions; the sweep then also fits the
            // 128-register cap).  The factor 2 is exact, so sum collects 2*(p1 + p2) and sq (2*(p1 + p2))^2:
            // publish() scales by kSumScale / kSqScale, exactly -- the published bits are those of the plain
            // formulation.  float keeps fmaxf (one FMNMX; d + |d| measured 2% slower there).
            // sq collects (2 * unit average)^2 (float) or (4 * unit average)^2 (double).
            // Dead lanes carry a huge negative FINITE price so that the positive part vanishes for every strike (no
            // mask travels with the prices; -inf would give inf - inf in d + |d|).  Only the lone unpaired path of an
            // odd R (one warp of the whole simulation) needs the factor 2 on its unit and takes the slow loop.
            const real dead = sizeof(real) == 4 ? (real)-1e30f : (real)-1e300;
            const real a1 = active ? sign * S1 : dead;
            const real a2 = paired ? sign * S2 : dead;
            const real sK[2] = {sign * s_strikes[lane], sign * s_strikes[lane + 32]};
            const bool lone = active && !paired;
            if (!__any_sync(0xffffffffu, lone)) {
                for (int src = 0; src < 32; ++src) {
                    const real s1 = __shfl_sync(0xffffffffu, a1, src);
                    const real s2 = __shfl_sync(0xffffffffu, a2, src);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double tot2 = (double)(positive_part(s1 - sK[h]) + positive_part(s2 - sK[h]));
                        sum[h] += tot2;
                        sq[h] = fma(tot2, tot2, sq[h]);
                    }
                }
            } else {
                const real f = lone ? (real)2 : (real)1;
                for (int src = 0; src < 32; ++src) {
                    const real s1 = __shfl_sync(0xffffffffu, a1, src);
                    const real s2 = __shfl_sync(0xffffffffu, a2, src);
                    const double twice = (double)__shfl_sync(0xffffffffu, f, src);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double tot2 = (double)(positive_part(s1 - sK[h]) + positive_part(s2 - sK[h]));
                        sum[h] += tot2;
                        sq[h] = fma(twice * tot2, twice * tot2, sq[h]);
                    }
                }
            }
        }
    }

    // Block reduction -> per-block partials -> the last block folds them in a fixed order.  Call once per thread,
    // by every thread of the block.
    // XCHG: the instantiation carries the fused cross-GPU step (kept out of the single-GPU kernels: its mere presence
    // changes ptxas' schedule of the time loop, -0.9% on the headline kernel)
    // scratch: blocks wider than kFusedMaxThreads (the 512-thread path-materialising kernel) hand in shared memory of
    // their own for the per-warp rows, [blockDim.x / 32][kMaxAcc] doubles
    template <bool XCHG = false>
    __device__ __forceinline__ void publish(int nK, ReduceWorkspace ws, double (*scratch)[kMaxAcc] = nullptr) const {
        const int nacc = 2 * nK;
        __shared__ double s_own[kFusedMaxWarps][kMaxAcc];
        __shared__ bool s_last;
        double (*s_acc)[kMaxAcc] = scratch != nullptr ? scratch : s_own;
        const unsigned lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
        if constexpr (!MULTI) {
            const double a = warp_sum(sum[0]), b = warp_sum(sq[0]);
            if (lane == 0) { s_acc[warp][0] = a; s_acc[warp][1] = b; }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                if (k < nK) { s_acc[warp][k] = kSumScale * sum[h]; s_acc[warp][nK + k] = kSqScale * sq[h]; }
            }
        }
        __syncthreads();
        for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
            double t = 0.0;
            for (int w = 0; w < nwarps; ++w) t += s_acc[w][a];
            ws.partials[(size_t)blockIdx.x * kMaxAcc + a] = t;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(ws.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const int nb = gridDim.x;
        if (nacc <= 2) {
            double t[2] = {0.0, 0.0};
            for (int b = threadIdx.x; b < nb; b += blockDim.x) {
                t[0] += __ldcg(ws.partials + (size_t)b * kMaxAcc);
                t[1] += __ldcg(ws.partials + (size_t)b * kMaxAcc + 1);
            }
            t[0] = warp_sum(t[0]);
            t[1] = warp_sum(t[1]);
            __syncthreads();
            if (lane == 0) { s_acc[warp][0] = t[0]; s_acc[warp][1] = t[1]; }
            __syncthreads();
            if (threadIdx.x < 2) {
                double r = 0.0;
                for (int w = 0; w < nwarps; ++w) r += s_acc[w][threadIdx.x];
                ws.result[threadIdx.x] = r;
            }
        } else {
            for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
                double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
                int b = 0;
                for (; b + 3 < nb; b += 4) {
                    t0 += __ldcg(ws.partials + (size_t)(b + 0) * kMaxAcc + a);
                    t1 += __ldcg(ws.partials + (size_t)(b + 1) * kMaxAcc + a);
                    t2 += __ldcg(ws.partials + (size_t)(b + 2) * kMaxAcc + a);
                    t3 += __ldcg(ws.partials + (size_t)(b + 3) * kMaxAcc + a);
                }
                for (; b < nb; ++b) t0 += __ldcg(ws.partials + (size_t)b * kMaxAcc + a);
                ws.result[a] = (t0 + t1) + (t2 + t3);
            }
        }
        if (threadIdx.x == 0) *ws.ticket = 0u;   // ready for the next launch on this stream
        if constexpr (XCHG) {
            if (ws.px != nullptr) peer_exchange_fold<double>(nacc, ws.result, ws.px, ws.epoch, ws.sig);
        }
    }
};

// strikes of the sweep in shared memory (lane-indexed reads); entries past nK repeat strike 0
template <typename real>
__device__ __forceinline__ void stage_strikes(real *s_strikes, const real *strikes, int nK) {
    if (threadIdx.x < kMaxStrikes) s_strikes[threadIdx.x] = strikes[threadIdx.x < nK ? threadIdx.x : 0];
    for (int t = threadIdx.x + blockDim.x; t < kMaxStrikes; t += blockDim.x) s_strikes[t] = strikes[t < nK ? t : 0];
    __syncthreads();
}

}  // namespace cb
