// Fused Heston kernel (sm_100a): Philox4x32-10 -> Box-Muller -> antithetic pair ->
// Euler-Maruyama steps -> call payoffs -> block reduction, all in registers.
//
// Replaces, for the B200 bundle, the reference's Python time loop and its ~21 array
// passes per step:
//   simulate_heston_model                        examples/stochastic_volatility/heston_utilities.py:58-96
//   randn_antithetic((R,2), 0) per step          cocos/numerics/random.py:190-216
//   compute_option_price_from_simulated_paths    heston_utilities.py:99-119
//
// One thread owns PPT antithetic pairs (path p and its partner p+ceil(R/2), which
// sees the negated normals) for all N-1 steps.  Nothing is read from or written to
// HBM inside the time loop (the path-materialising variant is heston_trajectory.cu).  Per pair-step:
// 5 MUFU (lg2, sin, cos, 2x sqrt(v*r^2)) against ~43 other issue slots: the MUFU (XU)
// pipe and the issue port are both close to their limits.
//
// Algebra (statistical mode, not the parity order): with a = 1-kappa*dt,
// b = kappa*v_bar*dt, w = sigma_v*(rho*dB0 + sqrt(1-rho^2)*dB1):
//   x' = x - 0.5*dt*v + sqrt(v)*dB0          (mu*dt*(N-1) + x0 added once at the end)
//   v' = max(a*v + b + sqrt(v)*w, eps_f32)
#include "kernels.h"
#include "heston_step.cuh"
#include "payoff_reduce.cuh"

namespace cb {

constexpr int kMaxThreads = kFusedMaxThreads;

// ---- the kernel -------------------------------------------------------------

// PAYOFF: 0 terminal (European), 1 arithmetic-average Asian, 2 knock-out barrier -- path functionals in registers
template <typename real, int PPT, bool MULTI, bool PIPE = true, int MINB = 1, int ROUNDS = 10, int PAYOFF = 0,
          bool XCHG = false>
__global__ void __launch_bounds__(kMaxThreads, MINB)
heston_fused_kernel(const __grid_constant__ FusedConsts<real> c, ReduceWorkspace ws,
                    real *__restrict__ x_out, real *__restrict__ v_out) {
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_iter = total * PPT;
    const uint64_t iters = (c.pair_count + per_iter - 1) / per_iter;   // same for every thread: warps stay converged
    const uint32_t lane = threadIdx.x & 31;

    __shared__ real s_strikes[kMaxStrikes];
    if (MULTI) stage_strikes(s_strikes, c.strikes, c.nK);
    PayoffAccumulator<real, MULTI> acc;
    const real x_shift = fma((real)c.steps, c.mu_dt, c.x0);
    const StepRegs<real> k(c);
    // 0x3F800000 in a register ptxas cannot constant-fold (steps < 2^31 always): keeps angle_float at one LOP3
    const uint32_t one_bits = 0x3F800000u | (c.steps >> 31);
    PhiloxKeys keys;   // round keys in registers: ptxas otherwise re-loads them from the constant bank every call
#pragma unroll
    for (int r = 0; r < 10; ++r) { keys.k0[r] = pin(c.keys.k0[r]); keys.k1[r] = pin(c.keys.k1[r]); }

    for (uint64_t it = 0; it < iters; ++it) {
        uint64_t local[PPT];
        bool active[PPT], paired[PPT];
        PairStream rng[PPT];
        real x1[PPT], v1[PPT], x2[PPT], v2[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            local[q] = (it * PPT + q) * total + gtid;      // a warp owns 32 consecutive pairs: coalesced rows
            active[q] = local[q] < c.pair_count;
            const uint64_t gp = c.pair_first + (active[q] ? local[q] : 0);   // idle lanes replay pair 0, masked below
            paired[q] = active[q] && gp < c.partnered;
            rng[q].init((uint32_t)gp, (uint32_t)(gp >> 32), c.subsequence, c.keys);
            x1[q] = x2[q] = (real)0;
            v1[q] = v2[q] = k.u0;            // v1/v2 hold u = c1*v inside the time loop
        }
        // last grid-stride round only: a warp that lies wholly past the end of the shard leaves its issue slots to the
        // warps that still have pairs (local[0] is the smallest index of the thread)
        if (!__any_sync(0xffffffffu, active[0])) continue;
        // one Euler step for every pair of this thread from one (radius, angle) word pair each
        // path functionals of the path-dependent payoffs: sum of S_t (Asian) or running extreme of x_t (barrier)
        real f1[PPT], f2[PPT];
        real run_shift = c.x0;                       // x0 + t*mu*dt: the drift the loop keeps out of x1/x2
#pragma unroll
        for (int q = 0; q < PPT; ++q) f1[q] = f2[q] = (PAYOFF == 2) ? (real)-3.0e38 : (real)0;
        auto advance = [&](const StepDraw (&d)[PPT]) {
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                euler_pair(x1[q], v1[q], x2[q], v2[q], d[q].t_radius, d[q].t_angle, k);
            }
            if constexpr (PAYOFF != 0) {
                run_shift += c.mu_dt;
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    if constexpr (PAYOFF == 1) {
                        f1[q] += exp_real(x1[q] + run_shift);
                        f2[q] += exp_real(x2[q] + run_shift);
                    } else {
                        f1[q] = plain_max(f1[q], c.barrier_sign * (x1[q] + run_shift));
                        f2[q] = plain_max(f2[q], c.barrier_sign * (x2[q] + run_shift));
                    }
                }
            }
        };

        // Three Euler steps per Philox call (split_draws): ceil(steps/3) calls per pair.
        const uint32_t calls = c.steps / 3, rest = c.steps - 3 * calls;
        if constexpr (PIPE) {
            // software pipeline: the Philox call of block b+1 (a 10-round dependent integer chain) is issued
            // next to the MUFU-heavy Box-Muller/Euler work of block b, so every warp feeds the XU queue and the
            // integer pipes at the same time instead of alternating between the two phases
            uint4 cur[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) cur[q] = rng[q].template block<ROUNDS>(0u, keys);
            uint64_t m1b = 0;                                   // M1 * (index of the block being prefetched)
#pragma unroll((MINB == 1 || MINB == 3 || sizeof(real) == 8) ? 2 : 1)
            for (uint32_t b = 0; b < calls; ++b) {
                uint4 nxt[PPT];
                StepDraw d0[PPT], d1[PPT], d2[PPT];
                m1b += kPhiloxM1;
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    nxt[q] = rng[q].template block_from_product<ROUNDS>(m1b, keys);
                    split_draws(cur[q], one_bits, d0[q], d1[q], d2[q]);
                }
                advance(d0);
                advance(d1);
                advance(d2);
#pragma unroll
                for (int q = 0; q < PPT; ++q) cur[q] = nxt[q];
            }
            if (rest) {
                StepDraw d0[PPT], d1[PPT], d2[PPT];
#pragma unroll
                for (int q = 0; q < PPT; ++q) split_draws(cur[q], one_bits, d0[q], d1[q], d2[q]);
                advance(d0);
                if (rest > 1) advance(d1);
            }
        } else {
            for (uint32_t b = 0; b < calls + (rest ? 1u : 0u); ++b) {
                StepDraw d0[PPT], d1[PPT], d2[PPT];
#pragma unroll
                for (int q = 0; q < PPT; ++q) split_draws(rng[q].template block<ROUNDS>(b, keys), one_bits, d0[q], d1[q], d2[q]);
                const uint32_t left = c.steps - 3 * b;
                advance(d0);
                if (left > 1) advance(d1);
                if (left > 2) advance(d2);
            }
        }

        // ---- terminal values and payoffs ----
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const real xT1 = x1[q] + x_shift, xT2 = x2[q] + x_shift;
            if (x_out != nullptr) {
                // u = c1*v back to v; a path sitting on the floor reports exactly eps_f32, as the reference's maximum() does
                const real vT1 = v1[q] == k.floor_u ? c.floor_v : v1[q] * k.inv_c1;
                const real vT2 = v2[q] == k.floor_u ? c.floor_v : v2[q] * k.inv_c1;
                if (active[q]) { x_out[local[q]] = xT1; v_out[local[q]] = vT1; }
                if (paired[q]) { x_out[c.pair_count + local[q]] = xT2; v_out[c.pair_count + local[q]] = vT2; }
            }
            real S1 = exp_real(xT1), S2 = exp_real(xT2);
            if constexpr (PAYOFF == 1) {                      // average of the N-1 monitored prices
                const real inv_n = (real)1 / (real)c.steps;
                S1 = f1[q] * inv_n;
                S2 = f2[q] * inv_n;
            } else if constexpr (PAYOFF == 2) {               // knocked out: a price that makes the payoff vanish
                const real dead = c.payoff_sign > (real)0 ? (real)0 : (real)3.0e38;
                if (f1[q] >= c.barrier_log) S1 = dead;
                if (f2[q] >= c.barrier_log) S2 = dead;
            }
            acc.add(S1, S2, active[q], paired[q], c.strikes[0], s_strikes, lane, c.payoff_sign);
        }
    }
    acc.template publish<XCHG>(c.nK, ws);
}

// ---- host side ---------------------------------------------------------------

// float strike sweep, one pair per thread: 85 registers uncapped = two resident 256-thread blocks; capped at 80 = three
template <typename real>
constexpr int kSweepMinBlocks = sizeof(real) == 4 ? 3 : 1;

// double, two pairs per thread: capped at 128 registers (two 256-thread blocks' worth) four 128-thread blocks are
// resident instead of three (+2% for one strike, +4% for the sweep); the strike sweep fits under the cap since its
// positive parts are d + |d| (122 registers, no spills; it takes 152 uncapped)
template <typename real>
constexpr int kTwoPairMinBlocks = sizeof(real) == 8 ? 2 : 1;

// pairs per thread of the kernels instantiated with the fused cross-GPU exchange = the host's default per type
template <typename real>
constexpr int kExchangePairs = sizeof(real) == 8 ? 2 : 1;

template <typename real, int PPT, bool MULTI>
constexpr int default_min_blocks() {
    if (PPT == 1 && MULTI) return kSweepMinBlocks<real>;
    if (PPT == 2) return kTwoPairMinBlocks<real>;
    return 1;
}

// The one table of instantiations: launch and occupancy query both go through it, so the grid is always sized for the
// kernel that actually runs.  nullptr = no such instantiation.
//   payoff 0 (European): every geometry without the exchange; with it, the default geometry of the type
//   payoff 1/2 (Asian / knock-out): one pair per thread; the exchange code is always compiled in and armed at run time
//   variant 1/2: tuning variants of the float single-strike kernel (cb_ctx_set_variant), never with the exchange
template <typename real, int PPT, bool MULTI>
static const void *plain_fn(bool xchg) {
    constexpr int MINB = default_min_blocks<real, PPT, MULTI>();
    if (!xchg) return (const void *)heston_fused_kernel<real, PPT, MULTI, true, MINB, 10, 0, false>;
    if constexpr (PPT == kExchangePairs<real>)
        return (const void *)heston_fused_kernel<real, PPT, MULTI, true, MINB, 10, 0, true>;
    return nullptr;
}

template <int PPT>
static const void *variant_fn(int variant) {
    switch (variant) {
        case 1: return (const void *)heston_fused_kernel<float, PPT, false, false, 1>;          // not software-pipelined
        case 2: return (const void *)heston_fused_kernel<float, PPT, false, true, (PPT == 4 ? 2 : 4)>;
        default: return nullptr;
    }
}

template <typename real>
static const void *fused_fn(int ppt, bool multi, int payoff, int variant, bool xchg) {
    if (payoff != 0) {
        if (ppt != 1 || variant != 0) return nullptr;
        if (payoff == 1)
            return multi ? (const void *)heston_fused_kernel<real, 1, true, true, 1, 10, 1, true>
                         : (const void *)heston_fused_kernel<real, 1, false, true, 1, 10, 1, true>;
        return multi ? (const void *)heston_fused_kernel<real, 1, true, true, 1, 10, 2, true>
                     : (const void *)heston_fused_kernel<real, 1, false, true, 1, 10, 2, true>;
    }
    if (variant != 0) {
        if constexpr (sizeof(real) == 4) {
            if (multi || xchg) return nullptr;
            return ppt == 1 ? variant_fn<1>(variant) : ppt == 2 ? variant_fn<2>(variant) : ppt == 4 ? variant_fn<4>(variant) : nullptr;
        }
        return nullptr;
    }
    switch (ppt) {
        case 1: return multi ? plain_fn<real, 1, true>(xchg) : plain_fn<real, 1, false>(xchg);
        case 2: return multi ? plain_fn<real, 2, true>(xchg) : plain_fn<real, 2, false>(xchg);
        case 4: return multi ? plain_fn<real, 4, true>(xchg) : plain_fn<real, 4, false>(xchg);
        default: return nullptr;
    }
}

template <typename real>
bool fused_exchange_supported(const LaunchShape &s, int payoff_kind) {
    return fused_fn<real>(s.pairs_per_thread, false, payoff_kind, s.variant, true) != nullptr;
}
template bool fused_exchange_supported<float>(const LaunchShape &, int);
template bool fused_exchange_supported<double>(const LaunchShape &, int);

template <typename real>
cudaError_t launch_heston_fused(const FusedConsts<real> &c, const LaunchShape &s, ReduceWorkspace ws,
                                real *d_x, real *d_v, cudaStream_t stream) {
    const void *fn = fused_fn<real>(s.pairs_per_thread, c.nK > 1, c.payoff_kind, s.variant, ws.px != nullptr);
    if (fn == nullptr) return cudaErrorInvalidValue;
    void *args[] = {(void *)&c, (void *)&ws, (void *)&d_x, (void *)&d_v};
    return cudaLaunchKernel(fn, dim3(s.blocks), dim3(s.threads), args, 0, stream);
}

template <typename real>
cudaError_t fused_occupancy(int threads, int ppt, bool multi, int payoff, int variant, bool xchg, int *blocks_per_sm,
                            int *regs) {
    const void *fn = fused_fn<real>(ppt, multi, payoff, variant, xchg);
    if (fn == nullptr) return cudaErrorInvalidValue;
    cudaFuncAttributes attr;
    cudaError_t e = cudaFuncGetAttributes(&attr, fn);
    if (e != cudaSuccess) return e;
    if (regs) *regs = attr.numRegs;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, threads, 0);
}

template cudaError_t launch_heston_fused<float>(const FusedConsts<float> &, const LaunchShape &, ReduceWorkspace,
                                                float *, float *, cudaStream_t);
template cudaError_t launch_heston_fused<double>(const FusedConsts<double> &, const LaunchShape &, ReduceWorkspace,
                                                 double *, double *, cudaStream_t);
template cudaError_t fused_occupancy<float>(int, int, bool, int, int, bool, int *, int *);
template cudaError_t fused_occupancy<double>(int, int, bool, int, int, bool, int *, int *);

}  // namespace cb
