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
// HBM inside the time loop unless the trajectory is requested.  Per pair-step:
// 5 MUFU (lg2, sin, cos, 2x sqrt(v*r^2)) against ~43 other issue slots: the MUFU (XU)
// pipe and the issue port are both close to their limits.
//
// Algebra (statistical mode, not the parity order): with a = 1-kappa*dt,
// b = kappa*v_bar*dt, w = sigma_v*(rho*dB0 + sqrt(1-rho^2)*dB1):
//   x' = x - 0.5*dt*v + sqrt(v)*dB0          (mu*dt*(N-1) + x0 added once at the end)
//   v' = max(a*v + b + sqrt(v)*w, eps_f32)
#include "kernels.h"
#include "payoff_reduce.cuh"

namespace cb {

constexpr int kMaxThreads = kFusedMaxThreads;

// ---- per-type math ---------------------------------------------------------

__device__ __forceinline__ float root(float v) { return mufu_sqrt(v); }
// double sqrt without the IEEE fix-up path: MUFU.RSQ64H seed (~2^-22) + ONE Newton step on 1/sqrt folded into the
// product,  g = p y,  e = 1 - g y,  sqrt p ~ g (1 + e/2)  =>  relative error 3 e^2/8 ~ 2^-45 -- seven orders below the
// float accuracy of the normals that multiply it (the parity kernel keeps __dsqrt_rn).  FP64 instructions cost two
// issue cycles each on B200, so the step count matters: 4 FP64 + 1 MUFU, no F2F (conversions share the XU pipe).
// p > 0 always: euler_pair keeps lg2 U away from zero.
__device__ __forceinline__ double root(double p) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    const double g = p * y;
    const double e = fma(-g, y, 1.0);
    return fma(g * 0.5, e, g);
}
// min without the NaN bookkeeping of fmin(double) (DSETP + 2 SEL instead of 6 instructions)
__device__ __forceinline__ float floor_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double floor_min(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ float exp_real(float x) { return mufu_ex2(x * 1.4426950408889634f); }
__device__ __forceinline__ double exp_real(double x) { return exp(x); }

// Launch constants the time loop reads every step.  The variance is carried as u = c1*v with
// c1 = -2 ln2 dt < 0, so that v*r^2 = v*(c1*lg2 U) is ONE multiply (u * lg2 U) per path and step.
template <typename real>
struct StepRegs {
    real drift_u;    // (-0.5 dt)/c1          x += drift_u*u + q*cos
    real decay;      // 1 - kappa dt          u' = min(decay*u + pull_u + q*g, floor_u)
    real pull_u;     // c1 * kappa v_bar dt
    float g0f, g1f;  // c1 * sigma_v * (rho, sqrt(1-rho^2))     g = g0*cos + g1*sin
    real floor_u;    // c1 * eps_f32 (u <= floor_u  <=>  v >= eps)
    real u0;         // c1 * v0
    real inv_c1;     // v = u * inv_c1
    float two_pi;
    __device__ __forceinline__ explicit StepRegs(const FusedConsts<real> &c) {
        const real c1 = (real)c.neg2ln2_dt;
        drift_u = c.drift_v / c1;
        decay = c.decay;
        pull_u = c1 * c.pull;
        g0f = (float)(c1 * c.w0);
        g1f = (float)(c1 * c.w1);
        floor_u = c1 * c.floor_v;
        u0 = c1 * c.v0;
        inv_c1 = (real)1 / c1;
        two_pi = kTwoPi;
    }
};

// The draws of one Euler step: two floats in [1,2) built from random mantissa bits (no I2F: it shares the XU pipe).
struct StepDraw {
    float t_radius;   // 23 random bits: U = 2 - t_radius in (0,1]
    float t_angle;    // 19 random bits: theta = 2 pi t_angle - 3 pi
};

// One Philox4x32-10 block (128 bits) feeds THREE Euler steps (126 bits used): the 32x32->64 multiplies of Philox
// (IMAD.WIDE: 4 issue cycles each, during which the sub-partition issues nothing else) are the most expensive
// instructions of the loop, so the block is used to the last bit.
//   step j (j = 0,1,2): radius <- w_j[31:9];  angle <- w_j[8:0] : w3[31-10j : 22-10j]
// (x & 0x007FFFF0) | one_bits as ONE LOP3: with two immediates ptxas emits two, so the exponent pattern
// 0x3F800000 travels in a register whose value ptxas cannot fold (see StepRegs::one_bits).
__device__ __forceinline__ float angle_float(uint32_t x, uint32_t one_bits) {
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x007FFFF0, %2, 0xEA;" : "=r"(r) : "r"(x), "r"(one_bits));   // (a & b) | c
    return __uint_as_float(r);
}

__device__ __forceinline__ void split_draws(const uint4 &o, uint32_t one_bits, StepDraw &d0, StepDraw &d1,
                                            StepDraw &d2) {
    d0.t_radius = mantissa_float(o.x);
    d1.t_radius = mantissa_float(o.y);
    d2.t_radius = mantissa_float(o.z);
    // funnel shift: upper word of (w_j : w3 << 10j) << 14 has w_j[8:0] : w3 bits in its low 23 bits
    d0.t_angle = angle_float(__funnelshift_l(o.w, o.x, 14), one_bits);
    d1.t_angle = angle_float(__funnelshift_l(o.w << 10, o.y, 14), one_bits);
    d2.t_angle = angle_float(__funnelshift_l(o.w << 20, o.z, 14), one_bits);
}

// One Euler step of an antithetic pair from one (radius, angle) draw.
//
// With r^2 = -2 dt ln U and (c, s) = (cos, sin)(theta) the reference needs sqrt(v)*dB0 = sqrt(v)*r*c and
// sqrt(v)*(rho dB0 + rho' dB1) = sqrt(v)*r*(rho c + rho' s).  Since sqrt(v)*r = sqrt(v*r^2), the radius never has to
// be formed: q = MUFU.SQRT(v * r^2) replaces MUFU.SQRT(r^2) and MUFU.SQRT(v) -- 5 MUFU per pair-step instead of 6.
// Per pair-step: 5 MUFU + 15 FP32 + 2 FMNMX.
template <typename real>
__device__ __forceinline__ void euler_pair(real &x1, real &u1, real &x2, real &u2, float t_radius, float t_angle,
                                           const StepRegs<real> &k) {
    float L_f = mufu_lg2(2.0f - t_radius);                                // lg2 U <= 0, U in (0,1]
    if constexpr (sizeof(real) == 8) L_f -= 1.17549435e-38f;              // U = 1: keeps the rsqrt seed of root() finite;
                                                                          // every other lg2 U (|.| >= 2^-24) is unchanged
    const real L = (real)L_f;
    const float theta = fmaf(t_angle, k.two_pi, -kThreePi);               // [-pi,pi)
    const float cs_f = mufu_cos(theta), sn_f = mufu_sin(theta);
    const real cs = (real)cs_f;
    const real g = (real)fmaf(cs_f, k.g0f, sn_f * k.g1f);                 // c1*sigma_v*(rho c + rho' s), float accuracy
                                                                          // is all the MUFU inputs have anyway
    const real q1 = root(u1 * L), q2 = root(u2 * L);                      // sqrt(v*r^2) = sqrt(v)*r*sqrt(dt)
    x1 = fma(q1, cs, fma(u1, k.drift_u, x1));
    x2 = fma(q2, -cs, fma(u2, k.drift_u, x2));
    u1 = floor_min(fma(q1, g, fma(u1, k.decay, k.pull_u)), k.floor_u);
    u2 = floor_min(fma(q2, -g, fma(u2, k.decay, k.pull_u)), k.floor_u);
}

// ---- the kernel -------------------------------------------------------------

// PAYOFF: 0 terminal (European), 1 arithmetic-average Asian, 2 knock-out barrier -- path functionals in registers
template <typename real, int PPT, bool MULTI, bool TRAJ, bool PIPE = true, int MINB = 1, int ROUNDS = 10, int PAYOFF = 0,
          bool XCHG = false>
__global__ void __launch_bounds__(kMaxThreads, MINB)
heston_fused_kernel(const __grid_constant__ FusedConsts<real> c, ReduceWorkspace ws,
                    real *__restrict__ x_out, real *__restrict__ v_out, real *__restrict__ traj) {
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_iter = total * PPT;
    const uint64_t iters = (c.pair_count + per_iter - 1) / per_iter;   // same for every thread: warps stay converged
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t row_len = 2 * c.pair_count;

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
        // trajectory mode: this thread's slots in the current row [first halves | partners]; a warp writes 128
        // contiguous bytes per half and step (predicated at the ragged edge), rows advance by row_len per step
        real *row1[PPT], *row2[PPT];
        real step_f = (real)0;
        if constexpr (TRAJ) {   // row 0: the initial log price
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                row1[q] = traj + local[q];
                row2[q] = traj + c.pair_count + local[q];
                if (active[q]) *row1[q] = c.x0;
                if (paired[q]) *row2[q] = c.x0;
            }
        }

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
                        f1[q] = fmax(f1[q], c.barrier_sign * (x1[q] + run_shift));
                        f2[q] = fmax(f2[q], c.barrier_sign * (x2[q] + run_shift));
                    }
                }
            }
            if constexpr (TRAJ) {
                step_f += (real)1;                                   // exact: steps < 2^24
                const real shift = fma(step_f, c.mu_dt, c.x0);       // the drift the loop keeps out of x1/x2
#pragma unroll
                for (int q = 0; q < PPT; ++q) {
                    row1[q] += row_len;
                    row2[q] += row_len;
                    if (active[q]) *row1[q] = x1[q] + shift;
                    if (paired[q]) *row2[q] = x2[q] + shift;
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
#pragma unroll((MINB == 1 || MINB == 3 || TRAJ || sizeof(real) == 8) ? 2 : 1)
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

// trajectory mode: capping registers for a third resident 256-thread block costs more re-materialisation (LDC, LEA)
// than the occupancy returns (374 vs 311 loop instructions): no cap, the host picks 128-thread blocks instead
template <typename real>
constexpr int kTrajMinBlocks = 1;

// float strike sweep, one pair per thread: 85 registers uncapped = two resident 256-thread blocks; capped at 80 = three
template <typename real>
constexpr int kSweepMinBlocks = sizeof(real) == 4 ? 3 : 1;

// double, two pairs per thread, one strike: capped at 128 registers (two 256-thread blocks' worth) four 128-thread
// blocks are resident instead of three (+2%); the strike sweep spills under the same cap and stays uncapped
template <typename real>
constexpr int kTwoPairMinBlocks = sizeof(real) == 8 ? 2 : 1;

template <typename real, int PPT, bool MULTI, bool TRAJ, bool PIPE = true, int MINB = 1>
static cudaError_t launch_one(const FusedConsts<real> &c, const LaunchShape &s, ReduceWorkspace ws,
                              real *d_x, real *d_v, real *d_traj, cudaStream_t stream) {
    heston_fused_kernel<real, PPT, MULTI, TRAJ, PIPE, MINB><<<s.blocks, s.threads, 0, stream>>>(c, ws, d_x, d_v, d_traj);
    return cudaGetLastError();
}

// tuning variants of the float single-strike kernel (variant = 1, 2), selected by cb_ctx_set_variant; all run the
// full Philox4x32-10
template <int PPT>
static const void *variant_fn(int variant) {
    switch (variant) {
        case 1: return (const void *)heston_fused_kernel<float, PPT, false, false, false, 1>;          // not software-pipelined
        case 2: return (const void *)heston_fused_kernel<float, PPT, false, false, true, (PPT == 4 ? 2 : 4)>;
        default: return (const void *)heston_fused_kernel<float, PPT, false, false, true, 1>;
    }
}
static const void *variant_fn(int ppt, int variant) {
    return ppt == 1 ? variant_fn<1>(variant) : ppt == 2 ? variant_fn<2>(variant) : variant_fn<4>(variant);
}

// pairs per thread of the kernels instantiated with the fused cross-GPU exchange = the host's default per type
template <typename real>
constexpr int kExchangePairs = sizeof(real) == 8 ? 2 : 1;

template <typename real>
bool fused_exchange_supported(const LaunchShape &s, int payoff_kind, bool traj) {
    return payoff_kind == 0 && !traj && s.variant == 0 && s.pairs_per_thread == kExchangePairs<real>;
}
template bool fused_exchange_supported<float>(const LaunchShape &, int, bool);
template bool fused_exchange_supported<double>(const LaunchShape &, int, bool);

template <typename real>
cudaError_t launch_heston_fused(const FusedConsts<real> &c, const LaunchShape &s, ReduceWorkspace ws,
                                real *d_x, real *d_v, real *d_traj, cudaStream_t stream) {
    const bool multi = c.nK > 1;
    if (c.payoff_kind != 0) {
        if (d_traj != nullptr) return cudaErrorInvalidValue;
        void *args[] = {(void *)&c, (void *)&ws, (void *)&d_x, (void *)&d_v, (void *)&d_traj};
        const void *fn = c.payoff_kind == 1
            ? (multi ? (const void *)heston_fused_kernel<real, 1, true, false, true, 1, 10, 1>
                     : (const void *)heston_fused_kernel<real, 1, false, false, true, 1, 10, 1>)
            : (multi ? (const void *)heston_fused_kernel<real, 1, true, false, true, 1, 10, 2>
                     : (const void *)heston_fused_kernel<real, 1, false, false, true, 1, 10, 2>);
        return cudaLaunchKernel(fn, dim3(s.blocks), dim3(s.threads), args, 0, stream);
    }
    if (ws.px != nullptr) {
        // the instantiations that carry the fused cross-GPU step: the default geometry of each type (capi.cu asks
        // fused_exchange_supported first and keeps the NCCL allreduce for everything else)
        if (!fused_exchange_supported<real>(s, c.payoff_kind, d_traj != nullptr)) return cudaErrorInvalidValue;
        void *args[] = {(void *)&c, (void *)&ws, (void *)&d_x, (void *)&d_v, (void *)&d_traj};
        constexpr int P = kExchangePairs<real>;
        const void *fn = multi
            ? (const void *)heston_fused_kernel<real, P, true, false, true, (P == 1 ? kSweepMinBlocks<real> : 1), 10, 0, true>
            : (const void *)heston_fused_kernel<real, P, false, false, true, (P == 2 ? kTwoPairMinBlocks<real> : 1), 10, 0, true>;
        return cudaLaunchKernel(fn, dim3(s.blocks), dim3(s.threads), args, 0, stream);
    }
    if (d_traj != nullptr)
        return multi ? launch_one<real, 1, true, true, true, kTrajMinBlocks<real>>(c, s, ws, d_x, d_v, d_traj, stream)
                     : launch_one<real, 1, false, true, true, kTrajMinBlocks<real>>(c, s, ws, d_x, d_v, d_traj, stream);
    if constexpr (sizeof(real) == 4) {
        if (!multi && s.variant > 0) {
            void *args[] = {(void *)&c, (void *)&ws, (void *)&d_x, (void *)&d_v, (void *)&d_traj};
            return cudaLaunchKernel(variant_fn(s.pairs_per_thread, s.variant), dim3(s.blocks), dim3(s.threads), args, 0,
                                    stream);
        }
    }
    switch (s.pairs_per_thread) {
        case 1:
            return multi ? launch_one<real, 1, true, false, true, kSweepMinBlocks<real>>(c, s, ws, d_x, d_v, nullptr, stream)
                         : launch_one<real, 1, false, false>(c, s, ws, d_x, d_v, nullptr, stream);
        case 2:
            return multi ? launch_one<real, 2, true, false>(c, s, ws, d_x, d_v, nullptr, stream)
                         : launch_one<real, 2, false, false, true, kTwoPairMinBlocks<real>>(c, s, ws, d_x, d_v, nullptr, stream);
        case 4:
            return multi ? launch_one<real, 4, true, false>(c, s, ws, d_x, d_v, nullptr, stream)
                         : launch_one<real, 4, false, false>(c, s, ws, d_x, d_v, nullptr, stream);
        default:
            return cudaErrorInvalidValue;
    }
}

template <typename real>
cudaError_t fused_occupancy(int threads, int ppt, bool multi, int variant, int *blocks_per_sm, int *regs) {
    const void *fn = nullptr;
    if (sizeof(real) == 4 && !multi && variant > 0 && ppt > 0) {
        fn = variant_fn(ppt, variant);
        cudaFuncAttributes a;
        cudaError_t e0 = cudaFuncGetAttributes(&a, fn);
        if (e0 != cudaSuccess) return e0;
        if (regs) *regs = a.numRegs;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, threads, 0);
    }
    if (ppt == 1) fn = multi ? (const void *)heston_fused_kernel<real, 1, true, false, true, kSweepMinBlocks<real>> : (const void *)heston_fused_kernel<real, 1, false, false>;
    else if (ppt == 2) fn = multi ? (const void *)heston_fused_kernel<real, 2, true, false> : (const void *)heston_fused_kernel<real, 2, false, false, true, kTwoPairMinBlocks<real>>;
    else if (ppt == 4) fn = multi ? (const void *)heston_fused_kernel<real, 4, true, false> : (const void *)heston_fused_kernel<real, 4, false, false>;
    else if (ppt == -1)   // trajectory mode
        fn = multi ? (const void *)heston_fused_kernel<real, 1, true, true, true, kTrajMinBlocks<real>>
                   : (const void *)heston_fused_kernel<real, 1, false, true, true, kTrajMinBlocks<real>>;
    else return cudaErrorInvalidValue;
    cudaFuncAttributes attr;
    cudaError_t e = cudaFuncGetAttributes(&attr, fn);
    if (e != cudaSuccess) return e;
    if (regs) *regs = attr.numRegs;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, threads, 0);
}

template cudaError_t launch_heston_fused<float>(const FusedConsts<float> &, const LaunchShape &, ReduceWorkspace,
                                                float *, float *, float *, cudaStream_t);
template cudaError_t launch_heston_fused<double>(const FusedConsts<double> &, const LaunchShape &, ReduceWorkspace,
                                                 double *, double *, double *, cudaStream_t);
template cudaError_t fused_occupancy<float>(int, int, bool, int, int *, int *);
template cudaError_t fused_occupancy<double>(int, int, bool, int, int *, int *);

}  // namespace cb
