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
// 6 MUFU (lg2, sqrt, sin, cos, 2x sqrt(v)) against ~41 other issue slots, so the
// kernel is bound by the MUFU pipe with the issue port close behind.
//
// Algebra (statistical mode, not the parity order): with a = 1-kappa*dt,
// b = kappa*v_bar*dt, w = sigma_v*(rho*dB0 + sqrt(1-rho^2)*dB1):
//   x' = x - 0.5*dt*v + sqrt(v)*dB0          (mu*dt*(N-1) + x0 added once at the end)
//   v' = max(a*v + b + sqrt(v)*w, eps_f32)
#include "kernels.h"

namespace cb {

constexpr int kMaxThreads = 256;
constexpr int kMaxWarps = kMaxThreads / 32;

// ---- per-type math ---------------------------------------------------------

__device__ __forceinline__ float root(float v) { return mufu_sqrt(v); }
__device__ __forceinline__ double root(double v) { return sqrt(v); }
__device__ __forceinline__ float exp_real(float x) { return mufu_ex2(x * 1.4426950408889634f); }
__device__ __forceinline__ double exp_real(double x) { return exp(x); }

// launch constants the time loop reads every step, held in registers (see pin())
template <typename real>
struct StepRegs {
    real drift_v, decay, pull, w0, w1, floor_v;
    float neg2ln2_dt, two_pi;
    __device__ __forceinline__ explicit StepRegs(const FusedConsts<real> &c)
        : drift_v(pin(c.drift_v)), decay(pin(c.decay)), pull(pin(c.pull)), w0(pin(c.w0)), w1(pin(c.w1)),
          floor_v(c.floor_v), neg2ln2_dt(c.neg2ln2_dt), two_pi(pin(kTwoPi)) {}
};

template <typename real>
__device__ __forceinline__ void euler_pair(real &x1, real &v1, real &x2, real &v2, real b0, real w,
                                           const StepRegs<real> &k) {
    const real s1 = root(v1), s2 = root(v2);
    x1 = fma(s1, b0, fma(v1, k.drift_v, x1));
    x2 = fma(s2, -b0, fma(v2, k.drift_v, x2));
    v1 = fmax(fma(s1, w, fma(v1, k.decay, k.pull)), k.floor_v);
    v2 = fmax(fma(s2, -w, fma(v2, k.decay, k.pull)), k.floor_v);
}

template <int N>
struct PairStore;   // vector store of N consecutive reals
template <>
struct PairStore<4> {
    static __device__ __forceinline__ void put(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void put(double *p, const double (&v)[4]) {
        reinterpret_cast<double2 *>(p)[0] = make_double2(v[0], v[1]);
        reinterpret_cast<double2 *>(p)[1] = make_double2(v[2], v[3]);
    }
};

// ---- the kernel -------------------------------------------------------------

template <typename real, int PPT, bool MULTI, bool TRAJ>
__global__ void __launch_bounds__(kMaxThreads)
heston_fused_kernel(const __grid_constant__ FusedConsts<real> c, ReduceWorkspace ws,
                    real *__restrict__ x_out, real *__restrict__ v_out, real *__restrict__ traj) {
    static_assert(!TRAJ || PPT == 4, "trajectory mode stores 4 consecutive pairs per thread");
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_iter = total * PPT;
    const uint64_t iters = (c.pair_count + per_iter - 1) / per_iter;   // same for every thread: warps stay converged
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t blocks2 = c.steps >> 1;
    const uint64_t row_len = 2 * c.pair_count;
    const bool vec_ok = TRAJ && (c.pair_count % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(traj) & (4 * sizeof(real) - 1)) == 0);

    __shared__ real s_strikes[kMaxStrikes];
    if (MULTI) {
        if (threadIdx.x < kMaxStrikes) s_strikes[threadIdx.x] = c.strikes[threadIdx.x < c.nK ? threadIdx.x : 0];
        __syncthreads();
    }

    double acc_sum[MULTI ? 2 : 1] = {};
    double acc_sq[MULTI ? 2 : 1] = {};
    const real x_shift = c.x0 + (real)c.steps * c.mu_dt;
    const StepRegs<real> k(c);
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
            local[q] = TRAJ ? (it * total + gtid) * PPT + q : (it * PPT + q) * total + gtid;
            active[q] = local[q] < c.pair_count;
            const uint64_t gp = c.pair_first + (active[q] ? local[q] : 0);   // idle lanes replay pair 0, masked below
            paired[q] = active[q] && gp < c.partnered;
            rng[q].init((uint32_t)gp, (uint32_t)(gp >> 32), c.subsequence, c.keys);
            x1[q] = x2[q] = (real)0;
            v1[q] = v2[q] = c.v0;
        }
        if constexpr (TRAJ) {   // row 0: the initial log price
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                if (active[q]) traj[local[q]] = c.x0;
                if (paired[q]) traj[c.pair_count + local[q]] = c.x0;
            }
        }

        // one Euler step for every pair of this thread from one (radius, angle) word pair each
        auto advance = [&](const uint32_t (&o_r)[PPT], const uint32_t (&o_a)[PPT], uint32_t step_no) {
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                float z0, z1;
                box_muller(o_r[q], o_a[q], k.neg2ln2_dt, k.two_pi, z0, z1);      // already scaled by sqrt(dt)
                const real b0 = (real)z0, b1 = (real)z1;
                const real w = fma(b0, k.w0, b1 * k.w1);
                euler_pair(x1[q], v1[q], x2[q], v2[q], b0, w, k);
            }
            if constexpr (TRAJ) {
                const real shift = c.x0 + (real)step_no * c.mu_dt;
                real *row = traj + (uint64_t)step_no * row_len;
                if (vec_ok && active[PPT - 1]) {
                    real a[PPT], b[PPT];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) { a[q] = x1[q] + shift; b[q] = x2[q] + shift; }
                    PairStore<PPT>::put(row + local[0], a);
                    if (paired[PPT - 1]) {
                        PairStore<PPT>::put(row + c.pair_count + local[0], b);
                    } else {
#pragma unroll
                        for (int q = 0; q < PPT; ++q) if (paired[q]) row[c.pair_count + local[q]] = b[q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        if (active[q]) row[local[q]] = x1[q] + shift;
                        if (paired[q]) row[c.pair_count + local[q]] = x2[q] + shift;
                    }
                }
            }
        };

#pragma unroll 2
        for (uint32_t b = 0; b < blocks2; ++b) {
            uint32_t r0[PPT], a0[PPT], r1[PPT], a1[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const uint4 o = rng[q].block(b, keys);
                r0[q] = o.x; a0[q] = o.y; r1[q] = o.z; a1[q] = o.w;
            }
            advance(r0, a0, 2 * b + 1);
            advance(r1, a1, 2 * b + 2);
        }
        if (c.steps & 1u) {
            uint32_t r0[PPT], a0[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const uint4 o = rng[q].block(blocks2, keys);
                r0[q] = o.x; a0[q] = o.y;
            }
            advance(r0, a0, c.steps);
        }

        // ---- terminal values and payoffs ----
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const real xT1 = x1[q] + x_shift, xT2 = x2[q] + x_shift;
            if (x_out != nullptr) {
                if (active[q]) { x_out[local[q]] = xT1; v_out[local[q]] = v1[q]; }
                if (paired[q]) { x_out[c.pair_count + local[q]] = xT2; v_out[c.pair_count + local[q]] = v2[q]; }
            }
            const real S1 = exp_real(xT1), S2 = exp_real(xT2);
            if constexpr (!MULTI) {
                const real K = c.strikes[0];
                const real p1 = active[q] ? fmax(S1 - K, (real)0) : (real)0;
                const real p2 = paired[q] ? fmax(S2 - K, (real)0) : (real)0;
                const double tot = (double)(p1 + p2);
                const double unit = paired[q] ? 0.5 * tot : tot;
                acc_sum[0] += tot;
                acc_sq[0] = fma(unit, unit, acc_sq[0]);
            } else {
                // lane-transposed sweep: lane l keeps strikes l and l+32 for the whole warp
                const real m1 = active[q] ? (real)1 : (real)0;
                const real m2 = paired[q] ? (real)1 : (real)0;
                for (int src = 0; src < 32; ++src) {
                    const real s1 = __shfl_sync(0xffffffffu, S1, src);
                    const real s2 = __shfl_sync(0xffffffffu, S2, src);
                    const real g1 = __shfl_sync(0xffffffffu, m1, src);
                    const real g2 = __shfl_sync(0xffffffffu, m2, src);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const real K = s_strikes[lane + 32 * h];
                        const real p1 = g1 * fmax(s1 - K, (real)0);
                        const real p2 = g2 * fmax(s2 - K, (real)0);
                        const double tot = (double)(p1 + p2);
                        const double unit = tot * (0.5 + 0.5 * (double)(g1 - g2));   // 0.5*tot if paired, tot if single
                        acc_sum[h] += tot;
                        acc_sq[h] = fma(unit, unit, acc_sq[h]);
                    }
                }
            }
        }
    }

    // ---- block reduction -> per-block partials -> last block folds them in a fixed order ----
    const int nacc = 2 * c.nK;
    __shared__ double s_acc[kMaxWarps][kMaxAcc];
    __shared__ bool s_last;
    const int warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    if (!MULTI) {
        const double a = warp_sum(acc_sum[0]), b = warp_sum(acc_sq[0]);
        if (lane == 0) { s_acc[warp][0] = a; s_acc[warp][1] = b; }
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            if (k < c.nK) { s_acc[warp][k] = acc_sum[h]; s_acc[warp][c.nK + k] = acc_sq[h]; }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nacc) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += s_acc[w][threadIdx.x];
        ws.partials[(size_t)blockIdx.x * kMaxAcc + threadIdx.x] = t;
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
}

// ---- host side ---------------------------------------------------------------

template <typename real, int PPT, bool MULTI, bool TRAJ>
static cudaError_t launch_one(const FusedConsts<real> &c, const LaunchShape &s, ReduceWorkspace ws,
                              real *d_x, real *d_v, real *d_traj, cudaStream_t stream) {
    heston_fused_kernel<real, PPT, MULTI, TRAJ><<<s.blocks, s.threads, 0, stream>>>(c, ws, d_x, d_v, d_traj);
    return cudaGetLastError();
}

template <typename real>
cudaError_t launch_heston_fused(const FusedConsts<real> &c, const LaunchShape &s, ReduceWorkspace ws,
                                real *d_x, real *d_v, real *d_traj, cudaStream_t stream) {
    const bool multi = c.nK > 1;
    if (d_traj != nullptr) return launch_one<real, 4, true, true>(c, s, ws, d_x, d_v, d_traj, stream);
    switch (s.pairs_per_thread) {
        case 1:
            return multi ? launch_one<real, 1, true, false>(c, s, ws, d_x, d_v, nullptr, stream)
                         : launch_one<real, 1, false, false>(c, s, ws, d_x, d_v, nullptr, stream);
        case 2:
            return multi ? launch_one<real, 2, true, false>(c, s, ws, d_x, d_v, nullptr, stream)
                         : launch_one<real, 2, false, false>(c, s, ws, d_x, d_v, nullptr, stream);
        case 4:
            return multi ? launch_one<real, 4, true, false>(c, s, ws, d_x, d_v, nullptr, stream)
                         : launch_one<real, 4, false, false>(c, s, ws, d_x, d_v, nullptr, stream);
        default:
            return cudaErrorInvalidValue;
    }
}

template <typename real>
cudaError_t fused_occupancy(int threads, int ppt, bool multi, int *blocks_per_sm, int *regs) {
    const void *fn = nullptr;
    if (ppt == 1) fn = multi ? (const void *)heston_fused_kernel<real, 1, true, false> : (const void *)heston_fused_kernel<real, 1, false, false>;
    else if (ppt == 2) fn = multi ? (const void *)heston_fused_kernel<real, 2, true, false> : (const void *)heston_fused_kernel<real, 2, false, false>;
    else if (ppt == 4) fn = multi ? (const void *)heston_fused_kernel<real, 4, true, false> : (const void *)heston_fused_kernel<real, 4, false, false>;
    else if (ppt == -4) fn = (const void *)heston_fused_kernel<real, 4, true, true>;
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
template cudaError_t fused_occupancy<float>(int, int, bool, int *, int *);
template cudaError_t fused_occupancy<double>(int, int, bool, int *, int *);

}  // namespace cb
