// Parity kernels: the reference's Euler-Maruyama update with INJECTED normals, in the
// reference's operation order with one rounding per operation (no FMA contraction
// except the dot product, see below), so that terminal values can be compared with
// the reference's NumPy branch bit for bit.
//
//   simulate_heston_model          examples/stochastic_volatility/heston_utilities.py:58-96
//   randn_antithetic layout        cocos/numerics/random.py:190-216
//   compute_option_price_from_...  heston_utilities.py:99-119   (payoff_sums_kernel)
//
// HBM-bound: one thread owns an antithetic pair and streams its 8 (16) bytes of
// normals per step as one coalesced float2 (double2) load; nothing else touches HBM.
// Algorithmic traffic: 4 B (f32) / 8 B (f64) per path-step.
#include "kernels.h"

namespace cb {

// np.dot((R,2),(2,)) in the reference is OpenBLAS sgemv/dgemv, which rounds as
// fma(b0, m0, rn(b1*m1)) (pinned by tests/golden/heston_injected_f32.npz).
struct OpsF32 {
    using real = float;
    using real2 = float2;
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
struct OpsF64 {
    using real = double;
    using real2 = double2;
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

template <typename Ops, typename Consts>
__device__ __forceinline__ void reference_step(typename Ops::real &x, typename Ops::real &v,
                                               typename Ops::real z0, typename Ops::real z1, const Consts &c) {
    using real = typename Ops::real;
    const real b0 = Ops::mul(z0, c.root_dt);                        // randn_antithetic(...) * sqrt_delta_t
    const real b1 = Ops::mul(z1, c.root_dt);
    const real root_v = Ops::sqrt(v);
    const real drift = Ops::mul(Ops::sub(c.mu, Ops::mul((real)0.5, v)), c.dt);
    const real x_next = Ops::add(Ops::add(x, drift), Ops::mul(root_v, b0));
    const real dot = Ops::fma(b0, c.m0, Ops::mul(b1, c.m1));
    const real pull = Ops::mul(Ops::mul(c.kappa, Ops::sub(c.v_bar, v)), c.dt);
    const real v_next = Ops::add(Ops::add(v, pull), Ops::mul(c.sigma_v, Ops::mul(root_v, dot)));
    v = v_next > (real)1.1920928955078125e-07f ? v_next : (real)1.1920928955078125e-07f;
    x = x_next;
}

template <typename Ops, typename Consts>
__global__ void __launch_bounds__(256)
heston_injected_kernel(const typename Ops::real2 *__restrict__ Z, const __grid_constant__ Consts c, uint64_t R,
                       uint32_t N, typename Ops::real *__restrict__ x_out, typename Ops::real *__restrict__ v_out) {
    using real = typename Ops::real;
    using real2 = typename Ops::real2;
    const uint64_t H = (R + 1) / 2, F = R / 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < H; p += stride) {
        real x1 = c.x0, v1 = c.v0, x2 = c.x0, v2 = c.v0;
        const real2 *zp = Z + p;
        const uint32_t steps = N - 1;
        uint32_t s = 0;
        // 4 independent loads in flight per thread (the loop-carried state is tiny)
        for (; s + 4 <= steps; s += 4) {
            real2 z[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) z[j] = __ldcs(zp + (uint64_t)(s + j) * H);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                reference_step<Ops>(x1, v1, z[j].x, z[j].y, c);
                reference_step<Ops>(x2, v2, -z[j].x, -z[j].y, c);
            }
        }
        for (; s < steps; ++s) {
            const real2 z = __ldcs(zp + (uint64_t)s * H);
            reference_step<Ops>(x1, v1, z.x, z.y, c);
            reference_step<Ops>(x2, v2, -z.x, -z.y, c);
        }
        x_out[p] = x1;
        v_out[p] = v1;
        if (p < F) {
            x_out[p + H] = x2;
            v_out[p + H] = v2;
        }
    }
}

static int grid_for(uint64_t work_items, int threads, int cap_blocks) {
    uint64_t b = (work_items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (uint64_t)cap_blocks) b = cap_blocks;
    return (int)b;
}

cudaError_t launch_heston_injected_f32(const float *d_Z, const InjectedConsts32 &c, uint64_t R, uint32_t N,
                                       float *d_x, float *d_v, cudaStream_t stream) {
    const int threads = 128;
    const int blocks = grid_for((R + 1) / 2, threads, 148 * 16);
    heston_injected_kernel<OpsF32, InjectedConsts32><<<blocks, threads, 0, stream>>>(
        reinterpret_cast<const float2 *>(d_Z), c, R, N, d_x, d_v);
    return cudaGetLastError();
}

cudaError_t launch_heston_injected_f64(const double *d_Z, const InjectedConsts64 &c, uint64_t R, uint32_t N,
                                       double *d_x, double *d_v, cudaStream_t stream) {
    const int threads = 128;
    const int blocks = grid_for((R + 1) / 2, threads, 148 * 16);
    heston_injected_kernel<OpsF64, InjectedConsts64><<<blocks, threads, 0, stream>>>(
        reinterpret_cast<const double2 *>(d_Z), c, R, N, d_x, d_v);
    return cudaGetLastError();
}

// ---- payoff sums over an array of terminal log prices (reference layout) ----------

struct StrikeList {
    int nK;
    double K[kMaxStrikes];
};

template <typename real>
__global__ void __launch_bounds__(256)
payoff_sums_kernel(const real *__restrict__ x, uint64_t R, const __grid_constant__ StrikeList strikes,
                   ReduceWorkspace ws) {
    const uint64_t H = (R + 1) / 2, F = R / 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __shared__ double s_acc[8][2];
    __shared__ bool s_last;
    for (int k = 0; k < strikes.nK; ++k) {
        const double K = strikes.K[k];
        double sum = 0.0, sq = 0.0;
        for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < H; p += stride) {
            const double a = fmax(exp((double)x[p]) - K, 0.0);
            double unit = a, tot = a;
            if (p < F) {
                const double b = fmax(exp((double)x[p + H]) - K, 0.0);
                tot = a + b;
                unit = 0.5 * tot;
            }
            sum += tot;
            sq = fma(unit, unit, sq);
        }
        sum = warp_sum(sum);
        sq = warp_sum(sq);
        __syncthreads();
        if (lane == 0) { s_acc[warp][0] = sum; s_acc[warp][1] = sq; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0.0;
            for (uint32_t w = 0; w < nwarps; ++w) t += s_acc[w][threadIdx.x];
            ws.partials[(size_t)blockIdx.x * kMaxAcc + threadIdx.x * strikes.nK + k] = t;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ws.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int a = threadIdx.x; a < 2 * strikes.nK; a += blockDim.x) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(ws.partials + (size_t)b * kMaxAcc + a);
        ws.result[a] = t;
    }
    if (threadIdx.x == 0) *ws.ticket = 0u;
}

template <typename real>
cudaError_t launch_call_payoff_sums(const real *d_x, uint64_t R, const double *strikes, int nK,
                                    ReduceWorkspace ws, int sm_count, cudaStream_t stream) {
    StrikeList sl;
    sl.nK = nK;
    for (int k = 0; k < kMaxStrikes; ++k) sl.K[k] = k < nK ? strikes[k] : 0.0;
    const int threads = 256;
    const int blocks = grid_for((R + 1) / 2, threads, sm_count * 4);
    payoff_sums_kernel<real><<<blocks, threads, 0, stream>>>(d_x, R, sl, ws);
    return cudaGetLastError();
}

template cudaError_t launch_call_payoff_sums<float>(const float *, uint64_t, const double *, int, ReduceWorkspace,
                                                    int, cudaStream_t);
template cudaError_t launch_call_payoff_sums<double>(const double *, uint64_t, const double *, int,
                                                     ReduceWorkspace, int, cudaStream_t);

}  // namespace cb
