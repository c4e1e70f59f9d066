// Heston terminal prices by conditional-Gaussian sampling of the log price (opt-in scheme, statistical mode only).
//
// Same Euler-Maruyama discretisation as the reference (examples/stochastic_volatility/heston_utilities.py:79-91) and
// the SAME LAW for every antithetic pair of terminal log prices, but half the normals.  Rotate the two normals of a
// step, w = rho z0 + rho' z1 and o = rho' z0 - rho z1 (independent N(0,1)).  The variance recursion sees w only, and
//   x_T = x0 + mu T - dt/2 sum v_t + rho sqrt(dt) sum sqrt(v_t) w_t + rho' sqrt(dt) G,    G = sum sqrt(v_t) o_t.
// Given the w-path, (G, G') of a pair (the partner runs on -w, -o) is bivariate normal with variances A = sum v_t,
// B = sum v'_t and covariance -C, C = sum sqrt(v_t v'_t): it is drawn ONCE at the end from two normals.  Per pair-step
// this needs one normal instead of two (one Philox block feeds SIX steps), 4 MUFU instead of 5 and 9 FP32 instead of 15.
// The reference's per-path values are NOT reproduced (there is no x_t inside the loop); prices, their standard
// errors and the joint law of antithetic partners are.  tests/test_gpu_conditional.py checks all three.
#include "kernels.h"
#include "payoff_reduce.cuh"

namespace cb {

struct CondRegs {
    float decay, pull, floor_v;
    float neg2ln2_s2dt;     // -2 ln2 sigma_v^2 dt: the Box-Muller radius carries sigma_v sqrt(dt)
    float two_pi;
};

// two consecutive steps from one (radius, angle) draw; w is already scaled by sigma_v*sqrt(dt)
struct CondState {
    float v1, v2, A, B, C, P1, P2;
};

__device__ __forceinline__ void cond_step(CondState &s, float w, const CondRegs &k) {
    const float s1 = mufu_sqrt(s.v1), s2 = mufu_sqrt(s.v2);
    s.A += s.v1;
    s.B += s.v2;
    s.C = fmaf(s1, s2, s.C);
    s.P1 = fmaf(s1, w, s.P1);
    s.P2 = fmaf(s2, w, s.P2);
    s.v1 = fmaxf(fmaf(s1, w, fmaf(s.v1, k.decay, k.pull)), k.floor_v);
    s.v2 = fmaxf(fmaf(s2, -w, fmaf(s.v2, k.decay, k.pull)), k.floor_v);
}

__device__ __forceinline__ void cond_draw(uint32_t radius_word, float t_angle, const CondRegs &k, float &wa, float &wb) {
    const float r = mufu_sqrt(mufu_lg2(2.0f - mantissa_float(radius_word)) * k.neg2ln2_s2dt);
    const float theta = fmaf(t_angle, k.two_pi, -kThreePi);
    wa = r * mufu_cos(theta);
    wb = r * mufu_sin(theta);
}

__device__ __forceinline__ float cond_angle(uint32_t lo, uint32_t hi, uint32_t one_bits) {
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x007FFFF0, %2, 0xEA;" : "=r"(r) : "r"(__funnelshift_l(lo, hi, 14)), "r"(one_bits));
    return __uint_as_float(r);
}

template <bool MULTI, bool XCHG>
__global__ void __launch_bounds__(kFusedMaxThreads)
heston_conditional_kernel(const __grid_constant__ FusedConsts<float> c, ReduceWorkspace ws, float sigma_v, float rho,
                          float dt, float *__restrict__ x_out, float *__restrict__ v_out) {
    const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t iters = (c.pair_count + total - 1) / total;
    const unsigned lane = threadIdx.x & 31;
    __shared__ float s_strikes[kMaxStrikes];
    if (MULTI) stage_strikes(s_strikes, c.strikes, c.nK);
    PayoffAccumulator<float, MULTI> acc;

    CondRegs k;
    k.decay = c.decay;
    k.pull = c.pull;
    k.floor_v = c.floor_v;
    k.neg2ln2_s2dt = c.neg2ln2_dt * sigma_v * sigma_v;
    k.two_pi = kTwoPi;
    const uint32_t one_bits = 0x3F800000u | (c.steps >> 31);
    PhiloxKeys keys;
#pragma unroll
    for (int r = 0; r < 10; ++r) { keys.k0[r] = pin(c.keys.k0[r]); keys.k1[r] = pin(c.keys.k1[r]); }
    const float rho_over_sigma = rho / sigma_v;
    const float orth = sqrtf(fmaxf(1.0f - rho * rho, 0.0f)) * sqrtf(dt);
    const float x_base = c.x0 + (float)c.steps * c.mu_dt;
    const uint32_t calls = c.steps / 6, rest = c.steps - 6 * calls;

    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t local = it * total + gtid;
        const bool active = local < c.pair_count;
        const uint64_t gp = c.pair_first + (active ? local : 0);
        const bool paired = active && gp < c.partnered;
        if (!__any_sync(0xffffffffu, active)) continue;      // last round: this warp lies wholly past the end of the shard
        PairStream rng;
        rng.init((uint32_t)gp, (uint32_t)(gp >> 32), c.subsequence, c.keys);
        CondState s{c.v0, c.v0, 0.f, 0.f, 0.f, 0.f, 0.f};

        auto six_steps = [&](const uint4 &o, uint32_t n_steps) {      // n_steps in 1..6
            float wa, wb;
            cond_draw(o.x, cond_angle(o.w, o.x, one_bits), k, wa, wb);
            cond_step(s, wa, k);
            if (n_steps > 1) cond_step(s, wb, k);
            if (n_steps > 2) {
                cond_draw(o.y, cond_angle(o.w << 10, o.y, one_bits), k, wa, wb);
                cond_step(s, wa, k);
                if (n_steps > 3) cond_step(s, wb, k);
            }
            if (n_steps > 4) {
                cond_draw(o.z, cond_angle(o.w << 20, o.z, one_bits), k, wa, wb);
                cond_step(s, wa, k);
                if (n_steps > 5) cond_step(s, wb, k);
            }
        };

        uint4 cur = rng.block(0u, keys);
        uint64_t m1b = 0;
        for (uint32_t b = 0; b < calls; ++b) {
            m1b += kPhiloxM1;
            const uint4 nxt = rng.block_from_product(m1b, keys);      // software-pipelined next block
            float wa, wb;
            cond_draw(cur.x, cond_angle(cur.w, cur.x, one_bits), k, wa, wb);
            cond_step(s, wa, k);
            cond_step(s, wb, k);
            cond_draw(cur.y, cond_angle(cur.w << 10, cur.y, one_bits), k, wa, wb);
            cond_step(s, wa, k);
            cond_step(s, wb, k);
            cond_draw(cur.z, cond_angle(cur.w << 20, cur.z, one_bits), k, wa, wb);
            cond_step(s, wa, k);
            cond_step(s, wb, k);
            cur = nxt;
        }
        if (rest) six_steps(cur, rest);

        // the orthogonal Gaussian parts of the pair: block 0xFFFFFFFF of the pair's counter is reserved for them
        const uint4 fin = rng.block(0xFFFFFFFFu, keys);
        float z1, z2;
        box_muller(fin.x, fin.y, -1.3862943611198906f, kTwoPi, z1, z2);
        const float rootA = sqrtf(s.A);
        const float G1 = rootA * z1;
        const float cross = s.A > 0.f ? s.C / rootA : 0.f;
        const float G2 = -cross * z1 + sqrtf(fmaxf(s.B - cross * cross, 0.f)) * z2;
        const float xT1 = x_base - 0.5f * dt * s.A + rho_over_sigma * s.P1 + orth * G1;
        const float xT2 = x_base - 0.5f * dt * s.B - rho_over_sigma * s.P2 + orth * G2;
        if (x_out != nullptr) {
            if (active) { x_out[local] = xT1; v_out[local] = s.v1; }
            if (paired) { x_out[c.pair_count + local] = xT2; v_out[c.pair_count + local] = s.v2; }
        }
        acc.add(mufu_ex2(xT1 * 1.4426950408889634f), mufu_ex2(xT2 * 1.4426950408889634f), active, paired, c.strikes[0],
                s_strikes, lane);
    }
    acc.template publish<XCHG>(c.nK, ws);
}

cudaError_t launch_heston_conditional(const FusedConsts<float> &c, float sigma_v, float rho, float dt, int threads,
                                      int blocks, ReduceWorkspace ws, float *d_x, float *d_v, cudaStream_t stream) {
    const bool x = ws.px != nullptr;
    if (c.nK > 1) {
        if (x) heston_conditional_kernel<true, true><<<blocks, threads, 0, stream>>>(c, ws, sigma_v, rho, dt, d_x, d_v);
        else heston_conditional_kernel<true, false><<<blocks, threads, 0, stream>>>(c, ws, sigma_v, rho, dt, d_x, d_v);
    } else {
        if (x) heston_conditional_kernel<false, true><<<blocks, threads, 0, stream>>>(c, ws, sigma_v, rho, dt, d_x, d_v);
        else heston_conditional_kernel<false, false><<<blocks, threads, 0, stream>>>(c, ws, sigma_v, rho, dt, d_x, d_v);
    }
    return cudaGetLastError();
}

cudaError_t conditional_occupancy(int threads, bool multi, int *blocks_per_sm) {
    const void *fn = multi ? (const void *)heston_conditional_kernel<true, false> : (const void *)heston_conditional_kernel<false, false>;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, threads, 0);
}

}  // namespace cb
