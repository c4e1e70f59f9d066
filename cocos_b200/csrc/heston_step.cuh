// Per-step arithmetic of the fused Heston kernels (sm_100a), shared by the pricing kernel (heston_fused.cu) and
// the path-materialising kernel (heston_trajectory.cu).
//
//   simulate_heston_model      examples/stochastic_volatility/heston_utilities.py:58-96
//
// Algebra (statistical mode, not the parity order): with a = 1-kappa*dt, b = kappa*v_bar*dt,
// w = sigma_v*(rho*dB0 + sqrt(1-rho^2)*dB1):
//   x' = x - 0.5*dt*v + sqrt(v)*dB0          (mu*dt per step is added on the way out)
//   v' = max(a*v + b + sqrt(v)*w, eps_f32)
#pragma once
#include "kernels.h"

namespace cb {

// ---- per-type math ---------------------------------------------------------

__device__ __forceinline__ float root(float v) { return mufu_sqrt(v); }
// double sqrt at the accuracy of everything else that multiplies it: ONE multiply on the MUFU.RSQ64H seed.
// rsqrt.approx.f64 reads only the HIGH word of p (20 mantissa bits) and is good to ~2^-22 -- the accuracy class of the
// MUFU lg2/sin/cos that make the normals this root is multiplied with, in the float64 kernel as in the float32 one
// (the float64 kernel keeps float64 STATE and accumulation; its normals are single-precision grade).
// Measured on B200 over 1.2e9 operands (tools/rsqrt_seed_error.cu; log-uniform p and the kernel's own p = u*lg2 U give
// the same figures): relative error of p*seed against sqrt(p): mean -1.3334e-7, rms 2.76e-7, |.| < 1e-6.  The mean
// is taken out for free by StepRegs<double>, which scales c1 (q = sqrt(u*lg2 U) carries sqrt(c1)) by 1 + 2.6668e-7;
// what is left is zero-mean noise of the size of the MUFU.SQRT error of the float32 kernel.  FP64 instructions cost
// two issue cycles each on B200: the Newton step this replaces (3 more FP64 per root, 6 of ~20 per pair-step) bought
// a 2^-45 root under 2^-22 normals and 9% of the kernel's time.
// The parity kernel keeps __dsqrt_rn.  p > 0 always: euler_pair keeps lg2 U away from zero.
__device__ __forceinline__ double root(double p) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
    return p * y;
}
// min without the NaN bookkeeping of fmin(double) (DSETP + 2 SEL instead of 6 instructions)
__device__ __forceinline__ float floor_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double floor_min(double a, double b) { return a < b ? a : b; }
// max without the NaN bookkeeping of fmax(double) (the operands are never NaN here)
__device__ __forceinline__ float plain_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double plain_max(double a, double b) { return a > b ? a : b; }
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
        // double: c1 = -2 ln2 dt from the double drift (not the float copy); the seed-only root() above is 1.3334e-7
        // too small on average -> c1 is scaled by 1 + 2 * 1.3334e-7 (q = sqrt(u * lg2 U) carries sqrt(c1))
        const real c1 = sizeof(real) == 8 ? (real)((double)c.drift_v * (4.0 * 0.6931471805599453) * (1.0 + 2.6668e-7))
                                          : (real)c.neg2ln2_dt;
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

// float -> double of a NORMAL, strictly NEGATIVE float without F2F: the conversion instruction shares the XU pipe with
// the MUFUs, and the float64 kernel has 5 MUFU + 2 RSQ64H + 3 F2F per pair-step on it (ncu: XU 75% busy, the busiest
// pipe).  Re-biasing the exponent with integer arithmetic is exact and costs LEA.HI + SHL: -3.4% kernel time
// (3.66 -> 3.54 ms on the C5 shard).  Only lg2 U has a known sign; the same for cos and g (sign and zero handling: four
// integer instructions each) was measured slower (3.75 ms) and is left to F2F.
template <typename real>
__device__ __forceinline__ real widen_negative(float f);
template <>
__device__ __forceinline__ float widen_negative<float>(float f) { return f; }
template <>
__device__ __forceinline__ double widen_negative<double>(float f) {
    const uint32_t b = __float_as_uint(f);       // 1 | E (8) | M (23)  ->  1 | E + 896 (11) | M << 29 (52)
    return __hiloint2double((int)((b >> 3) + 0xA8000000u), (int)(b << 29));
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
    const real L = widen_negative<real>(L_f);                             // exact; no F2F for double
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


}  // namespace cb
