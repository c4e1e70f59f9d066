// Counter-based RNG and the transforms shared by every kernel (sm_100a).
//
// Replaces ArrayFire's default engine, which the reference selects as
// PHILOX_4X32_10 (cocos/numerics/random.py:30-51; cocos/options/__init__.py:10-16)
// and calls through af.data.randn / af.data.randu (random.py:137,152).
//
// Stream layout (CPU model: oracle/philox_numpy.py, oracle/heston_oracle.c):
//   key     = (seed_lo, seed_hi)
//   counter = (index_lo, index_hi, block, subsequence)
// One Philox call costs 10 x (2 IMAD.WIDE.U32 + 2 LOP3): the round keys are
// launch-uniform, precomputed on the host and read from the constant bank.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cb {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// key schedule for the 10 rounds, filled by the host (philox_keys_from_seed)
struct PhiloxKeys {
    uint32_t k0[10];
    uint32_t k1[10];
};

static inline PhiloxKeys philox_keys_from_seed(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a;
        k.k1[r] = b;
        a += kPhiloxW0;
        b += kPhiloxW1;
    }
    return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const PhiloxKeys& k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;   // IMAD.WIDE.U32
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[r];   // one LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

// value-preserving moves the optimiser cannot see through: keeps launch constants
// and per-pair invariants in registers instead of re-deriving them in the time loop
__device__ __forceinline__ uint32_t pin(uint32_t v) { uint32_t r; asm("mov.b32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ float pin(float v) { float r; asm("mov.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ double pin(double v) { double r; asm("mov.f64 %0, %1;" : "=d"(r) : "d"(v)); return r; }

// Per-pair Philox stream with the loop-invariant part of rounds 0 and 1 hoisted.
// For counter (plo, phi, b, s) only c2 = b changes inside the time loop and b is
// warp-uniform, so round 0 is one vector LOP3 and round 1 saves one IMAD.WIDE:
// 36 vector instructions per call instead of 40.
struct PairStream {
    uint32_t phi;       // c1
    uint32_t c3_r1;     // lo(M0*plo): c3 entering round 1
    uint32_t p1hi_r1;   // hi(M1*c2) entering round 1, c2 = hi(M0*plo)^s^k1[0] (invariant)
    uint32_t p1lo_r1;   // lo(M1*c2) -> c1 after round 1

    __device__ __forceinline__ void init(uint32_t plo, uint32_t phi_, uint32_t subsequence, const PhiloxKeys& k) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * plo;
        const uint32_t c2 = (uint32_t)(p0 >> 32) ^ subsequence ^ k.k1[0];
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        phi = pin(phi_);
        c3_r1 = pin((uint32_t)p0);
        p1hi_r1 = pin((uint32_t)(p1 >> 32));
        p1lo_r1 = pin((uint32_t)p1);
    }

    template <int ROUNDS = 10>
    __device__ __forceinline__ uint4 block(uint32_t b, const PhiloxKeys& k) const {
        return block_from_product<ROUNDS>((uint64_t)kPhiloxM1 * b, k);
    }

    // u1 = M1 * block index, which a loop over consecutive blocks advances by one 64-bit addition of M1
    // (two issue slots) instead of recomputing it with an IMAD.WIDE (four)
    template <int ROUNDS = 10>
    __device__ __forceinline__ uint4 block_from_product(uint64_t u1, const PhiloxKeys& k) const {
        // round 0 (c0 = plo, c1 = phi, c2 = b, c3 = s): the products are invariant or warp-uniform
        uint32_t c0 = ((uint32_t)(u1 >> 32) ^ k.k0[0]) ^ phi;            // 1 LOP3
        uint32_t c1 = (uint32_t)u1;                                      // uniform
        uint32_t c3 = c3_r1;
        // round 1
        {
            const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
            const uint32_t n0 = p1hi_r1 ^ (c1 ^ k.k0[1]);
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[1];
            c1 = p1lo_r1;
            c3 = (uint32_t)p0;
            c0 = n0;
            // c2 below
            uint32_t c2 = n2;
#pragma unroll
            for (int r = 2; r < ROUNDS; ++r) {
                const uint64_t q0 = (uint64_t)kPhiloxM0 * c0;
                const uint64_t q1 = (uint64_t)kPhiloxM1 * c2;
                const uint32_t m0 = (uint32_t)(q1 >> 32) ^ c1 ^ k.k0[r];
                const uint32_t m2 = (uint32_t)(q0 >> 32) ^ c3 ^ k.k1[r];
                c1 = (uint32_t)q1;
                c3 = (uint32_t)q0;
                c0 = m0;
                c2 = m2;
            }
            return make_uint4(c0, c1, c2, c3);
        }
    }
};

// Philox4x32-10 for counters (c0, hi, 0, subsequence) with everything but c0 uniform: rounds 0 and 1 cost one
// multiply each.  Bit-identical to philox4x32_10(c0, hi, 0, subsequence, k).
struct UniformTail {
    uint32_t x0;        // subsequence ^ k1[0]           n2 of round 0 = hi(M0*c0) ^ x0
    uint32_t x1;        // hi(M0*A) ^ k1[1]              n2 of round 1 = lo(M0*c0) ^ x1,   A = hi ^ k0[0]
    uint32_t k01;       // k0[1]                         n0 of round 1 = hi(M1*n2) ^ k01   (c1 entering round 1 is 0)
    uint32_t x2;        // lo(M0*A) ^ k1[2]              n2 of round 2 = hi(M0*n0') ^ x2
    __device__ __forceinline__ void init(uint32_t hi, uint32_t subsequence, const PhiloxKeys &k) {
        const uint32_t A = hi ^ k.k0[0];
        const uint64_t pa = (uint64_t)kPhiloxM0 * A;
        x0 = pin(subsequence ^ k.k1[0]);
        x1 = pin((uint32_t)(pa >> 32) ^ k.k1[1]);
        k01 = pin(k.k0[1]);
        x2 = pin((uint32_t)pa ^ k.k1[2]);
    }
    __device__ __forceinline__ uint4 call(uint32_t c0, const PhiloxKeys &k) const {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;                    // round 0
        const uint32_t a2 = (uint32_t)(p0 >> 32) ^ x0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * a2;                    // round 1
        uint32_t c0r = (uint32_t)(p1 >> 32) ^ k01;
        uint32_t c2r = (uint32_t)p0 ^ x1;
        uint32_t c1r = (uint32_t)p1;
        uint32_t c3r;
        {                                                                // round 2: c3 entering it is uniform
            const uint64_t q0 = (uint64_t)kPhiloxM0 * c0r;
            const uint64_t q1 = (uint64_t)kPhiloxM1 * c2r;
            const uint32_t m0 = (uint32_t)(q1 >> 32) ^ c1r ^ k.k0[2];
            const uint32_t m2 = (uint32_t)(q0 >> 32) ^ x2;
            c1r = (uint32_t)q1;
            c3r = (uint32_t)q0;
            c0r = m0;
            c2r = m2;
        }
#pragma unroll
        for (int r = 3; r < 10; ++r) {
            const uint64_t q0 = (uint64_t)kPhiloxM0 * c0r;
            const uint64_t q1 = (uint64_t)kPhiloxM1 * c2r;
            const uint32_t m0 = (uint32_t)(q1 >> 32) ^ c1r ^ k.k0[r];
            const uint32_t m2 = (uint32_t)(q0 >> 32) ^ c3r ^ k.k1[r];
            c1r = (uint32_t)q1;
            c3r = (uint32_t)q0;
            c0r = m0;
            c2r = m2;
        }
        return make_uint4(c0r, c1r, c2r, c3r);
    }
};

// float in [1,2) from the high 23 bits: one LEA.HI (shift + add), no I2F (I2F
// shares the quarter-rate XU pipe with MUFU, the pipe this workload is bound by).
__device__ __forceinline__ float mantissa_float(uint32_t o) {
    return __uint_as_float((o >> 9) + 0x3F800000u);
}

__device__ __forceinline__ float u01_f32(uint32_t o) {          // [0,1), exact
    return __fadd_rn(mantissa_float(o), -1.0f);
}

__device__ __forceinline__ double u01_f64(uint32_t hi, uint32_t lo) {   // [0,1), 52 bits, exact
    const uint64_t bits = 0x3FF0000000000000ull | ((uint64_t)(hi & 0xFFFFFu) << 32) | lo;
    return __longlong_as_double((long long)bits) - 1.0;
}

// bare MUFU ops (ftz forms: no denormal fix-up code around the MUFU)
__device__ __forceinline__ float mufu_lg2(float x)  { float y; asm("lg2.approx.ftz.f32 %0, %1;"  : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x)  { float y; asm("ex2.approx.ftz.f32 %0, %1;"  : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsqrt(float x){ float y; asm("rsqrt.approx.ftz.f32 %0, %1;": "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x)  { float y; asm("sin.approx.ftz.f32 %0, %1;"  : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x)  { float y; asm("cos.approx.ftz.f32 %0, %1;"  : "=f"(y) : "f"(x)); return y; }

constexpr float kTwoPi   = 6.28318530717958647692f;
constexpr float kThreePi = 9.42477796076937971538f;

// Box-Muller: (o_r, o_a) -> scale * (r cos t, r sin t), r = sqrt(-2 ln U).
// neg2ln2_scale2 = -2 ln(2) * scale^2 folds the ln and the step scaling (sqrt(dt))
// into the one multiply that feeds MUFU.SQRT.  4 MUFU + 8 other instructions.
__device__ __forceinline__ void box_muller(uint32_t o_r, uint32_t o_a, float neg2ln2_scale2, float two_pi,
                                           float& z0, float& z1) {
    const float U = 2.0f - mantissa_float(o_r);                 // (0,1]
    const float r = mufu_sqrt(mufu_lg2(U) * neg2ln2_scale2);
    const float theta = fmaf(mantissa_float(o_a), two_pi, -kThreePi);   // [-pi,pi)
    z0 = r * mufu_cos(theta);
    z1 = r * mufu_sin(theta);
}

// ---- reductions -----------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

}  // namespace cb
