// Pipe-peak calibration kernels: the roofline denominators MEASURED_PEAKS.json lacks
// (FP32 FMA, MUFU, FP64, INT/Philox issue, HBM write), SURVEY.md section 8d step 0.
// Each thread runs 8 independent dependency chains so the pipe, not latency, is timed.
#include "kernels.h"

namespace cb {

constexpr int kChains = 8;
constexpr int kIters = 4096;

template <int KIND>
__device__ __forceinline__ float chain_op(float x, float a, float b) {
    if constexpr (KIND == 0) return fmaf(x, a, b);
    else if constexpr (KIND == 1) return mufu_lg2(x) + 6.0f;     // stays in [6,9]; the FADD is on another pipe
    else if constexpr (KIND == 2) return mufu_sqrt(x);
    else if constexpr (KIND == 3) return mufu_sin(x);
    else if constexpr (KIND == 4) return mufu_ex2(x) * 0.25f;
    else return mufu_rsqrt(x);
}

template <int KIND>
__global__ void __launch_bounds__(256)
peak_f32_kernel(float *out, float a, float b) {
    float x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < kIters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) x[i] = chain_op<KIND>(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;   // never true in practice; keeps the chains alive
}

__global__ void __launch_bounds__(256)
peak_f64_kernel(double *out, double a, double b) {
    double x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = 1.0 + 0.001 * (double)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < kIters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256)
peak_philox_kernel(uint32_t *out, const __grid_constant__ PhiloxKeys k) {
    uint32_t acc = 0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < kIters / 4; it += 2) {
        const uint4 a = philox4x32_10(tid, it, 0u, 1u, k);
        const uint4 b = philox4x32_10(tid, it + 1, 0u, 1u, k);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(256)
peak_alu_kernel(uint32_t *out, uint32_t a, uint32_t b) {
    uint32_t x[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = threadIdx.x * 2654435761u + i;
#pragma unroll 1
    for (int it = 0; it < kIters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) x[i] = (x[i] | a) ^ x[(i + 1) % kChains];   // one LOP3 each
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

// Integer-multiply experiments (what does Philox's 32x32->64 product cost next to other work?)
//   MODE 0: IMAD.WIDE.U32 only            MODE 1: IMAD.HI.U32 + IMAD (lo) pairs
//   MODE 2: IMAD.WIDE + 4 independent FFMA  MODE 3: IMAD.HI only   MODE 4: IMAD lo only
//   MODE 5: IMAD.WIDE + 4 independent LOP3  MODE 6: IMAD.HI+IMAD.LO + 4 FFMA
template <int MODE>
__global__ void __launch_bounds__(256)
peak_imad_kernel(uint32_t *out, uint32_t m0, float fa, float fb) {
    uint32_t x[kChains];
    float f[4];
    uint32_t l[4];
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = threadIdx.x * 2654435761u + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[i] = 1.0f + i; l[i] = threadIdx.x + i; }
#pragma unroll 1
    for (int it = 0; it < kIters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                if constexpr (MODE == 0 || MODE == 2 || MODE == 5) {
                    const uint64_t p = (uint64_t)x[i] * m0;
                    x[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;
                } else if constexpr (MODE == 1 || MODE == 6) {
                    x[i] = __umulhi(x[i], m0) ^ (x[i] * m0);
                } else if constexpr (MODE == 3) {
                    x[i] = __umulhi(x[i], m0);
                } else {
                    x[i] = x[i] * m0 + 1u;
                }
                if constexpr (MODE == 2 || MODE == 6) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) f[j] = fmaf(f[j], fa, fb);
                }
                if constexpr (MODE == 5) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) l[j] = (l[j] | m0) ^ l[(j + 1) & 3];
                }
            }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s ^= x[i];
    float t = f[0] + f[1] + f[2] + f[3];
    s ^= l[0] ^ l[1] ^ l[2] ^ l[3];
    if (s == 0x12345678u || t == 123.456f) out[0] = s;
}

// Blackwell packed FP32: fma.rn.f32x2 (SASS FFMA2).  MODE 0: FFMA2 chains only; MODE 1: each FFMA2 is followed by an
// independent FFMA; MODE 2: ... by an independent LOP3.  Tells whether an FFMA2 costs one issue slot or two.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(256)
peak_ffma2_kernel(float *out, float fa, float fb, uint32_t m0) {
    unsigned long long x[kChains];
    float f[kChains];
    uint32_t l[kChains];
    const unsigned long long a2 = pack2(fa, fa), b2 = pack2(fb, fb);
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        x[i] = pack2(1.0f + 0.001f * (threadIdx.x + i), 2.0f + 0.001f * i);
        f[i] = 1.0f + i;
        l[i] = threadIdx.x * 2654435761u + i;
    }
#pragma unroll 1
    for (int it = 0; it < kIters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                x[i] = ffma2(x[i], a2, b2);
                if constexpr (MODE == 1) f[i] = fmaf(f[i], fa, fb);
                if constexpr (MODE == 2) l[i] = (l[i] | m0) ^ l[(i + 1) % kChains];
            }
    }
    float s = 0.f;
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi + f[i];
        t ^= l[i];
    }
    if (s == 123.456f || t == 0x12345678u) out[0] = s;
}

// Pipe-model probes (round 2): per chain and trip W half Philox rounds (IMAD.WIDE.U32 + three-input XOR),
// F FFMA on three DISTINCT registers, L three-register LOP3 of a nonlinear feedback register, A FADD --
// every group on its own independent chains.  Timing the mixes tells which instructions share a pipe and what an
// IMAD.WIDE really costs next to FP32 and logic work (the Philox round is 2 WIDE + 2 LOP3).
template <int W, int F, int L, int A>
__global__ void __launch_bounds__(256)
peak_mix_kernel(uint32_t *out, uint32_t m0, float fa, float fb) {
    unsigned long long xw[kChains];
    float xf[kChains], yf[kChains], zf[kChains], xa[kChains];
    uint32_t xl[kChains], yl[kChains], zl[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        xw[i] = ((unsigned long long)(threadIdx.x * 2654435761u + i) << 32) | (threadIdx.x + 77u * i);
        xf[i] = 1.0f + 0.001f * (float)(threadIdx.x + i);
        yf[i] = fa + 1e-7f * (float)(i + threadIdx.x);
        zf[i] = fb * (float)(i + 1 + threadIdx.x);
        xa[i] = 0.5f * (float)(i + threadIdx.x);
        xl[i] = threadIdx.x * 40503u + i;
        yl[i] = m0 ^ (0x01010101u * (i + threadIdx.x));
        zl[i] = m0 + i * 977u + threadIdx.x;
    }
#pragma unroll 1
    for (int it = 0; it < kIters; it += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
#pragma unroll
                for (int w = 0; w < W; ++w) {    // half a Philox round: one IMAD.WIDE + one three-input XOR
                    const unsigned long long p = (unsigned long long)(uint32_t)xw[i] * m0;
                    xw[i] = (uint32_t)(p >> 32) ^ (uint32_t)p ^ yl[(i + 3) % kChains];
                }
#pragma unroll
                for (int f = 0; f < F; ++f) xf[i] = fmaf(xf[i], yf[i], zf[i]);
#pragma unroll
                for (int l = 0; l < L; ++l) {    // nonlinear feedback: ptxas cannot collapse consecutive steps
                    const uint32_t t = (xl[i] & yl[i]) ^ zl[i];
                    xl[i] = yl[i];
                    yl[i] = zl[i];
                    zl[i] = t;
                }
#pragma unroll
                for (int a = 0; a < A; ++a) xa[i] = xa[i] + yf[i];
            }
    }
    uint32_t s = 0;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        s ^= (uint32_t)xw[i] ^ xl[i] ^ yl[i] ^ zl[i];
        t += xf[i] + xa[i];
    }
    if (s == 0x12345678u || t == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256)
peak_hbm_write_kernel(float4 *out, size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float v = (float)threadIdx.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
        __stcs(out + i, make_float4(v, v, v, v));
}

__global__ void __launch_bounds__(256)
peak_hbm_copy_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
        __stcs(out + i, __ldcs(in + i));
}

cudaError_t launch_peak_kernel(int kind, int sm_count, void *d_scratch, size_t scratch_bytes, double *work_done,
                               cudaStream_t s) {
    const int threads = 256, grid = sm_count * 8;
    const double thread_ops = (double)grid * threads * kIters * kChains;
    switch (kind) {
        case 0: peak_f32_kernel<0><<<grid, threads, 0, s>>>((float *)d_scratch, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 1: peak_f32_kernel<1><<<grid, threads, 0, s>>>((float *)d_scratch, 0.f, 0.f); *work_done = thread_ops; break;
        case 2: peak_f32_kernel<2><<<grid, threads, 0, s>>>((float *)d_scratch, 0.f, 0.f); *work_done = thread_ops; break;
        case 3: peak_f32_kernel<3><<<grid, threads, 0, s>>>((float *)d_scratch, 0.f, 0.f); *work_done = thread_ops; break;
        case 4: peak_f32_kernel<4><<<grid, threads, 0, s>>>((float *)d_scratch, 0.f, 0.f); *work_done = thread_ops; break;
        case 5: peak_f32_kernel<5><<<grid, threads, 0, s>>>((float *)d_scratch, 0.f, 0.f); *work_done = thread_ops; break;
        case 6: peak_f64_kernel<<<grid, threads, 0, s>>>((double *)d_scratch, 1.000001, 1e-7); *work_done = thread_ops; break;
        case 7: {
            const PhiloxKeys k = philox_keys_from_seed(0x1234567ull);
            peak_philox_kernel<<<grid, threads, 0, s>>>((uint32_t *)d_scratch, k);
            *work_done = (double)grid * threads * (kIters / 4);   // Philox calls
            break;
        }
        case 8: {
            const size_t n_vec = scratch_bytes / 16;
            peak_hbm_write_kernel<<<sm_count * 16, threads, 0, s>>>((float4 *)d_scratch, n_vec);
            *work_done = (double)n_vec * 16.0;
            break;
        }
        case 9: peak_alu_kernel<<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0x9E3779B9u, 0x01010101u); *work_done = thread_ops; break;
        case 10: {
            const size_t n_vec = scratch_bytes / 32;
            peak_hbm_copy_kernel<<<sm_count * 16, threads, 0, s>>>((const float4 *)d_scratch,
                                                                   (float4 *)d_scratch + n_vec, n_vec);
            *work_done = (double)n_vec * 32.0;
            break;
        }
        case 11: peak_imad_kernel<0><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 12: peak_imad_kernel<1><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 13: peak_imad_kernel<2><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 14: peak_imad_kernel<3><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 15: peak_imad_kernel<4><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 16: peak_imad_kernel<5><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 17: peak_imad_kernel<6><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        case 18: peak_ffma2_kernel<0><<<grid, threads, 0, s>>>((float *)d_scratch, 1.000001f, 1e-7f, 0x9E3779B9u); *work_done = thread_ops; break;
        case 19: peak_ffma2_kernel<1><<<grid, threads, 0, s>>>((float *)d_scratch, 1.000001f, 1e-7f, 0x9E3779B9u); *work_done = thread_ops; break;
        case 20: peak_ffma2_kernel<2><<<grid, threads, 0, s>>>((float *)d_scratch, 1.000001f, 1e-7f, 0x9E3779B9u); *work_done = thread_ops; break;
#define CB_MIX(K, W, F, L, A) \
        case K: peak_mix_kernel<W, F, L, A><<<grid, threads, 0, s>>>((uint32_t *)d_scratch, 0xD2511F53u, 1.000001f, 1e-7f); *work_done = thread_ops; break;
        CB_MIX(21, 1, 0, 0, 0)   // IMAD.WIDE alone
        CB_MIX(22, 0, 1, 0, 0)   // FFMA, three distinct registers
        CB_MIX(23, 0, 0, 1, 0)   // LOP3, three registers
        CB_MIX(24, 0, 0, 0, 1)   // FADD
        CB_MIX(25, 1, 0, 1, 0)   // the Philox mix: WIDE + LOP3
        CB_MIX(26, 1, 0, 2, 0)
        CB_MIX(27, 1, 0, 4, 0)
        CB_MIX(28, 1, 1, 0, 0)   // WIDE + FFMA
        CB_MIX(29, 1, 2, 0, 0)
        CB_MIX(30, 1, 4, 0, 0)
        CB_MIX(31, 0, 1, 1, 0)   // FFMA + LOP3: two pipes
        CB_MIX(32, 0, 2, 1, 0)
        CB_MIX(33, 0, 1, 2, 0)
        CB_MIX(34, 1, 1, 1, 0)   // all three
        CB_MIX(35, 1, 2, 2, 0)
        CB_MIX(36, 0, 1, 0, 1)   // FFMA + FADD: same pipe?
        CB_MIX(37, 1, 0, 0, 1)   // WIDE + FADD
        CB_MIX(38, 1, 0, 0, 2)
#undef CB_MIX
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace cb
