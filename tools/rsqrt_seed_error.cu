// Mean and rms relative error of the seed-only double square root of the float64 Heston kernels,
//   q = p * rsqrt.approx.ftz.f64(p)      (MUFU.RSQ64H: reads the high word of p only)
// against sqrt(p), over p with a log-uniform mantissa and over the kernel's own operand distribution p = u * lg2 U.
// The mean goes into StepRegs<double> (csrc/heston_step.cuh) as the scaling of c1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/rsqrt_seed_error tools/rsqrt_seed_error.cu && /tmp/rsqrt_seed_error
#include <cstdint>
#include <cstdio>
#include <cmath>

__device__ __forceinline__ uint64_t mix(uint64_t z) {   // splitmix64
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void probe(int mode, uint64_t n_per_thread, double *sum, double *sumsq, double *minmax) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    double s = 0, s2 = 0, lo = 0, hi = 0;
    for (uint64_t i = 0; i < n_per_thread; ++i) {
        const uint64_t a = mix(tid * n_per_thread + i), b = mix(a);
        const double ua = (double)(a >> 11) * 0x1p-53, ub = ((double)(b >> 11) + 1.0) * 0x1p-53;
        double p;
        if (mode == 0) p = exp2(ua * 40.0 - 30.0);                    // log-uniform over 40 binades
        else p = (2.8e-3 * (0.002 + 0.04 * ua)) * (-log2(ub));        // |c1| v * |lg2 U| at dt ~ 1/500 .. typical v
        if (p <= 0) continue;
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
        const double e = p * y / sqrt(p) - 1.0;
        s += e; s2 += e * e; lo = fmin(lo, e); hi = fmax(hi, e);
    }
    atomicAdd(sum, s); atomicAdd(sumsq, s2);
    if (lo < -1e-6 || hi > 1e-6) { minmax[0] = lo; minmax[1] = hi; }
}

int main() {
    double *d;
    cudaMalloc(&d, 4 * sizeof(double));
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(d, 0, 4 * sizeof(double));
        const uint64_t per = 4096; const int blocks = 1184, threads = 256;
        probe<<<blocks, threads>>>(mode, per, d, d + 1, d + 2);
        double h[4];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double n = (double)per * blocks * threads;
        printf("mode %d (%s): n %.3g  mean rel err %.6e (= %.4f * 2^-22)  rms %.6e (= %.4f * 2^-22)  outliers [%g, %g]\n", mode,
               mode ? "p = |c1| v |lg2 U|" : "log-uniform", n, h[0] / n, h[0] / n / 0x1p-22, sqrt(h[1] / n), sqrt(h[1] / n) / 0x1p-22, h[2], h[3]);
    }
    return 0;
}
