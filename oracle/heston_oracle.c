/* TEST INFRASTRUCTURE ONLY -- never linked, loaded or called by the product.
 *
 * Plain-C restatement of the reference's Heston Monte-Carlo hot path with every
 * rounding made explicit, so that the CUDA parity kernel can be compared bit for
 * bit.  Built by oracle/Makefile into oracle/liboracle.so; loaded with ctypes by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline only.
 *
 * Pinned: tests/test_oracle_golden.py checks these functions bit-exactly against
 * tests/golden/ (outputs of the unmodified reference NumPy branch).
 *
 * Follows (paths relative to /root/reference):
 *   oracle_heston_injected_f32/_f64  examples/stochastic_volatility/heston_utilities.py:58-96
 *                                    antithetic layout cocos/numerics/random.py:190-216
 *   oracle_call_payoff_sums          heston_utilities.py:119
 *   oracle_philox4x32_10             published Philox4x32-10 (Salmon et al. SC'11); the reference
 *                                    selects it as ArrayFire's engine, cocos/numerics/random.py:30-51
 *   oracle_pi_inside                 examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py:24-31
 *                                    with rand_antithetic layout, cocos/numerics/random.py:219-245
 *
 * Compile with -ffp-contract=off: every product and sum below is one rounding,
 * exactly as NumPy evaluates the reference's array expressions.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    double T, mu, kappa, v_bar, sigma_v, rho, x0, v0;
} oracle_heston_params;

#define F32_EPS 1.1920928955078125e-07f

/* Z: [N-1][H][2] standard normals (H = ceil(R/2)); path i < H uses +Z[.][i],
 * path i >= H uses -Z[.][i-H].  x, v: [R] terminal values in the reference layout.
 *
 * np.dot((R,2) f32, (2,) f32) is evaluated by OpenBLAS sgemv; on the build
 * container it rounds as fmaf(z0, m0, rn(z1*m1)) (re-checked by
 * the golden test), which is what is written here. */
void oracle_heston_injected_f32(const float *Z, const oracle_heston_params *p,
                                uint64_t R, uint32_t N, float *x_out, float *v_out)
{
    const uint64_t H = (R + 1) / 2;
    const double dt_d = p->T / (double)(N - 1);
    const float dt = (float)dt_d;
    const float root_dt = (float)sqrt(dt_d);
    const float m0 = (float)p->rho;
    const float m1 = (float)sqrt(1.0 - p->rho * p->rho);
    const float mu = (float)p->mu, kappa = (float)p->kappa, v_bar = (float)p->v_bar;
    const float sigma_v = (float)p->sigma_v;

    for (uint64_t i = 0; i < R; ++i) {
        const float sign = (i < H) ? 1.0f : -1.0f;
        const uint64_t src = (i < H) ? i : i - H;
        float x = (float)p->x0, v = (float)p->v0;
        for (uint32_t s = 0; s + 1 < N; ++s) {
            const float *z = Z + ((uint64_t)s * H + src) * 2;
            const float b0 = (sign * z[0]) * root_dt;     /* negation is exact */
            const float b1 = (sign * z[1]) * root_dt;
            const float root_v = sqrtf(v);
            const float drift = (mu - 0.5f * v) * dt;
            const float x_next = (x + drift) + root_v * b0;
            const float dot = fmaf(b0, m0, b1 * m1);
            const float pull = (kappa * (v_bar - v)) * dt;
            const float v_next = (v + pull) + sigma_v * (root_v * dot);
            v = v_next > F32_EPS ? v_next : F32_EPS;      /* np.maximum; NaN cannot occur */
            x = x_next;
        }
        x_out[i] = x;
        v_out[i] = v;
    }
}

void oracle_heston_injected_f64(const double *Z, const oracle_heston_params *p,
                                uint64_t R, uint32_t N, double *x_out, double *v_out)
{
    const uint64_t H = (R + 1) / 2;
    const double dt = p->T / (double)(N - 1);
    const double root_dt = sqrt(dt);
    const double m0 = p->rho, m1 = sqrt(1.0 - p->rho * p->rho);
    const double eps = (double)F32_EPS;

    for (uint64_t i = 0; i < R; ++i) {
        const double sign = (i < H) ? 1.0 : -1.0;
        const uint64_t src = (i < H) ? i : i - H;
        double x = p->x0, v = p->v0;
        for (uint32_t s = 0; s + 1 < N; ++s) {
            const double *z = Z + ((uint64_t)s * H + src) * 2;
            const double b0 = (sign * z[0]) * root_dt;
            const double b1 = (sign * z[1]) * root_dt;
            const double root_v = sqrt(v);
            const double x_next = (x + (p->mu - 0.5 * v) * dt) + root_v * b0;
            const double dot = fma(b0, m0, b1 * m1);
            const double v_next = (v + (p->kappa * (p->v_bar - v)) * dt) + p->sigma_v * (root_v * dot);
            v = v_next > eps ? v_next : eps;
            x = x_next;
        }
        x_out[i] = x;
        v_out[i] = v;
    }
}

/* sum of call payoffs and sum of squared antithetic-unit averages, per strike */
void oracle_call_payoff_sums(const double *x_T, uint64_t R, const double *strikes, int nK,
                             double *sum, double *sumsq)
{
    const uint64_t H = (R + 1) / 2, F = R / 2;
    for (int k = 0; k < nK; ++k) {
        long double s = 0.0L, q = 0.0L;
        for (uint64_t i = 0; i < H; ++i) {
            double a = exp(x_T[i]) - strikes[k];
            a = a > 0.0 ? a : 0.0;
            double unit = a;
            s += a;
            if (i < F) {
                double b = exp(x_T[i + H]) - strikes[k];
                b = b > 0.0 ? b : 0.0;
                s += b;
                unit = 0.5 * (a + b);
            }
            q += (long double)unit * unit;
        }
        sum[k] = (double)s;
        sumsq[k] = (double)q;
    }
}

/* ---------------------------------------------------------------- Philox */

static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1;
    c[3] = (uint32_t)p0;
    c[0] = n0;
    c[2] = n2;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}

static inline float mantissa_float(uint32_t o)
{
    uint32_t bits = (o >> 9) | 0x3F800000u;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* raw words for indices [first, first+count): out[4*i..4*i+3] */
void oracle_philox_words(uint64_t first, uint64_t count, uint32_t block, uint32_t subsequence,
                         uint64_t seed, uint32_t *out)
{
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t idx = first + i;
        const uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), block, subsequence};
        oracle_philox4x32_10(ctr, key, out + 4 * i);
    }
}

/* exact count of points inside the quarter circle; see philox_numpy.pi_inside_count */
uint64_t oracle_pi_inside(uint64_t n, uint64_t seed, uint32_t subsequence, int antithetic)
{
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint64_t ndraw = antithetic ? (n + 1) / 2 : n;
    const uint64_t npart = antithetic ? n / 2 : 0;
    uint64_t inside = 0;
    for (uint64_t j = 0; j < ndraw; j += 2) {
        const uint64_t idx = j / 2;
        const uint32_t ctr[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), 0u, subsequence};
        uint32_t o[4];
        oracle_philox4x32_10(ctr, key, o);
        for (int h = 0; h < 2 && j + h < ndraw; ++h) {
            const float ux = mantissa_float(o[2 * h]) - 1.0f;
            const float uy = mantissa_float(o[2 * h + 1]) - 1.0f;
            const float xx = ux * ux, yy = uy * uy;
            inside += (xx + yy <= 1.0f);
            if (j + h < npart) {
                const float ax = 1.0f - ux, ay = 1.0f - uy;
                const float axx = ax * ax, ayy = ay * ay;
                inside += (axx + ayy <= 1.0f);
            }
        }
    }
    return inside;
}
