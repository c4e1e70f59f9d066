/* The C ABI without any Python: price the reference's European call (heston_pricing_benchmark.py:163-177) from plain C.
 *
 *   gcc -O2 -I include examples/c/heston_price.c -L cocos_b200 -lcocos_b200 -Wl,-rpath,$PWD/cocos_b200 -lm -o /tmp/heston_price
 *   /tmp/heston_price [paths] [time points]
 *
 * simulate_and_compute_option_price (examples/stochastic_volatility/heston_utilities.py:122-164) is
 * cb_heston_price_launch_f32 + cb_heston_price_finish; the standard error comes with it (sum of squared
 * antithetic-pair averages). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cocos_b200.h"

#define CHECK(call)                                                       \
    do {                                                                  \
        int rc__ = (call);                                                \
        if (rc__ != CB_OK) {                                              \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, cb_last_error()); \
            return 1;                                                     \
        }                                                                 \
    } while (0)

int main(int argc, char **argv) {
    const uint64_t R = argc > 1 ? strtoull(argv[1], NULL, 10) : 10000000ull;
    const uint32_t N = argc > 2 ? (uint32_t)strtoul(argv[2], NULL, 10) : 501u;
    const cb_heston_params p = {.T = 1.0, .mu = log(1.0319), .kappa = 6.21, .v_bar = 0.019, .sigma_v = 0.61,
                                .rho = -0.7, .x0 = 0.0, .v0 = 0.101 * 0.101};
    const double strike = 0.95;
    int devices = 0;
    CHECK(cb_device_count(&devices));
    char name[128];
    CHECK(cb_device_name(0, name, sizeof name));
    printf("%s, %d device(s): %s\n", cb_version(), devices, name);

    cb_ctx *ctx = NULL;
    CHECK(cb_ctx_create(0, NULL, &ctx));
    double sum = 0.0, sumsq = 0.0;
    float ms = 0.f;
    for (uint32_t launch = 0; launch < 3; ++launch) {          /* the first launch warms up */
        CHECK(cb_timer_start(ctx));
        CHECK(cb_heston_price_launch_f32(ctx, &p, R, N, &strike, 1, /*seed*/ 0, /*subsequence*/ launch,
                                         /*pair_first*/ 0, /*pair_count*/ (R + 1) / 2, /*comm*/ NULL,
                                         NULL, NULL, NULL));
        CHECK(cb_timer_stop(ctx, &ms));
        CHECK(cb_heston_price_finish(ctx, 1, &sum, &sumsq));
    }
    const double disc = exp(-p.mu * p.T), units = (double)((R + 1) / 2);
    const double price = disc * sum / (double)R;
    const double mean_unit = sum / (2.0 * units), var_unit = fmax(sumsq / units - mean_unit * mean_unit, 0.0);
    printf("%llu paths x %u steps: price %.6f +- %.1e   %.3f ms   %.3e path-steps/s\n", (unsigned long long)R, N - 1, price,
           disc * sqrt(var_unit / (units - 1.0)), ms, (double)R * (N - 1) / (ms * 1e-3));
    CHECK(cb_ctx_destroy(ctx));
    return 0;
}
