"""Heston call pricing on B200: the driver of examples/stochastic_volatility/heston_pricing_benchmark.py
(reference lines 154-271) rewritten for cocos_b200.  Same model parameters and sizes (2,000,000 paths, 500 time
points), same protocol: a small warm-up call, then timed pricing calls, on 1..G GPUs.

    python examples/stochastic_volatility/heston_pricing_benchmark.py [--paths 2000000] [--points 500]
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import cocos_b200.device as cd                                                    # noqa: E402
from cocos_b200 import CocosBundle                                                # noqa: E402
from cocos_b200.heston import (compute_option_price_from_simulated_paths, price_with_standard_error,  # noqa: E402
                               simulate_and_compute_option_price_gpu, simulate_heston_model)
from cocos_b200.multi_processing import ComputeDevicePool                         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=int, default=2_000_000)
    ap.add_argument("--points", type=int, default=500)
    args = ap.parse_args()

    # model and option parameters of the reference benchmark (heston_pricing_benchmark.py:163-177)
    model = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019)
    T, K, nT, R = 1.0, 0.95, args.points, args.paths
    cd.info()

    def two_step_price(paths):
        x, _ = simulate_heston_model(T=T, N=nT, R=paths, mu=model["r"], kappa=model["kappa"], v_bar=model["v_bar"],
                                     sigma_v=model["sigma_v"], rho=model["rho"], x0=model["x0"], v0=model["v0"],
                                     numerical_package_bundle=CocosBundle)
        return compute_option_price_from_simulated_paths(model["r"], T, K, x, CocosBundle)

    two_step_price(1000)                                   # warm-up, as the reference does (:105-118)
    t0 = time.perf_counter()
    price = two_step_price(R)
    CocosBundle.synchronize()
    dt = time.perf_counter() - t0
    print(f"simulate_heston_model + payoff : {price:.6f}   {dt * 1e3:8.3f} ms   {R * (nT - 1) / dt:.3e} path-steps/s")

    t0 = time.perf_counter()
    fused, se = price_with_standard_error(T=T, K=K, nT=nT, R=R, **model)
    dt = time.perf_counter() - t0
    print(f"fused pricing kernel           : {fused:.6f} +- {se:.1e}   {dt * 1e3:8.3f} ms   "
          f"{R * (nT - 1) / dt:.3e} path-steps/s")

    n_gpus = cd.ComputeDeviceManager.get_number_of_compute_devices()
    for g in range(1, n_gpus + 1):
        pool = ComputeDevicePool(compute_devices=list(range(g)))
        simulate_and_compute_option_price_gpu(T=T, K=K, nT=nT, R=20_000, compute_device_pool=pool, **model)   # :238-241
        t0 = time.perf_counter()
        p = simulate_and_compute_option_price_gpu(T=T, K=K, nT=nT, R=R, compute_device_pool=pool, **model)
        pool.sync()
        dt = time.perf_counter() - t0
        print(f"ComputeDevicePool, {g} GPU(s)      : {p:.6f}   {dt * 1e3:8.3f} ms   {R * (nT - 1) / dt:.3e} path-steps/s")
        pool.close()


if __name__ == "__main__":
    main()
