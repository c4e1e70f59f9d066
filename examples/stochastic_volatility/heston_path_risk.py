"""Quantifying risk from whole paths (the "quantifying risk" section of the reference's notebook,
notebooks/Getting Started.ipynb cells 46-64, taken from terminal values to full trajectories):
simulate the index level AND its variance at every date (the path-materialising mode of the fused kernel, 8 bytes per
path-step written with TMA tensor stores), then read value-at-risk, expected shortfall and the worst drawdown off them.

python examples/stochastic_volatility/heston_path_risk.py [paths] [dates]
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cocos_b200.numerics as cn                                    # noqa: E402
from cocos_b200.heston import simulate_heston_model                 # noqa: E402


def main():
    R = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000
    nT = int(sys.argv[2]) if len(sys.argv) > 2 else 253              # daily dates over one year
    cn.random.seed(7)
    t0 = time.perf_counter()
    x_T, v_T, x_path, v_path = simulate_heston_model(
        T=1.0, N=nT, R=R, mu=math.log(1.0319), kappa=6.21, v_bar=0.019, sigma_v=0.61, rho=-0.7, x0=0.0, v0=0.101 ** 2,
        return_trajectory=True, return_variance_trajectory=True)
    paths = x_path.to_numpy()[:, :R]                                  # (dates, paths): log index level
    variances = v_path.to_numpy()[:, :R]
    dt = time.perf_counter() - t0
    print(f"{R} paths x {nT - 1} steps simulated and {2 * paths.nbytes / 1e9:.2f} GB of (x, v) trajectories brought to the "
          f"host in {dt * 1e3:.0f} ms")
    level = np.exp(paths.astype(np.float64))
    ret = level[-1] - 1.0
    for q in (0.05, 0.01):
        var = -np.quantile(ret, q)
        es = -ret[ret <= -var].mean()
        print(f"1-year {100 * (1 - q):.0f}% value-at-risk {100 * var:.2f}%  expected shortfall {100 * es:.2f}%")
    running_max = np.maximum.accumulate(level, axis=0)
    drawdown = (1.0 - level / running_max).max(axis=0)
    print(f"worst drawdown along the path: median {100 * np.median(drawdown):.2f}%, 99th percentile "
          f"{100 * np.quantile(drawdown, 0.99):.2f}%")
    realised_vol = np.sqrt(variances.astype(np.float64).mean(axis=0))
    print(f"realised volatility per path: mean {100 * realised_vol.mean():.2f}%, 5%-95% band "
          f"{100 * np.quantile(realised_vol, 0.05):.2f}%-{100 * np.quantile(realised_vol, 0.95):.2f}% "
          f"(long-run level sqrt(v_bar) = {100 * math.sqrt(0.019):.2f}%)")
    assert np.array_equal(paths[-1], x_T.to_numpy())                  # the last row is the terminal state


if __name__ == "__main__":
    main()
