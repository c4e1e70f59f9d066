"""The notebook's "Quantifying Risk" workflow (notebooks/Getting Started.ipynb cells 46-83) on B200: simulate the
Heston model, then estimate the density of the terminal index level with a Gaussian KDE -- without leaving the device.

    python examples/kde/heston_terminal_price_density.py
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from cocos_b200.heston import simulate_heston_model                 # noqa: E402
import cocos_b200.numerics as cn                                    # noqa: E402
from cocos_b200.scientific import gaussian_kde                      # noqa: E402


def main():
    R, nT = 1_000_000, 500
    t0 = time.perf_counter()
    x, v = simulate_heston_model(T=1.0, N=nT, R=R, mu=math.log(1.0319), kappa=6.21, v_bar=0.019, sigma_v=0.61, rho=-0.7,
                                 x0=0.0, v0=0.101 ** 2)
    x._ctx.sync()
    print(f"simulated {R} paths x {nT - 1} steps in {(time.perf_counter() - t0) * 1e3:.2f} ms")
    prices = cn.exp(x)
    kde = gaussian_kde(prices)
    grid = np.linspace(0.4, 1.8, 1000)
    t0 = time.perf_counter()
    density = kde.evaluate(grid).to_numpy()
    print(f"KDE of {R} prices on {grid.size} points in {(time.perf_counter() - t0) * 1e3:.2f} ms; "
          f"integral {np.trapezoid(density, grid):.5f}, mode at {grid[density.argmax()]:.3f}")
    q = np.cumsum(density) * (grid[1] - grid[0])
    print(f"5% / 1% quantiles of the index level from the density: {grid[np.searchsorted(q, 0.05)]:.3f} / "
          f"{grid[np.searchsorted(q, 0.01)]:.3f}")


if __name__ == "__main__":
    main()
