"""A small invocation of every reduction-bearing kernel (for compute-sanitizer racecheck / memcheck)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import device, heston, monte_carlo_pi  # noqa: E402
from cocos_b200.numerics import random as rnd  # noqa: E402
from cocos_b200.scientific import gaussian_kde  # noqa: E402

P = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019, T=1.0)
device.init(device_id=0)
rnd.seed(1)
print("price", heston.price_with_standard_error(K=0.95, nT=33, R=70_001, **P))
print("sweep", heston.price_strike_sweep(strikes=[0.8 + 0.01 * i for i in range(40)], nT=17, R=30_000, **P)[0][:3])
print("f32 sweep", heston.price_strike_sweep(strikes=[0.9, 1.0], nT=17, R=30_000, dtype=np.float32, **P)[0])
print("conditional", heston.price_with_standard_error(K=0.95, nT=33, R=70_001, scheme="conditional", **P))
print("asian", heston.price_path_dependent(K=0.95, nT=17, R=30_000, kind="asian", **P))
x, v, traj = heston.simulate_heston_model(T=1.0, N=9, R=20_000, mu=P["r"], kappa=P["kappa"], v_bar=P["v_bar"],
                                          sigma_v=P["sigma_v"], rho=P["rho"], x0=0.0, v0=P["v0"], return_trajectory=True)
print("payoff from array", heston.compute_option_price_from_simulated_paths(P["r"], 1.0, 0.95, x))
print("pi", monte_carlo_pi.estimate_pi(2_000_001, antithetic=True))
k = gaussian_kde(rnd.randn(20_000))
print("kde", k.evaluate(np.linspace(-3, 3, 50)).to_numpy()[:3])
print("gamma", rnd.gamma(0.5, 1.0, 10_000).to_numpy().mean(), "randn_antithetic", rnd.randn_antithetic((1001, 2), 0).shape)
