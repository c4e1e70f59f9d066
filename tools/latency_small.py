"""End-to-end latency of small pricing calls through the Python entry point (BASELINE C1 size and neighbours)."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import device, heston  # noqa: E402

P = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019, T=1.0, K=0.95)
device.init(device_id=0)
for R, nT in ((1_000, 101), (100_000, 101), (1_000_000, 101), (2_000_000, 500)):
    for _ in range(20):
        heston.simulate_and_compute_option_price(nT=nT, R=R, **P)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        heston.simulate_and_compute_option_price(nT=nT, R=R, **P)
    dt = (time.perf_counter() - t0) / reps
    print(f"R={R:>9} nT={nT}: {dt * 1e6:8.1f} us per call  {R * (nT - 1) / dt:.3e} path-steps/s")
