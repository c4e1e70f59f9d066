"""Monte-Carlo estimate of pi on 1..G B200s: examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py
(reference lines 18-33, 185-188, 250-252: n = 1e9 points, 20 batches) for cocos_b200.

    python examples/monte_carlo_pi/monte_carlo_pi_multi_gpu_example.py [--points 1000000000]
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import cocos_b200.device as cd                                     # noqa: E402
from cocos_b200.monte_carlo_pi import estimate_pi                  # noqa: E402
from cocos_b200.multi_processing import ComputeDevicePool          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1_000_000_000)
    ap.add_argument("--batches", type=int, default=20)
    args = ap.parse_args()
    n, batches = args.points, args.batches
    estimate_pi(1_000_000)                                        # warm-up
    for label, kw in (("plain uniforms", {}), ("antithetic uniforms", {"antithetic": True})):
        t0 = time.perf_counter()
        pi = estimate_pi(n, batches=batches, **kw)
        dt = time.perf_counter() - t0
        print(f"1 GPU, {label:20s}: pi ~ {pi:.6f} (error {abs(pi - math.pi):.1e})  {dt * 1e3:8.2f} ms  {n / dt:.3e} points/s")
    n_gpus = cd.ComputeDeviceManager.get_number_of_compute_devices()
    for g in range(2, n_gpus + 1):
        pool = ComputeDevicePool(compute_devices=list(range(g)))
        # the reference's call pattern: whole batches mapped over the pool and averaged on the host
        t0 = time.perf_counter()
        pi = pool.map_reduce(f=lambda: estimate_pi(n=math.ceil(n / g)), reduction=lambda x, y: x + y / g,
                             initial_value=0.0, number_of_batches=g)
        dt = time.perf_counter() - t0
        print(f"{g} GPUs, map_reduce of batches   : pi ~ {pi:.6f}  {dt * 1e3:8.2f} ms  {n / dt:.3e} points/s")
        t0 = time.perf_counter()
        pi = estimate_pi(n, batches=1, antithetic=True, compute_device_pool=pool)      # sharded draws + one allreduce
        dt = time.perf_counter() - t0
        print(f"{g} GPUs, sharded + NCCL allreduce: pi ~ {pi:.6f}  {dt * 1e3:8.2f} ms  {n / dt:.3e} points/s")
        pool.close()


if __name__ == "__main__":
    main()
