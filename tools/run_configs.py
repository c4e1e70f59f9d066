"""BASELINE configs C4 (1B paths x 1000 steps f32) and C5 (100M x 252 f64, 64 strikes) on 1..8 GPUs.

python tools/run_configs.py                      (1 GPU)
python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py [--scale 1.0]
Prints one JSON line per config on rank 0: path-steps/s (max over ranks, CUDA events), price +- SE, and the
3-SE check against the reference prices in tests/golden/reference_values.json.
"""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the configured path counts")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from cocos_b200 import _lib, distributed as cdist, heston
    from cocos_b200.device import ComputeDeviceManager

    rank, world, local = cdist.env_rank_world()
    torch.cuda.set_device(local)
    ComputeDeviceManager.set_compute_device(local)
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = cdist.broadcast_unique_id()
    pricer = cdist.ShardedPricer(local, rank, world, uid)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_values.json")))
    q = golden["params"]
    p = _lib.HestonParams(q["T"], q["r"], q["kappa"], q["v_bar"], q["sigma_v"], q["rho"], q["x0"], q["v0"])
    configs = [
        ("C4 Heston 1B paths x 1000 steps f32 antithetic", "f32", int(1_000_000_000 * args.scale), 1001, [q["K"]], "nT1001"),
        ("C5 Heston 100M paths x 252 steps f64, 64 strikes", "f64", int(100_000_000 * args.scale), 253,
         golden["statistical"]["strikes"], "nT253"),
        ("C3 Heston 10M paths x 500 steps f32 (strong-scaled over the ranks)", "f32", 10_000_000, 501, [q["K"]], "nT501"),
    ]
    for name, sfx, R, nT, strikes, key in configs:
        def sync():
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            pricer.ctx.sync()
        nK = pricer.launch(p, R, nT, strikes, 0, 0, sfx)          # warm-up
        pricer.finish(nK)
        best = 1e30
        for rep in range(args.reps):
            sync()
            pricer.ctx.timer_start()
            nK = pricer.launch(p, R, nT, strikes, 0, 1 + rep, sfx)
            ms = pricer.ctx.timer_stop()
            if world > 1:
                import torch.distributed as dist
                t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            best = min(best, ms)
        sums, sumsq = pricer.finish(nK)
        price, se = heston._price_from_sums(q["r"], q["T"], R, sums, sumsq)
        ref = golden["statistical"][key]
        if len(strikes) == 1:
            ok = abs(price[0] - ref["price"]) <= 3 * math.hypot(se[0], ref["se"])
            shown = {"price": float(price[0]), "se": float(se[0]), "reference_price": ref["price"], "reference_se": ref["se"]}
        else:
            dev = np.abs(price - np.array(ref["sweep"]))
            ok = bool(np.all(dev <= 3 * np.hypot(se, 2 * ref["se"]) + 1e-7))
            shown = {"price_K0.70": float(price[0]), "price_K1.30": float(price[-1]), "max_abs_dev_vs_reference": float(dev.max())}
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "paths": R, "steps": nT - 1, "dtype": sfx, "ms": best,
                              "path_steps_per_s": R * (nT - 1) / best * 1e3, "within_3se_of_reference": bool(ok), **shown}),
                  flush=True)
    pricer.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
