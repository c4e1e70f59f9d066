"""Benchmark of the Heston Monte-Carlo hot path (BASELINE.json metric: path-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pricing pass: R antithetic Heston paths x (nT-1) Euler steps reduced to
a discounted call price.  Workload at N=1: BASELINE config C3 (10M paths x 500 steps,
float32, fused in-kernel Philox).  For N>1 (launched by torchrun, one rank per GPU) the
per-GPU work is fixed (weak scaling): the job simulates N x 10M paths, sharded by
antithetic pair range, and one NCCL allreduce per step combines the payoff sums.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events on the kernel's
stream, max over ranks) with everything resident; `e2e` is the same metric through the
public Python entry point (host scalars in, host float out, wall clock around it);
`roofline` is against the MUFU (SFU) pipe peak measured live by a calibration kernel;
`cpu_baseline` / `--impl reference` time the NumPy restatement of the reference's CPU
path (oracle/, test infrastructure -- the reference tree itself is not on the GPU box).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019,
              T=1.0, K=0.95)                      # heston_pricing_benchmark.py:163-173
PATHS_PER_GPU = 10_000_000
N_T = 501                                          # 500 Euler steps
METRIC = "heston_mc_path_steps_per_sec"
UNIT = "path-steps/s"
# algorithmic work per path-step: 24 FP32 flop, 3 SFU ops, 25 INT ops, 0 DRAM bytes
FLOP_PER_PATH_STEP, SFU_PER_PATH_STEP = 24, 3


# --------------------------------------------------------------------------- CPU arm

def _cpu_batch(args):
    """One batch of the reference's NumPy path (restated in oracle/heston_numpy.py)."""
    seed, R, nT = args
    import numpy as np
    from oracle import heston_numpy as hn
    np.random.seed(seed)
    return hn.price(nT=nT, R=R, **PARAMS)


class CpuArm:
    """map_reduce_multicore semantics (multi_core_batch_processing.py:10-94) over all host cores."""

    def __init__(self, batch_paths=100_000, nT=N_T):
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing
        self.cores = os.cpu_count() or 1
        self.batch_paths, self.nT = batch_paths, nT
        self.pool = ProcessPoolExecutor(max_workers=self.cores, mp_context=multiprocessing.get_context("spawn"))
        list(self.pool.map(_cpu_batch, [(i, 1000, 11) for i in range(self.cores)]))     # start the workers

    def step(self, batches, seed0=0):
        t0 = time.perf_counter()
        prices = list(self.pool.map(_cpu_batch, [(seed0 + i, self.batch_paths, self.nT) for i in range(batches)]))
        dt = time.perf_counter() - t0
        return batches * self.batch_paths * (self.nT - 1) / dt, dt, sum(prices) / len(prices)

    def describe(self, batches):
        return (f"{batches} batches x {self.batch_paths} paths x {self.nT - 1} steps, NumPy restatement of the "
                f"reference path in {self.cores} worker processes")

    def close(self):
        self.pool.shutdown()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    batches = max(arm.cores, 8)
    for w in range(min(args.warmup, 1)):
        arm.step(batches, seed0=10_000 + w)
    t0 = time.perf_counter()
    work = 0
    for k in range(args.steps):
        _, _, _ = arm.step(batches, seed0=1000 * (k + 1))
        work += batches * arm.batch_paths * (arm.nT - 1)
    total = time.perf_counter() - t0
    value = work / total
    arm.close()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3 Heston European call, 10M paths x 500 steps fp32 (CPU arm: bounded sample of "
                               f"{batches * arm.batch_paths} paths per step)",
                   "paths": batches * arm.batch_paths, "nT": arm.nT, "antithetic": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": arm.describe(batches) + f", {args.steps} steps", "cpu": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm

class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except OSError:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)

    def summary(self):
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [c for c in sm if c > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


def run_b200(args):
    import ctypes
    import torch
    from cocos_b200 import _lib, distributed as cdist, heston
    from cocos_b200.device import ComputeDeviceManager

    rank, world, local_rank = cdist.env_rank_world()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs one rank per GPU: launch with python -m torch.distributed.run "
                             f"--nproc-per-node {args.gpus} bench.py --gpus {args.gpus} ...")
    torch.cuda.set_device(local_rank)
    ComputeDeviceManager.set_compute_device(local_rank)
    unique_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        unique_id = cdist.broadcast_unique_id()
    pricer = cdist.ShardedPricer(local_rank, rank, world, unique_id)
    ctx = pricer.ctx
    L = _lib.lib()

    R = PATHS_PER_GPU * world
    p = _lib.HestonParams(PARAMS["T"], PARAMS["r"], PARAMS["kappa"], PARAMS["v_bar"], PARAMS["sigma_v"],
                          PARAMS["rho"], PARAMS["x0"], PARAMS["v0"])
    if args.threads or args.blocks_per_sm or args.ppt:
        ctx.set_launch(args.threads, args.blocks_per_sm, args.ppt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")   # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- pipe peaks, measured live (rank-local; rank 0's are reported)
    def peak(kind):
        ops, ms = ctypes.c_double(0), ctypes.c_float(0)
        _lib.check(L.cb_peak_measure(ctx.handle, kind, 3, ctypes.byref(ops), ctypes.byref(ms)))
        return ops.value
    mufu_peak = min(peak(k) for k in (1, 2, 3))          # lg2, sqrt, sin: thread-level MUFU ops/s
    ffma_peak = peak(0)

    # ---- device-resident timing: K launches (+ allreduce), CUDA events on the kernel's stream
    sub = 0
    for _ in range(max(args.warmup, 3)):
        nK = pricer.launch(p, R, N_T, [PARAMS["K"]], args.seed, sub)
        sub += 1
    sums, sumsq = pricer.finish(nK)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    dev_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()                                     # L2 flush between timed steps (untimed)
        torch.cuda.synchronize()
        ctx.timer_start()
        nK = pricer.launch(p, R, N_T, [PARAMS["K"]], args.seed, sub)
        dev_ms += ctx.timer_stop()
        sub += 1
    barrier()
    sums, sumsq = pricer.finish(nK)
    price, se = heston._price_from_sums(PARAMS["r"], PARAMS["T"], R, sums, sumsq)

    # ---- informational: the opt-in conditional-Gaussian scheme (same law of terminal prices, half the normals);
    #      NOT the headline: `value` above is the reference's step-by-step Euler scheme
    cond_ms = 0.0
    for i in range(3 + 5):
        pricer.ctx.timer_start()
        nK = pricer.launch(p, R, N_T, [PARAMS["K"]], args.seed, 1_000_000 + i, scheme="conditional")
        ms = pricer.ctx.timer_stop()
        if i >= 3:
            cond_ms += ms
    csums, csumsq = pricer.finish(nK)
    cond_price, cond_se = heston._price_from_sums(PARAMS["r"], PARAMS["T"], R, csums, csumsq)

    # ---- end to end: the public entry point, host scalars in, host float out, wall clock
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if world == 1:
            e2e_price = heston.simulate_and_compute_option_price(nT=N_T, R=R, **PARAMS)
        else:
            e2e_price = pricer.price(nT=N_T, R=R, seed=args.seed, subsequence=sub, **PARAMS)
            sub += 1
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop()
    clocks = sampler.summary()

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dev_ms, e2e_s, cond_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, cond_ms = float(t[0]), float(t[1]), float(t[2])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm()
        batches = 2 * arm.cores
        v, dt, cpu_price = arm.step(batches, seed0=777)
        t1 = time.perf_counter()
        _cpu_batch((778, arm.batch_paths, arm.nT))                       # one batch in this process: 1 core
        one_core = arm.batch_paths * (arm.nT - 1) / (time.perf_counter() - t1)
        cpu = {"value": v, "unit": UNIT, "cores": arm.cores, "kind": "port",
               "sample": arm.describe(batches) + f" ({dt:.1f} s)", "cpu": cpu_model(), "price": cpu_price,
               "value_1core": one_core}
        arm.close()

    if rank == 0:
        work = R * (N_T - 1) * args.steps
        value = work / (dev_ms * 1e-3)
        per_gpu = value / world
        sfu_roof = mufu_peak / SFU_PER_PATH_STEP
        fp32_roof = 2 * ffma_peak / FLOP_PER_PATH_STEP
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3 Heston European call, 10M paths x 500 steps fp32, fused in-kernel Philox"
                                   + (f", x{world} GPUs sharded by antithetic pair range, cross-GPU sum "
                                      + ("fused into the kernel over NVLink peer memory" if pricer.peer_exchange
                                         else "by 1 NCCL allreduce/step")
                                      if world > 1 else ", single B200"),
                       "paths": R, "paths_per_gpu": PATHS_PER_GPU, "nT": N_T, "antithetic": True,
                       "parallelism": f"pair-shard x{world}",
                       "collective": ("none" if world == 1 else
                                      "fused-peer-exchange" if pricer.peer_exchange else "nccl-allreduce"),
                       "l2": "256 MiB memset between timed steps (the kernel itself reads no memory input)",
                       "launch": {"threads": args.threads, "blocks_per_sm": args.blocks_per_sm, "ppt": args.ppt}},
            "price": float(price[0]), "price_se": float(se[0]),
            "e2e": {"value": work / e2e_s, "unit": UNIT, "h2d_bytes_per_step": ctypes.sizeof(_lib.HestonParams) + 8,
                    "d2h_bytes_per_step": 16, "price": e2e_price,
                    "api": "cocos_b200.heston.simulate_and_compute_option_price" if world == 1
                           else "cocos_b200.distributed.ShardedPricer.price"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "sfu", "achieved": per_gpu * SFU_PER_PATH_STEP / 1e9, "peak": mufu_peak / 1e9,
                         "unit": "GSFUop/s per GPU", "frac": per_gpu / sfu_roof,
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture
                         # (profiles/r01_ncu_fused_f32.md): the kernel's algorithmic DRAM traffic is 0 + 16 B of sums
                         "traffic": 81920,
                         "peak_source": "MUFU chain kernels (lg2/sqrt/sin, min) timed live on this GPU; "
                                        "MEASURED_PEAKS.json has no SFU/FP32 figure",
                         "kernel": "heston_fused_kernel<float>", "sfu_ops_per_path_step": SFU_PER_PATH_STEP,
                         "fp32_frac": per_gpu / fp32_roof, "fp32_peak_tflops": 2 * ffma_peak / 1e12,
                         "roof_path_steps_per_s": min(sfu_roof, fp32_roof)},
            "clocks": clocks,
            "extras": {"conditional_scheme": {"value": R * (N_T - 1) * 5 / (cond_ms * 1e-3), "unit": UNIT,
                                              "price": float(cond_price[0]), "price_se": float(cond_se[0]),
                                              "note": "opt-in scheme='conditional': log price sampled conditionally on "
                                                      "the variance path, same law as the Euler scheme, 5 launches"}},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    pricer.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--ppt", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
