"""Systematic error of the float32 / MUFU arithmetic of the fused kernel at the precision BASELINE config C4 claims.

python tools/bias_f32.py [paths_total=1e8] [nT=1001] [oracle_paths=1e6] [out.json]

(1) float32 kernel against the float64 kernel on the SAME Philox stream (same seed, subsequence, pair indices), in
    chunks of 1e7 paths: the paired per-path payoff difference d_i = payoff(x_T^f32) - payoff(x_T^f64) has a tiny
    variance, so its mean (the bias of the float32 state arithmetic and of MUFU.SQRT on v*r^2) is resolved far below
    the Monte-Carlo standard error of the price itself.  The kernels' own payoff sums (which include the float32
    exp of the payoff) are compared as well.
(2) float32 kernel against the oracle's float64 model of the same stream with exact libm Box-Muller
    (oracle/heston_numpy.py: fused_model) -- this also sees the MUFU.LG2 / SIN / COS approximation of the normals.
(3) float64 kernel against the same exact model (its normals and diffusion root are single-precision grade too).
Reported: mean difference +- its standard error (paired), in price units (discounted).
"""
import ctypes
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402
from cocos_b200.numerics._array import empty_on  # noqa: E402

P = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019, T=1.0, K=0.95)


def terminal(ctx, sfx, dtype, R, nT, seed, sub):
    L = _lib.lib()
    hp = _lib.HestonParams(P["T"], P["r"], P["kappa"], P["v_bar"], P["sigma_v"], P["rho"], P["x0"], P["v0"])
    k_arr, nK = _lib.strikes_array([P["K"]])
    H = (R + 1) // 2
    x, v = empty_on(ctx, (2 * H,), dtype), empty_on(ctx, (2 * H,), dtype)
    _lib.check(getattr(L, f"cb_heston_price_launch_{sfx}")(ctx.handle, ctypes.byref(hp), R, nT, k_arr, nK, seed, sub, 0, H,
                                                           None, x.data_ptr, v.data_ptr, None))
    s, q = (ctypes.c_double * 1)(), (ctypes.c_double * 1)()
    _lib.check(L.cb_heston_price_finish(ctx.handle, nK, s, q))
    return x.to_numpy()[:R].astype(np.float64), s[0]


def payoff(x):
    return np.maximum(np.exp(x) - P["K"], 0.0)


def paired(ctx, R_total, nT, chunk=10_000_000, seed=7):
    disc = math.exp(-P["r"] * P["T"])
    n = s1 = s2 = 0.0
    ksum32 = ksum64 = 0.0
    max_dx = 0.0
    for c in range(max(1, R_total // chunk)):
        x32, k32 = terminal(ctx, "f32", np.float32, chunk, nT, seed, c)
        x64, k64 = terminal(ctx, "f64", np.float64, chunk, nT, seed, c)
        d = payoff(x32) - payoff(x64)
        n += d.size
        s1 += d.sum()
        s2 += (d * d).sum()
        ksum32 += k32
        ksum64 += k64
        max_dx = max(max_dx, float(np.max(np.abs(x32 - x64))))
    mean = s1 / n
    se = math.sqrt(max(s2 / n - mean * mean, 0.0) / n)
    return {"paths": int(n), "nT": nT, "mean_payoff_difference_f32_minus_f64": disc * mean, "standard_error": disc * se,
            "kernel_sums_difference_per_path": disc * (ksum32 - ksum64) / n, "max_abs_x_difference": max_dx,
            "price_f64": disc * ksum64 / n}


def against_oracle(ctx, R, nT, seed=7, sub=900):
    from oracle import heston_numpy as hn
    disc = math.exp(-P["r"] * P["T"])
    x32, _ = terminal(ctx, "f32", np.float32, R, nT, seed, sub)
    t0 = time.perf_counter()
    xm, _ = hn.fused_model(P["x0"], P["v0"], P["r"], P["rho"], P["sigma_v"], P["kappa"], P["v_bar"], P["T"], nT, R,
                           seed=seed, subsequence=sub, dtype=np.float64)
    seconds = time.perf_counter() - t0
    xm = np.asarray(xm, dtype=np.float64)
    d = payoff(x32) - payoff(xm)
    res = {"paths": R, "nT": nT, "mean_payoff_difference_f32_minus_exact_model": disc * float(d.mean()),
           "standard_error": disc * float(d.std() / math.sqrt(d.size)),
           "max_abs_x_difference": float(np.max(np.abs(x32 - xm))), "oracle_seconds": seconds}
    # (3) the float64 kernel (float64 state, single-precision-grade normals and diffusion root) against the same model
    x64, _ = terminal(ctx, "f64", np.float64, R, nT, seed, sub)
    d64 = payoff(x64) - payoff(xm)
    res["f64_kernel"] = {"mean_payoff_difference_f64_kernel_minus_exact_model": disc * float(d64.mean()),
                         "standard_error": disc * float(d64.std() / math.sqrt(d64.size)),
                         "max_abs_x_difference": float(np.max(np.abs(x64 - xm))),
                         "mean_x_difference": float(np.mean(x64 - xm)),
                         "mean_abs_x_difference": float(np.mean(np.abs(x64 - xm)))}
    return res


def main():
    R_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    nT = int(sys.argv[2]) if len(sys.argv) > 2 else 1001
    R_oracle = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1_000_000
    out = sys.argv[4] if len(sys.argv) > 4 else "gpurun_out/bias_f32.json"
    device.init(device_id=0)
    ctx = device.current_context()
    res = {"params": P, "c4_standard_error": 1.9e-6,
           "f32_vs_f64_same_stream": paired(ctx, R_total, nT),
           "f32_vs_exact_float64_model": against_oracle(ctx, R_oracle, nT)}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
