"""Sweep the fused kernel's launch geometry on one B200 (path-steps/s per configuration).

python tools/tune_fused.py [R] [nT] [out.json]
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402

P = dict(T=1.0, mu=0.03140167, kappa=6.21, v_bar=0.019, sigma_v=0.61, rho=-0.7, x0=0.0, v0=0.101 ** 2)


def time_config(ctx, R, nT, sfx, strikes, threads, bps, ppt, reps=3, variant=0):
    L = _lib.lib()
    ctx.set_launch(threads, bps, ppt)
    ctx.set_variant(variant)
    p = _lib.HestonParams(P["T"], P["mu"], P["kappa"], P["v_bar"], P["sigma_v"], P["rho"], P["x0"], P["v0"])
    k_arr, nK = _lib.strikes_array(strikes)
    launch = getattr(L, f"cb_heston_price_launch_{sfx}")
    sums, sq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    best = 1e30
    for rep in range(reps + 1):
        ctx.timer_start()
        _lib.check(launch(ctx.handle, ctypes.byref(p), R, nT, k_arr, nK, 0, rep, 0, (R + 1) // 2, None, None, None, None))
        ms = ctx.timer_stop()
        if rep:
            best = min(best, ms)
    _lib.check(L.cb_heston_price_finish(ctx.handle, nK, sums, sq))
    return best, sums[0] / R


def main():
    R = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    nT = int(sys.argv[2]) if len(sys.argv) > 2 else 501
    out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/tune_fused.json"
    mode = sys.argv[4] if len(sys.argv) > 4 else "variants"
    device.init(device_id=0)
    ctx = device.current_context()
    rows = []

    def run(sfx, strikes, Rk, threads, ppt, bps, variant):
        try:
            ms, mean = time_config(ctx, Rk, nT, sfx, strikes, threads, bps, ppt, variant=variant)
        except Exception as e:          # noqa: BLE001
            print("skip", sfx, threads, ppt, bps, variant, e)
            return
        rate = Rk * (nT - 1) / (ms * 1e-3)
        rows.append(dict(dtype=sfx, nK=len(strikes), threads=threads, ppt=ppt, bps=bps, variant=variant, ms=ms,
                         rate=rate, mean_payoff=mean))
        print(f"{sfx} nK={len(strikes):2d} v={variant} thr={threads} ppt={ppt} bps={bps:2d}: {ms:8.3f} ms  "
              f"{rate:.3e} path-steps/s  mean payoff {mean:.5f}", flush=True)

    if mode == "variants":
        for variant in (0, 1, 2):
            for threads in (128, 256):
                for ppt in (1, 2, 4):
                    for bps in (0, 8, 12, 16, 32, 64):
                        if threads == 256 and bps in (12, 64):
                            continue
                        run("f32", [0.95], R, threads, ppt, bps, variant)
    else:
        for sfx, strikes in (("f32", [0.95]), ("f32", [0.7 + 0.6 * i / 63 for i in range(64)]), ("f64", [0.95]),
                             ("f64", [0.7 + 0.6 * i / 63 for i in range(64)])):
            Rk = R if sfx == "f32" else R // 4
            for threads in (128, 256):
                for ppt in (1, 2, 4):
                    for bps in (0, 8, 16, 32):
                        run(sfx, strikes, Rk, threads, ppt, bps, 0)
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
