"""Secondary kernels on one B200: achieved rates against their rooflines (HBM for the materialising kernels).

python tools/bench_aux.py [out.json]
"""
import ctypes
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device, heston, monte_carlo_pi  # noqa: E402
from cocos_b200.numerics import random as rnd  # noqa: E402
from cocos_b200.numerics._array import empty_on  # noqa: E402

P = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019, T=1.0)
HP = _lib.HestonParams(P["T"], P["r"], P["kappa"], P["v_bar"], P["sigma_v"], P["rho"], P["x0"], P["v0"])


def timed(ctx, fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        ctx.timer_start()
        fn()
        best = min(best, ctx.timer_stop())
    return best


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/aux.json"
    device.init(device_id=0)
    ctx = device.current_context()
    L = _lib.lib()
    res = {}

    def peak(kind):
        ops, ms = ctypes.c_double(0), ctypes.c_float(0)
        _lib.check(L.cb_peak_measure(ctx.handle, kind, 3, ctypes.byref(ops), ctypes.byref(ms)))
        return ops.value
    hbm_w, hbm_c, dfma, philox = peak(8), peak(10), peak(6), peak(7)
    res["peaks"] = {"hbm_write_GBs": hbm_w / 1e9, "hbm_copy_GBs": hbm_c / 1e9, "dfma_per_s": dfma,
                    "philox_calls_per_s": philox}

    # ---- random arrays (HBM-write bound)
    n = 1 << 30
    for name, dtype, item in (("randn_f32", np.float32, 4), ("rand_f32", np.float32, 4), ("randn_f64", np.float64, 8),
                              ("rand_f64", np.float64, 8)):
        m = n if item == 4 else n // 2
        buf = empty_on(ctx, (m,), dtype)
        fn = getattr(L, "cb_" + name)
        ms = timed(ctx, lambda: _lib.check(fn(ctx.handle, buf.data_ptr, m, 1, 0)))
        res[name] = {"elements": m, "ms": ms, "GBs": m * item / ms / 1e6, "frac_hbm_write": m * item / ms / 1e6 / (hbm_w / 1e9)}
        del buf
    R = 200_000_000
    buf = empty_on(ctx, (R, 2), np.float32)
    dims = (ctypes.c_uint64 * 2)(R, 2)
    ms = timed(ctx, lambda: _lib.check(L.cb_randn_antithetic_f32(ctx.handle, buf.data_ptr, dims, 2, 0, 1, 0)))
    res["randn_antithetic_f32_(2e8,2)"] = {"ms": ms, "GBs": R * 8 / ms / 1e6, "frac_hbm_write": R * 8 / ms / 1e6 / (hbm_w / 1e9)}
    del buf

    # ---- parity kernel (HBM-read bound): BASELINE C1 and a larger instance
    for Rr, N in ((100_000, 101), (4_000_000, 101)):
        H = (Rr + 1) // 2
        Z = empty_on(ctx, (N - 1, H, 2), np.float32)
        _lib.check(L.cb_randn_f32(ctx.handle, Z.data_ptr, Z.size, 3, 0))
        x, v = empty_on(ctx, (Rr,), np.float32), empty_on(ctx, (Rr,), np.float32)
        ms = timed(ctx, lambda: _lib.check(L.cb_heston_terminal_injected_f32(ctx.handle, Z.data_ptr, ctypes.byref(HP), Rr, N, x.data_ptr, v.data_ptr)))
        byts = Z.size * 4
        res[f"injected_f32_R{Rr}_N{N}"] = {"ms": ms, "path_steps_per_s": Rr * (N - 1) / ms * 1e3, "GBs_read": byts / ms / 1e6,
                                           "frac_hbm": byts / ms / 1e6 / (hbm_c / 1e9)}
        del Z, x, v

    # ---- trajectory (path-materialising) mode: 4 B written per path-step
    for Rr, N in ((4_000_000, 101), (2_000_000, 501)):
        rnd.seed(0)
        t0 = time.perf_counter()
        x, v, traj = heston.simulate_heston_model(T=P["T"], N=N, R=Rr, mu=P["r"], kappa=P["kappa"], v_bar=P["v_bar"],
                                                  sigma_v=P["sigma_v"], rho=P["rho"], x0=P["x0"], v0=P["v0"],
                                                  return_trajectory=True)
        ctx.sync()
        k_arr, nK = _lib.strikes_array([0.95])
        pairs = Rr // 2

        def go():
            _lib.check(L.cb_heston_price_launch_f32(ctx.handle, ctypes.byref(HP), Rr, N, k_arr, nK, 0, 5, 0, pairs, None,
                                                    x.data_ptr, v.data_ptr, traj.data_ptr))
        ms = timed(ctx, go)
        byts = N * Rr * 4
        res[f"trajectory_f32_R{Rr}_N{N}"] = {"ms": ms, "path_steps_per_s": Rr * (N - 1) / ms * 1e3, "GBs_written": byts / ms / 1e6,
                                             "frac_hbm_write": byts / ms / 1e6 / (hbm_w / 1e9)}
        del x, v, traj

    # ---- (x, v) path-materialising mode: 8 B written per path-step (the HBM-bound variant of SURVEY section 8d)
    for Rr, N in ((4_000_000, 101), (2_000_000, 253)):
        pairs = Rr // 2
        x, v = empty_on(ctx, (Rr,), np.float32), empty_on(ctx, (Rr,), np.float32)
        tx, tv = empty_on(ctx, (N, Rr), np.float32), empty_on(ctx, (N, Rr), np.float32)
        k_arr, nK = _lib.strikes_array([0.95])

        def go_xv():
            _lib.check(L.cb_heston_trajectory_launch_f32(ctx.handle, ctypes.byref(HP), Rr, N, k_arr, nK, 0, 5, 0, pairs,
                                                         None, x.data_ptr, v.data_ptr, tx.data_ptr, tv.data_ptr))
        ms = timed(ctx, go_xv)
        byts = 2 * N * Rr * 4
        res[f"trajectory_xv_f32_R{Rr}_N{N}"] = {"ms": ms, "path_steps_per_s": Rr * (N - 1) / ms * 1e3,
                                                "GBs_written": byts / ms / 1e6,
                                                "frac_hbm_write": byts / ms / 1e6 / (hbm_w / 1e9)}
        os.environ["COCOS_B200_NO_BULK"] = "1"
        ms = timed(ctx, go_xv)
        del os.environ["COCOS_B200_NO_BULK"]
        res[f"trajectory_xv_f32_R{Rr}_N{N}_direct_stores"] = {"ms": ms, "GBs_written": byts / ms / 1e6,
                                                              "frac_hbm_write": byts / ms / 1e6 / (hbm_w / 1e9)}
        del x, v, tx, tv

    # ---- fused variants: strike sweeps and float64 state (C5 shape per GPU: 12.5M paths x 252 steps, 64 strikes)
    strikes64 = [0.70 + 0.60 * i / 63 for i in range(64)]
    for name, sfx, Rr, N, ks in (("fused_f32_1strike_C3", "f32", 10_000_000, 501, [0.95]),
                                 ("fused_f32_64strikes", "f32", 10_000_000, 501, strikes64),
                                 ("fused_f32_C4_shard_1000steps", "f32", 10_000_000, 1001, [0.95]),
                                 ("fused_f64_1strike", "f64", 12_500_000, 253, [0.95]),
                                 ("fused_f64_64strikes_C5_shard", "f64", 12_500_000, 253, strikes64)):
        k_arr, nK = _lib.strikes_array(ks)
        launch = getattr(L, f"cb_heston_price_launch_{sfx}")

        def go():
            _lib.check(launch(ctx.handle, ctypes.byref(HP), Rr, N, k_arr, nK, 0, 9, 0, (Rr + 1) // 2, None, None, None, None))
        ms = timed(ctx, go)
        res[name] = {"ms": ms, "path_steps_per_s": Rr * (N - 1) / ms * 1e3}

    # ---- pi (BASELINE config 2)
    rnd.seed(0)
    for anti in (True, False):
        nn = 1_000_000_000
        ndraw = (nn + 1) // 2 if anti else nn

        def go():
            _lib.check(L.cb_pi_count_launch(ctx.handle, nn, 0, 1, int(anti), 0, ndraw, None))
        ms = timed(ctx, go)
        inside = ctypes.c_uint64(0)
        _lib.check(L.cb_pi_count_finish(ctx.handle, ctypes.byref(inside)))
        calls = ndraw / 2                      # one Philox4x32-10 block feeds two draws
        res[f"pi_1e9_{'antithetic' if anti else 'plain'}"] = {"ms": ms, "points_per_s": nn / ms * 1e3, "pi": 4 * inside.value / nn,
                                                              "philox_calls_per_s": calls / ms * 1e3,
                                                              "frac_philox_call_peak": calls / ms * 1e3 / philox}

    # ---- derived distributions (one fused kernel each) and the Gaussian KDE (MUFU.EX2 per pair)
    mufu = min(peak(4), peak(1))
    n_d = 1 << 28
    buf = empty_on(ctx, (n_d,), np.float32)
    for name, kind, a, b in (("uniform", 0, 2.0, 5.0), ("exponential", 1, 1.5, 0.0), ("normal", 2, 0.0, 1.0),
                             ("lognormal", 3, 0.0, 0.5), ("logistic", 4, 5.0, 2.0), ("wald", 5, 1.0, 1.0),
                             ("gamma_7.5", 6, 7.5, 1.0), ("gamma_0.5", 6, 0.5, 1.0), ("beta_2_5", 7, 2.0, 5.0)):
        ms = timed(ctx, lambda: _lib.check(L.cb_random_distribution_f32(ctx.handle, buf.data_ptr, n_d, kind, a, b, 1, 0)), reps=3)
        res["dist_" + name] = {"elements": n_d, "ms": ms, "elements_per_s": n_d / ms * 1e3, "GBs": n_d * 4 / ms / 1e6}
    del buf
    from cocos_b200.scientific.kde import gaussian_kde
    from cocos_b200.numerics._array import ndarray as dev_array
    rnd.seed(1)
    x, _ = heston.simulate_heston_model(T=P["T"], N=101, R=2_000_000, mu=P["r"], kappa=P["kappa"], v_bar=P["v_bar"],
                                        sigma_v=P["sigma_v"], rho=P["rho"], x0=P["x0"], v0=P["v0"])
    ctx.sync()
    import cocos_b200.numerics as cn
    prices = cn.exp(x)
    kde = gaussian_kde(prices)
    for m in (1000, 100_000):
        grid = np.linspace(0.3, 2.0, m)
        xi = empty_on(ctx, (m, 1), np.float32)
        w = np.ascontiguousarray((grid[:, None] * kde._matrix[0, 0]).astype(np.float32))
        _lib.check(L.cb_memcpy_h2d(ctx.handle, xi.data_ptr, w.ctypes.data, w.nbytes))
        out = empty_on(ctx, (m,), np.float32)
        ms = timed(ctx, lambda: _lib.check(L.cb_kde_evaluate_f32(ctx.handle, kde._whitened.data_ptr, None, kde.n, 1, xi.data_ptr, m,
                                                                  kde.normalization_constant / kde.n, out.data_ptr)), reps=3)
        pairs = kde.n * m
        res[f"kde_n2e6_m{m}"] = {"ms": ms, "pairs_per_s": pairs / ms * 1e3, "frac_mufu_peak": pairs / ms * 1e3 / mufu,
                                 "integral": float(np.trapezoid(out.to_numpy(), grid))}

    # ---- CPU pi baseline: the reference's estimate_pi semantics (NumPy, one core), 1e8 points in 20 batches
    from oracle import heston_numpy as hn
    np.random.seed(0)
    t0 = time.perf_counter()
    pi_cpu = hn.estimate_pi(100_000_000, batches=20)
    dt = time.perf_counter() - t0
    res["pi_cpu_numpy_1core_1e8"] = {"seconds": dt, "points_per_s": 1e8 / dt, "pi": pi_cpu}

    for k, v in res.items():
        print(k, json.dumps(v))
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
