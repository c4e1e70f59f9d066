"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/* from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Every array below is an output of the reference's own NumPy branch
(oracle/reference_runner.py explains the import stubs); nothing here comes from
the restatements, so the fixtures pin them.
"""
import json
import math
import os

import numpy as np

from .reference_runner import InjectedNormals, load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

PARAMS = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21,
              v_bar=0.019, T=1.0, K=0.95)          # heston_pricing_benchmark.py:163-173


def _simulate(ref, R, N, Z, params=PARAMS):
    with InjectedNormals(normals=[Z[s].astype(np.float64) for s in range(N - 1)]):
        return ref.simulate_heston_model(
            T=params["T"], N=N, R=R, mu=params["r"], kappa=params["kappa"], v_bar=params["v_bar"],
            sigma_v=params["sigma_v"], rho=params["rho"], x0=params["x0"], v0=params["v0"],
            numerical_package_bundle=ref.NumpyBundle)


def heston_injected(ref):
    """Terminal (x, v) of the reference for injected half-normals, float32."""
    rng = np.random.default_rng(20260101)
    cases = {}
    # np.dot((R,2),(2,)) is OpenBLAS sgemv: its rounding depends on the code path the row count selects
    # (here: fma(z1,m1,rn(z0*m0)) with scalar tails below ~16k rows, fma(z0,m0,rn(z1*m1)) above, where every
    # BASELINE config lives).  "large" pins the latter bit for bit; the small cases pin it to 1e-6.
    for name, (R, N) in {"even": (1000, 20), "odd": (1001, 13), "two": (2, 5), "one": (1, 3),
                         "long": (64, 501), "large": (20481, 6)}.items():
        H = (R + 1) // 2
        Z = rng.standard_normal((N - 1, H, 2)).astype(np.float32)
        x, v = _simulate(ref, R, N, Z)
        assert x.dtype == np.float32 and v.dtype == np.float32
        price = ref.compute_option_price_from_simulated_paths(
            r=PARAMS["r"], T=PARAMS["T"], K=PARAMS["K"], x_simulated=x,
            numerical_package_bundle=ref.NumpyBundle)
        cases[name + "_RN"] = np.array([R, N], dtype=np.int64)
        cases[name + "_Z"] = Z
        cases[name + "_x"] = x
        cases[name + "_v"] = v
        cases[name + "_price"] = np.array(price, dtype=np.float64)
    # a second parameter set: Feller satisfied, positive correlation, non-zero x0
    alt = dict(PARAMS, x0=0.25, v0=0.09, rho=0.3, sigma_v=0.2, kappa=2.0, v_bar=0.06, T=2.0, r=0.01)
    R, N = 500, 33
    Z = rng.standard_normal((N - 1, (R + 1) // 2, 2)).astype(np.float32)
    x, v = _simulate(ref, R, N, Z, alt)
    cases["alt_RN"] = np.array([R, N], dtype=np.int64)
    cases["alt_Z"], cases["alt_x"], cases["alt_v"] = Z, x, v
    cases["alt_params"] = np.array([alt[k] for k in
                                    ("T", "r", "kappa", "v_bar", "sigma_v", "rho", "x0", "v0")])
    np.savez_compressed(os.path.join(OUT, "heston_injected_f32.npz"), **cases)


def seeded_prices(ref):
    """Prices of the reference entry point with numpy's own seeded generator."""
    out = {}
    for seed in (0, 1, 2):
        np.random.seed(seed)
        out[f"C1_seed{seed}"] = ref.simulate_and_compute_option_price(
            nT=101, R=100000, numerical_package_bundle=ref.NumpyBundle, **PARAMS)
    np.random.seed(7)
    out["odd_R_seed7"] = ref.simulate_and_compute_option_price(
        nT=17, R=4999, numerical_package_bundle=ref.NumpyBundle, **PARAMS)
    return out


def statistical_prices(ref):
    """Reference price and pair-based SE at 1e6 paths for the 3-SE acceptance tests."""
    out = {}
    disc = math.exp(-PARAMS["r"] * PARAMS["T"])
    strikes = np.linspace(0.70, 1.30, 64)            # C5 strike grid
    for nT in (101, 253, 501, 1001):
        np.random.seed(1000 + nT)
        units, pay_sum, n_paths = [], 0.0, 0
        sweep = np.zeros(64)
        batches = 10 if nT <= 501 else 5
        for _ in range(batches):
            R = 100000
            x, _ = ref.simulate_heston_model(
                T=PARAMS["T"], N=nT, R=R, mu=PARAMS["r"], kappa=PARAMS["kappa"],
                v_bar=PARAMS["v_bar"], sigma_v=PARAMS["sigma_v"], rho=PARAMS["rho"],
                x0=PARAMS["x0"], v0=PARAMS["v0"], numerical_package_bundle=ref.NumpyBundle)
            S = np.exp(x.astype(np.float64))
            pay = np.maximum(S - PARAMS["K"], 0.0)
            units.append(0.5 * (pay[:R // 2] + pay[R // 2:]))
            pay_sum += pay.sum()
            n_paths += R
            sweep += np.maximum(S[:, None] - strikes[None, :], 0.0).sum(axis=0)
        units = np.concatenate(units)
        out[f"nT{nT}"] = dict(price=disc * pay_sum / n_paths,
                              se=disc * units.std(ddof=1) / math.sqrt(units.size),
                              paths=n_paths,
                              sweep=(disc * sweep / n_paths).tolist())
    out["strikes"] = strikes.tolist()
    return out


def antithetic_layouts(ref):
    rng = np.random.default_rng(3)
    arrays = {}
    for tag, shape, axis in (("a", (5, 2), 0), ("b", (2, 9), 1), ("c", (3, 4, 5), 2), ("d", (7,), 0),
                             ("e", (2, 3, 2, 4), 1)):
        half = list(shape)
        half[axis] = math.ceil(shape[axis] / 2)
        z = rng.standard_normal(half)
        u = rng.random(half)
        with InjectedNormals(normals=[z], uniforms=[u]):
            arrays[tag + "_randn"] = ref.randn_antithetic(shape, axis, np.float32, np)
            arrays[tag + "_rand"] = ref.rand_antithetic(shape, axis, np.float32, np)
        arrays[tag + "_z"], arrays[tag + "_u"] = z, u
        arrays[tag + "_axis"] = np.array(axis)
    np.savez_compressed(os.path.join(OUT, "antithetic_layouts.npz"), **arrays)


def helper_kats(ref):
    def sl(t):
        return [[s.start, s.stop, s.step] for s in t]
    kats = {
        # the reference's own expectations: cocos/tests/test_numerics/test_random/test_antithetic_slices.py:4-15
        "antithetic_slices": [
            {"shape": [2, 10], "axis": 0, "expect": sl(ref.get_antithetic_slices((2, 10), 0))},
            {"shape": [2, 10], "axis": 1, "expect": sl(ref.get_antithetic_slices((2, 10), 1))},
            {"shape": [2, 9], "axis": 1, "expect": sl(ref.get_antithetic_slices((2, 9), 1))},
            {"shape": [7], "axis": 0, "expect": sl(ref.get_antithetic_slices((7,), 0))},
        ],
        # cocos/tests/test_multiprocessing/test_generate_slices.py:10-20
        "slices_with_batch_size": [
            {"n": 10, "batch_size": b, "expect": [list(t) for t in ref.generate_slices_with_batch_size(10, b)]}
            for b in (2, 3, 4)],
        "slices_with_number_of_batches": [
            {"n": 10, "number_of_batches": b,
             "expect": [list(t) for t in ref.generate_slices_with_number_of_batches(10, b)]}
            for b in (2, 3, 4)],
    }
    a, k, n = ref.extract_arguments(args_list=None, kwargs_list=[{"q": 1}, {"q": 2}], number_of_batches=None)
    kats["extract_kwargs_only"] = {"args": [list(x) for x in a], "kwargs": k, "n": n}
    a, k, n = ref.extract_arguments(args_list=[[1], [2], [3]], kwargs_list=None, number_of_batches=None)
    kats["extract_args_only"] = {"args": [list(x) for x in a], "kwargs": k, "n": n}
    a, k, n = ref.extract_arguments(number_of_batches=4)
    kats["extract_count_only"] = {"args": [list(x) for x in a], "kwargs": k, "n": n}
    return kats


def _blas_info():
    try:
        import threadpoolctl
        return [{k: d.get(k) for k in ("internal_api", "version", "architecture", "num_threads")}
                for d in threadpoolctl.threadpool_info() if d.get("user_api") == "blas"]
    except Exception:           # noqa: BLE001
        return []


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    heston_injected(ref)
    antithetic_layouts(ref)
    doc = {
        "generated_by": "oracle/make_golden.py against /root/reference (cocos 0.2.2, NumPy branch)",
        "numpy": np.__version__,
        "blas": _blas_info(),
        "params": PARAMS,
        "seeded_prices": seeded_prices(ref),
        "helper_kats": helper_kats(ref),
        "statistical": statistical_prices(ref),
        # Random123 known-answer vectors for Philox4x32-10 (ctr, key, out), hex words
        "philox_kat": [
            {"ctr": ["00000000"] * 4, "key": ["00000000"] * 2,
             "out": ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]},
            {"ctr": ["ffffffff"] * 4, "key": ["ffffffff"] * 2,
             "out": ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]},
            {"ctr": ["243f6a88", "85a308d3", "13198a2e", "03707344"], "key": ["a4093822", "299f31d0"],
             "out": ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]},
        ],
    }
    with open(os.path.join(OUT, "reference_values.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
