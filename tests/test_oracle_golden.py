"""The oracle against the reference's own outputs (tests/golden/, made by oracle/make_golden.py).

These pin oracle/heston_numpy.py, oracle/heston_oracle.c and oracle/philox_numpy.py; the GPU
parity tests then compare the CUDA path with the pinned oracle.  CPU only.
"""
import math

import numpy as np
import pytest

from oracle import c_oracle, heston_numpy as hn, philox_numpy as px
from conftest import HESTON_CASES, heston_case, rel_err_in_price

# float32 tolerance of the north_star's correctness check (1): 1e-6 relative on terminal prices
TOL_F32 = 1e-6


def _same_blas(golden_values):
    """True when np.dot runs the OpenBLAS build/path the fixtures were generated with."""
    try:
        import threadpoolctl
        here = [{k: d.get(k) for k in ("internal_api", "version", "architecture", "num_threads")}
                for d in threadpoolctl.threadpool_info() if d.get("user_api") == "blas"]
    except Exception:           # noqa: BLE001
        return False
    return here == golden_values.get("blas")


@pytest.mark.parametrize("name", HESTON_CASES)
def test_numpy_restatement_against_reference(golden_heston, golden_values, ref_params, name):
    """Same NumPy expressions as the reference: identical bits under the same BLAS, 1e-6 under any."""
    R, N, Z, x_ref, v_ref, p = heston_case(golden_heston, name, ref_params)
    x, v = hn.heston_terminal(R=R, N=N, dtype=np.float32, half_normals=lambda s: Z[s].astype(np.float64), **p)
    assert x.dtype == np.float32 and v.dtype == np.float32
    assert rel_err_in_price(x, x_ref) <= TOL_F32
    np.testing.assert_allclose(v, v_ref, rtol=2e-5, atol=1e-9)
    if _same_blas(golden_values):
        assert np.array_equal(x, x_ref) and np.array_equal(v, v_ref)


@pytest.mark.parametrize("name", HESTON_CASES)
def test_c_restatement_against_reference(golden_heston, ref_params, name):
    """The C restatement fixes np.dot's rounding to fma(z0,m0,rn(z1*m1)), the path OpenBLAS takes above
    ~16k rows: bit-exact on the large case (and the tiny ones), within 1e-6 where sgemv used another path."""
    R, N, Z, x_ref, v_ref, p = heston_case(golden_heston, name, ref_params)
    x, v = c_oracle.heston_injected(Z, R, N, **p)
    assert rel_err_in_price(x, x_ref) <= TOL_F32
    np.testing.assert_allclose(v, v_ref, rtol=2e-5, atol=1e-9)
    if name in ("large", "two", "one"):
        assert np.array_equal(x, x_ref) and np.array_equal(v, v_ref)


def test_c_f64_matches_numpy_f64(golden_heston, ref_params):
    R, N, Z, _, _, p = heston_case(golden_heston, "odd", ref_params)
    Z64 = Z.astype(np.float64)
    xc, vc = c_oracle.heston_injected(Z64, R, N, **p)
    xn, vn = hn.heston_terminal(R=R, N=N, dtype=np.float64, half_normals=lambda s: Z64[s], **p)
    np.testing.assert_allclose(xc, xn, rtol=0, atol=1e-13)
    np.testing.assert_allclose(vc, vn, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name", ("even", "odd", "two", "one", "long", "large"))
def test_payoff_matches_reference_price(golden_heston, ref_params, name):
    x = golden_heston[name + "_x"]
    want = float(golden_heston[name + "_price"])
    got = hn.call_price(ref_params["r"], ref_params["T"], ref_params["K"], x)
    assert got == want
    s, _ = c_oracle.payoff_sums(x, [ref_params["K"]])
    assert math.exp(-ref_params["r"] * ref_params["T"]) * s[0] / len(x) == pytest.approx(want, rel=2e-6)


def test_seeded_prices_reproduce(golden_values, ref_params):
    """Same numpy seed, same draws: the restatement returns the reference's price exactly."""
    exact = _same_blas(golden_values)
    for seed in (0, 1, 2):
        np.random.seed(seed)
        got = hn.price(nT=101, R=100000, **ref_params)
        want = golden_values["seeded_prices"][f"C1_seed{seed}"]
        assert got == want if exact else got == pytest.approx(want, rel=1e-6)
    np.random.seed(7)
    got, want = hn.price(nT=17, R=4999, **ref_params), golden_values["seeded_prices"]["odd_R_seed7"]
    assert got == want if exact else got == pytest.approx(want, rel=1e-6)


def test_antithetic_layouts(golden_antithetic):
    g = golden_antithetic
    for tag in "abcde":
        axis = int(g[tag + "_axis"])
        want_n, want_u = g[tag + "_randn"], g[tag + "_rand"]
        z, u = g[tag + "_z"], g[tag + "_u"]
        got_n = hn.normal_antithetic(want_n.shape, axis, np.float32, draw=lambda *s: z)
        got_u = hn.uniform_antithetic(want_u.shape, axis, np.float32, draw=lambda *s: u)
        assert np.array_equal(got_n, want_n) and got_n.dtype == want_n.dtype
        assert np.array_equal(got_u, want_u) and got_u.dtype == want_u.dtype


def test_helper_kats(golden_values):
    k = golden_values["helper_kats"]
    for case in k["antithetic_slices"]:
        got = hn.antithetic_slices(case["shape"], case["axis"])
        assert [[s.start, s.stop, s.step] for s in got] == case["expect"]
    for case in k["slices_with_batch_size"]:
        assert [list(t) for t in hn.slices_with_batch_size(case["n"], case["batch_size"])] == case["expect"]
    for case in k["slices_with_number_of_batches"]:
        assert [list(t) for t in hn.slices_with_number_of_batches(case["n"], case["number_of_batches"])] \
            == case["expect"]
    a, kw, n = hn.normalise_batches(kwargs_list=[{"q": 1}, {"q": 2}])
    assert ([list(x) for x in a], kw, n) == (k["extract_kwargs_only"]["args"], k["extract_kwargs_only"]["kwargs"], 2)
    a, kw, n = hn.normalise_batches(number_of_batches=4)
    assert ([list(x) for x in a], kw, n) == (k["extract_count_only"]["args"], k["extract_count_only"]["kwargs"], 4)
    with pytest.raises(ValueError):
        hn.normalise_batches()


def test_reference_slice_expectations_verbatim():
    """Expected values written in the reference's own tests (test_antithetic_slices.py:4-15,
    test_generate_slices.py:10-20)."""
    assert hn.antithetic_slices((2, 10), 0) == (slice(0, 1, 1), slice(0, 10, 1))
    assert hn.antithetic_slices((2, 10), 1) == (slice(0, 2, 1), slice(0, 5, 1))
    assert hn.antithetic_slices((2, 9), 1) == (slice(0, 2, 1), slice(0, 4, 1))
    assert hn.slices_with_batch_size(10, 3) == ((0, 3), (3, 6), (6, 9), (9, 10))
    assert hn.slices_with_number_of_batches(10, 4) == ((0, 3), (3, 6), (6, 9), (9, 10))


def test_philox_known_answers(golden_values):
    for kat in golden_values["philox_kat"]:
        ctr = [int(w, 16) for w in kat["ctr"]]
        key = [int(w, 16) for w in kat["key"]]
        out = [int(w, 16) for w in kat["out"]]
        assert [int(w) for w in px.philox4x32_10(*ctr, *key)] == out
        assert [int(w) for w in c_oracle.philox(ctr, key)] == out


def test_philox_numpy_equals_c():
    seed, sub = 0x0123456789ABCDEF, 77
    idx = np.array([0, 1, 2, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 5], dtype=np.uint64)
    for block in (0, 3):
        o = np.stack(px.philox_block(idx, block, sub, seed), axis=1)
        for row, i in zip(o, idx):
            assert np.array_equal(row, c_oracle.philox_words(int(i), 1, block, sub, seed)[0])


def test_pi_models_agree():
    for n in (1, 2, 3, 1000, 99999):
        for anti in (False, True):
            assert px.pi_inside_count(n, 5, 9, anti) == c_oracle.pi_inside(n, 5, 9, anti)
    est = 4 * c_oracle.pi_inside(4_000_000, 1, 0, True) / 4_000_000
    assert abs(est - math.pi) < 5e-3


def test_box_muller_model_is_standard_normal():
    from scipy import stats
    z = px.randn_f32_model(400_000, seed=3, subsequence=1)
    assert stats.kstest(z, "norm").pvalue > 0.01           # the reference's KS criterion (test_normal_rng.py)
    u = px.rand_f32(400_000, seed=3, subsequence=2)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 0.01


def test_fused_model_prices_within_3se_of_reference(golden_values, ref_params):
    """The Philox/Box-Muller stream drives the reference update to the reference's price (check 2)."""
    stat = golden_values["statistical"]["nT101"]
    x, _ = hn.fused_model(nT=101, R=200_000, seed=42, subsequence=0,
                          **{k: ref_params[k] for k in ("x0", "v0", "r", "rho", "sigma_v", "kappa", "v_bar", "T")})
    price, se = hn.call_price_and_se(ref_params["r"], ref_params["T"], ref_params["K"], x)
    assert abs(price - stat["price"]) <= 3 * math.hypot(se, stat["se"])
