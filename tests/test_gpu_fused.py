"""Fused kernel (in-kernel Philox + Box-Muller + Euler + payoff reduction) against the oracle.

(a) per path, against the oracle's model of the SAME stream (Philox bits -> float64 Box-Muller -> the
    reference update in float32/float64): tight tolerance, far below Monte-Carlo noise;
(b) correctness check (2) of the north_star: prices within 3 standard errors of the reference's own
    prices (tests/golden/reference_values.json "statistical", made from the unmodified reference);
(c) size-independent properties at BASELINE sizes: shards add up, geometry does not change a path,
    launches are reproducible.
"""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gpu_util import fused

pytestmark = pytest.mark.gpu

STRIKES64 = [0.70 + 0.60 * i / 63 for i in range(64)]


@pytest.fixture(scope="module")
def ctx(gpu_device):
    c = gpu_device.current_context()
    yield c
    c.set_launch(0, 0, 0)


@pytest.fixture(scope="module")
def P(ref_params):
    q = ref_params
    return dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
                x0=q["x0"], v0=q["v0"])


def model(P, R, N, seed, sub, dtype, pair_offset=0):
    from oracle import heston_numpy as hn
    return hn.fused_model(P["x0"], P["v0"], P["mu"], P["rho"], P["sigma_v"], P["kappa"], P["v_bar"], P["T"], N, R,
                          seed, sub, pair_offset=pair_offset, dtype=dtype)


@pytest.mark.parametrize("R,N", [(20_001, 64), (4096, 101), (1, 2), (2, 3), (7, 6), (1025, 17)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_paths_match_stream_model(ctx, P, R, N, dtype):
    from oracle import heston_numpy as hn
    sums, sumsq, x, v = fused(ctx, P, R, N, [0.95], seed=99, sub=3, dtype=dtype, outputs=True)
    xm, vm = model(P, R, N, 99, 3, dtype)
    # float32 Box-Muller on MUFU vs float64: ~1e-6 per normal, accumulated over N-1 steps
    tol = 2e-5 if N <= 101 else 1e-4
    np.testing.assert_allclose(x[:R], xm, rtol=0, atol=tol)
    np.testing.assert_allclose(v[:R], vm, rtol=5e-3, atol=5e-6)
    ws, wq = hn.payoff_sums(xm, [0.95])
    assert sums[0] == pytest.approx(ws[0], rel=1e-4, abs=1e-4)
    assert sumsq[0] == pytest.approx(wq[0], rel=2e-4, abs=1e-4)
    # the kernel's own sums agree with its own terminal values
    os_, oq = hn.payoff_sums(x[:R], [0.95])
    assert sums[0] == pytest.approx(os_[0], rel=1e-5, abs=1e-6)
    assert sumsq[0] == pytest.approx(oq[0], rel=1e-5, abs=1e-6)


def test_antithetic_partners(ctx, P):
    """One Euler step: partners see -z, so x1 + x2 = 2*(x0 + (mu - v0/2) dt) up to rounding."""
    R = 10_000
    _, _, x, v = fused(ctx, P, R, 2, [0.95], seed=1, sub=0, outputs=True)
    H = R // 2
    np.testing.assert_allclose(x[:H] + x[H:R], 2 * (P["x0"] + (P["mu"] - 0.5 * P["v0"]) * P["T"]), atol=2e-6)
    assert np.std(x[:H] - x[H:R]) > 0.05


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_multi_strike_equals_single_strike(ctx, P, dtype):
    R, N = 50_001, 33
    s64, q64 = fused(ctx, P, R, N, STRIKES64, seed=5, sub=1, dtype=dtype)
    for k in (0, 17, 31, 32, 63):
        s1, q1 = fused(ctx, P, R, N, [STRIKES64[k]], seed=5, sub=1, dtype=dtype)
        assert s64[k] == pytest.approx(s1[0], rel=1e-9)
        assert q64[k] == pytest.approx(q1[0], rel=1e-9)
    s3, _ = fused(ctx, P, R, N, STRIKES64[:3], seed=5, sub=1, dtype=dtype)
    np.testing.assert_allclose(s3, s64[:3], rtol=1e-9)
    assert np.all(np.diff(s64) < 0)                       # call payoffs decrease in the strike


@pytest.mark.parametrize("threads,bps,ppt", [(128, 0, 1), (128, 0, 2), (256, 0, 4), (64, 3, 2), (256, 1, 1)])
def test_geometry_does_not_change_paths(ctx, P, threads, bps, ppt):
    R, N = 30_003, 21
    ctx.set_launch(0, 0, 0)
    s0, q0, x0, v0 = fused(ctx, P, R, N, [0.95], seed=8, sub=2, outputs=True)
    ctx.set_launch(threads, bps, ppt)
    s1, q1, x1, v1 = fused(ctx, P, R, N, [0.95], seed=8, sub=2, outputs=True)
    ctx.set_launch(0, 0, 0)
    assert np.array_equal(x0[:R], x1[:R]) and np.array_equal(v0[:R], v1[:R])     # per-path bits identical
    assert s1[0] == pytest.approx(s0[0], rel=1e-12) and q1[0] == pytest.approx(q0[0], rel=1e-12)


def test_reproducible_and_stream_separation(ctx, P):
    a = fused(ctx, P, 100_000, 51, [0.95], seed=3, sub=7)
    b = fused(ctx, P, 100_000, 51, [0.95], seed=3, sub=7)
    assert a[0][0] == b[0][0] and a[1][0] == b[1][0]       # fixed-order reduction: identical bits
    c = fused(ctx, P, 100_000, 51, [0.95], seed=3, sub=8)
    d = fused(ctx, P, 100_000, 51, [0.95], seed=4, sub=7)
    assert c[0][0] != a[0][0] and d[0][0] != a[0][0]


@pytest.mark.parametrize("R,parts", [(100_001, 2), (100_000, 8), (999, 4), (5, 8)])
def test_shards_add_up(ctx, P, R, parts):
    """Device-count invariance: disjoint pair ranges reproduce the whole simulation path by path."""
    from cocos_b200.multi_processing.utilities import shard_range
    N = 25
    H, F = (R + 1) // 2, R // 2
    s, q, x, v = fused(ctx, P, R, N, STRIKES64[:5], seed=21, sub=4, outputs=True)
    tot_s, tot_q = np.zeros(5), np.zeros(5)
    for g in range(parts):
        first, count = shard_range(H, parts, g)
        out = fused(ctx, P, R, N, STRIKES64[:5], seed=21, sub=4, first=first, count=count, outputs=count > 0)
        tot_s += out[0]
        tot_q += out[1]
        if count:
            xs = out[2]
            assert np.array_equal(xs[:count], x[first:first + count])
            n_partner = max(0, min(first + count, F) - first)
            assert np.array_equal(xs[count:count + n_partner], x[H + first:H + first + n_partner])
    np.testing.assert_allclose(tot_s, s, rtol=1e-12)
    np.testing.assert_allclose(tot_q, q, rtol=1e-12)


def test_c3_sums_reproduce_the_committed_golden(ctx, golden_values):
    """The fixed job of the device-count-invariance checks (bench.py extras.invariance): one GPU reproduces the sums
    in tests/golden/fused_c3_sums.json bit for bit (fixed-order reduction), and the same job cut into 2, 4 and 8
    contiguous pair shards -- what 2/4/8 GPUs simulate -- adds up to them to float64 summation order."""
    import json
    from cocos_b200.multi_processing.utilities import shard_range
    g = json.load(open(os.path.join(GOLDEN, "fused_c3_sums.json")))
    q = golden_values["params"]
    p = dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"], x0=q["x0"],
             v0=q["v0"])
    R, nT = g["R"], g["nT"]
    s, sq = fused(ctx, p, R, nT, [g["K"]], seed=g["seed"], sub=g["subsequence"])
    assert float(s[0]).hex() == g["sum_hex"] and float(sq[0]).hex() == g["sumsq_hex"]
    H = (R + 1) // 2
    for parts in (2, 4, 8):
        tot_s = tot_q = 0.0
        for r in range(parts):
            first, count = shard_range(H, parts, r)
            a, b = fused(ctx, p, R, nT, [g["K"]], seed=g["seed"], sub=g["subsequence"], first=first, count=count)
            tot_s += a[0]
            tot_q += b[0]
        assert abs(tot_s - g["sum"]) <= 1e-12 * g["sum"] and abs(tot_q - g["sumsq"]) <= 1e-12 * g["sumsq"]


def test_trajectory_mode(ctx, P):
    R, N = 8_000, 41
    s, q, x, v, traj = fused(ctx, P, R, N, [0.95], seed=6, sub=0, outputs=True, traj=True)
    H = R // 2
    assert traj.shape == (N, R)
    assert np.all(traj[0] == np.float32(P["x0"]))
    assert np.array_equal(traj[-1], x[:R])                 # last row is the terminal value
    s2, q2, x2, v2 = fused(ctx, P, R, N, [0.95], seed=6, sub=0, outputs=True)
    assert np.array_equal(x2[:R], x[:R])                   # same paths as the non-materialising kernel
    incr = np.diff(traj.astype(np.float64), axis=0)
    assert 0.5 * math.sqrt(P["v_bar"] / (N - 1)) < incr.std() < 3 * math.sqrt(P["v0"] / (N - 1))
    # ragged: odd R, pair count not a multiple of 4 -> scalar-store path
    R = 4_999
    out = fused(ctx, P, R, 9, [0.95], seed=6, sub=1, outputs=True, traj=True)
    xr, tr = out[2], out[4]
    assert np.array_equal(tr[-1][:R], xr[:R])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("R,N", [(8_192, 41), (8_000, 14), (5_000, 8), (4_999, 9), (1_000_000, 5)])
def test_trajectory_with_variance(ctx, P, dtype, R, N):
    """(x, v) path-materialising mode: TMA tensor stores (aligned shards) and direct stores (ragged shards) write the
    same rows; the last rows are the terminal state bit for bit; x rows equal the x-only launch."""
    H = (R + 1) // 2
    s, q, x, v, tx, tv = fused(ctx, P, R, N, [0.95], seed=8, sub=2, dtype=dtype, outputs=True, traj_v=True)
    assert tx.shape == (N, 2 * H) and tv.shape == (N, 2 * H) and tx.dtype == dtype
    eps = np.float32(1.1920929e-07)
    assert np.all(tx[0, :R] == dtype(P["x0"])) and np.all(tv[0, :R] == dtype(P["v0"]))
    assert np.array_equal(tx[-1][:R], x[:R])
    # the variance rows clamp with max(u/c1, eps): equal to the terminal v except for ulp effects at the floor
    np.testing.assert_allclose(tv[-1][:R], v[:R], rtol=3e-7 if dtype == np.float32 else 1e-15)
    assert np.all(tv[:, :R] >= eps)
    s1, q1, x1, v1, tx1 = fused(ctx, P, R, N, [0.95], seed=8, sub=2, dtype=dtype, outputs=True, traj=True)
    assert np.array_equal(tx1[:, :R], tx[:, :R]) and np.array_equal(x1[:R], x[:R])
    np.testing.assert_allclose(s1, s, rtol=1e-13)
    # same stream through the non-materialising kernel
    s2, q2, x2, v2 = fused(ctx, P, R, N, [0.95], seed=8, sub=2, dtype=dtype, outputs=True)
    assert np.array_equal(x2[:R], x[:R]) and np.array_equal(v2[:R], v[:R])
    # variance rows against a host replay of the recursion from the stored x increments is not possible (two normals
    # per step), so check the law instead: E[v_t] relaxes from v0 towards v_bar
    if R >= 100_000:
        m = tv[:, :R].astype(np.float64).mean(axis=1)
        assert abs(m[0] - P["v0"]) < 1e-9 and m[-1] > m[0]


@pytest.mark.parametrize("R,N", [(65_536, 33), (8, 20), (72, 7), (520, 2), (131_080, 12)])
def test_trajectory_bulk_equals_direct(ctx, P, monkeypatch, R, N):
    """COCOS_B200_NO_BULK=1 forces the direct-store variant on a shard the TMA variant would take: identical output
    (also for shards narrower than the 256-column box, boxes cut by the last column, and fewer steps than a stage).
    The payoff sums are added up in block order and the two variants use different block widths (512 / 256 threads):
    they agree to float64 summation order, every path value bit for bit."""
    a = fused(ctx, P, R, N, [0.9, 1.0], seed=9, sub=1, outputs=True, traj_v=True)
    monkeypatch.setenv("COCOS_B200_NO_BULK", "1")
    b = fused(ctx, P, R, N, [0.9, 1.0], seed=9, sub=1, outputs=True, traj_v=True)
    for u, w in zip(a[:2], b[:2]):
        np.testing.assert_allclose(u, w, rtol=1e-13)
    for u, w in zip(a[2:], b[2:]):
        assert np.array_equal(u, w)


def test_trajectory_mode_f64(ctx, P):
    R, N = 4_000, 25
    s, q, x, v, traj = fused(ctx, P, R, N, [0.95], seed=6, sub=3, dtype=np.float64, outputs=True, traj=True)
    assert traj.dtype == np.float64 and traj.shape == (N, R)
    assert np.array_equal(traj[-1], x[:R]) and np.all(traj[0] == P["x0"])
    x32 = fused(ctx, P, R, N, [0.95], seed=6, sub=3, dtype=np.float32, outputs=True)[2]
    np.testing.assert_allclose(x[:R], x32[:R], atol=2e-5)           # same stream, float64 vs float32 state


@pytest.mark.parametrize("nT,R", [(101, 4_000_000), (253, 2_000_000), (501, 2_000_000), (1001, 1_000_000)])
def test_price_within_3se_of_reference(gpu_device, golden_values, ref_params, nT, R):
    """Correctness check (2): in-kernel RNG price vs the reference's NumPy price, 3 combined SE."""
    from cocos_b200 import heston
    from cocos_b200.numerics import random as rnd
    rnd.seed(20261019)
    price, se = heston.price_with_standard_error(nT=nT, R=R, **ref_params)
    ref = golden_values["statistical"][f"nT{nT}"]
    assert 0.3 * ref["se"] * math.sqrt(ref["paths"] / R) < se < 3 * ref["se"] * math.sqrt(ref["paths"] / R)
    assert abs(price - ref["price"]) <= 3 * math.hypot(se, ref["se"]), (price, se, ref["price"], ref["se"])
    plain = heston.simulate_and_compute_option_price(nT=nT, R=R, **ref_params)
    assert abs(plain - ref["price"]) <= 3 * math.hypot(se, ref["se"])


def test_strike_sweep_f64_within_3se(gpu_device, golden_values, ref_params):
    """BASELINE config C5 shape: float64 state, 64 strikes sharing paths, nT = 253."""
    from cocos_b200 import heston
    from cocos_b200.numerics import random as rnd
    rnd.seed(5)
    kw = {k: ref_params[k] for k in ("x0", "v0", "r", "rho", "sigma_v", "kappa", "v_bar", "T")}
    prices, ses = heston.price_strike_sweep(strikes=golden_values["statistical"]["strikes"], nT=253, R=2_000_000,
                                            dtype=np.float64, **kw)
    ref = np.array(golden_values["statistical"]["nT253"]["sweep"])
    ref_se = golden_values["statistical"]["nT253"]["se"]
    assert prices.shape == (64,)
    assert np.all(np.abs(prices - ref) <= 3 * np.hypot(ses, 2.0 * ref_se) + 1e-7)
    rnd.seed(5)
    p32, _ = heston.price_strike_sweep(strikes=golden_values["statistical"]["strikes"], nT=253, R=2_000_000,
                                       dtype=np.float32, **kw)
    np.testing.assert_allclose(p32, prices, rtol=0, atol=5e-5)        # same stream, float32 vs float64 state


def test_full_size_c3_properties(ctx, P, golden_values, ref_params):
    """BASELINE config C3 (10M paths x 500 steps): halves add up; price within 3 SE of the reference."""
    R, N = 10_000_000, 501
    H = R // 2
    whole = fused(ctx, P, R, N, [0.95], seed=0, sub=0)
    a = fused(ctx, P, R, N, [0.95], seed=0, sub=0, first=0, count=H // 2)
    b = fused(ctx, P, R, N, [0.95], seed=0, sub=0, first=H // 2, count=H - H // 2)
    assert a[0][0] + b[0][0] == pytest.approx(whole[0][0], rel=1e-12)
    assert a[1][0] + b[1][0] == pytest.approx(whole[1][0], rel=1e-12)
    from cocos_b200.heston import _price_from_sums
    price, se = _price_from_sums(ref_params["r"], ref_params["T"], R, whole[0], whole[1])
    ref = golden_values["statistical"]["nT501"]
    assert abs(price[0] - ref["price"]) <= 3 * math.hypot(se[0], ref["se"])


def test_argument_errors(ctx, P):
    with pytest.raises(ValueError):
        fused(ctx, P, 100, 1, [0.95], 0, 0)                       # N < 2
    with pytest.raises(ValueError):
        fused(ctx, P, 0, 10, [0.95], 0, 0)                        # R < 1
    with pytest.raises(ValueError):
        fused(ctx, P, 100, 10, [0.95], 0, 0, first=40, count=20)  # pair range outside [0, 50)
    with pytest.raises(ValueError):
        fused(ctx, P, 100, 10, [1.0] * 65, 0, 0)
    from cocos_b200 import heston
    from oracle import heston_numpy as hn
    with pytest.raises(TypeError):
        heston.simulate_and_compute_option_price(nT=10, R=10, numerical_package_bundle=object, **hn.REFERENCE_PARAMS)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pair_indices_beyond_32_bits(ctx, P, dtype):
    """Shards far into a huge simulation: the 64-bit pair index feeds both Philox counter words."""
    first = 2 ** 32 - 6                       # straddles the 2^32 boundary of the low counter word
    count, N = 12, 20
    R = 2 ** 40
    s, q, x, v = fused(ctx, P, R, N, [0.95], seed=77, sub=9, first=first, count=count, dtype=dtype, outputs=True)
    xm, vm = model(P, 2 * count, N, 77, 9, dtype, pair_offset=first)
    np.testing.assert_allclose(x[:2 * count], xm, rtol=0, atol=2e-5)
    np.testing.assert_allclose(v[:2 * count], vm, rtol=5e-3, atol=5e-6)
    far = fused(ctx, P, R, N, [0.95], seed=77, sub=9, first=2 ** 38 + 3, count=count, dtype=dtype, outputs=True)
    xf, _ = model(P, 2 * count, N, 77, 9, dtype, pair_offset=2 ** 38 + 3)
    np.testing.assert_allclose(far[2][:2 * count], xf, rtol=0, atol=2e-5)


def test_tiny_and_degenerate_shapes(ctx, P):
    """R = 1 (a lone unpaired path), N = 2 (one step), empty shards, with every output mode."""
    s, q, x, v, traj = fused(ctx, P, 1, 2, [0.95], seed=1, sub=1, outputs=True, traj=True)
    assert traj.shape == (2, 2) and traj[0, 0] == np.float32(P["x0"]) and traj[1, 0] == x[0]
    assert s[0] == pytest.approx(max(math.exp(float(x[0])) - 0.95, 0.0), rel=1e-5) and q[0] == pytest.approx(s[0] ** 2, rel=1e-5)
    empty = fused(ctx, P, 10, 5, [0.9, 1.0], seed=1, sub=1, first=5, count=0)
    assert np.all(empty[0] == 0) and np.all(empty[1] == 0)
    odd = fused(ctx, P, 3, 4, [0.95], seed=1, sub=1, outputs=True)
    xm, _ = model(P, 3, 4, 1, 1, np.float32)
    np.testing.assert_allclose(odd[2][:3], xm, atol=1e-5)


def test_multicore_entry_point_is_one_fused_launch(gpu_device, golden_values, ref_params):
    """simulate_and_compute_option_price_multicore (heston_utilities.py:167-214): same signature, same estimator
    (number_of_cores * ceil(R / number_of_cores) paths), no CPU branch."""
    from cocos_b200 import heston
    from cocos_b200.numerics import random as rnd
    from cocos_b200.numerics.numerical_package_bundle import NumericalPackageBundle
    rnd.seed(77)
    price = heston.simulate_and_compute_option_price_multicore(nT=501, R=1_999_999, number_of_cores=7, **ref_params)
    rnd.seed(77)
    same = heston.simulate_and_compute_option_price(nT=501, R=math.ceil(1_999_999 / 7) * 7, **ref_params)
    assert price == same
    ref = golden_values["statistical"]["nT501"]
    assert abs(price - ref["price"]) <= 3 * math.hypot(ref["se"], ref["se"] / 1.4)

    class NumpyLike(NumericalPackageBundle):
        pass
    with pytest.raises(TypeError):
        heston.simulate_and_compute_option_price_multicore(nT=11, R=100, numerical_package_bundle=NumpyLike, **ref_params)
