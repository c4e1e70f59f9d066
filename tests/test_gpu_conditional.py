"""Opt-in conditional-Gaussian scheme: same law of terminal prices and of antithetic pairs as the Euler kernel
(hence as the reference), half the normals.  Checked (a) per path against the oracle's model of the same Philox
stream, (b) against the reference's prices (3 SE), (c) in law against the Euler kernel."""
import ctypes
import math

import numpy as np
import pytest
from scipy import stats

from gpu_util import fused, hparams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gpu_device):
    return gpu_device.current_context()


@pytest.fixture(scope="module")
def P(ref_params):
    q = ref_params
    return dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
                x0=q["x0"], v0=q["v0"])


def conditional(ctx, p, R, N, strikes, seed, sub, first=0, count=None, outputs=False):
    from cocos_b200 import _lib
    from cocos_b200.numerics._array import empty_on
    H = (R + 1) // 2
    count = H - first if count is None else count
    k_arr, nK = _lib.strikes_array(strikes)
    hp = hparams(p)
    dx = dv = None
    if outputs:
        dx, dv = empty_on(ctx, (2 * max(count, 1),), np.float32), empty_on(ctx, (2 * max(count, 1),), np.float32)
    _lib.check(_lib.lib().cb_heston_price_conditional_launch_f32(
        ctx.handle, ctypes.byref(hp), R, N, k_arr, nK, seed, sub, first, count, None,
        dx.data_ptr if dx is not None else None, dv.data_ptr if dv is not None else None))
    sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    _lib.check(_lib.lib().cb_heston_price_finish(ctx.handle, nK, sums, sumsq))
    out = [np.array(sums), np.array(sumsq)]
    if dx is not None:
        out += [dx.to_numpy(), dv.to_numpy()]
    return out


@pytest.mark.parametrize("R,N", [(20_001, 64), (4096, 101), (1, 2), (7, 8), (1000, 14)])
def test_paths_match_stream_model(ctx, P, R, N):
    from oracle import heston_numpy as hn
    s, q, x, v = conditional(ctx, P, R, N, [0.95], seed=5, sub=2, outputs=True)
    xm, vm = hn.conditional_model(P["x0"], P["v0"], P["mu"], P["rho"], P["sigma_v"], P["kappa"], P["v_bar"], P["T"], N, R,
                                  5, 2)
    np.testing.assert_allclose(x[:R], xm, rtol=0, atol=1e-4)
    np.testing.assert_allclose(v[:R], vm, rtol=5e-3, atol=5e-6)
    ws, wq = hn.payoff_sums(x[:R], [0.95])
    assert s[0] == pytest.approx(ws[0], rel=1e-5, abs=1e-6) and q[0] == pytest.approx(wq[0], rel=1e-5, abs=1e-6)


@pytest.mark.parametrize("nT,R", [(101, 4_000_000), (501, 2_000_000), (1001, 1_000_000)])
def test_price_within_3se_of_reference(gpu_device, golden_values, ref_params, nT, R):
    from cocos_b200 import heston
    from cocos_b200.numerics import random as rnd
    rnd.seed(77)
    price, se = heston.price_with_standard_error(nT=nT, R=R, scheme="conditional", **ref_params)
    ref = golden_values["statistical"][f"nT{nT}"]
    assert abs(price - ref["price"]) <= 3 * math.hypot(se, ref["se"]), (price, se, ref)
    rnd.seed(77)
    euler_price, euler_se = heston.price_with_standard_error(nT=nT, R=R, **ref_params)
    assert se == pytest.approx(euler_se, rel=0.03)          # same pair law -> same standard error
    assert abs(price - euler_price) <= 4 * math.hypot(se, euler_se)


def test_same_law_as_euler_kernel(ctx, P):
    """Two-sample KS on terminal log prices, partner correlation, strike sweep: Euler vs conditional."""
    R, N = 2_000_000, 101
    strikes = [0.8, 0.95, 1.1]
    se_, qe, xe, ve = fused(ctx, P, R, N, strikes, seed=1, sub=0, outputs=True)
    sc, qc, xc, vc = conditional(ctx, P, R, N, strikes, seed=2, sub=0, outputs=True)
    H = R // 2
    assert stats.ks_2samp(xe[:H:4], xc[:H:4]).pvalue > 0.01
    assert stats.ks_2samp(ve[:H:4], vc[:H:4]).pvalue > 0.01
    floor = np.float32(1.1920929e-07)                       # Feller is violated: a visible atom sits on the floor
    assert abs(np.mean(ve[:R] == floor) - np.mean(vc[:R] == floor)) < 2e-3 and np.mean(ve[:R] == floor) > 0.01
    for xs in (xe, xc):
        assert abs(xs[:R].mean() - xe[:R].mean()) < 5e-4
    corr_e = np.corrcoef(xe[:H], xe[H:R])[0, 1]
    corr_c = np.corrcoef(xc[:H], xc[H:R])[0, 1]
    assert corr_e < -0.5 and abs(corr_e - corr_c) < 0.01     # antithetic partners: same (negative) dependence
    np.testing.assert_allclose(sc / R, se_ / R, atol=6e-4)
    np.testing.assert_allclose(qc / H, qe / H, rtol=0.02)


def test_shards_and_pool(ctx, P, gpu_device, ref_params):
    from cocos_b200 import heston
    from cocos_b200.multi_processing import ComputeDevicePool
    from cocos_b200.numerics import random as rnd
    R, N = 100_001, 31
    H = (R + 1) // 2
    whole = conditional(ctx, P, R, N, [0.95], 9, 1, outputs=True)
    a = conditional(ctx, P, R, N, [0.95], 9, 1, first=0, count=20_000, outputs=True)
    b = conditional(ctx, P, R, N, [0.95], 9, 1, first=20_000, count=H - 20_000)
    assert a[0][0] + b[0][0] == pytest.approx(whole[0][0], rel=1e-12)
    assert np.array_equal(a[2][:20_000], whole[2][:20_000])
    pool = ComputeDevicePool()
    rnd.seed(4)
    p_pool = heston.price_with_standard_error(nT=N, R=R, scheme="conditional", compute_device_pool=pool, **ref_params)
    rnd.seed(4)
    p_one = heston.price_with_standard_error(nT=N, R=R, scheme="conditional", **ref_params)
    assert p_pool[0] == pytest.approx(p_one[0], rel=1e-12)
    pool.close()
    with pytest.raises(ValueError):
        heston.simulate_and_compute_option_price(nT=N, R=R, scheme="milstein", **ref_params)
    with pytest.raises(ValueError):
        conditional(ctx, dict(P, sigma_v=0.0), 100, 10, [0.95], 0, 0)
