"""Path-dependent payoffs inside the fused kernel against the oracle: the same Philox
stream drives the reference update in NumPy with the whole path recorded, then textbook payoff definitions."""
import ctypes
import math

import numpy as np
import pytest

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


def payoff_launch(ctx, p, R, N, strikes, seed, sub, kind, put=False, barrier=0.0, dtype=np.float32):
    from cocos_b200 import _lib
    sfx = "f32" if dtype == np.float32 else "f64"
    k_arr, nK = _lib.strikes_array(strikes)
    hp = hparams(p)
    spec = _lib.PayoffSpec(kind, int(put), float(barrier))
    fn = getattr(_lib.lib(), f"cb_heston_payoff_launch_{sfx}")
    _lib.check(fn(ctx.handle, ctypes.byref(hp), R, N, k_arr, nK, seed, sub, 0, (R + 1) // 2, None, ctypes.byref(spec)))
    sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    _lib.check(_lib.lib().cb_heston_price_finish(ctx.handle, nK, sums, sumsq))
    return np.array(sums), np.array(sumsq)


def oracle_path(P, R, N, seed, sub, dtype):
    from oracle import heston_numpy as hn
    rec = []
    hn.heston_terminal(T=P["T"], N=N, R=R, mu=P["mu"], kappa=P["kappa"], v_bar=P["v_bar"], sigma_v=P["sigma_v"],
                       rho=P["rho"], x0=P["x0"], v0=P["v0"], dtype=dtype,
                       half_normals=hn.philox_half_normals(R, seed, sub), record=rec)
    return np.array(rec)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind,name,barrier", [(1, "asian", None), (2, "up_and_out", 1.25), (3, "down_and_out", 0.8),
                                              (0, "european", None)])
@pytest.mark.parametrize("put", [False, True])
def test_sums_match_oracle(ctx, P, dtype, kind, name, barrier, put):
    from oracle import heston_numpy as hn
    R, N = 20_001, 41
    strikes = [0.9, 0.95, 1.05]
    xs = oracle_path(P, R, N, 13, 4, dtype)
    want_s, want_q = hn.path_dependent_sums(xs, strikes, name, put, barrier)
    s, q = payoff_launch(ctx, P, R, N, strikes, 13, 4, kind, put, barrier or 0.0, dtype)
    # a path within float rounding of the barrier may flip: allow a few payoffs of slack on the knock-outs
    slack = 3.0 if kind >= 2 else 0.0
    np.testing.assert_allclose(s, want_s, rtol=2e-4, atol=slack + 1e-3)
    np.testing.assert_allclose(q, want_q, rtol=5e-4, atol=slack + 1e-3)
    one_s, one_q = payoff_launch(ctx, P, R, N, strikes[1:2], 13, 4, kind, put, barrier or 0.0, dtype)
    # single- and multi-strike kernels are separate instantiations: equal up to float32 rounding of the path functionals
    assert one_s[0] == pytest.approx(s[1], rel=1e-6) and one_q[0] == pytest.approx(q[1], rel=1e-6)


def test_limits_and_orderings(ctx, P, gpu_device, ref_params):
    R, N = 400_000, 101
    euro = fused(ctx, P, R, N, [0.95], seed=3, sub=0)
    same = payoff_launch(ctx, P, R, N, [0.95], 3, 0, kind=0)
    assert same[0][0] == euro[0][0] and same[1][0] == euro[1][0]               # kind 0 = the pricing kernel
    wide = payoff_launch(ctx, P, R, N, [0.95], 3, 0, kind=2, barrier=1e6)
    assert wide[0][0] == pytest.approx(euro[0][0], rel=1e-12)                  # unreachable barrier = European
    dead = payoff_launch(ctx, P, R, N, [0.95], 3, 0, kind=2, barrier=0.5)
    assert dead[0][0] == 0.0                                                   # barrier below every path
    asian = payoff_launch(ctx, P, R, N, [0.95], 3, 0, kind=1)
    uo = payoff_launch(ctx, P, R, N, [0.95], 3, 0, kind=2, barrier=1.2)
    assert 0 < asian[0][0] < euro[0][0] and 0 < uo[0][0] < euro[0][0]
    # put-call parity on the same paths: C - P = E[S_T] - K (undiscounted sums), exact per path
    put = payoff_launch(ctx, P, R, N, [0.95], 3, 0, kind=0, put=True)
    x = fused(ctx, P, R, N, [0.95], seed=3, sub=0, outputs=True)[2][:R]
    assert euro[0][0] - put[0][0] == pytest.approx(np.exp(x.astype(np.float64)).sum() - 0.95 * R, rel=1e-5)
    from cocos_b200 import heston
    from cocos_b200.numerics import random as rnd
    rnd.seed(8)
    price, se = heston.price_path_dependent(nT=101, R=R, kind="asian", **ref_params)
    assert 0.0 < price < 0.1034 and 0 < se < 2e-4
    rnd.seed(8)
    prices, ses = heston.price_path_dependent(nT=101, R=R, kind="down_and_out", option="put", barrier=0.7,
                                              **dict(ref_params, K=[0.9, 1.0, 1.1]))
    assert prices.shape == (3,) and np.all(np.diff(prices) > 0)
    with pytest.raises(ValueError):
        heston.price_path_dependent(nT=11, R=10, kind="up_and_out", **ref_params)
    with pytest.raises(ValueError):
        heston.price_path_dependent(nT=11, R=10, kind="lookback", **ref_params)
