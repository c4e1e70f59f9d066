"""Gaussian KDE against SciPy, the truth the reference's own test uses
(cocos/tests/test_scientific/test_gaussian_kde.py:33-59: n = 1000 points, 101 / 101^2 grid, np.allclose)."""
import numpy as np
import pytest
import scipy.stats as ss

pytestmark = pytest.mark.gpu
N, GRID = 1000, 101
TOL = dict(rtol=2e-5, atol=1e-8)          # float32 data + MUFU.EX2 on the device, float64 in SciPy


@pytest.fixture(scope="module")
def kde_mod(gpu_device):
    from cocos_b200.scientific import kde
    return kde


def grid_2d():
    g = np.linspace(-5.0, 5.0, GRID)
    gx, gy = np.meshgrid(g, g)
    return np.hstack((gx.reshape((-1, 1)), gy.reshape((-1, 1)))).T


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("bw", [None, "silverman", 0.3, lambda info: 0.25])
def test_against_scipy(kde_mod, d, bw):
    rng = np.random.default_rng(d)
    pts = rng.standard_normal((d, N)) * np.arange(1, d + 1)[:, None] + rng.standard_normal((d, 1))
    if d > 1:
        pts[1] += 0.5 * pts[0]                               # correlated coordinates: a full covariance
    xi = np.linspace(-5, 5, GRID)[None, :] if d == 1 else (grid_2d() if d == 2 else rng.uniform(-4, 4, (3, 5000)))
    mine = kde_mod.gaussian_kde(pts.squeeze(), bw_method=bw)
    ref = ss.gaussian_kde(pts, bw_method=(bw if not callable(bw) else 0.25))
    assert (mine.d, mine.n) == (d, N)
    np.testing.assert_allclose(mine.factor, ref.factor, rtol=1e-12)
    np.testing.assert_allclose(mine.covariance, ref.covariance, rtol=1e-5)
    np.testing.assert_allclose(mine.evaluate(xi).to_numpy(), ref.evaluate(xi), **TOL)
    np.testing.assert_allclose(mine(xi[:, :7]).to_numpy(), ref(xi[:, :7]), **TOL)
    np.testing.assert_allclose(mine.logpdf(xi[:, 40:60]), ref.logpdf(xi[:, 40:60]), rtol=1e-4, atol=1e-4)


def test_weights_and_errors(kde_mod):
    rng = np.random.default_rng(5)
    pts = rng.standard_normal(N)
    w = rng.uniform(0.1, 2.0, N)
    xi = np.linspace(-5, 5, GRID)
    mine = kde_mod.gaussian_kde(pts, weights=w)
    ref = ss.gaussian_kde(pts, weights=w)
    np.testing.assert_allclose(mine.neff, ref.neff, rtol=1e-6)
    np.testing.assert_allclose(mine.evaluate(xi).to_numpy(), ref.evaluate(xi), **TOL)
    with pytest.raises(ValueError):
        kde_mod.gaussian_kde(pts, weights=w[:-1])
    with pytest.raises(ValueError):
        kde_mod.gaussian_kde(pts, bw_method="nope")
    with pytest.raises(ValueError):
        kde_mod.gaussian_kde(np.array([1.0]))
    with pytest.raises(ValueError):
        mine.evaluate(np.zeros((2, 5)))
    with pytest.raises(NotImplementedError):
        kde_mod.gaussian_kde(pts, gpu=False)
    assert mine.evaluate(np.zeros((1, 0))).shape == (0,)


def test_batched_and_multi_device_equal_direct(kde_mod):
    from cocos_b200.multi_processing import ComputeDevicePool
    rng = np.random.default_rng(6)
    pts = rng.standard_normal((2, 5000))
    xi = rng.uniform(-3, 3, (2, 3001))
    k = kde_mod.gaussian_kde(pts)
    direct = k.evaluate(xi).to_numpy()
    batched = k.evaluate_in_batches(xi, maximum_number_of_elements_per_batch=2_000_000)    # 200 points per batch
    assert batched.shape == (3001,) and np.array_equal(batched, direct)
    np.testing.assert_array_equal(kde_mod.evaluate_gaussian_kde_in_batches(k, xi, 5_000_000), direct)
    pool = ComputeDevicePool()
    multi = k.evaluate_in_batches_on_multiple_devices(xi, 2_000_000, pool)
    np.testing.assert_allclose(multi, direct, rtol=1e-6, atol=1e-12)
    pool.close()


def test_heston_terminal_prices_workflow(kde_mod, ref_params):
    """The notebook's workflow: simulate on the device, estimate the density of exp(x_T) without leaving it."""
    from cocos_b200 import heston
    from cocos_b200.numerics import random as rnd
    from cocos_b200.numerics._array import ndarray
    q = ref_params
    rnd.seed(3)
    x, _ = heston.simulate_heston_model(T=q["T"], N=101, R=200_000, mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"],
                                        sigma_v=q["sigma_v"], rho=q["rho"], x0=q["x0"], v0=q["v0"])
    gpu_device_ctx = x._ctx
    gpu_device_ctx.sync()
    import cocos_b200.numerics as cn
    prices = cn.exp(x)                                       # cn.exp(x_T): the one elementwise step of the workflow
    assert isinstance(prices, ndarray)
    k = kde_mod.gaussian_kde(prices)
    grid = np.linspace(0.4, 1.8, 400)
    dens = k.evaluate(grid).to_numpy()
    ref = ss.gaussian_kde(prices.to_numpy().astype(np.float64)).evaluate(grid)
    np.testing.assert_allclose(dens, ref, rtol=2e-4, atol=1e-7)
    assert abs(np.trapezoid(dens, grid) - 1.0) < 1e-3
