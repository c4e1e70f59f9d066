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


# ---- integrals and resampling (cocos/scientific/kde.py:816-1025): SciPy is the truth, as in the reference's own test

def _datasets():
    rng = np.random.default_rng(5)
    one = np.concatenate([rng.normal(0.0, 1.0, 3000), rng.normal(3.0, 0.5, 2000)])
    two = rng.multivariate_normal([0.0, 1.0], [[1.0, 0.6], [0.6, 2.0]], size=4000).T
    w = rng.uniform(0.2, 1.0, one.size)
    return one, two, w


def test_integrate_gaussian_box_and_kde_match_scipy(kde_mod):
    one, two, w = _datasets()
    for data, weights in ((one, None), (one, w), (two, None)):
        k = kde_mod.gaussian_kde(data, weights=weights)
        ref = ss.gaussian_kde(data, weights=weights)
        d = k.d
        mean = np.full(d, 0.3)
        cov = np.eye(d) * 0.7 + 0.1
        assert k.integrate_gaussian(mean, cov) == pytest.approx(ref.integrate_gaussian(mean, cov), rel=2e-4)
        other_data = data[..., ::3] + 0.25
        ko, ro = kde_mod.gaussian_kde(other_data), ss.gaussian_kde(other_data)
        assert k.integrate_kde(ko) == pytest.approx(ref.integrate_kde(ro), rel=2e-4)
        assert ko.integrate_kde(k) == pytest.approx(ref.integrate_kde(ro), rel=2e-4)        # symmetric
        if d == 1:
            for low, high in ((-1.0, 2.0), (-np.inf, 0.5), (2.5, np.inf), (-40.0, 40.0), (6.0, 7.0)):
                assert k.integrate_box_1d(low, high) == pytest.approx(ref.integrate_box_1d(low, high), rel=1e-5, abs=1e-9)
        else:
            with pytest.raises(ValueError):
                k.integrate_box_1d(0.0, 1.0)
    k1 = kde_mod.gaussian_kde(one)
    with pytest.raises(ValueError):
        k1.integrate_gaussian([0.0, 1.0], np.eye(2))
    with pytest.raises(ValueError):
        k1.integrate_kde(kde_mod.gaussian_kde(two))


def test_resample_follows_the_estimated_density(kde_mod):
    """Samples from the estimate: the right shape, reproducible with a seed, mean/covariance of the mixture
    (data moments + kernel covariance), and in 1-D a KS test against the estimate's own CDF (integrate_box_1d)."""
    one, two, w = _datasets()
    for data, weights in ((one, None), (one, w), (two, None)):
        k = kde_mod.gaussian_kde(data, weights=weights)
        size = 400_000
        s1 = k.resample(size, seed=11).to_numpy()
        assert s1.shape == (k.d, size)
        assert np.array_equal(s1, k.resample(size, seed=11).to_numpy())
        assert not np.array_equal(s1, k.resample(size, seed=12).to_numpy())
        assert k.resample().shape == (k.d, int(k.neff))
        data2 = np.atleast_2d(data)
        wn = k.weights.astype(np.float64) / np.sum(k.weights)
        mean = data2 @ wn
        cov = (data2 - mean[:, None]) @ np.diag(wn) @ (data2 - mean[:, None]).T + k.covariance
        assert np.all(np.abs(s1.mean(axis=1) - mean) < 6 * np.sqrt(np.diag(np.atleast_2d(cov)) / size))
        np.testing.assert_allclose(np.cov(s1), np.squeeze(cov), rtol=0.02, atol=0.01)
        if k.d == 1:
            ref = ss.gaussian_kde(data, weights=weights)
            sample = s1[0, :20_000].astype(np.float64)
            grid = np.linspace(sample.min() - 1, sample.max() + 1, 2001)
            cdf = np.array([ref.integrate_box_1d(-np.inf, g) for g in grid])
            stat, pvalue = ss.kstest(sample, lambda x: np.interp(x, grid, cdf))
            assert pvalue > 0.001, (stat, pvalue)
