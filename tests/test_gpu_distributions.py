"""Derived distributions: the reference's KS criterion (alpha = 0.01, its parameters and
sample sizes, cocos/tests/test_numerics/test_statistics/test_rng/*.py) plus element-wise closeness to the oracle's
model of the same Philox stream (bit-exact where the arithmetic is exact)."""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu
ALPHA = 0.01


@pytest.fixture(scope="module")
def rnd(gpu_device):
    import cocos_b200.numerics as cn
    return cn.random


def ks_ok(x, dist, args):
    frozen = getattr(stats, dist)(*args)        # frozen distribution: scipy's 'norm' fast path rejects loc/scale args
    return stats.kstest(np.asarray(x, dtype=np.float64), frozen.cdf).pvalue >= ALPHA


@pytest.mark.parametrize("a,b", [(0, 1), (2, 5), (3, 12), (5, 10), (1, 2), (10, 50), (20, 25)])
def test_uniform(rnd, a, b):                                   # test_uniform_rng.py
    from oracle import distributions_numpy as dn
    rnd.seed(1)
    u = rnd.uniform(a, b, 1_000_000).to_numpy()
    assert u.dtype == np.float32 and u.min() >= a and u.max() <= b
    assert ks_ok(u, "uniform", (a, b - a))
    assert np.array_equal(u, dn.simple("uniform", u.size, a, b, 1, 0))       # exact float32 arithmetic
    with pytest.raises(ValueError):
        rnd.uniform(2.0, 1.0, 10)


@pytest.mark.parametrize("mu,sigma", [(0.0, 0.2), (0.0, 1.0), (0.0, 5.0), (-2.0, 0.5)])
def test_normal_and_lognormal(rnd, mu, sigma):                 # test_normal_rng.py, test_lognormal_rng.py
    from oracle import distributions_numpy as dn
    rnd.seed(2)
    x = rnd.normal(mu, sigma, 5_000_000).to_numpy()
    assert ks_ok(x, "norm", (mu, sigma))
    model = dn.simple("normal", 200_000, mu, sigma, 2, 0)
    assert np.mean(np.abs(x[:200_000] - model) > 2e-5 * max(sigma, 1.0) + 1e-5 * np.abs(model)) <= 2e-5
    y = rnd.lognormal(mu, min(sigma, 1.5), 5_000_000).to_numpy()
    assert ks_ok(y, "lognorm", (min(sigma, 1.5), 0.0, np.exp(mu)))
    z = rnd.standard_normal(1_000_000).to_numpy()
    assert ks_ok(z, "norm", (0.0, 1.0))


@pytest.mark.parametrize("scale", [0.5, 1.0, 1.5])
def test_exponential(rnd, scale):                              # test_exponential_rng.py
    from oracle import distributions_numpy as dn
    rnd.seed(3)
    x = rnd.exponential(scale, 2_000_000).to_numpy()
    assert x.min() >= 0.0 and ks_ok(x, "expon", (0.0, scale))
    np.testing.assert_allclose(x[:100_000], dn.simple("exponential", 100_000, scale, 0.0, 3, 0), rtol=2e-6, atol=3e-7)
    assert ks_ok(rnd.standard_exponential(1_000_000).to_numpy(), "expon", (0.0, 1.0))


@pytest.mark.parametrize("loc,scale", [(5.0, np.sqrt(3.0) / np.pi), (5.0, 2.0), (9.0, 3.0), (9.0, 4.0), (6.0, 2.0),
                                        (2.0, 1.0)])
def test_logistic(rnd, loc, scale):                            # test_logistic_rng.py
    from oracle import distributions_numpy as dn
    rnd.seed(4)
    x = rnd.logistic(loc, scale, 1_000_000).to_numpy()
    assert np.all(np.isfinite(x)) and ks_ok(x, "logistic", (loc, scale))
    np.testing.assert_allclose(x[:100_000], dn.simple("logistic", 100_000, loc, scale, 4, 0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("a,b", [(1, 2), (2, 2), (3, 2), (5, 1), (9, 0.5), (7.5, 1), (0.5, 1)])
def test_gamma(rnd, a, b):                                     # test_gamma_rng.py
    from oracle import distributions_numpy as dn
    rnd.seed(5)
    x = rnd.gamma(a, b, 1_500_000).to_numpy()
    assert x.min() > 0.0 and ks_ok(x, "gamma", (a, 0.0, b))
    model = dn.gamma(50_000, a, b, 5, 0)
    assert np.mean(np.abs(x[:50_000] - model) > 1e-4 * (1 + np.abs(model))) <= 1e-3     # rare accept/reject flips
    assert ks_ok(rnd.standard_gamma(a, 500_000).to_numpy(), "gamma", (a, 0.0, 1.0))


@pytest.mark.parametrize("df", [1, 2, 3, 4, 5, 6, 9])
def test_chisquare(rnd, df):                                   # test_chisquare_rng.py (with a valid scipy scale)
    rnd.seed(6)
    assert ks_ok(rnd.chisquare(df, 2_000_000).to_numpy(), "chi2", (df,))


@pytest.mark.parametrize("a,b", [(0.5, 0.5), (5.0, 1.0), (1.0, 3.0), (2.0, 2.0), (2.0, 5.0)])
def test_beta(rnd, a, b):                                      # test_beta_rng.py (n = 10000 there)
    from oracle import distributions_numpy as dn
    rnd.seed(7)
    x = rnd.beta(a, b, 1_000_000).to_numpy()
    assert 0.0 <= x.min() and x.max() <= 1.0 and ks_ok(x, "beta", (a, b))
    model = dn.beta(20_000, a, b, 7, 0)
    assert np.mean(np.abs(x[:20_000] - model) > 1e-4) <= 2e-3


@pytest.mark.parametrize("mean,scale", [(1.0, 1.0), (1.0, 0.2), (1.0, 3.0), (3.0, 1.0), (3.0, 0.2)])
def test_wald(rnd, mean, scale):                               # test_wald_rng.py (with the valid invgauss mapping)
    from oracle import distributions_numpy as dn
    rnd.seed(8)
    x = rnd.wald(mean, scale, 1_000_000).to_numpy()
    assert x.min() > 0.0 and ks_ok(x, "invgauss", (mean / scale, 0.0, scale))
    model = dn.wald(50_000, mean, scale, 8, 0)
    assert np.mean(np.abs(x[:50_000] - model) > 1e-3 * (1 + np.abs(model))) <= 2e-3     # cancellation near y -> 0


def test_randint_and_choice(rnd, gpu_device):
    from oracle import distributions_numpy as dn
    from gpu_util import to_device
    rnd.seed(9)
    i = rnd.randint(3, 17, size=(1000, 50))
    assert i.shape == (1000, 50) and i.dtype == np.int32
    iv = i.to_numpy()
    assert iv.min() == 3 and iv.max() == 16
    assert np.array_equal(iv.reshape(-1), dn.randint(50_000, 3, 17, 9, 0))
    counts = np.bincount(iv.reshape(-1) - 3, minlength=14)
    assert stats.chisquare(counts).pvalue > 1e-3
    j = rnd.randint(10, size=7, dtype=np.int64)
    assert j.dtype == np.int64 and j.shape == (7,) and 0 <= j.to_numpy().min() and j.to_numpy().max() < 10
    assert rnd.randint(5).shape == (1,)
    src = np.linspace(0.0, 1.0, 101, dtype=np.float32)
    c = rnd.choice(to_device(gpu_device.current_context(), src), size=(200, 10)).to_numpy()
    assert c.shape == (200, 10) and np.all(np.isin(c, src))
    c64 = rnd.choice(to_device(gpu_device.current_context(), src.astype(np.float64)), size=33).to_numpy()
    assert c64.shape == (33,) and np.all(np.isin(c64, src.astype(np.float64)))
    with pytest.raises(ValueError):
        rnd.choice(to_device(gpu_device.current_context(), src), 5, replace=False)
    with pytest.raises(ValueError):
        rnd.choice(to_device(gpu_device.current_context(), src), 5, p=[0.5, 0.5])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_multivariate_normal(rnd, dtype):
    from oracle import philox_numpy as px
    mean = np.array([1.0, -2.0, 0.5], dtype=dtype)
    A = np.array([[2.0, 0.3, 0.1], [0.3, 1.0, -0.4], [0.1, -0.4, 0.5]], dtype=dtype)
    rnd.seed(10)
    x = rnd.multivariate_normal(mean, A, size=(1000, 200)).to_numpy()
    assert x.shape == (1000, 200, 3) and x.dtype == dtype
    flat = x.reshape(-1, 3).astype(np.float64)
    np.testing.assert_allclose(flat.mean(axis=0), mean, atol=0.02)
    np.testing.assert_allclose(np.cov(flat.T), A, atol=0.03)
    z = (px.randn_f32_model if dtype == np.float32 else px.randn_f64_model)(30_000, 10, 0).reshape(-1, 3)
    want = z @ np.linalg.cholesky(A.astype(np.float64)).T + mean
    tol = 1e-4 if dtype == np.float32 else 1e-10
    assert np.mean(np.abs(flat[:10_000] - want) > tol) <= 1e-4
    with pytest.raises(ValueError):
        rnd.multivariate_normal(mean, A[:2, :2], size=4)
    with pytest.raises(ValueError):
        rnd.multivariate_normal(mean.astype(np.float32), A.astype(np.float64), size=4)


def test_size_conventions(rnd):
    rnd.seed(11)
    assert isinstance(rnd.normal(), float)                       # size=None -> Python scalar (asscalar)
    assert rnd.normal(size=5).shape == (5,)
    assert rnd.gamma(2.0, 1.0, size=(3, 4)).shape == (3, 4)
    with pytest.raises(TypeError):
        rnd.uniform(size="many")
