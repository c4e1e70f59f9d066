"""Correctness check (1): injected normals, per-path terminal values against the reference.

CUDA parity kernel (through the C ABI) vs (a) the golden outputs of the unmodified reference,
(b) the C oracle on the same seeded inputs, at sizes up to BASELINE config C1 (100k x 100).
Bar: float32 within 1e-6 relative on terminal prices (the kernel is in fact bit-identical to the
oracle); float64 within 1e-12.
"""
import numpy as np
import pytest

from conftest import HESTON_CASES, heston_case, rel_err_in_price
from gpu_util import injected, injected_host

pytestmark = pytest.mark.gpu
TOL_F32, TOL_F64 = 1e-6, 1e-12


@pytest.fixture(scope="module")
def ctx(gpu_device):
    return gpu_device.current_context()


@pytest.mark.parametrize("name", HESTON_CASES)
def test_against_reference_golden(ctx, golden_heston, ref_params, name):
    R, N, Z, x_ref, v_ref, p = heston_case(golden_heston, name, ref_params)
    x, v = injected(ctx, Z, R, N, p)
    assert rel_err_in_price(x, x_ref) <= TOL_F32
    np.testing.assert_allclose(v, v_ref, rtol=2e-5, atol=1e-9)
    if name in ("large", "two", "one"):          # the BLAS path every BASELINE config takes: identical bits
        assert np.array_equal(x, x_ref) and np.array_equal(v, v_ref)


@pytest.mark.parametrize("name", HESTON_CASES)
def test_bit_identical_to_c_oracle(ctx, golden_heston, ref_params, name):
    from oracle import c_oracle
    R, N, Z, _, _, p = heston_case(golden_heston, name, ref_params)
    xo, vo = c_oracle.heston_injected(Z, R, N, **p)
    x, v = injected(ctx, Z, R, N, p)
    assert np.array_equal(x, xo) and np.array_equal(v, vo)
    xh, vh = injected_host(ctx, Z, R, N, p)      # host-buffer entry point: same kernel behind copies
    assert np.array_equal(xh, xo) and np.array_equal(vh, vo)


@pytest.mark.parametrize("R,N", [(100_000, 101), (100_001, 101), (3, 2), (12_345, 7)])
def test_config_c1_size_f32(ctx, ref_params, R, N):
    """BASELINE config C1 (100k paths x 100 steps f32) and ragged neighbours vs the C oracle."""
    from oracle import c_oracle
    q = ref_params
    p = dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
             x0=q["x0"], v0=q["v0"])
    Z = np.random.default_rng(R + N).standard_normal((N - 1, (R + 1) // 2, 2)).astype(np.float32)
    xo, vo = c_oracle.heston_injected(Z, R, N, **p)
    x, v = injected(ctx, Z, R, N, p)
    assert rel_err_in_price(x, xo) <= TOL_F32
    assert np.array_equal(x, xo) and np.array_equal(v, vo)
    H = (R + 1) // 2                              # antithetic layout: partner of i is i + ceil(R/2)
    assert x.shape == (R,) and np.all(v >= np.float32(1.1920929e-07))
    if R >= 2 and N == 2:                         # one step: x1 + x2 = 2*(x0 + drift) exactly up to rounding
        np.testing.assert_allclose(x[:R // 2] + x[H:], 2 * (q["x0"] + (q["r"] - 0.5 * q["v0"]) * q["T"]), atol=1e-6)


@pytest.mark.parametrize("R,N", [(50_001, 64), (1000, 253)])
def test_f64_extension(ctx, ref_params, R, N):
    """float64 state (BASELINE config C5's precision): 1e-12 relative against the C oracle."""
    from oracle import c_oracle
    q = ref_params
    p = dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
             x0=q["x0"], v0=q["v0"])
    Z = np.random.default_rng(R * N).standard_normal((N - 1, (R + 1) // 2, 2))
    xo, vo = c_oracle.heston_injected(Z, R, N, **p)
    x, v = injected(ctx, Z, R, N, p)
    assert x.dtype == np.float64
    assert rel_err_in_price(x, xo) <= TOL_F64
    np.testing.assert_allclose(v, vo, rtol=1e-11, atol=1e-18)


def test_argument_errors(ctx, ref_params):
    q = ref_params
    p = dict(T=q["T"], mu=q["r"], kappa=q["kappa"], v_bar=q["v_bar"], sigma_v=q["sigma_v"], rho=q["rho"],
             x0=q["x0"], v0=q["v0"])
    Z = np.zeros((1, 1, 2), np.float32)
    with pytest.raises(ValueError):
        injected(ctx, Z, 1, 1, p)                # N must be >= 2
    with pytest.raises(ValueError):
        injected(ctx, Z, 1, 2, dict(p, rho=1.5))
