"""Random-array kernels against the Philox/Box-Muller oracle (bit-exact where the arithmetic is exact)
and the reference's distributional criterion (KS, alpha = 0.01, cocos/tests/.../test_rng/*)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cn(gpu_device):
    import cocos_b200.numerics as cn
    return cn


def assert_close_to_box_muller_model(z, model):
    """MUFU lg2/sqrt/sin/cos against the float64 model of the same bits.

    lg2.approx has ~2^-22 ABSOLUTE error, so for U within a few ulps of 1 (radius r ~ 1e-3, probability ~1e-6 per
    draw) r = sqrt(-2 ln U) is off by up to ~1e-4; everywhere else the error is a few 1e-6.
    """
    diff = np.abs(z.astype(np.float64) - model)
    assert diff.max() < 1e-3
    assert np.mean(diff > 2e-5 + 1e-5 * np.abs(model)) <= 2e-5
    if diff.size >= 1000:
        assert np.median(diff) < 3e-7


def test_philox_known_answers_on_device(cn, golden_values):
    zero = cn.random.philox_words(1, first_index=0, block=0, subsequence=0, seed_value=0).to_numpy()[0]
    assert [f"{w:08x}" for w in zero] == golden_values["philox_kat"][0]["out"]
    ones = cn.random.philox_words(1, first_index=2 ** 64 - 1, block=0xFFFFFFFF, subsequence=0xFFFFFFFF,
                                  seed_value=2 ** 64 - 1).to_numpy()[0]
    assert [f"{w:08x}" for w in ones] == golden_values["philox_kat"][1]["out"]
    pi = cn.random.philox_words(1, first_index=0x85a308d3243f6a88, block=0x13198a2e, subsequence=0x03707344,
                                seed_value=0x299f31d0a4093822).to_numpy()[0]
    assert [f"{w:08x}" for w in pi] == golden_values["philox_kat"][2]["out"]


def test_philox_words_match_oracle(cn):
    from oracle import c_oracle
    seed, sub, block, first, n = 0xDEADBEEFCAFEF00D, 9, 4, 2 ** 32 - 50, 4099
    got = cn.random.philox_words(n, first_index=first, block=block, subsequence=sub, seed_value=seed).to_numpy()
    assert np.array_equal(got, c_oracle.philox_words(first, n, block, sub, seed))


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 100_003])
def test_rand_bit_exact(cn, n):
    from oracle import philox_numpy as px
    cn.random.seed(123)
    u = cn.random.rand(n).to_numpy()
    assert u.dtype == np.float32 and np.array_equal(u, px.rand_f32(n, 123, 0))
    u64 = cn.random.rand(n, dtype=np.float64).to_numpy()
    assert u64.dtype == np.float64 and np.array_equal(u64, px.rand_f64(n, 123, 1))
    assert cn.random.get_seed() == 123 and cn.random.get_state() == (123, 2)


@pytest.mark.parametrize("n", [2, 5, 6, 7, 191, 192, 193, 4096, 200_001])
def test_randn_close_to_model(cn, n):
    from oracle import philox_numpy as px
    cn.random.seed(5)
    z = cn.random.randn(n).to_numpy()
    model = px.randn_f32_model(n, 5, 0)
    assert_close_to_box_muller_model(z, model)
    z64 = cn.random.randn(n, dtype=np.float64).to_numpy()
    np.testing.assert_allclose(z64, px.randn_f64_model(n, 5, 1), rtol=1e-12, atol=1e-13)


def test_seed_semantics(cn):
    cn.random.seed(None)
    assert cn.random.get_seed() == 0                     # random.py:63-64
    a = cn.random.randn(1000).to_numpy()
    b = cn.random.randn(1000).to_numpy()
    assert not np.array_equal(a, b)                      # successive calls advance the stream
    cn.random.seed(0)
    assert np.array_equal(cn.random.randn(1000).to_numpy(), a)
    m = cn.random.rand(3, 4, 5, 2)
    assert m.shape == (3, 4, 5, 2) and m.dtype == np.float32


def test_distribution_ks(cn):
    from scipy import stats
    cn.random.seed(2024)
    z = cn.random.randn(5_000_000).to_numpy()            # n as in test_normal_rng.py
    assert stats.kstest(z, "norm").pvalue > 0.01
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3
    u = cn.random.rand(1_000_000).to_numpy()             # n as in test_uniform_rng.py
    assert stats.kstest(u, "uniform").pvalue > 0.01
    assert u.min() >= 0.0 and u.max() < 1.0
    z64 = cn.random.randn(1_000_000, dtype=np.float64).to_numpy()
    assert stats.kstest(z64, "norm").pvalue > 0.01


SHAPES = [((5, 2), 0), ((2, 9), 1), ((3, 4, 5), 2), ((7,), 0), ((2, 3, 2, 4), 1), ((100_001, 2), 0),
          ((100_000, 2), 0), ((1, 1), 0), ((4, 6), 0), ((6, 4, 2), 1)]


@pytest.mark.parametrize("shape,axis", SHAPES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_antithetic_layout(cn, shape, axis, dtype):
    """(z, -z[:floor(n/2)]) along the axis, draws in draw-shape order: random.py:190-245."""
    from oracle import heston_numpy as hn, philox_numpy as px
    half = list(shape)
    half[axis] = (shape[axis] + 1) // 2
    n_draw = int(np.prod(half))
    cn.random.seed(77)
    zn = cn.random.randn_antithetic(shape, axis, dtype).to_numpy()
    un = cn.random.rand_antithetic(shape, axis, dtype, num_pack=cn).to_numpy()
    assert zn.shape == tuple(shape) and zn.dtype == dtype
    if dtype == np.float32:
        draw_n, draw_u = px.randn_pairs_f32_model(n_draw, 77, 0), px.rand_f32(n_draw, 77, 1)
    else:
        draw_n, draw_u = px.randn_f64_model(n_draw, 77, 0), px.rand_f64(n_draw, 77, 1)
    want_n = hn.normal_antithetic(shape, axis, dtype, draw=lambda *s: draw_n.reshape(s))
    want_u = hn.uniform_antithetic(shape, axis, dtype, draw=lambda *s: draw_u.reshape(s))
    if dtype == np.float32:
        assert_close_to_box_muller_model(zn, want_n)
    else:
        np.testing.assert_allclose(zn, want_n, rtol=1e-12, atol=1e-13)
    assert np.array_equal(un, want_u)                    # uniforms and 1-u are exact
    # the reflection is exact on the device output itself
    sl = hn.antithetic_slices(shape, axis)
    first = zn[sl]
    second = np.take(zn, range(half[axis], shape[axis]), axis=axis)
    assert np.array_equal(second, -first)


def test_antithetic_errors(cn):
    with pytest.raises(ValueError):
        cn.random.randn_antithetic((2, 2, 2, 2, 2), 0)
    with pytest.raises(ValueError):
        cn.random.randn_antithetic((), 0)
    with pytest.raises(ValueError):
        cn.random.rand_antithetic((4, 2), 2)
    with pytest.raises(ValueError):
        cn.random.rand_antithetic((4, 2), -1)
    with pytest.raises(TypeError):
        cn.random.randn_antithetic((4, 2))               # the reference's default None raises TypeError too
    with pytest.raises(NotImplementedError):
        cn.random.randn_antithetic((4, 2), 0, num_pack=np)
    with pytest.raises(TypeError):
        cn.random.randn(4, dtype=np.int32)


@pytest.mark.parametrize("n", [1, 7, 1000, 70_001])
def test_fills_into_unaligned_output_write_the_same_elements(cn, n):
    """The vectorised fast paths need a 16-byte aligned output; a pointer that is only element-aligned (a view into a
    larger buffer) takes the scalar-store kernels and must produce the same elements, neighbours untouched."""
    import ctypes
    from cocos_b200 import _lib
    from cocos_b200.device import current_context
    from cocos_b200.numerics._array import empty_on
    from oracle import philox_numpy as px
    ctx = current_context()
    L = _lib.lib()
    for sfx, dtype, model in (("rand_f32", np.float32, px.rand_f32), ("rand_f64", np.float64, px.rand_f64)):
        item = np.dtype(dtype).itemsize
        buf = cn.array(np.full(n + 2, -7.0, dtype=dtype))
        _lib.check(getattr(L, "cb_" + sfx)(ctx.handle, ctypes.c_void_p(buf.data_ptr + item), n, 11, 3))
        got = buf.to_numpy()
        assert got[0] == -7.0 and got[-1] == -7.0
        assert np.array_equal(got[1:-1], model(n, 11, 3))
    for sfx, dtype, model, tol in (("randn_f32", np.float32, px.randn_f32_model, None),
                                   ("randn_f64", np.float64, px.randn_f64_model, 1e-12)):
        item = np.dtype(dtype).itemsize
        buf = cn.array(np.full(n + 2, -7.0, dtype=dtype))
        aligned = empty_on(ctx, (n,), dtype)
        _lib.check(getattr(L, "cb_" + sfx)(ctx.handle, ctypes.c_void_p(buf.data_ptr + item), n, 11, 3))
        _lib.check(getattr(L, "cb_" + sfx)(ctx.handle, ctypes.c_void_p(aligned.data_ptr), n, 11, 3))
        got = buf.to_numpy()
        assert got[0] == -7.0 and got[-1] == -7.0
        assert np.array_equal(got[1:-1], aligned.to_numpy())          # same bits as the aligned fast path
    for kind in (0, 2, 4):                                            # uniform, normal, logistic
        buf = cn.array(np.full(n + 2, -7.0, dtype=np.float32))
        aligned = empty_on(ctx, (n,), np.float32)
        _lib.check(L.cb_random_distribution_f32(ctx.handle, ctypes.c_void_p(buf.data_ptr + 4), n, kind, 0.5, 2.0, 11, 3))
        _lib.check(L.cb_random_distribution_f32(ctx.handle, ctypes.c_void_p(aligned.data_ptr), n, kind, 0.5, 2.0, 11, 3))
        got = buf.to_numpy()
        assert got[0] == -7.0 and got[-1] == -7.0 and np.array_equal(got[1:-1], aligned.to_numpy())
