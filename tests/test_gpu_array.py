"""The device array handle (cocos_b200/numerics/_array.py): storage from cb_malloc/cb_free, contiguous views,
astype / transpose / exp through the C ABI, zero-copy export -- and no torch anywhere in the product path."""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(gpu_device):
    return gpu_device.current_context()


def test_views_are_pointer_arithmetic(ctx):
    import cocos_b200.numerics as cn
    host = np.arange(4 * 6 * 5, dtype=np.float32).reshape(4, 6, 5)
    a = cn.array(host)
    assert a.shape == host.shape and a.dtype == np.float32 and a.size == host.size and len(a) == 4
    np.testing.assert_array_equal(a.to_numpy(), host)
    for key in (2, -1, slice(1, 3), (1, 2), (3, slice(2, 5)), (slice(1, 4),), (0, 5, slice(1, 3)), Ellipsis,
                (1, Ellipsis), (slice(None), slice(None), slice(None))):
        np.testing.assert_array_equal(a[key].to_numpy(), host[key])
    assert a[1].data_ptr == a.data_ptr + 6 * 5 * 4                     # a view, not a copy
    np.testing.assert_array_equal(a.reshape(-1, 5).to_numpy(), host.reshape(-1, 5))
    np.testing.assert_array_equal(a.reshape(24, 5)[3:7].to_numpy(), host.reshape(24, 5)[3:7])
    for bad in ((slice(None), 1), (slice(0, 2), slice(0, 2)), slice(0, 4, 2)):
        with pytest.raises(NotImplementedError):
            a[bad]
    with pytest.raises(IndexError):
        a[4]
    with pytest.raises(ValueError):
        a.reshape(7, -1)
    with pytest.raises(NotImplementedError):
        a + a                                                          # elementwise arithmetic is not part of the path


def test_astype_transpose_exp(ctx):
    import cocos_b200.numerics as cn
    rng = np.random.default_rng(0)
    host = rng.standard_normal((37, 129)).astype(np.float32)
    a = cn.array(host)
    np.testing.assert_array_equal(a.astype(np.float64).to_numpy(), host.astype(np.float64))
    np.testing.assert_array_equal(a.astype(np.float64).astype(np.float32).to_numpy(), host)
    np.testing.assert_array_equal(cn.array((host * 100).astype(np.int32)).astype(np.float32).to_numpy(),
                                  (host * 100).astype(np.int32).astype(np.float32))
    np.testing.assert_array_equal(a.T.to_numpy(), host.T)
    np.testing.assert_array_equal(a.astype(np.float64).transpose().to_numpy(), host.astype(np.float64).T)
    np.testing.assert_allclose(cn.exp(a).to_numpy(), np.exp(host), rtol=2e-6)
    np.testing.assert_allclose(cn.log(cn.exp(a.astype(np.float64))).to_numpy(), host.astype(np.float64), atol=1e-14)
    with pytest.raises(TypeError):
        a.astype(np.float16)
    empty = cn.array(np.zeros((0, 3), np.float32))
    assert empty.to_numpy().shape == (0, 3) and empty.astype(np.float64).shape == (0, 3)


def test_allocator_reuses_blocks_and_frees_on_gc(ctx):
    """cb_malloc/cb_free are stream-ordered: a tight allocate/fill/drop loop neither leaks nor synchronises."""
    from cocos_b200.numerics import random as rnd
    rnd.seed(1)
    total = 0.0
    for _ in range(200):
        z = rnd.randn(1 << 20)
        total += float(z[:1].to_numpy()[0])
        del z
    assert np.isfinite(total)


def test_cuda_array_interface_and_no_torch_in_the_product(ctx):
    import cocos_b200.numerics as cn
    from cocos_b200 import heston, monte_carlo_pi  # noqa: F401
    from cocos_b200.scientific import kde  # noqa: F401
    a = cn.array(np.arange(10, dtype=np.float64))
    iface = a[2:7].__cuda_array_interface__
    assert iface["shape"] == (5,) and iface["typestr"] == "<f8" and iface["data"][0] == a.data_ptr + 16
    if "torch" not in sys.modules:           # nothing above may have pulled it in
        import cocos_b200
        assert "torch" not in sys.modules, cocos_b200
    import torch                              # a consumer may adopt the memory without a copy
    t = torch.as_tensor(a, device="cuda")
    assert t.data_ptr() == a.data_ptr and float(t.sum()) == 45.0
