"""``cocos_b200.numerics``: what ``B200Bundle.module()`` returns.

Only the names the Heston / pi hot path touches exist: the device array handle and
the ``random`` module (cocos/numerics/__init__.py:7-188 exports ~200 more, which are
out of scope).
"""
from . import random                      # noqa: F401
from ._array import ndarray               # noqa: F401


def asnumpy(a):
    return a.to_numpy() if isinstance(a, ndarray) else a


def array(obj, dtype=None) -> ndarray:
    """Host data -> device array (the reference's ``cn.array``, cocos/numerics/_array.py): a plain H2D copy."""
    import numpy as np
    from .. import _lib
    from ..device import current_context
    from ._array import empty_on
    if isinstance(obj, ndarray):
        return obj if dtype is None or np.dtype(dtype) == obj.dtype else obj.astype(dtype)
    host = np.ascontiguousarray(obj, dtype=dtype)
    if host.dtype == np.float64 and dtype is None and not isinstance(obj, np.ndarray):
        host = host.astype(np.float32)                    # Python lists become float32, the device default
    ctx = current_context()
    out = empty_on(ctx, host.shape, host.dtype)
    if host.nbytes:
        _lib.check(_lib.lib().cb_memcpy_h2d(ctx.handle, out.data_ptr, host.ctypes.data, host.nbytes))
    return out


def _unary(a: ndarray, op: int) -> ndarray:
    import numpy as np
    from .. import _lib
    from ._array import empty_on
    if not isinstance(a, ndarray):
        raise TypeError("expected a cocos_b200 device array (no CPU branch)")
    if a.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("exp/log take float32 or float64 device arrays")
    out = empty_on(a._ctx, a.shape, a.dtype)
    _lib.check(_lib.lib().cb_unary(a._ctx.handle, out.data_ptr, a.data_ptr, a.size, op, int(a.dtype.itemsize == 8)))
    return out


def exp(a: ndarray) -> ndarray:
    """Elementwise exp on the device (``cn.exp``): log prices -> prices before the density estimate."""
    return _unary(a, 0)


def log(a: ndarray) -> ndarray:
    """Elementwise natural logarithm on the device (``cn.log``)."""
    return _unary(a, 1)
