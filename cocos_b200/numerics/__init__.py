"""``cocos_b200.numerics``: what ``B200Bundle.module()`` returns.

Only the names the Heston / pi hot path touches exist: the device array handle and
the ``random`` module (cocos/numerics/__init__.py:7-188 exports ~200 more, which are
out of scope).
"""
from . import random                      # noqa: F401
from ._array import ndarray               # noqa: F401


def asnumpy(a):
    return a.to_numpy() if isinstance(a, ndarray) else a
