"""TEST INFRASTRUCTURE ONLY -- never imported by the product (cocos_b200/).

Runs the UNMODIFIED reference (michaelnowotny/cocos, mounted at /root/reference)
on its own NumPy branch so that golden vectors can be generated from it.

The reference's GPU branch needs ArrayFire, which is absent here; its NumPy
branch never touches ArrayFire, but the modules import it at the top
(cocos/numerics/random.py:1, cocos/device.py:1) and use ``af.broadcast`` as an
import-time decorator (cocos/numerics/_array.py:321).  We therefore put inert
stand-ins for the missing third-party modules into ``sys.modules`` before
importing the reference.  Nothing of the reference is copied: it is imported
from where it lies.  This file only works in the build container (the GPU box
has no /root/reference); tests use the fixtures it wrote to tests/golden/.
"""
import sys
import types
from unittest import mock

REFERENCE_ROOT = "/root/reference"

_STUBBED = (
    "arrayfire", "arrayfire.library", "arrayfire.arith", "arrayfire.array",
    "arrayfire.util", "arrayfire.random", "arrayfire.data",
    "matplotlib", "matplotlib.pyplot", "loky", "pathos", "pathos.pools",
    "contexttimer", "cached_property", "numexpr",
)


def reference_available() -> bool:
    import os
    return os.path.isdir(REFERENCE_ROOT + "/cocos")


def install_stubs() -> None:
    """Insert MagicMock modules for the reference's absent dependencies."""
    for name in _STUBBED:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    af = sys.modules["arrayfire"]
    # used as a decorator at import time: must hand the function back
    af.broadcast = lambda f, *a, **k: f
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference() -> types.SimpleNamespace:
    """Import the reference's hot-path callables (NumPy branch)."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted at " + REFERENCE_ROOT)
    install_stubs()
    from cocos.numerics.random import (            # noqa: E402
        randn_antithetic, rand_antithetic, get_antithetic_slices)
    from cocos.numerics.numerical_package_bundle import NumpyBundle
    from cocos.multi_processing.utilities import (
        generate_slices_with_batch_size, generate_slices_with_number_of_batches,
        _extract_arguments_and_number_of_batches)
    from examples.stochastic_volatility.heston_utilities import (
        simulate_heston_model, compute_option_price_from_simulated_paths,
        simulate_and_compute_option_price)
    return types.SimpleNamespace(
        randn_antithetic=randn_antithetic,
        rand_antithetic=rand_antithetic,
        get_antithetic_slices=get_antithetic_slices,
        NumpyBundle=NumpyBundle,
        generate_slices_with_batch_size=generate_slices_with_batch_size,
        generate_slices_with_number_of_batches=generate_slices_with_number_of_batches,
        extract_arguments=_extract_arguments_and_number_of_batches,
        simulate_heston_model=simulate_heston_model,
        compute_option_price_from_simulated_paths=compute_option_price_from_simulated_paths,
        simulate_and_compute_option_price=simulate_and_compute_option_price,
    )


class InjectedNormals:
    """Context manager: numpy.random.randn / rand pop pre-generated arrays.

    The reference draws ``numpy.random.randn(ceil(R/2), 2)`` once per time step
    (cocos/numerics/random.py:203-206); we hand it our arrays in that order.
    """

    def __init__(self, normals=None, uniforms=None):
        self._normals = list(normals) if normals is not None else None
        self._uniforms = list(uniforms) if uniforms is not None else None

    def __enter__(self):
        import numpy
        self._saved = (numpy.random.randn, numpy.random.rand)
        if self._normals is not None:
            src = self._normals

            def randn(*shape):
                z = src.pop(0)
                assert tuple(z.shape) == tuple(shape), (z.shape, shape)
                return z
            numpy.random.randn = randn
        if self._uniforms is not None:
            usrc = self._uniforms

            def rand(*shape):
                u = usrc.pop(0)
                assert tuple(u.shape) == tuple(shape), (u.shape, shape)
                return u
            numpy.random.rand = rand
        return self

    def __exit__(self, *exc):
        import numpy
        numpy.random.randn, numpy.random.rand = self._saved
        return False
