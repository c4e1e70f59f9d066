"""TEST INFRASTRUCTURE ONLY -- never imported by the product (cocos_b200/).

Runs the UNMODIFIED reference (michaelnowotny/cocos, mounted at /root/reference)
on its own NumPy branch so that golden vectors can be generated from it.

The reference's GPU branch needs ArrayFire, which is absent here; its NumPy
branch never touches ArrayFire, but the modules import it at the top
(cocos/numerics/random.py:1, cocos/device.py:1) and use ``af.broadcast`` as an
import-time decorator (cocos/numerics/_array.py:321).  We therefore put inert
stand-ins for the missing third-party modules into ``sys.modules`` before
importing the reference.  The reference is imported from where it lies (/root/reference in the build
container) or, on the GPU box, from the byte-for-byte staged copy under
``oracle/_ref/`` (git-ignored, written by oracle/build_ref.py; bench.py's CPU arm
is its only user there).  Tests use the fixtures this file wrote to tests/golden/.
"""
import os
import sys
import types
from unittest import mock

_STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if os.path.isdir("/root/reference/cocos") else _STAGED_ROOT

_STUBBED = (
    "arrayfire", "arrayfire.library", "arrayfire.arith", "arrayfire.array",
    "arrayfire.util", "arrayfire.random", "arrayfire.data",
    "matplotlib", "matplotlib.pyplot", "loky", "pathos", "pathos.pools",
    "contexttimer", "cached_property", "numexpr",
)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cocos")) and \
        os.path.isfile(os.path.join(REFERENCE_ROOT, "examples", "stochastic_volatility", "heston_utilities.py"))


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
        raise RuntimeError("reference tree neither mounted at /root/reference nor staged under " + _STAGED_ROOT)
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


# ---------------------------------------------------------------------------------------------------------------
# The reference's CPU multi-core pricing (heston_utilities.py:167-214): map_reduce_multicore over
# number_of_cores = 5 * cpu_count() equal batches of ceil(R / number_of_cores) paths, the batch prices folded as
# x + y / number_of_cores.  loky/pathos (the reference's process pools) are absent from this image, so the batches
# are dispatched with concurrent.futures; every batch runs the reference's own, unmodified
# simulate_and_compute_option_price on its NumpyBundle.

_REF = None


def reference_batch_price(args):
    """Worker: one batch of the unmodified reference entry point (NumPy bundle), seeded per batch."""
    global _REF
    seed, kwargs = args
    if _REF is None:
        _REF = load_reference()
    import numpy
    numpy.random.seed(seed)
    return _REF.simulate_and_compute_option_price(numerical_package_bundle=_REF.NumpyBundle, **kwargs)


def reference_multicore_price(executor, R, number_of_batches, seed0, **kwargs):
    """simulate_and_compute_option_price_multicore semantics on a concurrent.futures executor.
    Returns (price, paths actually simulated = number_of_batches * ceil(R / number_of_batches))."""
    import math
    per_batch = math.ceil(R / number_of_batches)
    jobs = [(seed0 + i, dict(kwargs, R=per_batch)) for i in range(number_of_batches)]
    price = 0.0
    for p in executor.map(reference_batch_price, jobs):        # ordered: the reference folds in list order
        price = price + p / number_of_batches
    return price, per_batch * number_of_batches
