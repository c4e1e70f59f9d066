"""``cocos_b200.numerics.random``: the hot-path slice of cocos/numerics/random.py.

Same names and argument meaning as the reference for
  seed, get_seed                       random.py:58-73
  get_antithetic_slices                random.py:80-100
  verify_shape_and_antithetic_dimension random.py:103-120
  rand, randn                          random.py:127-153
  rand_with_dtype, randn_with_dtype    random.py:172-187
  randn_antithetic, rand_antithetic    random.py:190-245
and, one fused kernel each on the same generator,
  randint, choice, uniform, exponential, standard_exponential, gamma, standard_gamma, chisquare, beta, wald,
  normal, standard_normal, lognormal, logistic, multivariate_normal        random.py:252-531

Generator: Philox4x32-10 (the reference's default engine, random.py:30-51), key =
seed, one fresh 32-bit ``subsequence`` per kernel launch.  Host state is (seed,
launch counter), per thread: ``ComputeDevicePool`` gives each batch its own derived
seed so that results do not depend on scheduling (the reference has no stream
separation between its worker processes at all).
"""
import ctypes
import math
import threading
import typing as tp

import numpy as np

from .. import _lib
from ..device import current_context
from ._array import empty_on, ndarray

_MASK64 = (1 << 64) - 1


class _GeneratorState(threading.local):
    def __init__(self):
        self.seed = 0
        self.counter = 0


_state = _GeneratorState()


def seed(seed: tp.Optional[int] = None):
    """Seed the generator (None -> 0, as the reference does) and rewind the launch counter."""
    if seed is None:
        seed = 0
    _state.seed = int(seed) & _MASK64
    _state.counter = 0


def get_seed() -> int:
    return _state.seed


def get_state() -> tp.Tuple[int, int]:
    """(seed, launches issued so far)."""
    return _state.seed, _state.counter


def set_state(seed_value: int, counter: int = 0):
    _state.seed = int(seed_value) & _MASK64
    _state.counter = int(counter)


def next_subsequence() -> int:
    """Reserve the Philox subsequence of the next kernel launch."""
    s = _state.counter
    _state.counter += 1
    return s & 0xFFFFFFFF


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK64
    return x ^ (x >> 31)


def derive_seed(parent_seed: int, epoch: int, index: int) -> int:
    """Seed of batch ``index`` of the ``epoch``-th fan-out under ``parent_seed``."""
    return splitmix64(parent_seed ^ splitmix64((epoch << 32) ^ index ^ 0xC0C05B200))


# ---------------------------------------------------------------------------
# antithetic helpers (pure Python, pinned by the reference's own test)
# ---------------------------------------------------------------------------

def get_antithetic_slices(shape: tp.Sequence[int], antithetic_dimension: int) -> tp.Tuple[slice, ...]:
    """Slices selecting the first floor(d/2) entries along the antithetic axis."""
    return tuple(slice(0, math.floor(d / 2), 1) if axis == antithetic_dimension else slice(0, d, 1)
                 for axis, d in enumerate(shape))


def verify_shape_and_antithetic_dimension(shape: tp.Sequence[int], antithetic_dimension: tp.Optional[int] = None):
    if len(shape) > 4:
        raise ValueError('arrays with more than 4 axes are not supported')
    if len(shape) < 1:
        raise ValueError('array must have at least one axis')
    # antithetic_dimension=None raises TypeError here, exactly like the reference (random.py:118)
    if antithetic_dimension < 0 or antithetic_dimension > len(shape) - 1:
        raise ValueError(f'antithetic dimension must be None or between 0 and {len(shape)}')


# ---------------------------------------------------------------------------
# array draws
# ---------------------------------------------------------------------------

_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def _suffix(dtype) -> str:
    try:
        return _SUFFIX[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise TypeError(f"dtype {dtype} is not supported (float32 and float64 are)") from None


def _dims(d0, d1, d2, d3) -> tp.Tuple[int, ...]:
    dims = [d for d in (d0, d1, d2, d3) if d is not None]
    if any(int(d) < 0 for d in dims):
        raise ValueError("negative dimensions are not allowed")
    return tuple(int(d) for d in dims)


def _fill(kind: str, shape: tp.Tuple[int, ...], dtype) -> ndarray:
    ctx = current_context()
    out = empty_on(ctx, shape, dtype)
    fn = getattr(_lib.lib(), f"cb_{kind}_{_suffix(dtype)}")
    _lib.check(fn(ctx.handle, out.data_ptr, out.size, _state.seed, next_subsequence()))
    return out


def rand(d0: int, d1: tp.Optional[int] = None, d2: tp.Optional[int] = None, d3: tp.Optional[int] = None,
         dtype: np.generic = np.float32) -> ndarray:
    """Uniform [0,1) values in a given shape (up to 4 axes)."""
    return _fill("rand", _dims(d0, d1, d2, d3), dtype)


def randn(d0: int, d1: tp.Optional[int] = None, d2: tp.Optional[int] = None, d3: tp.Optional[int] = None,
          dtype: np.generic = np.float32) -> ndarray:
    """Standard normal values in a given shape (up to 4 axes)."""
    return _fill("randn", _dims(d0, d1, d2, d3), dtype)


def _reject_host_package(num_pack):
    if num_pack is np:
        raise NotImplementedError(
            "cocos_b200 has no CPU branch: pass num_pack=cocos_b200.numerics (or omit it); the NumPy "
            "restatement of this path lives in oracle/ as test infrastructure only")


def rand_with_dtype(shape: tp.Sequence[int], dtype: np.generic = np.float32, num_pack=None) -> ndarray:
    _reject_host_package(num_pack)
    return rand(*shape, dtype=dtype)


def randn_with_dtype(shape: tp.Sequence[int], dtype: np.generic = np.float32, num_pack=None) -> ndarray:
    _reject_host_package(num_pack)
    return randn(*shape, dtype=dtype)


def _antithetic(kind: str, shape, antithetic_dimension, dtype, num_pack) -> ndarray:
    _reject_host_package(num_pack)
    shape = tuple(int(s) for s in shape)
    verify_shape_and_antithetic_dimension(shape, antithetic_dimension)
    ctx = current_context()
    out = empty_on(ctx, shape, dtype)
    dims = (ctypes.c_uint64 * len(shape))(*shape)
    fn = getattr(_lib.lib(), f"cb_{kind}_antithetic_{_suffix(dtype)}")
    _lib.check(fn(ctx.handle, out.data_ptr, dims, len(shape), int(antithetic_dimension), _state.seed,
                  next_subsequence()))
    return out


def randn_antithetic(shape: tp.Sequence[int], antithetic_dimension: tp.Optional[int] = None,
                     dtype: np.generic = np.float32, num_pack=None) -> ndarray:
    """(z, -z[:floor(n/2)]) along ``antithetic_dimension``: ceil(n/2) draws, the rest reflected."""
    return _antithetic("randn", shape, antithetic_dimension, dtype, num_pack)


def rand_antithetic(shape: tp.Sequence[int], antithetic_dimension: tp.Optional[int] = None,
                    dtype: np.generic = np.float32, num_pack=None) -> ndarray:
    """(u, 1-u[:floor(n/2)]) along ``antithetic_dimension``."""
    return _antithetic("rand", shape, antithetic_dimension, dtype, num_pack)


def philox_words(n_blocks: int, first_index: int = 0, block: int = 0, subsequence: int = 0,
                 seed_value: tp.Optional[int] = None) -> ndarray:
    """Raw Philox4x32-10 output blocks, shape (n_blocks, 4) uint32 (known-answer tests)."""
    ctx = current_context()
    out = empty_on(ctx, (int(n_blocks), 4), np.uint32)
    s = _state.seed if seed_value is None else int(seed_value) & _MASK64
    _lib.check(_lib.lib().cb_philox_words(ctx.handle, out.data_ptr, int(n_blocks), int(first_index), int(block), s,
                                          int(subsequence)))
    return out


# ---------------------------------------------------------------------------
# derived distributions (random.py:252-531): one fused kernel each, float32 like the reference
# ---------------------------------------------------------------------------

SIZE_TYPE = tp.Optional[tp.Union[int, tp.Sequence[int]]]
_UNIFORM, _EXPONENTIAL, _NORMAL, _LOGNORMAL, _LOGISTIC, _WALD, _GAMMA, _BETA = range(8)


def _count_and_shape(size: SIZE_TYPE) -> tp.Tuple[int, tp.Optional[tp.Tuple[int, ...]]]:
    """Element count and final shape for the reference's ``size`` convention (_draw_and_reshape, random.py:306-326):
    None -> one draw returned as a Python scalar, int -> 1-D, tuple/list -> that shape."""
    if size is None or (not isinstance(size, (list, tuple)) and not size):
        return 1, None
    if isinstance(size, (int, np.integer)):
        return int(size), (int(size),)
    if isinstance(size, (list, tuple)):
        return int(np.prod(size)), tuple(int(s) for s in size)
    raise TypeError("size must be either of type int or tuple")


def _finish(out: ndarray, shape, scalar_if_none: bool):
    if shape is None:
        return out.item() if scalar_if_none else out
    return out.reshape(shape) if len(shape) != 1 else out


def _distribution(kind: int, a: float, b: float, size: SIZE_TYPE):
    n, shape = _count_and_shape(size)
    ctx = current_context()
    out = empty_on(ctx, (n,), np.float32)
    _lib.check(_lib.lib().cb_random_distribution_f32(ctx.handle, out.data_ptr, n, kind, float(a), float(b), _state.seed,
                                                     next_subsequence()))
    return _finish(out, shape, scalar_if_none=(size is None))


def uniform(low: float = 0.0, high: float = 1.0, size: SIZE_TYPE = None):
    """Draw samples from a uniform distribution on [low, high)."""
    if high < low:
        raise ValueError("high must not be less than low")
    return _distribution(_UNIFORM, low, high, size)


def exponential(scale: float = 1.0, size: SIZE_TYPE = None, antithethic: bool = False):
    return _distribution(_EXPONENTIAL, scale, 0.0, size)


def standard_exponential(size: SIZE_TYPE = None):
    return exponential(size=size)


def normal(loc: float = 0.0, scale: float = 1.0, size: SIZE_TYPE = None):
    return _distribution(_NORMAL, loc, scale, size)


def standard_normal(size: SIZE_TYPE = None):
    return _distribution(_NORMAL, 0.0, 1.0, size)


def lognormal(mean: float = 0.0, sigma: float = 1.0, size: SIZE_TYPE = None):
    return _distribution(_LOGNORMAL, mean, sigma, size)


def logistic(loc: float = 0.0, scale: float = 1.0, size: SIZE_TYPE = None):
    return _distribution(_LOGISTIC, loc, scale, size)


def wald(mean: float, scale: float, size: SIZE_TYPE = None):
    return _distribution(_WALD, mean, scale, size)


def gamma(shape: float, scale: float = 1.0, size: SIZE_TYPE = None):
    """Marsaglia-Tsang rejection sampling, one thread per value (the reference loops over array passes)."""
    return _distribution(_GAMMA, shape, scale, size)


def standard_gamma(shape: float, size: SIZE_TYPE = None):
    return gamma(shape, size=size)


def chisquare(df, size: SIZE_TYPE = None):
    return gamma(df / 2.0, 2.0, size)


def beta(a: float, b: float, size: SIZE_TYPE = None):
    return _distribution(_BETA, a, b, size)


def randint(low: int, high: tp.Optional[int] = None, size: tp.Optional[tp.Union[tp.Tuple[int, ...], int]] = None,
            dtype: np.generic = np.int32) -> ndarray:
    """Random integers in [low, high) (or [0, low) when ``high`` is omitted) of the given shape."""
    if not high:
        high, low = low, 0
    if not size:
        size = (1,)
    elif isinstance(size, (int, np.integer)):
        size = (int(size),)
    n = int(np.prod(size))
    ctx = current_context()
    is64 = np.dtype(dtype) != np.dtype(np.int32)
    if is64 and np.dtype(dtype) != np.dtype(np.int64):
        raise TypeError("randint produces int32 or int64 on the device")
    out = empty_on(ctx, (n,), np.int64 if is64 else np.int32)
    _lib.check(_lib.lib().cb_randint(ctx.handle, out.data_ptr, int(is64), n, int(low), int(high), _state.seed,
                                     next_subsequence()))
    return out.reshape(tuple(int(s) for s in size))


def choice(a: ndarray, size: tp.Optional[tp.Union[tp.Tuple[int, ...], int]] = None, replace: bool = True,
           p=None) -> ndarray:
    """Uniform draws with replacement from a device array."""
    if p:
        raise ValueError('p != None is not supported')
    if not replace:
        raise ValueError('replace=False is not supported')
    if not isinstance(a, ndarray):
        raise TypeError("a must be a cocos_b200 device array")
    index = randint(0, a.size, size=size)
    ctx = current_context()
    flat = a.reshape(-1)
    out = empty_on(ctx, index.shape, a.dtype)
    _lib.check(_lib.lib().cb_gather(ctx.handle, out.data_ptr, flat.data_ptr, index.data_ptr, index.size,
                                    a.dtype.itemsize))
    return out


def multivariate_normal(mean, cov, size) -> ndarray:
    """Rows of N(mean, cov): randn(n, d) @ cholesky(cov).T + mean, shape size + (d,).

    (The reference computes ``z @ cholesky(cov).T`` and drops ``mean`` -- random.py:523-531; NumPy's semantics,
    which its signature promises, are implemented here.)  The d x d factorisation is parameter preparation on the host.
    """
    mean = np.asarray(mean.to_numpy() if isinstance(mean, ndarray) else mean)
    cov = np.asarray(cov.to_numpy() if isinstance(cov, ndarray) else cov)
    d = len(mean)
    if not isinstance(size, (list, tuple)):
        size = [size]
    if not cov.shape == (d, d):
        raise ValueError('mean and cov must be a arrays with shapes (d, ) and (d, d) respectively')
    if mean.dtype != cov.dtype:
        raise ValueError('the dtypes of mean and cov mjust match')
    sfx = _suffix(cov.dtype)
    rows = int(np.prod(size))
    chol = np.ascontiguousarray(np.linalg.cholesky(cov), dtype=cov.dtype)
    mean = np.ascontiguousarray(mean, dtype=cov.dtype)
    ctx = current_context()
    out = empty_on(ctx, (rows, d), cov.dtype)
    fn = getattr(_lib.lib(), f"cb_multivariate_normal_{sfx}")
    _lib.check(fn(ctx.handle, out.data_ptr, rows, d, chol.ctypes.data, mean.ctypes.data, _state.seed, next_subsequence()))
    return out.reshape(tuple(int(s) for s in size) + (d,))
