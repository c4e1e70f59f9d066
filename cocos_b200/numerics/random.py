"""``cocos_b200.numerics.random``: the hot-path slice of cocos/numerics/random.py.

Same names and argument meaning as the reference for
  seed, get_seed                       random.py:58-73
  get_antithetic_slices                random.py:80-100
  verify_shape_and_antithetic_dimension random.py:103-120
  rand, randn                          random.py:127-153
  rand_with_dtype, randn_with_dtype    random.py:172-187
  randn_antithetic, rand_antithetic    random.py:190-245
The derived distributions (random.py:252-531) are out of scope.

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
