"""Device array handle returned by the random/Heston entry points.

The reference's ``cocos.numerics.ndarray`` wraps an ``af.Array`` and forwards ~200
NumPy-named operations to ArrayFire (cocos/numerics/_array.py:150-803).  That
general façade is out of scope: on the hot path every array
operation is fused into the CUDA kernels.  What remains is an owning handle with
shape/dtype, host conversion (the reference converts through NumPy as well,
_array.py:163-171), contiguous views, ``astype`` and a 2-D transpose.

The only native dependency is libcocos_b200.so: storage comes from ``cb_malloc`` /
``cb_free`` (the device's stream-ordered pool, ordered on the owning context's stream),
conversions from ``cb_cast`` / ``cb_transpose``.  ``__cuda_array_interface__`` lets
PyTorch, CuPy or Numba adopt the memory without a copy.
"""
import ctypes

import numpy as np

from .. import _lib

# element kinds of cb_cast
_KIND = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
         np.dtype(np.uint32): 4}


def _dtype(dtype):
    try:
        dt = np.dtype(dtype)
    except TypeError:
        dt = None
    if dt not in _KIND:
        raise TypeError(f"dtype {dtype} is not supported on the device (float32/float64, int32/int64, uint32)")
    return dt


class _Buffer:
    """Owner of one device allocation; freed (stream-ordered) when the last view goes away."""

    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().cb_malloc(ctx.handle, self.nbytes, ctypes.byref(ptr)))
        self.ptr = ptr.value or 0

    def __del__(self):
        try:
            if self.ptr and getattr(self.ctx, "handle", None):
                _lib.lib().cb_free(self.ctx.handle, ctypes.c_void_p(self.ptr))
        except Exception:       # noqa: BLE001  (interpreter shutdown)
            pass
        self.ptr = 0


def empty_on(ctx, shape, dtype):
    """Uninitialised device array on the context's device, ordered on its stream."""
    dt = _dtype(dtype)
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    return ndarray(_Buffer(ctx, n * dt.itemsize), 0, shape, dt, ctx)


class ndarray:
    """C-contiguous device array: (buffer, byte offset, shape, dtype)."""
    __array_priority__ = 100

    def __init__(self, buffer, offset, shape, dtype, ctx):
        self._buf, self._offset, self._shape, self._dt, self._ctx = buffer, int(offset), tuple(shape), np.dtype(dtype), ctx

    # ---- metadata
    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        return int(np.prod(self._shape, dtype=np.int64)) if self._shape else 1

    @property
    def dtype(self):
        return self._dt

    @property
    def itemsize(self):
        return self._dt.itemsize

    @property
    def nbytes(self):
        return self.size * self._dt.itemsize

    @property
    def data_ptr(self):
        return self._buf.ptr + self._offset

    @property
    def device_id(self):
        return self._ctx.device

    @property
    def __cuda_array_interface__(self):
        """Zero-copy hand-over to PyTorch / CuPy / Numba (version 3; the consumer must order itself after the
        context's stream, exported here)."""
        return {"shape": self._shape, "typestr": self._dt.str, "data": (self.data_ptr, False), "version": 3,
                "strides": None, "stream": self._ctx.stream_ptr or 1}

    def __len__(self):
        if not self._shape:
            raise TypeError("len() of a 0-d array")
        return self._shape[0]

    # ---- host conversion
    def to_numpy(self):
        out = np.empty(self._shape, dtype=self._dt)
        if out.size:
            _lib.check(_lib.lib().cb_memcpy_d2h(self._ctx.handle, out.ctypes.data, ctypes.c_void_p(self.data_ptr),
                                                out.nbytes))
        else:
            self._ctx.sync()
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.to_numpy().reshape(()))

    def item(self):
        return self.to_numpy().item()

    # ---- views (pure pointer arithmetic; only what stays C-contiguous)
    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = next(j for j, k in enumerate(key) if k is Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        if len(key) > self.ndim:
            raise IndexError("too many indices for array")
        key = key + (slice(None),) * (self.ndim - len(key))
        offset, shape, lead, sliced = self._offset, [], True, False
        inner = self._dt.itemsize
        strides = []
        for s in reversed(self._shape):
            strides.append(inner)
            inner *= s
        strides.reverse()
        for dim, (k, n, st) in enumerate(zip(key, self._shape, strides)):
            if isinstance(k, (int, np.integer)):
                k = int(k)
                if k < 0:
                    k += n
                if not 0 <= k < n:
                    raise IndexError(f"index {k} is out of bounds for axis {dim} with size {n}")
                if sliced:
                    raise NotImplementedError("an integer index after a sliced axis would not be contiguous")
                offset += k * st
            elif isinstance(k, slice):
                start, stop, step = k.indices(n)
                if step != 1:
                    raise NotImplementedError("only unit-step slices are views of a device array")
                count = max(stop - start, 0)
                full = start == 0 and count == n
                if not lead and not full:
                    raise NotImplementedError("only the leading sliced axis may be partial (the view must stay contiguous)")
                offset += start * st
                shape.append(count)
                sliced = True
                if not full:
                    lead = False
            else:
                raise TypeError("device arrays are indexed with integers and unit-step slices")
        return ndarray(self._buf, offset, tuple(shape), self._dt, self._ctx)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = [int(s) for s in shape]
        if shape.count(-1) > 1:
            raise ValueError("can only specify one unknown dimension")
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64)) if len(shape) > 1 else 1
            if known == 0 or self.size % known:
                raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
            shape[shape.index(-1)] = self.size // known
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
        return ndarray(self._buf, self._offset, tuple(shape), self._dt, self._ctx)

    def is_contiguous(self):
        return True

    # ---- conversions on the device (cb_cast / cb_transpose)
    def astype(self, dtype):
        """Element-type conversion on the device (ndarray.astype, cocos/numerics/_array.py:221-228)."""
        dt = _dtype(dtype)
        out = empty_on(self._ctx, self._shape, dt)
        _lib.check(_lib.lib().cb_cast(self._ctx.handle, ctypes.c_void_p(out.data_ptr), _KIND[dt],
                                      ctypes.c_void_p(self.data_ptr), _KIND[self._dt], self.size))
        return out

    def transpose(self):
        """A new (cols, rows) array for a 2-D one; 0-d and 1-d arrays are returned as they are."""
        if self.ndim < 2:
            return self
        if self.ndim != 2:
            raise NotImplementedError("transpose is implemented for 2-D device arrays")
        rows, cols = self._shape
        out = empty_on(self._ctx, (cols, rows), self._dt)
        _lib.check(_lib.lib().cb_transpose(self._ctx.handle, ctypes.c_void_p(out.data_ptr), ctypes.c_void_p(self.data_ptr),
                                           rows, cols, self._dt.itemsize))
        return out

    T = property(transpose)

    def copy_to(self, ctx):
        """A copy on another context's device (NVLink peer copy) or on the same one."""
        out = empty_on(ctx, self._shape, self._dt)
        if self.size:
            if ctx is self._ctx:
                _lib.check(_lib.lib().cb_memcpy_d2d(ctx.handle, ctypes.c_void_p(out.data_ptr),
                                                    ctypes.c_void_p(self.data_ptr), self.nbytes))
            else:
                _lib.check(_lib.lib().cb_memcpy_peer(ctx.handle, ctypes.c_void_p(out.data_ptr), self._ctx.handle,
                                                     ctypes.c_void_p(self.data_ptr), self.nbytes))
        return out

    def __repr__(self):
        return f"cocos_b200.ndarray(shape={self.shape}, dtype={self.dtype}, device={self._ctx.device})"

    def _no_arith(self, *_):
        raise NotImplementedError(
            "cocos_b200 implements the Heston Monte-Carlo hot path as fused kernels; elementwise array "
            "arithmetic is not part of it (use simulate_heston_model / estimate_pi, or .to_numpy()).")

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __neg__ = _no_arith
    __le__ = __lt__ = __ge__ = __gt__ = _no_arith
