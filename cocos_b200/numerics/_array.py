"""Device array handle returned by the random/Heston entry points.

The reference's ``cocos.numerics.ndarray`` wraps an ``af.Array`` and forwards ~200
NumPy-named operations to ArrayFire (cocos/numerics/_array.py:150-803).  That
general façade is out of scope: on the hot path every array
operation is fused into the CUDA kernels.  What remains is an owning handle with
shape/dtype, host conversion (the reference converts through NumPy as well,
_array.py:163-171) and indexing views.  Device memory comes from torch's caching
allocator; the kernels only ever see the raw pointer.
"""
import numpy as np

_TORCH_DTYPES = None


def torch_dtype(dtype):
    global _TORCH_DTYPES
    import torch
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                         np.dtype(np.uint32): torch.uint32, np.dtype(np.int64): torch.int64,
                         np.dtype(np.int32): torch.int32}
    try:
        return _TORCH_DTYPES[np.dtype(dtype)]
    except KeyError:
        raise TypeError(f"dtype {dtype} is not supported on the device (float32/float64, int32/int64)") from None


def empty_on(ctx, shape, dtype):
    """Uninitialised device array on the context's device, ordered on its stream."""
    import torch
    with torch.cuda.stream(ctx.torch_stream()):
        t = torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(dtype),
                        device=torch.device("cuda", ctx.device))
    return ndarray(t, ctx)


class ndarray:
    __array_priority__ = 100

    def __init__(self, tensor, ctx):
        self._t = tensor
        self._ctx = ctx

    # ---- metadata
    @property
    def tensor(self):
        """The backing torch tensor (call cocos_b200.device.sync() before using it on another stream)."""
        return self._t

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def size(self):
        return self._t.numel()

    @property
    def dtype(self):
        return np.dtype(str(self._t.dtype).replace("torch.", ""))

    @property
    def data_ptr(self):
        return self._t.data_ptr()

    @property
    def device_id(self):
        return self._ctx.device

    def __len__(self):
        return self.shape[0]

    # ---- host conversion
    def to_numpy(self):
        self._ctx.sync()
        return self._t.cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.to_numpy().reshape(()))

    def item(self):
        return self.to_numpy().item()

    # ---- views (pure pointer arithmetic)
    def __getitem__(self, key):
        return ndarray(self._t[key], self._ctx)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return ndarray(self._t.reshape(shape), self._ctx)

    def astype(self, dtype):
        """Type conversion on the device (storage plumbing, done by the allocator's library)."""
        import torch
        self._ctx.sync()
        with torch.cuda.stream(self._ctx.torch_stream()):
            return ndarray(self._t.to(torch_dtype(dtype)), self._ctx)

    def __repr__(self):
        return f"cocos_b200.ndarray(shape={self.shape}, dtype={self.dtype}, device={self._ctx.device})"

    def _no_arith(self, *_):
        raise NotImplementedError(
            "cocos_b200 implements the Heston Monte-Carlo hot path as fused kernels; elementwise array "
            "arithmetic is not part of it (use simulate_heston_model / estimate_pi, or .to_numpy()).")

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __neg__ = _no_arith
    __le__ = __lt__ = __ge__ = __gt__ = _no_arith
