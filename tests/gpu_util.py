"""Helpers for the -m gpu tests: every call goes through the C ABI (ctypes) of libcocos_b200.so."""
import ctypes

import numpy as np

from cocos_b200 import _lib
from cocos_b200.numerics._array import empty_on


def hparams(p):
    return _lib.HestonParams(p["T"], p["mu"], p["kappa"], p["v_bar"], p["sigma_v"], p["rho"], p["x0"], p["v0"])


def to_device(ctx, a):
    a = np.ascontiguousarray(a)
    d = empty_on(ctx, a.shape, a.dtype)
    _lib.check(_lib.lib().cb_memcpy_h2d(ctx.handle, d.data_ptr, a.ctypes.data, a.nbytes))
    return d


def injected(ctx, Z, R, N, p):
    """cb_heston_terminal_injected_{f32,f64} on device buffers -> numpy (x, v)."""
    sfx = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[Z.dtype]
    dZ = to_device(ctx, Z)
    dx, dv = empty_on(ctx, (R,), Z.dtype), empty_on(ctx, (R,), Z.dtype)
    hp = hparams(p)
    fn = getattr(_lib.lib(), f"cb_heston_terminal_injected_{sfx}")
    _lib.check(fn(ctx.handle, dZ.data_ptr, ctypes.byref(hp), R, N, dx.data_ptr, dv.data_ptr))
    return dx.to_numpy(), dv.to_numpy()


def injected_host(ctx, Z, R, N, p):
    x, v = np.empty(R, np.float32), np.empty(R, np.float32)
    Z = np.ascontiguousarray(Z, dtype=np.float32)
    hp = hparams(p)
    _lib.check(_lib.lib().cb_heston_terminal_injected_host_f32(ctx.handle, Z.ctypes.data, ctypes.byref(hp), R, N,
                                                               x.ctypes.data, v.ctypes.data))
    return x, v


def fused(ctx, p, R, N, strikes, seed, sub, first=0, count=None, dtype=np.float32, outputs=False, traj=False,
          comm=None, traj_v=False):
    """cb_heston_price_launch (cb_heston_trajectory_launch with traj_v) + finish -> (sums, sumsq[, x, v[, traj[, traj_v]]])
    as numpy."""
    sfx = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dtype)]
    H = (R + 1) // 2
    count = H - first if count is None else count
    k_arr, nK = _lib.strikes_array(strikes)
    hp = hparams(p)
    dx = dv = dt = dtv = None
    traj = traj or traj_v
    if outputs or traj:
        dx, dv = empty_on(ctx, (2 * max(count, 1),), dtype), empty_on(ctx, (2 * max(count, 1),), dtype)
    if traj:
        dt = empty_on(ctx, (N, 2 * max(count, 1)), dtype)
    if traj_v:
        dtv = empty_on(ctx, (N, 2 * max(count, 1)), dtype)
        fn = getattr(_lib.lib(), f"cb_heston_trajectory_launch_{sfx}")
        _lib.check(fn(ctx.handle, ctypes.byref(hp), R, N, k_arr, nK, seed, sub, first, count, comm, dx.data_ptr,
                      dv.data_ptr, dt.data_ptr, dtv.data_ptr))
    else:
        fn = getattr(_lib.lib(), f"cb_heston_price_launch_{sfx}")
        _lib.check(fn(ctx.handle, ctypes.byref(hp), R, N, k_arr, nK, seed, sub, first, count, comm,
                      dx.data_ptr if dx is not None else None, dv.data_ptr if dv is not None else None,
                      dt.data_ptr if dt is not None else None))
    sums, sumsq = (ctypes.c_double * nK)(), (ctypes.c_double * nK)()
    _lib.check(_lib.lib().cb_heston_price_finish(ctx.handle, nK, sums, sumsq))
    out = [np.array(sums), np.array(sumsq)]
    if dx is not None:
        out += [dx.to_numpy(), dv.to_numpy()]
    if dt is not None:
        out.append(dt.to_numpy())
    if dtv is not None:
        out.append(dtv.to_numpy())
    return out


def shard_to_reference_layout(x_shard, count, first, R):
    """[pairs | partners] of a shard -> (values of paths first..first+count, values of their partners)."""
    H, F = (R + 1) // 2, R // 2
    n_partner = max(0, min(first + count, F) - first)
    return x_shard[:count], x_shard[count:count + n_partner]


def pi_count(ctx, n, seed, sub, antithetic, first=0, count=None):
    ndraw = (n + 1) // 2 if antithetic else n
    count = ndraw - first if count is None else count
    _lib.check(_lib.lib().cb_pi_count_launch(ctx.handle, n, seed, sub, int(antithetic), first, count, None))
    inside = ctypes.c_uint64(0)
    _lib.check(_lib.lib().cb_pi_count_finish(ctx.handle, ctypes.byref(inside)))
    return inside.value
