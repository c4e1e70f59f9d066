"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so (heston_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class HestonParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in
                ("T", "mu", "kappa", "v_bar", "sigma_v", "rho", "x0", "v0")]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "heston_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u64, u32, vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p
        pp = ctypes.POINTER(HestonParams)
        L.oracle_heston_injected_f32.argtypes = [vp, pp, u64, u32, vp, vp]
        L.oracle_heston_injected_f32.restype = None
        L.oracle_heston_injected_f64.argtypes = [vp, pp, u64, u32, vp, vp]
        L.oracle_heston_injected_f64.restype = None
        L.oracle_call_payoff_sums.argtypes = [vp, u64, vp, ctypes.c_int, vp, vp]
        L.oracle_call_payoff_sums.restype = None
        L.oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.oracle_philox4x32_10.restype = None
        L.oracle_philox_words.argtypes = [u64, u64, u32, u32, u64, vp]
        L.oracle_philox_words.restype = None
        L.oracle_pi_inside.argtypes = [u64, u64, u32, ctypes.c_int]
        L.oracle_pi_inside.restype = u64
        _LIB = L
    return _LIB


def _params(T, mu, kappa, v_bar, sigma_v, rho, x0, v0):
    return HestonParams(T, mu, kappa, v_bar, sigma_v, rho, x0, v0)


def heston_injected(Z, R, N, T, mu, kappa, v_bar, sigma_v, rho, x0, v0):
    """Z: (N-1, ceil(R/2), 2) float32 or float64; returns terminal (x, v) of length R."""
    Z = np.ascontiguousarray(Z)
    H = (R + 1) // 2
    assert Z.shape == (N - 1, H, 2), (Z.shape, (N - 1, H, 2))
    x = np.empty(R, dtype=Z.dtype)
    v = np.empty(R, dtype=Z.dtype)
    p = _params(T, mu, kappa, v_bar, sigma_v, rho, x0, v0)
    fn = {np.dtype(np.float32): lib().oracle_heston_injected_f32,
          np.dtype(np.float64): lib().oracle_heston_injected_f64}[Z.dtype]
    fn(Z.ctypes.data, ctypes.byref(p), R, N, x.ctypes.data, v.ctypes.data)
    return x, v


def payoff_sums(x_T, strikes):
    x_T = np.ascontiguousarray(x_T, dtype=np.float64)
    strikes = np.ascontiguousarray(np.atleast_1d(strikes), dtype=np.float64)
    s = np.empty(len(strikes))
    q = np.empty(len(strikes))
    lib().oracle_call_payoff_sums(x_T.ctypes.data, len(x_T), strikes.ctypes.data, len(strikes),
                                  s.ctypes.data, q.ctypes.data)
    return s, q


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.empty(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def philox_words(first, count, block, subsequence, seed):
    o = np.empty((count, 4), dtype=np.uint32)
    lib().oracle_philox_words(first, count, block, subsequence, seed, o.ctypes.data)
    return o


def pi_inside(n, seed, subsequence, antithetic=True):
    return int(lib().oracle_pi_inside(n, seed, subsequence, int(bool(antithetic))))
