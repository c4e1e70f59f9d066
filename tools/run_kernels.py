"""Launch each secondary kernel a few times (for ncu): python tools/run_kernels.py [reps]

pi (antithetic, 1e9 points), rand/randn f32 (2^28 elements), the path-materialising Heston kernel with x only and with
x and v (4M paths x 100 steps), the strike-sweep and float64 pricing kernels.
"""
import ctypes
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402
from cocos_b200.numerics._array import empty_on  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
device.init(device_id=0)
ctx = device.current_context()
L = _lib.lib()
HP = _lib.HestonParams(1.0, math.log(1.0319), 6.21, 0.019, 0.61, -0.7, 0.0, 0.101 ** 2)
k_arr, nK = _lib.strikes_array([0.95])


def timed(label, fn, work, unit):
    for rep in range(reps):
        ctx.timer_start()
        fn(rep)
        ms = ctx.timer_stop()
    print(f"{label}: {ms:.3f} ms  {work / ms * 1e3:.3e} {unit}")


n = 1_000_000_000
timed("pi antithetic 1e9", lambda r: _lib.check(L.cb_pi_count_launch(ctx.handle, n, 0, r, 1, 0, (n + 1) // 2, None)), n, "points/s")
m = 1 << 28
buf = empty_on(ctx, (m,), np.float32)
timed("rand f32 2^28", lambda r: _lib.check(L.cb_rand_f32(ctx.handle, buf.data_ptr, m, 1, r)), 4 * m, "B/s")
timed("randn f32 2^28", lambda r: _lib.check(L.cb_randn_f32(ctx.handle, buf.data_ptr, m, 1, r)), 4 * m, "B/s")
del buf
R, N = 4_000_000, 101
pairs = R // 2
x, v = empty_on(ctx, (R,), np.float32), empty_on(ctx, (R,), np.float32)
tx, tv = empty_on(ctx, (N, R), np.float32), empty_on(ctx, (N, R), np.float32)
timed("trajectory x", lambda r: _lib.check(L.cb_heston_trajectory_launch_f32(
    ctx.handle, ctypes.byref(HP), R, N, k_arr, nK, 0, r, 0, pairs, None, x.data_ptr, v.data_ptr, tx.data_ptr, None)),
    R * (N - 1), "path-steps/s")
timed("trajectory x,v", lambda r: _lib.check(L.cb_heston_trajectory_launch_f32(
    ctx.handle, ctypes.byref(HP), R, N, k_arr, nK, 0, r, 0, pairs, None, x.data_ptr, v.data_ptr, tx.data_ptr,
    tv.data_ptr)), R * (N - 1), "path-steps/s")
ctx.sync()
del tx, tv
R, N = 12_500_000, 253                      # one GPU's shard of C5
s64, n64 = _lib.strikes_array(list(np.linspace(0.8, 1.2, 64)))
timed("pricing f64, 1 strike", lambda r: _lib.check(L.cb_heston_price_launch_f64(
    ctx.handle, ctypes.byref(HP), R, N, k_arr, nK, 0, r, 0, R // 2, None, None, None, None)), R * (N - 1), "path-steps/s")
timed("pricing f64, 64 strikes", lambda r: _lib.check(L.cb_heston_price_launch_f64(
    ctx.handle, ctypes.byref(HP), R, N, s64, n64, 0, r, 0, R // 2, None, None, None, None)), R * (N - 1), "path-steps/s")
ctx.sync()
