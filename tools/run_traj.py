"""Path-materialising mode once (for ncu): python tools/run_traj.py [R] [N] [reps]"""
import ctypes
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402
from cocos_b200.numerics._array import empty_on  # noqa: E402

R = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 101
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 0
device.init(device_id=0)
ctx = device.current_context()
if threads:
    ctx.set_launch(threads, 0, 0)
L = _lib.lib()
HP = _lib.HestonParams(1.0, math.log(1.0319), 6.21, 0.019, 0.61, -0.7, 0.0, 0.101 ** 2)
pairs = (R + 1) // 2
x, v = empty_on(ctx, (2 * pairs,), np.float32), empty_on(ctx, (2 * pairs,), np.float32)
traj = empty_on(ctx, (N, 2 * pairs), np.float32)
k_arr, nK = _lib.strikes_array([0.95])
for rep in range(reps):
    ctx.timer_start()
    _lib.check(L.cb_heston_price_launch_f32(ctx.handle, ctypes.byref(HP), R, N, k_arr, nK, 0, rep, 0, pairs, None,
                                            x.data_ptr, v.data_ptr, traj.data_ptr))
    ms = ctx.timer_stop()
    print(f"trajectory R={R} N={N}: {ms:.3f} ms  {R * (N - 1) / ms * 1e3:.3e} path-steps/s  "
          f"{N * R * 4 / ms / 1e6:.0f} GB/s written")
