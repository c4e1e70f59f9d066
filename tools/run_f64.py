"""float64 strike sweep once (for ncu): python tools/run_f64.py [R] [N] [reps]"""
import ctypes
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402

R = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 253
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
device.init(device_id=0)
ctx = device.current_context()
L = _lib.lib()
HP = _lib.HestonParams(1.0, math.log(1.0319), 6.21, 0.019, 0.61, -0.7, 0.0, 0.101 ** 2)
k_arr, nK = _lib.strikes_array([0.70 + 0.60 * i / 63 for i in range(64)])
for rep in range(reps):
    ctx.timer_start()
    _lib.check(L.cb_heston_price_launch_f64(ctx.handle, ctypes.byref(HP), R, N, k_arr, nK, 0, rep, 0, (R + 1) // 2, None,
                                            None, None, None))
    ms = ctx.timer_stop()
    print(f"f64 64 strikes R={R} N={N}: {ms:.3f} ms  {R * (N - 1) / ms * 1e3:.3e} path-steps/s")
