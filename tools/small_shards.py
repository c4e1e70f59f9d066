"""Pricing pass on small shards (what one of 8 GPUs gets when C3 is strong-scaled) against block size: python tools/small_shards.py"""
import ctypes, sys, os, math
sys.path.insert(0, os.getcwd())
from cocos_b200 import _lib, device
device.init(device_id=0)
ctx = device.current_context(); L=_lib.lib()
hp=_lib.HestonParams(1.0, math.log(1.0319), 6.21, 0.019, 0.61, -0.7, 0.0, 0.101**2)
k_arr,nK=_lib.strikes_array([0.95])
def t(R,nT,threads,bps=0,reps=7):
    ctx.set_launch(threads,bps,0)
    best=1e9
    for r in range(reps):
        ctx.timer_start()
        _lib.check(L.cb_heston_price_launch_f32(ctx.handle, ctypes.byref(hp), R, nT, k_arr, nK, 0, r, 0, (R+1)//2, None, None, None, None))
        best=min(best, ctx.timer_stop())
    return best
for R in (1_250_000, 2_500_000, 5_000_000, 10_000_000, 300_000):
    print(R, ' '.join(f"thr{th}: {t(R,501,th)*1e3:8.1f} us" for th in (256,128,64)), f" ideal {3765*R/1e7:.1f}")
