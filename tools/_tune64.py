import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["x"]
from tools.tune_fused import time_config
from cocos_b200 import device
device.init(device_id=0)
ctx = device.current_context()
K64 = [0.7 + 0.6 * i / 63 for i in range(64)]
for sfx, strikes, R, nT in (("f64", K64, 12_500_000, 253), ("f64", [0.95], 12_500_000, 253), ("f32", K64, 10_000_000, 501)):
    for threads in (64, 128, 256):
        for ppt in (1, 2, 4):
            for bps in (0,):
                try:
                    ms, mean = time_config(ctx, R, nT, sfx, strikes, threads, bps, ppt, reps=3)
                    print(f"{sfx} nK={len(strikes):2d} thr={threads:3d} ppt={ppt}: {ms:7.3f} ms {R*(nT-1)/ms*1e3:.3e}", flush=True)
                except Exception as e:
                    print("skip", sfx, threads, ppt, str(e)[:60])
