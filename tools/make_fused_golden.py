"""Writes tests/golden/fused_c3_sums.json: the payoff sums of the fused kernel on ONE B200 for the fixed job the
device-count-invariance checks use (bench.py extras.invariance, tests/test_gpu_multi.py): R = 10M paths, nT = 501,
seed 0, subsequence 424242, K = 0.95.  Run on a B200:  python tools/make_fused_golden.py [out.json]
The launch is bit-reproducible (fixed-order reduction), so any 1-GPU run must reproduce these numbers exactly and
any G-GPU run to float64 summation order (1e-12 relative)."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402
import ctypes  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fused_c3_sums.json"
device.init(device_id=0)
ctx = device.current_context()
L = _lib.lib()
P = dict(x0=0.0, v0=0.101 ** 2, r=math.log(1.0319), rho=-0.7, sigma_v=0.61, kappa=6.21, v_bar=0.019, T=1.0, K=0.95)
hp = _lib.HestonParams(P["T"], P["r"], P["kappa"], P["v_bar"], P["sigma_v"], P["rho"], P["x0"], P["v0"])
R, nT, seed, sub = 10_000_000, 501, 0, 424242
k_arr, nK = _lib.strikes_array([P["K"]])
vals = []
for rep in range(2):
    _lib.check(L.cb_heston_price_launch_f32(ctx.handle, ctypes.byref(hp), R, nT, k_arr, nK, seed, sub, 0, (R + 1) // 2,
                                            None, None, None, None))
    s, q = (ctypes.c_double * 1)(), (ctypes.c_double * 1)()
    _lib.check(L.cb_heston_price_finish(ctx.handle, nK, s, q))
    vals.append((s[0], q[0]))
assert vals[0] == vals[1], vals
res = {"R": R, "nT": nT, "seed": seed, "subsequence": sub, "K": P["K"], "sum": vals[0][0], "sumsq": vals[0][1],
       "sum_hex": float(vals[0][0]).hex(), "sumsq_hex": float(vals[0][1]).hex(),
       "price": math.exp(-P["r"] * P["T"]) * vals[0][0] / R,
       "made_by": "tools/make_fused_golden.py on one NVIDIA B200 (heston_fused_kernel<float>, default geometry)"}
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
