"""Tail of the in-kernel normals: observed against expected counts of |z| > t over >= 1e9 draws (chunked).

python tools/tail_counts.py [draws=2^30] [out.json]

The float32 Box-Muller of the fused kernels takes its radius from 23 random bits, U = k/2^23 (k = 1..2^23), so
r = sqrt(-2 ln U) <= sqrt(46 ln 2) = 5.6467 and the mass beyond (2^-23 per radius draw) sits AT the cap; the angle has
19 bits.  The array generator cb_randn_f32 uses the same draw (csrc/rng_kernels.cu: randn6_draws), so its output is
the stream the Heston kernel consumes.  Expected counts are n * erfc(t / sqrt 2).
"""
import ctypes
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402
from cocos_b200.numerics._array import empty_on  # noqa: E402

THRESHOLDS = (3.0, 4.0, 4.5, 5.0, 5.5, 5.6467, 5.65)


def main():
    total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/tail_counts.json"
    device.init(device_id=0)
    ctx = device.current_context()
    L = _lib.lib()
    chunk = 1 << 27
    buf = empty_on(ctx, (chunk,), np.float32)
    counts = {t: 0 for t in THRESHOLDS}
    n = 0
    s1 = s2 = s4 = 0.0
    zmax = 0.0
    for c in range(max(1, total // chunk)):
        _lib.check(L.cb_randn_f32(ctx.handle, buf.data_ptr, chunk, 20261020, c))
        z = np.abs(buf.to_numpy())
        zmax = max(zmax, float(z.max()))
        for t in THRESHOLDS:
            counts[t] += int(np.count_nonzero(z > t))
        z64 = z.astype(np.float64)
        s2 += float(np.dot(z64, z64))
        s4 += float(np.sum(z64 ** 4))
        n += chunk
    rows = []
    for t in THRESHOLDS:
        expected = n * math.erfc(t / math.sqrt(2.0))
        sd = math.sqrt(expected) if expected > 0 else 0.0
        rows.append({"threshold": t, "observed": counts[t], "expected": expected,
                     "z_score": (counts[t] - expected) / sd if sd > 0 else None})
    res = {"draws": n, "max_abs_z": zmax, "radius_cap": math.sqrt(46 * math.log(2.0)), "second_moment": s2 / n,
           "fourth_moment": s4 / n, "tails": rows}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
