"""Measure the pipe peaks the roofline needs (FP32 FMA, MUFU, FP64, Philox issue, HBM write/copy).

Run on a B200:  python tools/measure_peaks.py [out.json]
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402

KINDS = {0: "ffma", 1: "mufu_lg2", 2: "mufu_sqrt", 3: "mufu_sin", 4: "mufu_ex2", 5: "mufu_rsq", 6: "dfma",
         7: "philox_calls", 8: "hbm_write_bytes", 9: "lop3", 10: "hbm_copy_bytes",
         11: "imad_wide", 12: "imad_wide_ffma_pairs", 13: "imad_wide_lop3_pairs", 14: "imad_hi", 15: "imad_lo",
         18: "ffma2_packed", 19: "ffma2_ffma_pairs", 20: "ffma2_lop3_pairs"}


def measure(ctx, kind, repeats=5):
    ops, ms = ctypes.c_double(0), ctypes.c_float(0)
    _lib.check(_lib.lib().cb_peak_measure(ctx.handle, kind, repeats, ctypes.byref(ops), ctypes.byref(ms)))
    return ops.value, ms.value


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/peaks.json"
    device.init(device_id=0)
    ctx = device.current_context()
    dev = device.ComputeDeviceManager.get_current_compute_device()
    res = {"gpu": dev.name, "compute": dev.compute_version, "unit": "thread-level ops per second (FMA = 1 op)"}
    for kind, name in KINDS.items():
        ops, ms = measure(ctx, kind)
        res[name] = ops
        res[name + "_ms"] = ms
        print(f"{name:18s} {ops:.4e} /s   ({ms:.3f} ms)")
    res["fp32_tflops"] = 2 * res["ffma"] / 1e12
    res["fp64_tflops"] = 2 * res["dfma"] / 1e12
    res["mufu_gops_min"] = min(res[k] for k in ("mufu_lg2", "mufu_sqrt", "mufu_sin", "mufu_ex2", "mufu_rsq")) / 1e9
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
