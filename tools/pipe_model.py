"""Pipe-model probes on one B200: times the peak_mix_kernel<W,F,L,A> instantiations (csrc/peaks.cu) and prints, next
to each, the instruction mix ptxas actually emitted for its loop and the sub-partition cycles one loop trip took.

python tools/pipe_model.py [out.json]

Reading the table: cycles per trip ~ max over pipes of (instructions on the pipe x its issue interval) when the
pipes overlap, ~ the sum when they do not.  The pairs (W only, W+L, W+2L...) and (F only, F+L, F+A) separate the two.
"""
import ctypes
import json
import os
import re
import subprocess
import sys
from collections import Counter

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cocos_b200 import _lib, device  # noqa: E402

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cocos_b200", "libcocos_b200.so")
MIXES = {21: (1, 0, 0, 0), 22: (0, 1, 0, 0), 23: (0, 0, 1, 0), 24: (0, 0, 0, 1), 25: (1, 0, 1, 0), 26: (1, 0, 2, 0),
         27: (1, 0, 4, 0), 28: (1, 1, 0, 0), 29: (1, 2, 0, 0), 30: (1, 4, 0, 0), 31: (0, 1, 1, 0), 32: (0, 2, 1, 0),
         33: (0, 1, 2, 0), 34: (1, 1, 1, 0), 35: (1, 2, 2, 0), 36: (0, 1, 0, 1), 37: (1, 0, 0, 1), 38: (1, 0, 0, 2)}
K_ITERS, CHAINS, UNROLL = 4096, 8, 4


def loop_mix(text, name):
    for f in re.split(r"\n\s*Function : ", text)[1:]:
        if not f.startswith(name):
            continue
        lines = []
        for l in f.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
        best = None
        for i, (addr, ins) in enumerate(lines):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", ins)
            if m and int(m.group(1), 16) < addr:
                j = next(k for k, (a, _) in enumerate(lines) if a == int(m.group(1), 16))
                body = [x for _, x in lines[j:i + 1]]
                if best is None or len(body) > len(best):
                    best = body
        ops = Counter()
        for x in best or []:
            o = re.sub(r"^@!?U?P\d\s+", "", x).split()[0]
            ops["IMAD.WIDE" if o.startswith("IMAD.WIDE") else o.split(".")[0]] += 1
        return ops
    return Counter()


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pipe_model.json"
    device.init(device_id=0)
    ctx = device.current_context()
    text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    res = {}
    clock = float(os.environ.get("SM_GHZ", "1.965")) * 1e9
    print(f"{'kind':>4s} {'W F L A':>8s} {'ms':>8s} {'cyc/trip':>9s}  loop mix (one trip = {UNROLL} x {CHAINS} chain steps)")
    for kind, mix in MIXES.items():
        ops, ms = ctypes.c_double(0), ctypes.c_float(0)
        _lib.check(_lib.lib().cb_peak_measure(ctx.handle, kind, 5, ctypes.byref(ops), ctypes.byref(ms)))
        # one warp runs K_ITERS/UNROLL trips; 8 blocks x 8 warps per SM = 16 warps per sub-partition
        trips = K_ITERS / UNROLL
        cycles_per_trip = ms.value * 1e-3 * clock / (trips * 16)
        name = "_ZN2cb15peak_mix_kernelILi%dELi%dELi%dELi%dEEE" % mix
        mixc = loop_mix(text, name)
        body = {k: v for k, v in mixc.items() if k in ("IMAD.WIDE", "IMAD", "FFMA", "FADD", "FMUL", "LOP3", "IADD3", "VIADD", "MOV", "SHF")}
        res[str(kind)] = {"mix": mix, "ms": ms.value, "cycles_per_trip": cycles_per_trip, "loop": dict(mixc)}
        print(f"{kind:4d} {' '.join(map(str, mix)):>8s} {ms.value:8.4f} {cycles_per_trip:9.1f}  {body}")
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
