"""Issue-slot model of the fused kernels, computed from the SASS of the built library (no GPU needed).

    python tools/issue_model.py [clock_GHz]

B200 calibration (tools/measure_peaks.py, kinds 0/6/11-17): per warp instruction and SM sub-partition an
IMAD.WIDE.U32 / IMAD.HI holds the issue port for 4 cycles, an FP64 instruction for 2, everything else for 1; nothing
else issues meanwhile.  cycles per pair-step = (weighted instruction count of the time loop) / (pair-steps per loop
trip); path-steps/s = 2 * 32 lanes * 4 sub-partitions * 148 SMs * clock / cycles.  DESIGN.md section 5.1 compares
the prediction with the measured rates (profiles/r01_aux_kernels_1gpu.json, profiles/r01_bench_c3_1gpu.json).
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cocos_b200", "libcocos_b200.so")
CLOCK = float(sys.argv[1]) * 1e9 if len(sys.argv) > 1 else 1.955e9
SMSP = 148 * 4

# kernel (substring of the mangled name), MUFU per pair-step, pairs per thread, label, measured key
KERNELS = [
    ("heston_fused_kernelIfLi1ELb0ELb1ELi1ELi10ELi0ELb0E", 5, "f32 Euler, 1 strike (headline)", "fused_f32_1strike_C3"),
    ("heston_fused_kernelIfLi1ELb1ELb1ELi3ELi10ELi0ELb0E", 5, "f32 Euler, strike sweep", "fused_f32_64strikes"),
    ("heston_fused_kernelIdLi2ELb0ELb1ELi2ELi10ELi0ELb0E", 5, "f64 Euler, 1 strike", "fused_f64_1strike"),
    ("heston_fused_kernelIdLi2ELb1ELb1ELi2ELi10ELi0ELb0E", 5, "f64 Euler, strike sweep", "fused_f64_64strikes_C5_shard"),
    ("heston_conditional_kernelILb0ELb0E", 4, "f32 conditional scheme", None),
]
FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}


def loops(text, pat):
    for f in re.split(r"\n\s*Function : ", text)[1:]:
        name = f.split("\n", 1)[0].strip()
        if pat not in name:
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
                if any("MUFU" in x for x in body) and (best is None or len(body) < len(best)):
                    best = body
        return best
    return None


def main():
    text = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "profiles", "r02_aux_kernels_1gpu.json")))
    except OSError:
        pass
    print(f"clock {CLOCK / 1e9:.3f} GHz, {SMSP} sub-partitions; weights: IMAD.WIDE/HI 4, FP64 2, other 1")
    print(f"{'kernel':34s} {'instr':>6s} {'WIDE':>5s} {'FP64':>5s} {'MUFU':>5s} {'cyc/pair-step':>14s} {'model path-steps/s':>19s} {'measured':>10s}")
    for pat, mufu_per_step, label, key in KERNELS:
        body = loops(text, pat)
        if body is None:
            print(f"{label:34s}  (not in the library)")
            continue
        ops = [re.sub(r"^@!?U?P\d\s+", "", x).split()[0] for x in body]
        wide = sum(o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI") for o in ops)
        fp64 = sum(o.split(".")[0] in FP64 for o in ops)
        mufu = sum(o.startswith("MUFU") for o in ops)
        cycles = len(ops) + 3 * wide + fp64
        steps = mufu / mufu_per_step                      # pair-steps per loop trip
        cps = cycles / steps
        rate = 2 * 32 * SMSP * CLOCK / cps
        got = measured.get(key, {}).get("path_steps_per_s") if key else None
        print(f"{label:34s} {len(ops):6d} {wide:5d} {fp64:5d} {mufu:5d} {cps:14.1f} {rate:19.3e} "
              f"{(f'{got:.3e}' if got else '-'):>10s}")
    print("MUFU floor: 8 cycles per MUFU warp instruction on the XU pipe (16 lanes/clk/SM): 40 cycles per Euler pair-step")


if __name__ == "__main__":
    main()
