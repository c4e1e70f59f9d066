"""Summarise one kernel launch of an .ncu-rep as markdown:
python tools/ncu_summary.py report.ncu-rep "title" [kernel-name substring [occurrence]] > out.md
(default: the first launch in the report; occurrence counts matching launches from 0)"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = sys.argv[3] if len(sys.argv) > 3 else None
nth = int(sys.argv[4]) if len(sys.argv) > 4 else 0
name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
matches = [r for r in rows[2:] if want is None or (name_col is not None and want in r[name_col])]
d = matches[nth]
m = {h: (d[i], units[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# ncu: {title}\n")
print(f"Kernel: `{d[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}`\n")
print("| metric | value | unit |\n|---|---|---|")
for k in keys:
    if k in m:
        print(f"| `{k}` | {m[k][0]} | {m[k][1]} |")
st = [(h.split("stalled_")[1], float(d[i])) for i, h in enumerate(hdr)
      if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and d[i] not in ("", "n/a")]
st = sorted([s for s in st if s[1] > 0], key=lambda t: -t[1])
tot = sum(v for _, v in st) or 1
print("\nWarp-state samples (all samples):\n\n| stall | samples | share |\n|---|---|---|")
for s_, v in st[:8]:
    print(f"| {s_} | {int(v)} | {100 * v / tot:.1f}% |")
