"""Instruction histogram of the hottest loop (most MUFU, shortest) of a kernel in libcocos_b200.so.

python tools/sass_loop.py <substring of the mangled kernel name> [--dump]
"""
import re
import subprocess
import sys
from collections import Counter

so = "cocos_b200/libcocos_b200.so"
pat = sys.argv[1]
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", text)
for f in funcs[1:]:
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
            body = lines[j:i + 1]
            nm = sum("MUFU" in x for _, x in body)
            if nm and (best is None or len(body) < len(best)):
                best = body
    print(name)
    if best is None:
        print("  no MUFU loop found")
        continue
    nm = sum("MUFU" in x for _, x in best)
    print(f"  loop of {len(best)} instructions, {nm} MUFU")
    print("  ", Counter(re.sub(r"^@!?U?P\d\s+", "", x).split()[0].split(".")[0] for _, x in best).most_common())
    if "--dump" in sys.argv:
        for a, x in best:
            print(f"    {a:04x}  {x}")
