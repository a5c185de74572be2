"""
SASS opcode summary per kernel of the built library (cuobjdump -sass): how many DMMA (FP64 tensor), UBLKCP (TMA bulk copy),
SYNCS (mbarrier), LDS/STS (shared), LDL/STL (local = spills), LDG/STG, BAR and total instructions each kernel has.

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bild_b200", "libbild_b200.so")
COLS = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "UBLKCP", "SYNCS", "LDS", "STS", "LDL", "STL", "LDG", "STG", "SHFL", "BAR", "WARPSYNC", "UCGABAR_ARV"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
        kernels[cur]["_total"] += 1


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("bildk::", "").replace("void ", "")
    except Exception:
        return name


arch = re.search(r"arch = (sm_\w+)", sass)
print(f"SASS opcode counts per kernel of bild_b200/libbild_b200.so ({arch.group(1) if arch else '?'}; static instruction counts, `cuobjdump -sass`)\n")
print("| kernel | regs | stack B | total | " + " | ".join(COLS) + " |")
print("|---|---|---|---|" + "---|" * len(COLS))
for name, c in sorted(kernels.items(), key=lambda kv: demangle(kv[0])):
    r, s = usage.get(name, (0, 0))
    print(f"| {demangle(name)} | {r} | {s} | {c['_total']} | " + " | ".join(str(c.get(k, 0)) for k in COLS) + " |")
tot = Counter()
for c in kernels.values():
    tot.update(c)
print(f"\nlibrary totals: {tot['_total']} instructions, DMMA {tot['DMMA']}, UBLKCP {tot['UBLKCP']}, SYNCS {tot['SYNCS']}, LDL {tot['LDL']}, STL {tot['STL']}; "
      "no UTCMMA / TMEM instructions: tcgen05 has no f64 kind, DMMA.8x8x4 is the FP64 tensor path.")
