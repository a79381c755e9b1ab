#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export of the dense stream kernel: instruction mix per sampled row and the hottest
SASS lines by stall samples.  Usage: scripts/ncu_source_summary.py [csv] > profiles/ncu/<name>_source_summary.md"""
import csv
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "ncu", "stream_bf16_source.csv")
allrows = list(csv.reader(open(path)))
kernel = allrows[0][1]
rows = allrows[2:]
tot_inst = sum(int(r[5]) for r in rows)
tot_samp = sum(int(r[4]) for r in rows)
nrows = max(int(r[5]) for r in rows if "FFMA2" in r[1])
mix, stall = Counter(), Counter()
for r in rows:
    t = r[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    mix[op] += int(r[5])
    stall[op] += int(r[4])
print(f"# ncu source page, summarised (`{os.path.relpath(path, ROOT)}`)\n")
print(f"Kernel: `{kernel}`\n")
print(f"* warp instructions executed: {tot_inst:,}; sampled rows processed (executions of any row-body `FFMA2`): {nrows:,} "
      f"-> **{tot_inst / nrows:.0f} warp instructions per row**")
print(f"* stall samples: {tot_samp:,}\n")
print("## Instruction mix per row (top 20 opcodes)\n\n| opcode | executed per row | share of stall samples |\n|---|---|---|")
for op, n in mix.most_common(20):
    print(f"| `{op}` | {n / nrows:.2f} | {100 * stall[op] / tot_samp:.1f} % |")
print("\n## Hottest SASS lines by stall samples (top 25)\n\n| samples | share | executions per row | instruction |\n|---|---|---|---|")
for r in sorted(rows, key=lambda r: -int(r[4]))[:25]:
    print(f"| {int(r[4])} | {100 * int(r[4]) / tot_samp:.1f} % | {int(r[5]) / nrows:.2f} | `{r[1].strip()[:90]}` |")
