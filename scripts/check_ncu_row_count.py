#!/usr/bin/env python
"""Cross-check an ncu source-page capture of the dense round kernel against the CPU sampler.

Every sampled row is consumed by exactly one warp, which executes each `FFMA2` of the row body exactly once, so the "Instructions
Executed" column of any FFMA2 line of `profiles/ncu/stream_bf16_source.csv` is the number of rows the captured launch processed.
The capture (`scripts/gpu_runs/ncu_stream.sh`: `bench.py --gpus 1 --steps 3 --warmup 3`, `-s 1 -c 1`) is the TIMED launch of the
pipelined kernel = worker-local rounds 3, 4, 5; the sampler (`core/sampler.py`, the same counter hash the kernel evaluates in
line) gives the exact number of rows those rounds select out of the 8.1 M.  Equality means: no chunk lost, none processed
twice, across dynamic chunk scheduling, the three chunk tiers, round-parity buffers and the pull-ahead."""
import csv
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asyncframework_b200.core.sampler import hash_u24, threshold24  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "ncu", "stream_bf16_source.csv")
rounds = [int(r) for r in (sys.argv[2].split(",") if len(sys.argv) > 2 else "3,4,5".split(","))]
rows = list(csv.reader(open(path)))[2:]
ffma2 = sorted({int(r[5]) for r in rows if "FFMA2" in r[1]})
executed = ffma2[-1]
idx = np.arange(8_100_000, dtype=np.int64)
thr = threshold24(0.1)
expected = sum(int((hash_u24(42 + r + 1, idx) < thr).sum()) for r in rounds)
print(f"kernel: {rows and path}")
print(f"rows processed by the captured launch (FFMA2 executions per instruction): {executed}")
print(f"rows selected by the sampler for rounds {rounds} (seed 42, rate 0.1, 8.1 M rows): {expected}")
print("MATCH" if executed == expected else f"MISMATCH ({executed - expected:+d})")
sys.exit(0 if executed == expected else 1)
