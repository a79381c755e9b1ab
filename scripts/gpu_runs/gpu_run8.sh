#!/bin/bash
# 1-GPU: full gpu test-suite, bench, eval bench, ncu captures (dense round, eval_tc), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest8.log
cat gpurun_out/pytest8.log
timeout 300 python bench.py --steps 400 --warmup 40 > gpurun_out/r8_n1.json 2> gpurun_out/r8_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r8_n1.json").read().strip().splitlines()[-1])
print("N=1 it/s=%.1f ms=%.4f GB/s=%.0f e2e=%.1f"%(d["value"],d["ms_per_step"],d["hbm_gbps_per_gpu"],d["e2e"]["value"]), d["last_round_phases_us_rank0_worker0"])
PY
timeout 300 python benchmarks/eval_bench.py --rows 2000000 --T 16 32 64 > gpurun_out/eval_bench.log 2>&1; grep EVAL_RESULT gpurun_out/eval_bench.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_round|eval_tc|split_w" -c 30 --csv --log-file gpurun_out/launches8.csv python bench.py --steps 10 --warmup 3 --rows 2000000 > gpurun_out/ncu_l8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_tc -c 1 -o gpurun_out/prof_eval_tc python benchmarks/eval_bench.py --rows 1000000 --T 32 --only tc > gpurun_out/ncu_e8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_round -s 6 -c 1 -o gpurun_out/prof_dense2 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/ncu_f8.log 2>&1
ls -la gpurun_out/*.ncu-rep
