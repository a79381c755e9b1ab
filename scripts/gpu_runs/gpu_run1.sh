#!/bin/bash
# First GPU pass: numerics tests, smoke, 1-GPU bench, launch list, one full ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest1.log
timeout 120 python __graft_entry__.py > gpurun_out/smoke1.log 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/launches1.csv python bench.py --steps 10 --warmup 3 --rows 2000000 > gpurun_out/ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_round -s 4 -c 2 -o gpurun_out/prof_dense1 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/ncu_f.log 2>&1
cat gpurun_out/pytest1.log gpurun_out/smoke1.log gpurun_out/bench1.json
tail -3 gpurun_out/bench1.err
