#!/bin/bash
# ncu of the stream kernel in the N = 8 per-GPU regime (1 GPU, 1/8 of the rows): where do the ~10 us per round go?
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_stream -s 1 -c 1 -f -o /tmp/ncu_small \
    python bench.py --gpus 1 --steps 24 --warmup 3 --no-e2e --no-baseline --rows 1012500 > gpurun_out/ncu_small.log 2>&1
echo "[ncu small] rc=$? $(tail -1 gpurun_out/ncu_small.log | cut -c1-200)"
ncu -i /tmp/ncu_small.ncu-rep --page raw --csv > gpurun_out/ncu_small_raw.csv 2>/dev/null
ncu -i /tmp/ncu_small.ncu-rep --page source --csv > gpurun_out/ncu_small_source.csv 2>/dev/null
ncu -i /tmp/ncu_small.ncu-rep --page details > gpurun_out/ncu_small_details.txt 2>/dev/null
