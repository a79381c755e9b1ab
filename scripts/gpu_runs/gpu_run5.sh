#!/bin/bash
# 8-GPU scaling pass
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest5.log
cat gpurun_out/pytest5.log
for n in 1 2 4 8; do
  timeout 300 python bench.py --gpus $n --steps 400 --warmup 40 > gpurun_out/sc_ours_n${n}.json 2> gpurun_out/sc_ours_n${n}.err
done
timeout 300 python bench.py --gpus 8 --steps 400 --warmup 40 --workers-per-gpu 1 > gpurun_out/sc_ours_n8_w1.json 2> gpurun_out/sc_ours_n8_w1.err
timeout 300 python bench.py --gpus 8 --steps 400 --warmup 40 --workers-per-gpu 2 --graph 40 --no-e2e > gpurun_out/sc_ours_n8_graph.json 2> gpurun_out/sc_ours_n8_graph.err
timeout 300 python bench.py --gpus 8 --steps 400 --warmup 40 --algo sgd_sync > gpurun_out/sc_sync_n8.json 2> gpurun_out/sc_sync_n8.err
timeout 300 python bench.py --gpus 8 --steps 400 --warmup 40 --algo asaga > gpurun_out/sc_asaga_n8.json 2> gpurun_out/sc_asaga_n8.err
for n in 1 8; do
  timeout 300 python bench.py --gpus $n --steps 50 --warmup 5 --impl nccl > gpurun_out/sc_nccl_n${n}.json 2> gpurun_out/sc_nccl_n${n}.err
done
for f in gpurun_out/sc_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["impl"], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "GB/s/gpu=%.0f"%d["hbm_gbps_per_gpu"], "e2e=", (round(d["e2e"]["value"],1) if "e2e" in d else None), "wpg=",d["config"].get("workers_per_gpu"), d["clocks"]["reasons"] if d.get("clocks") else None)
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/sc_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -12 "$f"; fi; done
