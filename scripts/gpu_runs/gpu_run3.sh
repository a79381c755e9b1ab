#!/bin/bash
# usage: gpu_run3.sh "<list of N>" ; per N: ours (1 and 2 workers/GPU), ours sync, NCCL baseline
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest3.log
  cat gpurun_out/pytest3.log
fi
for n in $1; do
  for w in 1 2; do
    timeout 300 python bench.py --gpus $n --steps 300 --warmup 30 --workers-per-gpu $w > gpurun_out/s_ours_n${n}_w${w}.json 2> gpurun_out/s_ours_n${n}_w${w}.err
  done
  timeout 300 python bench.py --gpus $n --steps 300 --warmup 30 --algo sgd_sync > gpurun_out/s_sync_n${n}.json 2> gpurun_out/s_sync_n${n}.err
  timeout 300 python bench.py --gpus $n --steps 50 --warmup 5 --impl nccl > gpurun_out/s_nccl_n${n}.json 2> gpurun_out/s_nccl_n${n}.err
done
for f in gpurun_out/s_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["impl"], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "GB/s/gpu=%.0f"%d["hbm_gbps_per_gpu"], "e2e=", (round(d["e2e"]["value"],1) if "e2e" in d else None), d["config"].get("workers_per_gpu"), d["clocks"]["reasons"] if d.get("clocks") else None)
except Exception as e:
    print("ERR", e)
PY
done
for f in gpurun_out/s_*.err; do if [ -s "$f" ]; then echo "== $f"; tail -5 "$f"; fi; done
