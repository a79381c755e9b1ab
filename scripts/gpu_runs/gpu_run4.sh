#!/bin/bash
# pipe-variant validation: tests (under timeout), then N=1/2 A/B of simple vs pipe at 1 and 2 workers/GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -15 > gpurun_out/pytest4.log
cat gpurun_out/pytest4.log
AFB_DENSE_VARIANT=pipe timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mg_worker.py > gpurun_out/mg4.log 2>&1
grep MG_RESULT gpurun_out/mg4.log | cut -c1-600
for n in $1; do
  for v in simple pipe; do
    for w in 1 2; do
      timeout 300 python bench.py --gpus $n --steps 300 --warmup 30 --workers-per-gpu $w --variant $v --no-e2e > gpurun_out/v_${v}_n${n}_w${w}.json 2> gpurun_out/v_${v}_n${n}_w${w}.err
    done
  done
done
for f in gpurun_out/v_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "GB/s/gpu=%.0f"%d["hbm_gbps_per_gpu"], d["config"].get("workers_per_gpu"), d["config"].get("dense_variant"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/v_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -8 "$f"; fi; done
