#!/bin/bash
# ring-variant validation + N=8 A/B (stagger, variants)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -12 > gpurun_out/pytest6.log
cat gpurun_out/pytest6.log
N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mg_worker.py > gpurun_out/mg6.log 2>&1
grep -o "MG_RESULT.*" gpurun_out/mg6.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[len('MG_RESULT '):])
    print({k:v for k,v in d.items() if k.endswith('_k') or k.endswith('steps_per_rank') or 'versioned' in k or 'relerr' in k})
"
tail -3 gpurun_out/mg6.log | cut -c1-300
run() { name=$1; shift; timeout 300 python bench.py --steps 400 --warmup 40 --no-e2e "$@" > gpurun_out/r6_${name}.json 2> gpurun_out/r6_${name}.err; }
run n1_simple --gpus 1 --variant simple
run n1_ring --gpus 1 --variant ring
run n1_ring_g148 --gpus 1 --variant ring --workers-per-gpu 2
run nN_simple_w2 --gpus $N --variant simple --workers-per-gpu 2
run nN_simple_w2_st --gpus $N --variant simple --workers-per-gpu 2 --stagger-us 40
run nN_ring_w1 --gpus $N --variant ring --workers-per-gpu 1
run nN_ring_w2 --gpus $N --variant ring --workers-per-gpu 2
run nN_ring_w2_st --gpus $N --variant ring --workers-per-gpu 2 --stagger-us 40
run nN_ring_w3_st --gpus $N --variant ring --workers-per-gpu 3 --stagger-us 40
for f in gpurun_out/r6_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "GB/s/gpu=%.0f"%d["hbm_gbps_per_gpu"], "wpg=",d["config"].get("workers_per_gpu"), d["config"].get("dense_variant"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/r6_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -8 "$f"; fi; done
