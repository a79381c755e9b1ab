#!/bin/bash
# 8-GPU: scaling of the headline (N=1,2,4,8) + multi-GPU test + BASELINE configs 3-5 (short)
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -5 > gpurun_out/pytest9.log; cat gpurun_out/pytest9.log
for n in 1 2 4 $N; do
  timeout 300 python bench.py --gpus $n --steps 400 --warmup 40 > gpurun_out/s9_ours_n${n}.json 2> gpurun_out/s9_ours_n${n}.err
done
timeout 300 python bench.py --gpus $N --steps 400 --warmup 40 --workers-per-gpu 1 > gpurun_out/s9_ours_n${N}_w1.json 2> gpurun_out/s9_ours_n${N}_w1.err
timeout 300 python bench.py --gpus $N --steps 400 --warmup 40 --algo sgd_sync > gpurun_out/s9_sync_n${N}.json 2> gpurun_out/s9_sync_n${N}.err
timeout 300 python bench.py --gpus $N --steps 60 --warmup 10 --impl nccl > gpurun_out/s9_nccl_n${N}.json 2> gpurun_out/s9_nccl_n${N}.err
timeout 300 python bench.py --gpus 1 --steps 60 --warmup 10 --impl nccl > gpurun_out/s9_nccl_n1.json 2> gpurun_out/s9_nccl_n1.err
for f in gpurun_out/s9_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["impl"], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "GB/s/gpu=%.0f"%d["hbm_gbps_per_gpu"], "e2e=", (round(d["e2e"]["value"],1) if d.get("e2e") else None), "wpg=",d["config"].get("workers_per_gpu"))
    if d.get("last_round_phases_us_rank0_worker0"): print("    phases:", {k:round(v,1) for k,v in d["last_round_phases_us_rank0_worker0"].items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/s9_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -8 "$f"; fi; done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
timeout 300 $TR benchmarks/run_config.py --config epsilon_asaga --iters 600 --eval-every 100 --target-frac 0.0 --out gpurun_out/configs.jsonl 2>&1 | grep -E "CONFIG_RESULT|Error|error" | cut -c1-600
timeout 300 $TR benchmarks/run_config.py --config rcv1_asgd --iters 600 --eval-every 100 --target-frac 0.0 --out gpurun_out/configs.jsonl 2>&1 | grep -E "CONFIG_RESULT|Error|error" | cut -c1-600
timeout 600 $TR benchmarks/run_config.py --config straggler --iters 400 --eval-every 100 --target-frac 0.0 --out gpurun_out/configs.jsonl 2>&1 | grep -E "straggler\]|Error|error" | cut -c1-300
