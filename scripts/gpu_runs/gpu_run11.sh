#!/bin/bash
# N-GPU validation + A/B of the prologue/ticket changes
mkdir -p gpurun_out
N=$1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest11.log
cat gpurun_out/pytest11.log
run() { name=$1; shift; timeout 300 python bench.py --steps 400 --warmup 40 "$@" > gpurun_out/r11_${name}.json 2> gpurun_out/r11_${name}.err; }
run n1 --gpus 1
run nN_w1 --gpus $N --workers-per-gpu 1
run nN_w2 --gpus $N --workers-per-gpu 2
run nN_sync --gpus $N --algo sgd_sync
if [ "$N" = "8" ]; then
  run n2 --gpus 2
  run n4 --gpus 4
  AFB_PDL=0 run nN_w1_nopdl --gpus $N --workers-per-gpu 1
  AFB_SEED_MODE=k run nN_w1_seedk --gpus $N --workers-per-gpu 1
fi
for f in gpurun_out/r11_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "e2e=%.1f"%d["e2e"]["value"], "wpg=",d["config"].get("workers_per_gpu"))
    ph=d.get("last_round_phases_us_rank0_worker0")
    if ph: print("    ", {k:(round(v,1) if not isinstance(v,list) else v) for k,v in ph.items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/r11_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -8 "$f"; fi; done
if [ "$N" = "8" ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
  timeout 400 $TR benchmarks/run_config.py --config straggler --iters 300 --delays 0.0 0.6 1.0 --out gpurun_out/configs2.jsonl 2>&1 | grep -E "straggler\]|Error|error" | cut -c1-300
fi
