#!/bin/bash
# 2-GPU: full test-suite + PDL on/off A/B at N=1,2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest10.log
cat gpurun_out/pytest10.log
run() { name=$1; shift; timeout 300 python bench.py --steps 400 --warmup 40 "$@" > gpurun_out/r10_${name}.json 2> gpurun_out/r10_${name}.err; }
AFB_PDL=1 run n1_pdl --gpus 1
AFB_PDL=0 run n1_nopdl --gpus 1
AFB_PDL=1 run n2_w1_pdl --gpus 2 --workers-per-gpu 1
AFB_PDL=0 run n2_w1_nopdl --gpus 2 --workers-per-gpu 1
AFB_PDL=1 run n2_w2_pdl --gpus 2 --workers-per-gpu 2
AFB_PDL=0 run n2_w2_nopdl --gpus 2 --workers-per-gpu 2
AFB_PDL=1 run n2_sync_pdl --gpus 2 --algo sgd_sync
for f in gpurun_out/r10_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "e2e=%.1f"%d["e2e"]["value"], "wpg=",d["config"].get("workers_per_gpu"), "loss/row=%.4f"%d["e2e"]["last_loss_per_row"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/r10_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -8 "$f"; fi; done
