#!/bin/bash
mkdir -p gpurun_out
N=$1
run() { name=$1; shift; timeout 300 python bench.py --steps 400 --warmup 40 "$@" > gpurun_out/r7_${name}.json 2> gpurun_out/r7_${name}.err; }
run n1 --gpus 1
run nN_w1 --gpus $N --workers-per-gpu 1
run nN_w2 --gpus $N --workers-per-gpu 2
for f in gpurun_out/r7_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "n=",d["n_gpus"], "it/s=%.1f"%d["value"], "ms=%.4f"%d["ms_per_step"], "e2e=%.1f"%d["e2e"]["value"], "wpg=",d["config"].get("workers_per_gpu"))
    print("   phases(us):", {k:round(v,2) for k,v in (d.get("last_round_phases_us_rank0_worker0") or {}).items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/r7_*.err; do if grep -q -i "error\|Traceback" "$f"; then echo "== $f"; tail -8 "$f"; fi; done
bash scripts/sanitize.sh
