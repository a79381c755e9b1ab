#!/bin/bash
# Round 2, last 1-GPU pass: GPU suite on the final build, end-to-end step A/B (copies on side streams vs on the launching
# stream), launch-setup caching effect on the 20-step protocol, one ncu source-page capture of the final bf16 stream kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f2_gpu_suite.log 2>&1
echo "[gpu suite] rc=$? $(tail -1 gpurun_out/f2_gpu_suite.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { name=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/f2_${name}.json 2> gpurun_out/f2_${name}.err; \
        python - "gpurun_out/f2_${name}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    vb = d.get("vs_baseline") or {}
    e = d.get("e2e") or {}
    print(f"{sys.argv[2]:24s} it/s={d['value']:.1f} ms={d['ms_per_step']:.4f} e2e={e.get('value')} e2e_ms={e.get('ms_per_step')} variant={d['config']['dense_variant']} launches={d['gpu_launches']} ratio={vb.get('value_ratio')} e2e_ratio={vb.get('e2e_ratio')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json', '.err')).read()[-600:])
PY
}
run bf16_20 --gpus 1 --steps 20 --warmup 5
run bf16_20_b --gpus 1 --steps 20 --warmup 5 --no-baseline
AFB_E2E_OVERLAP=0 run bf16_20_samestream --gpus 1 --steps 20 --warmup 5 --no-baseline
run n8shard_20 --gpus 1 --steps 20 --warmup 5 --no-baseline --rows 1012500
AFB_E2E_OVERLAP=0 run n8shard_20_samestream --gpus 1 --steps 20 --warmup 5 --no-baseline --rows 1012500
run n8shard_400 --gpus 1 --steps 400 --warmup 40 --no-baseline --rows 1012500
AFB_E2E_OVERLAP=0 run n8shard_400_samestream --gpus 1 --steps 400 --warmup 40 --no-baseline --rows 1012500
DTYPES=bf16 bash scripts/gpu_runs/ncu_stream.sh
