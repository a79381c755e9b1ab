#!/bin/bash
# quick 1-GPU loop: stream tests + a handful of bench points
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -x -q -p no:cacheprovider > gpurun_out/q1_stream_tests.log 2>&1
echo "[stream tests] rc=$? $(tail -1 gpurun_out/q1_stream_tests.log)"
run() { name=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/q1_${name}.json 2> gpurun_out/q1_${name}.err; \
        python - "gpurun_out/q1_${name}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    vb = d.get("vs_baseline") or {}
    print(f"{sys.argv[2]:24s} it/s={d['value']:.1f} ms={d['ms_per_step']:.4f} GB/s={d['hbm_gbps_per_gpu']:.0f} e2e={(d.get('e2e') or {}).get('value')} variant={d['config']['dense_variant']} dtype={d['dtype']} launches={d['gpu_launches']} ratio={vb.get('value_ratio')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json', '.err')).read()[-600:])
PY
}
run bf16_400 --gpus 1 --steps 400 --warmup 40 --no-baseline
run fp32_400 --gpus 1 --steps 400 --warmup 40 --dtype fp32 --no-baseline
run bf16_20 --gpus 1 --steps 20 --warmup 5 --no-baseline
run fp32_400_serial --gpus 1 --steps 400 --warmup 40 --dtype fp32 --no-baseline --no-pipeline
for extra in "$@"; do eval "$extra"; done
