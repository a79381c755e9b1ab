#!/bin/bash
# Round 2, final 1-GPU evidence pass: whole GPU suite, compute-sanitizer, ncu captures, bench A/B table.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/f1_gpu_suite.log 2>&1
echo "[gpu suite] rc=$? $(tail -1 gpurun_out/f1_gpu_suite.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/sanitize.sh
# ---- ncu: one capture per hot kernel (CSV pages produced here; .ncu-rep kept only when small)
cap() { name=$1; regex=$2; shift 2; timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s 1 -c 1 -f -o /tmp/ncu_$name "$@" > gpurun_out/ncu_${name}.log 2>&1; \
        echo "[ncu $name] rc=$? $(tail -1 gpurun_out/ncu_${name}.log | cut -c1-160)"; \
        ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null; \
        ncu -i /tmp/ncu_$name.ncu-rep --page details > gpurun_out/ncu_${name}_details.txt 2>/dev/null; }
cap stream_bf16 dense_stream python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-baseline
cap stream_fp32 dense_stream python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-baseline --dtype fp32
cap stream_bf16_n8shard dense_stream python bench.py --gpus 1 --steps 24 --warmup 3 --no-e2e --no-baseline --rows 1012500
cap csr_stream csr_stream python scripts/gpu_runs/ncu_targets.py csr
cap push_apply push_apply python scripts/gpu_runs/ncu_targets.py csr_split
cap eval_tc eval_tc_kernel python scripts/gpu_runs/ncu_targets.py eval
# ---- bench A/B (N = 1)
run() { name=$1; shift; timeout 300 python bench.py "$@" > gpurun_out/f1_${name}.json 2> gpurun_out/f1_${name}.err; \
        python - "gpurun_out/f1_${name}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    vb = d.get("vs_baseline") or {}
    print(f"{sys.argv[2]:24s} it/s={d['value']:.1f} ms={d['ms_per_step']:.4f} GB/s={d['hbm_gbps_per_gpu']:.0f} e2e={(d.get('e2e') or {}).get('value')} variant={d['config']['dense_variant']} dtype={d['dtype']} launches={d['gpu_launches']} ratio={vb.get('value_ratio')} e2e_ratio={vb.get('e2e_ratio')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json', '.err')).read()[-600:])
PY
}
run bf16_20 --gpus 1 --steps 20 --warmup 5
run bf16_400 --gpus 1 --steps 400 --warmup 40 --no-baseline
run fp32_400 --gpus 1 --steps 400 --warmup 40 --dtype fp32 --no-baseline
run bf16_400_serial --gpus 1 --steps 400 --warmup 40 --no-baseline --no-pipeline
run fp32_400_simple --gpus 1 --steps 400 --warmup 40 --dtype fp32 --no-baseline --variant simple
AFB_TMA_MASK=0xFF run fp32_400_tma --gpus 1 --steps 400 --warmup 40 --dtype fp32 --no-baseline
AFB_TMA_MASK=0x00 run bf16_400_lsu --gpus 1 --steps 400 --warmup 40 --no-baseline
AFB_TMA_MASK=0x55 run bf16_400_alt --gpus 1 --steps 400 --warmup 40 --no-baseline
run n8shard_bf16_400 --gpus 1 --steps 400 --warmup 40 --no-baseline --rows 1012500
