#!/bin/bash
# multi-GPU validation: mg_worker tests + bench (driver protocol 20/5 and a long run) on N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -p no:cacheprovider > gpurun_out/mg${N}_tests.log 2>&1
echo "[mg tests N=$N] rc=$? $(tail -1 gpurun_out/mg${N}_tests.log)"
grep -o "MG_RESULT.*" gpurun_out/mg${N}_tests.log | head -1 | cut -c1-3000
mp() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
         bench.py --gpus $N "$@" > gpurun_out/mg${N}_${name}.json 2> gpurun_out/mg${N}_${name}.err; \
       python - "gpurun_out/mg${N}_${name}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    vb = d.get("vs_baseline") or {}
    print(f"{sys.argv[2]:20s} N={d['n_gpus']} it/s={d['value']:.1f} ms={d['ms_per_step']:.4f} GB/s={d['hbm_gbps_per_gpu']:.0f} e2e={(d.get('e2e') or {}).get('value')} variant={d['config']['dense_variant']} dtype={d['dtype']} ratio={vb.get('value_ratio')} e2e_ratio={vb.get('e2e_ratio')} base={vb.get('variants')} phases={d.get('last_round_phases_us_rank0_worker0')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
}
mp bf16_20 --steps 20 --warmup 5
mp bf16_400 --steps 400 --warmup 40 --no-baseline
mp fp32_20 --steps 20 --warmup 5 --dtype fp32 --no-baseline
for extra in "${@:2}"; do eval "$extra"; done
