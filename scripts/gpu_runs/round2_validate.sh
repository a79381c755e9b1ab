#!/bin/bash
# First GPU session of round 2: validate the experimental kernels (csrc/experimental/*) one by one under hard timeouts,
# then A/B them against the default path.  Usage (from the repo root, on the box):
#     bash scripts/gpu_runs/round2_validate.sh [N_GPUS]
# Every step has its own timeout so a hanging persistent kernel cannot eat the session; results land in gpurun_out/.
N=${1:-1}
mkdir -p gpurun_out
export AFB_TEST_EXPERIMENTAL=1
# 1. correctness, one test function at a time (a spin-wait bug shows up as a timeout of THAT test only)
for t in test_persistent_rounds_match_single_round_launches test_persistent_launch_refuses_non_resident_grid \
         test_pipelined_rounds_match_the_simulator_semantics test_persistent_csr_round_matches_the_split_round \
         test_bf16_row_storage_kernel_matches_fp32_kernel_on_bf16_exact_data \
         test_tensor_core_evaluation_with_loss_epilogue_matches_the_exact_kernel; do
  timeout 240 python -m pytest tests/test_gpu_experimental.py -x -q -k "$t" > gpurun_out/exp_$t.log 2>&1
  echo "[$t] rc=$? $(tail -1 gpurun_out/exp_$t.log)"
done
unset AFB_TEST_EXPERIMENTAL
# 2. regression: the default suite
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_default.log; cat gpurun_out/pytest_default.log
# 3. A/B on one GPU
run() { name=$1; shift; timeout 300 python bench.py --steps 400 --warmup 40 "$@" > gpurun_out/r2_${name}.json 2> gpurun_out/r2_${name}.err; \
        python - "gpurun_out/r2_${name}.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:28s} it/s={d['value']:.1f} ms={d['ms_per_step']:.4f} GB/s={d['hbm_gbps_per_gpu']:.0f} e2e={(d.get('e2e') or {}).get('value')} variant={d['config']['dense_variant']} dtype={d['dtype']}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run n1_default --gpus 1
run n1_persist --gpus 1 --variant persist
run n1_pipelined --gpus 1 --variant pipelined
run n1_bf16 --gpus 1 --dtype bf16
if [ "$N" -gt 1 ]; then
  mp() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
           bench.py --gpus $N --steps 400 --warmup 40 "$@" > gpurun_out/r2_${name}.json 2> gpurun_out/r2_${name}.err; \
         tail -1 gpurun_out/r2_${name}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), 'it/s', d['config']['dense_variant'], d['dtype'])" || echo "$name FAILED"; }
  mp nN_default_w2
  mp nN_default_w1 --workers-per-gpu 1
  mp nN_persist_w2 --variant persist
  mp nN_persist_w1 --variant persist --workers-per-gpu 1
  mp nN_pipelined --variant pipelined
  mp nN_bf16_w2 --dtype bf16
  AFB_DENSE_VARIANT=persist timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
      benchmarks/run_config.py --config rcv1_asgd --iters 600 --eval-every 100 --target-frac 0 --out gpurun_out/r2_rcv1_persist.jsonl 2>&1 | grep CONFIG_RESULT | cut -c1-300
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
      benchmarks/time_to_loss.py --target 0.60 --out gpurun_out/r2_time_to_loss.jsonl 2>&1 | grep TIME_TO_LOSS | cut -c1-400
fi
