#!/bin/bash
# BASELINE experiment matrix on N GPUs (results appended to gpurun_out/exp_N*.jsonl):
#   straggler sweep (ASAGA vs sync SAGA, delay intensity 0/0.3/0.6/1.0), production-cluster mode (32 workers, beta 0.7,
#   delay -1), epsilon ASAGA, rcv1 ASGD (CSR), time-to-loss ASGD vs the NCCL arm.
N=${1:-2}
IT=${2:-300}
mkdir -p gpurun_out
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 "$@"; }
tr benchmarks/run_config.py --config straggler --iters $IT --out gpurun_out/exp_N${N}_straggler.jsonl 2>&1 | grep -E "^\[straggler\]|Error|error" | cut -c1-300
tr benchmarks/run_config.py --config production --iters $IT --out gpurun_out/exp_N${N}_production.jsonl 2>&1 | grep -E "^\[production\]|Error|error" | cut -c1-300
tr benchmarks/run_config.py --config epsilon_asaga --iters 600 --eval-every 200 --target-frac 0 --out gpurun_out/exp_N${N}_epsilon.jsonl 2>&1 | grep -E "CONFIG_RESULT|Error|error" | cut -c1-700
tr benchmarks/run_config.py --config rcv1_asgd --iters 600 --eval-every 200 --target-frac 0 --out gpurun_out/exp_N${N}_rcv1.jsonl 2>&1 | grep -E "CONFIG_RESULT|Error|error" | cut -c1-700
tr benchmarks/time_to_loss.py --target 0.60 --eval-every 50 --max-iters 1500 --out gpurun_out/exp_N${N}_ttl.jsonl 2>&1 | grep -E "TIME_TO_LOSS|Error|error" | cut -c1-600
