#!/bin/bash
# 2-GPU pass: single-GPU tests again, multi-GPU fused checks, N=1/2 bench for ours and the NCCL baseline.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest2.log
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mg_worker.py > gpurun_out/mg2.log 2>&1
for n in 1 2; do
  timeout 300 python bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/bench_ours_$n.json 2> gpurun_out/bench_ours_$n.err
  timeout 300 python bench.py --gpus $n --steps 50 --warmup 5 --impl nccl > gpurun_out/bench_nccl_$n.json 2> gpurun_out/bench_nccl_$n.err
done
timeout 300 python bench.py --gpus 2 --steps 200 --warmup 20 --algo sgd_sync > gpurun_out/bench_ours_sync_2.json 2> gpurun_out/bench_ours_sync_2.err
cat gpurun_out/pytest2.log; tail -5 gpurun_out/mg2.log; cat gpurun_out/bench_*.json; tail -3 gpurun_out/*.err
