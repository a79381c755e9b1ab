#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (single GPU): memcheck, racecheck, synccheck.
# Output summaries go to gpurun_out/sanitizer_*.txt (copied into profiles/ afterwards).
mkdir -p gpurun_out
SEL='test_dense_grad_full and 784 and logistic or test_csr_grad and logistic or test_asgd_round_single_gpu_semantics or test_asaga_round_single_gpu_semantics or test_level1'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py -x -q -k "$SEL" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit=$?" >> gpurun_out/sanitizer_$tool.txt
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit=" gpurun_out/sanitizer_$tool.txt | tail -4
done
AFB_TEST_TC=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_eval.py -x -q -k "tcgen05 and 16-784" > gpurun_out/sanitizer_memcheck_tc.txt 2>&1
echo "exit=$?" >> gpurun_out/sanitizer_memcheck_tc.txt
grep -E "ERROR SUMMARY|passed|failed|exit=" gpurun_out/sanitizer_memcheck_tc.txt | tail -3
