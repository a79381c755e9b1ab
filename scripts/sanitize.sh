#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (single GPU): memcheck, racecheck, synccheck.
# FULL logs go to gpurun_out/sanitizer_*.txt (copied into profiles/sanitizer/ afterwards).
mkdir -p gpurun_out
# default round kernels (TMA-ring stream kernel: serial + pipelined, fp32 + bf16), one-kernel CSR round, fused features
SEL='test_persistent_stream_rounds_match_single_round_launches and (784 or 200) or test_pipelined_rounds_match_the_simulator_semantics and 784 or test_bf16_row_storage_matches_fp32_rows_on_bf16_exact_data and 784 or test_persistent_csr_round_matches_the_split_round and 6000 or test_regularisation_is_fused_into_the_apply or test_enable_mask_skips_rounds_and_ssp_gate_opens'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 50 \
      python -m pytest tests/test_gpu_stream.py -x -q -p no:cacheprovider -k "$SEL" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit=$?" >> gpurun_out/sanitizer_$tool.txt
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit=" gpurun_out/sanitizer_$tool.txt | tail -4
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_stream.py -x -q -p no:cacheprovider -k "test_tensor_core_evaluation_with_loss_epilogue_matches_fp64 and 20" > gpurun_out/sanitizer_memcheck_tc.txt 2>&1
echo "exit=$?" >> gpurun_out/sanitizer_memcheck_tc.txt
grep -E "ERROR SUMMARY|passed|failed|exit=" gpurun_out/sanitizer_memcheck_tc.txt | tail -3
