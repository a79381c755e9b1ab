// Objective evaluation on the tensor cores for the NON-squared losses (logistic / hinge / half-squared): the same
// tcgen05 / TMEM / TMA GEMM as csrc/kernels/eval_tc.cu (shared template, afb_eval_tc_impl.cuh) with the loss applied in the
// TMEM -> register epilogue through afb_common.cuh::loss_coeff.  Lives in libafb_experimental.so until it has been run on
// hardware (AFB_EVAL_TC_LOSS=1 selects it in ops/eval.py; tests/test_gpu_experimental.py compares it with the exact
// CUDA-core kernel).  The margin is computed in tf32 x (hi + lo) like the squared-error kernel: exact on integer-pixel /
// bf16-representable features, ~1e-3 relative otherwise.
#include "afb_eval_tc_impl.cuh"

extern "C" {

int afb_eval_loss_tc(const float* X, const float* y, long long n, long long ld, int d, const float* W, int T, int loss_id,
                     double* out, float* scratch, void* stream) {
  (void)d;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (loss_id) {
    case afb::LOSS_LS: return afb::eval_tc_all<afb::LOSS_LS>(X, y, n, ld, W, T, out, scratch, s);
    case afb::LOSS_LOGISTIC: return afb::eval_tc_all<afb::LOSS_LOGISTIC>(X, y, n, ld, W, T, out, scratch, s);
    case afb::LOSS_HINGE: return afb::eval_tc_all<afb::LOSS_HINGE>(X, y, n, ld, W, T, out, scratch, s);
    case afb::LOSS_LS_HALF: return afb::eval_tc_all<afb::LOSS_LS_HALF>(X, y, n, ld, W, T, out, scratch, s);
    default: return static_cast<int>(cudaErrorInvalidValue);
  }
}

}  // extern "C"
