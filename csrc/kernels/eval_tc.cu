// Objective evaluation GEMM on the 5th-gen tensor cores (K12, the one genuinely GEMM-shaped op):
//     out[t] = Σ_i loss(x_i · w_t, y_i)        for T parameter snapshots in ONE pass over the shard
// (SparkASGDThread.scala:386-401 evaluates every snapshot in one pass over the data; here that pass is a GEMM).
// D[128 rows, N] = X_tile[128, K] · B[N, K]ᵀ with kind::tf32 UMMA, accumulators in TMEM.
//
//  * B holds each snapshot twice: w_hi = tf32(w) and w_lo = w − w_hi (N = 2·T) so the parameter side
//    keeps ~fp32 accuracy; X is consumed as tf32 (exact for integer-pixel / bf16-representable data).
//  * Warp-specialised persistent CTA (192 threads): warp 0 = TMA producer (X and B tiles of 32 fp32
//    columns, 128B swizzle, 4-stage mbarrier ring), warp 1 = TMEM allocator + single-thread
//    tcgen05.mma issuer (4 x K=8 MMAs per stage, commit -> frees the stage), warps 2-5 = epilogue
//    (tcgen05.ld 32x32b.x16 of their 32-lane quarter; margin = hi + lo, then the loss — squared error, logistic
//    (log1pExp, Gradient.scala:188-196), hinge, half-squared — applied per column in registers through
//    afb_common.cuh::loss_coeff and accumulated).
//  * Two TMEM accumulator buffers: the epilogue of tile i overlaps the MMAs of tile i+1.
//  * K tail (ld not a multiple of 32) and row tail are zero-filled by TMA out-of-bounds handling.
// The kernel body is the shared template csrc/include/afb_eval_tc_impl.cuh.
#include "afb_eval_tc_impl.cuh"

extern "C" {

// W: [T][ld] snapshots (row-major, zero padded to ld); scratch: >= 64*ld floats; out: double[T] (+=).
int afb_eval_mse_tc(const float* X, const float* y, long long n, long long ld, int d, const float* W, int T,
                    double* out, float* scratch, void* stream) {
  (void)d;
  return afb::eval_tc_all<afb::LOSS_LS>(X, y, n, ld, W, T, out, scratch, static_cast<cudaStream_t>(stream));
}

// Same GEMM, loss selected at run time (the BASELINE objective is logistic).
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
