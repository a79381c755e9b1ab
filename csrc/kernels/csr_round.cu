// Fused CSR round kernel (sm_100a): same phases as dense_round.cu with a sparse gradient body.
// rcv1_full shape: d = 47 236, ~74 nnz/row.  A warp owns a row: lanes stride over the row's
// non-zeros, gather w[idx] (w_local is 189 KB => L2/L1 resident), shuffle-reduce the dot, then
// scatter coeff*val into the worker-local g with fire-and-forget L2 reductions (K1-K4, K7, K11).
// The reference pays an O(d) dense allocation + dense add per *sample* here;
// this kernel touches only the nnz.
#include "afb_round_impl.cuh"

namespace afb {

constexpr int kCsrThreads = 512;
constexpr int kCsrChunkRows = 64;

__global__ void __launch_bounds__(kCsrThreads, 2) csr_round_kernel(const RoundParams p) {
  __shared__ RoundShared<kCsrThreads> sh;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  WorkerCtl* wc = p.wctl;
  const unsigned int round = wc->round;
  const bool saga = (p.hist != nullptr);

  if (p.algo != ALGO_GRAD_ONLY) {
    if (round_begin<kCsrThreads>(p, sh, round) != PROCEED_RUN) return;
  }
  if (p.seed_mode == 3 && !*reinterpret_cast<volatile int*>(&wc->proceed)) return;   // split round that was stopped at the gate
  const unsigned long long key = (p.algo == ALGO_GRAD_ONLY && p.seed_mode != 3)
                                     ? mix64(static_cast<unsigned long long>(p.seed))
                                     : *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
  const float* __restrict__ w = p.w_local;
  float loss_acc = 0.f;
  unsigned int cnt = 0;
  const int num_chunks = (p.n_rows + kCsrChunkRows - 1) / kCsrChunkRows;

  unsigned int chunk = 0;
  if (lane == 0) chunk = atomicAdd(&wc->chunk_ctr, 1u);
  chunk = __shfl_sync(0xffffffffu, chunk, 0);
  while (chunk < static_cast<unsigned int>(num_chunks)) {
    unsigned int nextc = 0;
    if (lane == 0) nextc = atomicAdd(&wc->chunk_ctr, 1u);        // requested one ahead
    const int base = static_cast<int>(chunk) * kCsrChunkRows;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = base + 32 * half + lane;
      bool sel = false;
      if (r < p.n_rows) {
        unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r]) : static_cast<unsigned long long>(p.row_base + r);
        sel = hash_u24(key, g) < static_cast<unsigned int>(p.thr24);
      }
      unsigned int m = __ballot_sync(0xffffffffu, sel);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int row = base + 32 * half + b;
        const long long s = p.indptr[row], e = p.indptr[row + 1];
        float dot = 0.f;
        for (long long k = s + lane; k < e; k += 32) dot = fmaf(p.values[k], w[p.indices[k]], dot);
        const float margin = warp_sum(dot);
        float coeff, loss;
        loss_coeff_rt(p.loss_id, margin, __ldg(p.y + row), coeff, loss);
        float eff = coeff;
        if (saga) {
          const float h = p.hist[row];
          eff = coeff - h;
          if (lane == 0) {
            if (p.hist_prev) { p.hist_prev[row] = h; p.hist_tag[row] = static_cast<int>(round); }
            p.hist[row] = coeff;
          }
        }
        for (long long k = s + lane; k < e; k += 32) red_add_gpu_f32(p.g_local + p.indices[k], eff * p.values[k]);
        loss_acc += loss;
        ++cnt;
      }
    }
    chunk = __shfl_sync(0xffffffffu, nextc, 0);
  }
  round_finish<kCsrThreads>(p, sh, round, loss_acc, cnt);
}

}  // namespace afb

extern "C" {

int afb_csr_round_launch(const afb::RoundParams* params, int grid, void* stream) {
  const afb::RoundParams& p = *params;
  if (p.ld % 4 != 0 || p.ld <= 0) return static_cast<int>(cudaErrorInvalidValue);
  afb::csr_round_kernel<<<grid, afb::kCsrThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
