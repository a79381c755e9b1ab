// EXPERIMENTAL — compiled into its own library (libafb_experimental.so), never loaded by the default path.
// Status: builds for sm_100a and is modelled by the round simulator; NOT yet validated on hardware.
//
// Persistent multi-round variant of the fused dense round (csrc/kernels/dense_round.cu): ONE launch runs
// `rounds` consecutive rounds of one logical worker.  Motivation (profiles/README.md, run11 phase table):
// at 8 GPUs a round is ~50 us of streaming plus 10-13 us of inter-kernel gap (launch latency + grid
// drain/ramp).  Keeping the grid resident removes that gap:
//   * every CTA loops over rounds; between rounds it hashes its first row chunk and puts its first row
//     fetch in flight (immutable inputs only), then spins on `WorkerCtl.round` — the word the last CTA of
//     the previous round publishes after its NVLink tail — instead of waiting for a new grid to launch;
//   * all per-round protocol state (tickets, chunk counter, g_local) is reset by that last CTA BEFORE it
//     publishes the round word, exactly as the one-round kernel does before it exits, so the protocol code
//     (afb_round_impl.cuh) is shared unchanged.
// Requirement: the whole grid must be co-resident (a CTA that is not scheduled can never finish round r,
// and the resident CTAs would spin forever) — the launcher checks occupancy x SM count / `share` >= grid.
#include "afb_dense_rows.cuh"

namespace afb {

// One round of the persistent loop.  Kept out of line so the register allocation of the hot row loop is the
// one-round kernel's (128 registers at NV4 <= 8, no spills): only `round` / `wait` cross the call.
// Returns false when the worker has been stopped (every CTA of the grid sees the same answer).
template <int NV4, int kThreads>
__device__ __noinline__ bool persist_one_round(const RoundParams& p, RoundShared<kThreads>& sh, float* s_dyn,
                                               const unsigned int round, const bool wait) {
  float* s_g = s_dyn;
  float* s_w = s_dyn + p.ld;
  const bool SAGA = (p.hist != nullptr);
  const int LOSS = p.loss_id;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  const bool early = (p.seed_mode == 1 || p.seed_mode == 2);
  {
    float4 acc[NV4];
#pragma unroll
    for (int c = 0; c < NV4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float loss_acc = 0.f;
    unsigned int cnt = 0;
    RowScheduler sch;
    sched_init(sch, p.n_rows, gridDim.x * (kThreads / 32), blockIdx.x * (kThreads / 32) + (tid >> 5));
    RowRegs<NV4> cur, nxt;
    int rc = -1;
    unsigned long long key = 0;

    // prologue on immutable inputs only (X, y, gidx): overlaps the previous round's NVLink tail
    if (early) {
      unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
      if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
      key = mix64(seed_eff);
      const unsigned int c0 = sch.next_chunk;
      sch.next_chunk = 0xffffffffu;
      if (c0 < static_cast<unsigned int>(sch.num_chunks)) {
        load_chunk_masks(sch, p, c0, lane, key);
        if (sch.m0) { int b = __ffs(sch.m0) - 1; sch.m0 &= sch.m0 - 1; rc = sch.base + b; }
        else if (sch.m1) { int b = __ffs(sch.m1) - 1; sch.m1 &= sch.m1 - 1; rc = sch.base + 32 + b; }
        if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, false);     // hist is mutable: read after the round barrier
      }
    }
    // round barrier: the previous round (tail included) is published
    if (wait) {
      if (tid == 0) {
        while (ld_acquire_gpu_u32(&wc->round) != round) __nanosleep(20);
      }
      __syncthreads();
    }

    RoundPull pl;
    round_begin_issue<kThreads>(p, sh, round, pl);
    if (early) {
      if (SAGA && rc >= 0) cur.h = p.hist[rc];
      sch.next_chunk = fetch_chunk(&wc->chunk_ctr, lane, sch.static_chunks);
      if (rc < 0) {
        rc = next_row(sch, p, &wc->chunk_ctr, lane, key);
        if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, SAGA);
      }
    }
    if (!round_begin_complete<kThreads>(p, sh, round, pl)) return false;     // stopped: every CTA sees proceed == 0

    if (!early) key = *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
    for (int i = tid; i < p.ld; i += kThreads) { s_g[i] = 0.f; s_w[i] = p.w_local[i]; }
    __syncthreads();
    if (!early) {
      rc = next_row(sch, p, &wc->chunk_ctr, lane, key);
      if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, SAGA);
    }
    while (rc >= 0) {
      int rn = next_row(sch, p, &wc->chunk_ctr, lane, key);
      if (rn >= 0) load_row<NV4>(nxt, p, rn, lane, nvec, SAGA);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int c = 0; c < NV4; ++c) {
        int v = lane + 32 * c;
        if (v < nvec) {
          const float4 wv = *reinterpret_cast<const float4*>(s_w + 4 * v);
          d0 = fmaf(cur.v[c].x, wv.x, d0);
          d1 = fmaf(cur.v[c].y, wv.y, d1);
          d2 = fmaf(cur.v[c].z, wv.z, d2);
          d3 = fmaf(cur.v[c].w, wv.w, d3);
        }
      }
      float margin = warp_sum((d0 + d1) + (d2 + d3));
      float coeff, loss;
      loss_coeff_rt(LOSS, margin, cur.y, coeff, loss);
      float eff = coeff;
      if (SAGA) {
        eff = coeff - cur.h;
        if (lane == 0) p.hist[rc] = coeff;          // write-through only: rollback rounds are not persistent
      }
#pragma unroll
      for (int c = 0; c < NV4; ++c) {
        acc[c].x = fmaf(eff, cur.v[c].x, acc[c].x);
        acc[c].y = fmaf(eff, cur.v[c].y, acc[c].y);
        acc[c].z = fmaf(eff, cur.v[c].z, acc[c].z);
        acc[c].w = fmaf(eff, cur.v[c].w, acc[c].w);
      }
      loss_acc += loss;
      ++cnt;
      cur = nxt;
      rc = rn;
    }

#pragma unroll
    for (int c = 0; c < NV4; ++c) {
      int v = lane + 32 * c;
      if (v < nvec) {
        atomicAdd(&s_g[4 * v + 0], acc[c].x);
        atomicAdd(&s_g[4 * v + 1], acc[c].y);
        atomicAdd(&s_g[4 * v + 2], acc[c].z);
        atomicAdd(&s_g[4 * v + 3], acc[c].w);
      }
    }
    __syncthreads();
    for (int v = tid; v < nvec; v += kThreads) {
      float4 gv = *reinterpret_cast<float4*>(&s_g[4 * v]);
      red_add_gpu_f32x4(p.g_local + 4 * v, gv);
    }
    round_finish<kThreads>(p, sh, round, loss_acc, cnt);
    __syncthreads();                                 // sh / s_dyn are reused by the next round
  }
  return true;
}

template <int NV4, int kThreads>
__global__ void __launch_bounds__(kThreads, (kThreads == 256 && NV4 <= 8) ? 2 : 1)
dense_round_persist_kernel(const __grid_constant__ RoundParams p, const int rounds) {
  __shared__ RoundShared<kThreads> sh;
  extern __shared__ __align__(16) float s_dyn[];
  // No CTA of this grid can have finished a round before every CTA has started (the last-CTA election
  // needs all of them), so this read is the same value in every CTA.
  const unsigned int round0 = *reinterpret_cast<volatile unsigned int*>(&p.wctl->round);
  for (int it = 0; it < rounds; ++it)
    if (!persist_one_round<NV4, kThreads>(p, sh, s_dyn, round0 + static_cast<unsigned int>(it), it > 0)) return;
}

template <int NV4, int kThreads>
static cudaError_t launch_persist(const RoundParams& p, int rounds, int grid, int share, cudaStream_t stream) {
  size_t smem = 2 * static_cast<size_t>(p.ld) * sizeof(float);
  auto k = dense_round_persist_kernel<NV4, kThreads>;
  cudaError_t e;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 0, per_sm = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, smem)) != cudaSuccess) return e;
  if (share < 1) share = 1;
  if (static_cast<long long>(grid) * share > static_cast<long long>(per_sm) * sms) return cudaErrorLaunchOutOfResources;
  k<<<grid, kThreads, smem, stream>>>(p, rounds);
  return cudaGetLastError();
}

}  // namespace afb

extern "C" {

int afb_dense_round_persist_max_ld() { return 2048; }

// `share` = number of persistent grids that co-run on this GPU (logical workers per GPU): the launch is refused
// unless grid x share fits in the resident-CTA capacity, because a non-resident CTA would deadlock the round barrier.
int afb_dense_round_persist_launch(const afb::RoundParams* params, int rounds, int grid, int share, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.ld % 4 != 0 || p.ld > 2048 || p.ld <= 0 || rounds <= 0) return static_cast<int>(cudaErrorInvalidValue);
  if (p.algo == ALGO_GRAD_ONLY || p.seed_mode == 3 || p.hist_prev != nullptr || p.pdl != 0) return static_cast<int>(cudaErrorInvalidValue);
  int nv4 = static_cast<int>((p.ld / 4 + 31) / 32);
  cudaError_t e;
  switch (nv4) {
    case 1: e = launch_persist<1, 256>(p, rounds, grid, share, s); break;
    case 2: e = launch_persist<2, 256>(p, rounds, grid, share, s); break;
    case 3: case 4: e = launch_persist<4, 256>(p, rounds, grid, share, s); break;
    case 5: case 6: case 7: e = launch_persist<7, 256>(p, rounds, grid, share, s); break;
    case 8: e = launch_persist<8, 256>(p, rounds, grid, share, s); break;
    case 9: case 10: case 11: case 12: e = launch_persist<12, 256>(p, rounds, grid, share, s); break;
    default: e = launch_persist<16, 256>(p, rounds, grid, share, s); break;
  }
  return static_cast<int>(e);
}

}  // extern "C"
