// Fused dense round kernel (sm_100a):  pull(w) -> sample -> x·w -> loss coeff -> g += c·x
//                                      -> (SAGA history) -> block/grid reduce -> push-and-apply.
//
// One launch = one round of one worker.  Hot-op inventory covered: K1 (dot),
// K2 (residual/loss), K3 (gradient accumulate), K4 (SAGA correction), K5/K6 (SGD/SAGA apply),
// K7 (history write), K9 (gradient push, in-kernel NVLink reductions into rank-0 HBM),
// K10 (parameter pull, P2P load by one elected CTA overlapped with the other CTAs' first row
// fetches), K11 (Bernoulli sampling, evaluated on the fly per row: no index list, no extra pass).
//
// Every sampled row is read from HBM exactly once (the reference materialises a dense d-vector
// per sample).  A warp owns a row: lane l holds columns {4(l+32c)..+3}; dot via
// shuffle tree, axpy into per-lane register accumulators; rows are double-buffered in registers
// so two row fetches (2 x ld*4 B) are in flight per warp.  Work distribution is dynamic (64-row
// chunks from a global counter, next chunk requested one ahead) so binomial sampling noise does
// not create a tail.  The last CTA to finish (ticket) performs the cross-GPU part: logical-clock
// atomics, staleness gate, Hogwild `red.add` of the scaled gradient straight into the server's w
// (and alpha-bar for SAGA), result-ring record, worker-status update.
#include "afb_dense_rows.cuh"

namespace afb {

template <int NV4, int kThreads>
__global__ void __launch_bounds__(kThreads, (kThreads == 256 && NV4 <= 8) ? 2 : 1) dense_round_kernel(const RoundParams p) {
  __shared__ RoundShared<kThreads> sh;
  extern __shared__ __align__(16) float s_dyn[];
  float* s_g = s_dyn;                            // [ld] CTA-level gradient accumulator
  float* s_w = s_dyn + p.ld;                     // [ld] the pulled parameter vector (read per row via LDS.128)
  const bool SAGA = (p.hist != nullptr);
  const int LOSS = p.loss_id;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  // Programmatic dependent launch: this grid may start while the PREVIOUS round's last CTA is still in
  // its NVLink tail.  Everything before `griddepcontrol.wait` touches only immutable inputs (X, y, gidx)
  // and uses the host-counted round index; all protocol state is read after the wait.
  const bool pdl = (p.pdl != 0);
  const unsigned int seed_round = pdl ? static_cast<unsigned int>(p.round_host) : wc->round;   // sampling-seed input only

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

  // When the sampling seed does not depend on the server's k (seed_mode 1: seed + round + 1, or a
  // fixed seed), every warp can hash its first chunk and put its first row fetch in flight BEFORE the
  // parameter pull completes: the NVLink round trip of phase 0 overlaps the first HBM round trip.
  const bool early = (p.algo != ALGO_GRAD_ONLY) && (p.seed_mode == 1 || p.seed_mode == 2);
  if (early) {
    unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
    if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(seed_round) + 1ULL;
    key = mix64(seed_eff);
    // statically assigned first chunk only: no atomics on the chunk counter before the dependency wait
    const unsigned int c0 = sch.next_chunk;
    sch.next_chunk = 0xffffffffu;                 // "not fetched yet"
    if (c0 < static_cast<unsigned int>(sch.num_chunks)) {
      load_chunk_masks(sch, p, c0, lane, key);
      if (sch.m0) { int b = __ffs(sch.m0) - 1; sch.m0 &= sch.m0 - 1; rc = sch.base + b; }
      else if (sch.m1) { int b = __ffs(sch.m1) - 1; sch.m1 &= sch.m1 - 1; rc = sch.base + 32 + b; }
      if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, SAGA && !pdl);   // hist is mutable: not before the wait
    }
  }
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned int round = *reinterpret_cast<volatile unsigned int*>(&wc->round);   // protocol round (device truth)

  // ------------------------------------------------------------------ phase 0a: elect the puller, put the pull in flight
  RoundPull pl;
  round_begin_issue<kThreads>(p, sh, round, pl);

  // ... while the NVLink round trip is in flight: first dynamic chunk id + first row fetch
  if (early) {
    if (pdl && SAGA && rc >= 0) cur.h = p.hist[rc];
    sch.next_chunk = fetch_chunk(&wc->chunk_ctr, lane, sch.static_chunks);
    if (rc < 0) {
      rc = next_row(sch, p, &wc->chunk_ctr, lane, key);
      if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, SAGA);
    }
  }

  // ------------------------------------------------------------------ phase 0b: w_local / ts / seed published
  if (round_begin_complete<kThreads>(p, sh, round, pl) != PROCEED_RUN) return;

  // ------------------------------------------------------------------ phase 1: fused gradient
  if (p.seed_mode == 3 && !*reinterpret_cast<volatile int*>(&wc->proceed)) return;   // split round that was stopped at the gate
  if (!early) {
    key = (p.algo == ALGO_GRAD_ONLY && p.seed_mode != 3)
              ? mix64(static_cast<unsigned long long>(p.seed))
              : *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
  }
  for (int i = tid; i < p.ld; i += kThreads) { s_g[i] = 0.f; s_w[i] = p.w_local[i]; }
  __syncthreads();
  if (!early) {
    rc = next_row(sch, p, &wc->chunk_ctr, lane, key);
    if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, SAGA);
  }
  while (rc >= 0) {
    int rn = next_row(sch, p, &wc->chunk_ctr, lane, key);
    if (rn >= 0) load_row<NV4>(nxt, p, rn, lane, nvec, SAGA);   // prefetch: second row in flight
    // K1: dot
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
    // K2 (+K4): coefficient
    float coeff, loss;
    loss_coeff_rt(LOSS, margin, cur.y, coeff, loss);
    float eff = coeff;
    if (SAGA) {
      eff = coeff - cur.h;
      if (lane == 0) {                       // K7: history write (row is owned by this worker)
        if (p.hist_prev) { p.hist_prev[rc] = cur.h; p.hist_tag[rc] = static_cast<int>(round); }
        p.hist[rc] = coeff;
      }
    }
    // K3: axpy into register accumulators
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

  // ------------------------------------------------------------------ phase 2: CTA reduce -> g_local
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
  // this CTA's part of the round is done: let the next round's grid start its prologue (first-row
  // prefetch) while the last CTA of this grid runs the NVLink tail
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  round_finish<kThreads>(p, sh, round, loss_acc, cnt);
}

// SAGA rollback (finite taw only): undo the history writes of a rejected round.
__global__ void saga_rollback_kernel(float* hist, const float* hist_prev, const int* hist_tag, int n_rows,
                                     const WorkerCtl* wc) {
  if (wc->last_accepted != 0) return;
  const int tag = static_cast<int>(wc->round) - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x)
    if (hist_tag[i] == tag) hist[i] = hist_prev[i];
}

template <int NV4, int kThreads>
static cudaError_t launch_nv(const RoundParams& p, int grid, cudaStream_t stream) {
  size_t smem = 2 * static_cast<size_t>(p.ld) * sizeof(float);
  auto k = dense_round_kernel<NV4, kThreads>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (p.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p);
  }
  k<<<grid, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace afb

extern "C" {

// Max row width handled with register accumulators: NV4 <= 16  =>  ld <= 2048 floats.
int afb_dense_round_max_ld() { return 2048; }
// resident CTAs per SM of the dense kernel for a given row width (grid sizing)
int afb_dense_round_ctas_per_sm(long long ld) { return ((ld / 4 + 31) / 32) <= 8 ? 2 : 1; }

int afb_dense_round_launch(const afb::RoundParams* params, int grid, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.ld % 4 != 0 || p.ld > 2048 || p.ld <= 0) return static_cast<int>(cudaErrorInvalidValue);
  int nv4 = static_cast<int>((p.ld / 4 + 31) / 32);
  cudaError_t e;
  // <= 8 float4 per lane: 256 threads x 2 CTAs/SM (128 regs) — two co-resident CTAs let two logical
  // workers (or two rounds) share an SM so one worker's NVLink pull/push latency hides behind the
  // other's HBM streaming; wider rows: 256 threads x 1 CTA/SM (255 regs) so the accumulators + two
  // in-flight rows stay in registers.
  switch (nv4) {
    case 1: e = launch_nv<1, 256>(p, grid, s); break;
    case 2: e = launch_nv<2, 256>(p, grid, s); break;
    case 3: case 4: e = launch_nv<4, 256>(p, grid, s); break;
    case 5: case 6: case 7: e = launch_nv<7, 256>(p, grid, s); break;
    case 8: e = launch_nv<8, 256>(p, grid, s); break;
    case 9: case 10: case 11: case 12: e = launch_nv<12, 256>(p, grid, s); break;
    default: e = launch_nv<16, 256>(p, grid, s); break;
  }
  return static_cast<int>(e);
}

int afb_saga_rollback_launch(float* hist, const float* hist_prev, const int* hist_tag, int n_rows,
                             const afb::WorkerCtl* wc, void* stream) {
  afb::saga_rollback_kernel<<<296, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, hist_prev, hist_tag, n_rows, wc);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
