// EXPERIMENTAL — compiled into libafb_experimental.so, never loaded by the default path.
// Status: builds for sm_100a; protocol logic checked by the CTA-level interleaving model
// (asyncframework_b200/ops/round_model.py, mode "pipelined"); NOT yet validated on hardware.
//
// Pipelined persistent ASGD rounds: ONE launch runs `rounds` rounds of ONE worker that owns every SM of its GPU, and
// round r+1's pull + gradient phase overlap round r's NVLink tail.  Motivation (benchmarks/round_timeline_model.py,
// fitted to the measured phase timestamps of run11): at 8 GPUs a round streams ~50 us and carries ~29 us of serial
// work (inter-kernel gap, pull, fill/drain, tail); hiding all of it behind the next round's streaming predicts
// ~18 k iterations/s instead of the measured 12.8 k.
//
// How (mirrors `WorkerModel.pipelined_cta` of the model line by line):
//   * every piece of per-round state is double-buffered by round parity: tickets, chunk counter, g_local, the pulled
//     w / clock / seed, the early clock ticket (`PipeBuf`);
//   * a CTA may start round r once round r-2 has been PUBLISHED (`WorkerCtl.round >= r-1`): its parity buffers are clean;
//   * the last CTA to finish round r runs the tail (staleness gate decision taken earlier by the first finisher, Hogwild
//     reductions into the server, result record, reset of buffer r&1, publication of `round = r+1`) while every other
//     CTA is already streaming round r+1; tails are published in round order.
// Semantics: the worker computes round r+1 on parameters that do not contain its own round-r update (one own update in
// flight; the logical clock does not count it).  ASGD only (SAGA's per-row history would be read and written by two
// overlapping rounds), no bucket gate, no delay injection.
#include "afb_dense_rows.cuh"

namespace afb {

struct alignas(16) PipeBuf {
  unsigned int chunk_ctr;
  unsigned int done_ctas;
  unsigned int pull_ticket;
  unsigned int pull_flag;            // == round+1 once w / ts / proceed of this buffer are valid
  int proceed;
  unsigned int batch_count;
  float loss_sum;
  unsigned int tk_flag;              // == round+1 once the early clock ticket is valid
  unsigned long long ts;
  unsigned long long k_read;
  unsigned long long t_start_ns;
  long long tk_staleness;
  long long tk_k_after;
  int tk_accept;
  float tk_step;
  unsigned int tk_clock_old;
  unsigned int pad;
};
struct PipeState { PipeBuf b[2]; };

template <int NV4, int kThreads>
__device__ __noinline__ bool pipelined_one_round(const RoundParams& p, RoundShared<kThreads>& sh, float* s_dyn,
                                                 const unsigned int round, const int it, const int rounds,
                                                 PipeState* ps, float* g2, float* w2) {
  constexpr int kWarps = kThreads / 32;
  float* s_g = s_dyn;
  float* s_w = s_dyn + p.ld;
  const int LOSS = p.loss_id;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  PipeBuf* S = &ps->b[round & 1u];
  float* g_loc = g2 + static_cast<long long>(round & 1u) * p.ld;
  float* w_loc = w2 + static_cast<long long>(round & 1u) * p.ld;

  float4 acc[NV4];
#pragma unroll
  for (int c = 0; c < NV4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  float loss_acc = 0.f;
  unsigned int cnt = 0;
  RowScheduler sch;
  sched_init(sch, p.n_rows, gridDim.x * kWarps, blockIdx.x * kWarps + warp);
  RowRegs<NV4> cur, nxt;
  int rc = -1;

  // ---- prologue on immutable inputs: sampling key of this round, first (static) chunk, first row fetch
  unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
  if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
  const unsigned long long key = mix64(seed_eff);
  {
    const unsigned int c0 = sch.next_chunk;
    sch.next_chunk = 0xffffffffu;
    if (c0 < static_cast<unsigned int>(sch.num_chunks)) {
      load_chunk_masks(sch, p, c0, lane, key);
      if (sch.m0) { int b = __ffs(sch.m0) - 1; sch.m0 &= sch.m0 - 1; rc = sch.base + b; }
      else if (sch.m1) { int b = __ffs(sch.m1) - 1; sch.m1 &= sch.m1 - 1; rc = sch.base + 32 + b; }
      if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, false);
    }
  }
  // ---- buffer reuse: round r-2 (same parity) must be published before anything of buffer r&1 is touched
  if (it >= 2) {
    if (tid == 0) {
      while (static_cast<int>(ld_acquire_gpu_u32(&wc->round) - (round - 1u)) < 0) __nanosleep(20);
    }
    __syncthreads();
  }
  // ---- puller election; the pull's loads are issued and stay in flight
  if (tid == 0) sh.flag = atomicAdd(&S->pull_ticket, 1u);
  __syncthreads();
  const bool puller = (sh.flag == 0);
  float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long ck = 0;
  if (puller) {
    if (tid == 0) wc->dbg_ns[0] = globaltimer_ns();
    if (tid < nvec) wv = ld_relaxed_sys_f32x4(p.w_server + 4 * tid);
    if (tid == 0) ck = ld_acquire_sys_u64(&p.sctl->clock_k);
  }
  // ---- first dynamic chunk id while the NVLink round trip is in flight
  sch.next_chunk = fetch_chunk(&S->chunk_ctr, lane, sch.static_chunks);
  if (rc < 0) {
    rc = next_row(sch, p, &S->chunk_ctr, lane, key);
    if (rc >= 0) load_row<NV4>(cur, p, rc, lane, nvec, false);
  }
  // ---- pull completion / wait for it
  if (puller) {
    if (tid == 0) {
      int proceed = 1;
      if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) proceed = 0;
      const unsigned long long now = globaltimer_ns();
      if (proceed && it == 0) {
        red_add_release_sys_s32(&p.sctl->avail_count, -1);      // busy for the whole persistent launch
        WorkerSlot* me = p.status + p.worker_id;
        me->available = 0;
        me->busy_since_ns = now;
      }
      if (proceed) (p.status + p.worker_id)->heartbeat_ns = now;
      S->ts = ck & 0xffffffffULL;
      S->k_read = ck >> 32;
      S->proceed = proceed;
      S->t_start_ns = now;
    }
    if (tid < nvec) *reinterpret_cast<float4*>(w_loc + 4 * tid) = wv;
    for (int v = tid + kThreads; v < nvec; v += kThreads)
      *reinterpret_cast<float4*>(w_loc + 4 * v) = ld_relaxed_sys_f32x4(p.w_server + 4 * v);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      wc->dbg_ns[2] = globaltimer_ns();
      st_release_gpu_u32(&S->pull_flag, round + 1u);
    }
  } else {
    if (tid == 0) {
      while (ld_acquire_gpu_u32(&S->pull_flag) != round + 1u) __nanosleep(32);
    }
    __syncthreads();
  }
  if (!*reinterpret_cast<volatile int*>(&S->proceed)) {
    // stopped (k >= max_iters or stop flag): the last CTA to arrive resets the buffer and publishes, in round order
    __syncthreads();
    if (tid == 0) {
      unsigned int t = atomicAdd(&S->done_ctas, 1u);
      if (t == gridDim.x - 1) {
        while (ld_acquire_gpu_u32(&wc->round) != round) __nanosleep(20);     // the previous round's tail first
        S->done_ctas = 0; S->pull_ticket = 0; S->chunk_ctr = 0;
        if (it > 0) {                                                        // this launch had made the worker busy
          (p.status + p.worker_id)->available = 1;
          fence_sys();
          red_add_release_sys_s32(&p.sctl->avail_count, 1);
        }
        wc->last_accepted = -1;
        __threadfence();
        st_release_gpu_u32(&wc->round, round + 1u);
      }
    }
    return false;
  }

  // ---- phase 1: fused gradient (identical to dense_round_kernel, ASGD: no history)
  for (int i = tid; i < p.ld; i += kThreads) { s_g[i] = 0.f; s_w[i] = w_loc[i]; }
  __syncthreads();
  while (rc >= 0) {
    int rn = next_row(sch, p, &S->chunk_ctr, lane, key);
    if (rn >= 0) load_row<NV4>(nxt, p, rn, lane, nvec, false);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int c = 0; c < NV4; ++c) {
      int v = lane + 32 * c;
      if (v < nvec) {
        const float4 w4 = *reinterpret_cast<const float4*>(s_w + 4 * v);
        d0 = fmaf(cur.v[c].x, w4.x, d0);
        d1 = fmaf(cur.v[c].y, w4.y, d1);
        d2 = fmaf(cur.v[c].z, w4.z, d2);
        d3 = fmaf(cur.v[c].w, w4.w, d3);
      }
    }
    float margin = warp_sum((d0 + d1) + (d2 + d3));
    float coeff, loss;
    loss_coeff_rt(LOSS, margin, cur.y, coeff, loss);
#pragma unroll
    for (int c = 0; c < NV4; ++c) {
      acc[c].x = fmaf(coeff, cur.v[c].x, acc[c].x);
      acc[c].y = fmaf(coeff, cur.v[c].y, acc[c].y);
      acc[c].z = fmaf(coeff, cur.v[c].z, acc[c].z);
      acc[c].w = fmaf(coeff, cur.v[c].w, acc[c].w);
    }
    loss_acc += loss;
    ++cnt;
    cur = nxt;
    rc = rn;
  }
  // ---- phase 2: CTA reduce -> g_local[parity]
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
    red_add_gpu_f32x4(g_loc + 4 * v, gv);
  }
  if (lane == 0) { sh.red[warp] = loss_acc; sh.cnt[warp] = cnt; }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f; unsigned int n = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { l += sh.red[w]; n += sh.cnt[w]; }
    atomicAdd(&S->loss_sum, l);
    atomicAdd(&S->batch_count, n);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.flag = atomicAdd(&S->done_ctas, 1u);
  __syncthreads();
  const unsigned int my_ticket = sh.flag;
  __syncthreads();                           // sh.flag is reused below

  // ---- early logical-clock ticket by the first CTA to finish (its NVLink round trip overlaps the drain)
  if (my_ticket == 0 && tid == 0) {
    const unsigned long long ts = S->ts;
    unsigned long long old, k_old;
    long long staleness;
    int accept;
    if (p.taw >= (1LL << 31)) {
      old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, (1ULL << 32) | 1ULL);
      k_old = old >> 32;
      staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
      accept = (static_cast<long long>(k_old) < p.max_iters) ? 1 : 0;
    } else {
      old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL);
      staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
      accept = (staleness <= p.taw) ? 1 : 0;
      k_old = old >> 32;
      if (accept) {
        unsigned long long o2 = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL << 32);
        k_old = o2 >> 32;
        if (static_cast<long long>(k_old) >= p.max_iters) accept = 0;
      }
    }
    S->tk_accept = accept;
    S->tk_step = p.gamma * rsqrtf(static_cast<float>(k_old / static_cast<unsigned long long>(p.num_workers)) + 1.f);
    S->tk_clock_old = static_cast<unsigned int>(old & 0xffffffffULL);
    S->tk_staleness = staleness;
    S->tk_k_after = static_cast<long long>(k_old) + (accept ? 1 : 0);
    __threadfence();
    st_release_gpu_u32(&S->tk_flag, round + 1u);
  }
  if (my_ticket != gridDim.x - 1) return true;      // everybody else: on to round r+1

  // ---- tail of round `round` (this CTA only), overlapping the other CTAs' round r+1
  if (tid == 0) {
    wc->dbg_ns[3] = globaltimer_ns();
    while (ld_acquire_gpu_u32(&wc->round) != round) __nanosleep(20);        // tails publish in round order
    while (ld_acquire_gpu_u32(&S->tk_flag) != round + 1u) __nanosleep(32);
    wc->dbg_ns[4] = globaltimer_ns();
    sh.flag = static_cast<unsigned int>(*reinterpret_cast<volatile int*>(&S->tk_accept));
    sh.bcast[0] = *reinterpret_cast<volatile float*>(&S->tk_step);
  }
  __threadfence();
  __syncthreads();
  const int accept = static_cast<int>(sh.flag);
  const float step = sh.bcast[0];
  if (accept) {
    const float a = -step * p.inv_par_recs;
    for (int v = tid; v < nvec; v += kThreads) {
      float4 g = __ldcg(reinterpret_cast<const float4*>(g_loc + 4 * v));
      red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x, a * g.y, a * g.z, a * g.w));
    }
  }
  for (int v = tid; v < nvec; v += kThreads)
    *reinterpret_cast<float4*>(g_loc + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long now = 0, task_ns = 0;
  unsigned int clock_old = 0;
  ResultRec* rec = nullptr;
  float round_loss = 0.f;
  unsigned int round_batch = 0;
  if (tid == 0) {
    round_loss = *reinterpret_cast<volatile float*>(&S->loss_sum);
    round_batch = *reinterpret_cast<volatile unsigned int*>(&S->batch_count);
    now = globaltimer_ns();
    wc->dbg_ns[5] = now;
    task_ns = now - S->t_start_ns;
    clock_old = *reinterpret_cast<volatile unsigned int*>(&S->tk_clock_old);
    rec = p.ring + (clock_old % static_cast<unsigned int>(p.ring_size));
    rec->worker = p.worker_id; rec->accepted = accept;
    rec->ts = static_cast<long long>(S->ts); rec->staleness = S->tk_staleness;
    rec->k_after = S->tk_k_after; rec->batch = round_batch;
    rec->loss_sum = round_loss; rec->step = accept ? step : 0.f;
    rec->t_start_ns = S->t_start_ns; rec->t_end_ns = now;
    WorkerSlot* me = p.status + p.worker_id;
    const unsigned long long ntasks = wc->rounds_done + 1ULL;
    const float task_ms = static_cast<float>(task_ns) * 1e-6f;
    wc->avg_task_ms += (task_ms - wc->avg_task_ms) / static_cast<float>(ntasks);
    wc->last_finish_ns = now;
    me->staleness = S->tk_staleness;
    me->num_tasks = static_cast<long long>(ntasks);
    me->last_task_ms = task_ms;
    me->avg_task_ms = wc->avg_task_ms;
    me->heartbeat_ns = now; me->last_finish_ns = now;
    if (it == rounds - 1) me->available = 1;                                 // the persistent launch ends with this round
  }
  fence_sys();                               // ONE fence: reductions + record + status performed before publication
  __syncthreads();
  if (tid == 0) {
    wc->dbg_ns[6] = globaltimer_ns();
    *reinterpret_cast<volatile unsigned long long*>(&rec->seq) = static_cast<unsigned long long>(clock_old) + 1ULL;
    if (it == rounds - 1) red_add_release_sys_s32(&p.sctl->avail_count, 1);
    wc->last_accepted = accept; wc->last_batch = round_batch; wc->last_loss = round_loss;
    wc->last_step = accept ? step : 0.f; wc->last_task_ns = task_ns;
    wc->last_staleness = S->tk_staleness; wc->last_k_after = S->tk_k_after;
    wc->ts = S->ts; wc->k_read = S->k_read;
    wc->rounds_done += 1ULL;
    S->done_ctas = 0; S->pull_ticket = 0; S->chunk_ctr = 0; S->loss_sum = 0.f; S->batch_count = 0;
    wc->dbg_ns[7] = globaltimer_ns();
    __threadfence();
    st_release_gpu_u32(&wc->round, round + 1u);
  }
  __syncthreads();
  return true;
}

template <int NV4, int kThreads>
__global__ void __launch_bounds__(kThreads, (kThreads == 256 && NV4 <= 8) ? 2 : 1)
dense_round_pipelined_kernel(const __grid_constant__ RoundParams p, const int rounds, PipeState* ps, float* g2, float* w2) {
  __shared__ RoundShared<kThreads> sh;
  extern __shared__ __align__(16) float s_dyn[];
  // same value in every CTA: no round can be published before every CTA of the grid has started (last-CTA election)
  const unsigned int round0 = *reinterpret_cast<volatile unsigned int*>(&p.wctl->round);
  for (int it = 0; it < rounds; ++it) {
    if (!pipelined_one_round<NV4, kThreads>(p, sh, s_dyn, round0 + static_cast<unsigned int>(it), it, rounds, ps, g2, w2)) return;
    __syncthreads();                         // sh / s_dyn are reused by the next round
  }
}

template <int NV4, int kThreads>
static cudaError_t launch_pipelined(const RoundParams& p, int rounds, int grid, PipeState* ps, float* g2, float* w2, cudaStream_t stream) {
  size_t smem = 2 * static_cast<size_t>(p.ld) * sizeof(float);
  auto k = dense_round_pipelined_kernel<NV4, kThreads>;
  cudaError_t e;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 0, per_sm = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, smem)) != cudaSuccess) return e;
  if (static_cast<long long>(grid) > static_cast<long long>(per_sm) * sms) return cudaErrorLaunchOutOfResources;   // must be co-resident
  k<<<grid, kThreads, smem, stream>>>(p, rounds, ps, g2, w2);
  return cudaGetLastError();
}

}  // namespace afb

extern "C" {

int afb_pipe_state_bytes() { return static_cast<int>(sizeof(afb::PipeState)); }

// ASGD only, no bucket gate, no delay injection, worker-local sampling seed (seed_mode 1 or 2); the grid must be fully
// co-resident and is expected to be the ONLY round kernel on its GPU (one logical worker per GPU).
// `pipe_state` = afb_pipe_state_bytes() zeroed bytes, `g2` / `w2` = [2][ld] floats (g2 zeroed), all on the worker's GPU.
int afb_dense_round_pipelined_launch(const afb::RoundParams* params, int rounds, int grid, void* pipe_state, float* g2, float* w2,
                                     void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.ld % 4 != 0 || p.ld > 2048 || p.ld <= 0 || rounds <= 0) return static_cast<int>(cudaErrorInvalidValue);
  if (p.algo != ALGO_ASGD || p.hist != nullptr || p.gate_threshold != 0 || p.delay_mode != 0 || p.pdl != 0 ||
      !(p.seed_mode == 1 || p.seed_mode == 2))
    return static_cast<int>(cudaErrorInvalidValue);
  PipeState* ps = static_cast<PipeState*>(pipe_state);
  int nv4 = static_cast<int>((p.ld / 4 + 31) / 32);
  cudaError_t e;
  switch (nv4) {
    case 1: e = launch_pipelined<1, 256>(p, rounds, grid, ps, g2, w2, s); break;
    case 2: e = launch_pipelined<2, 256>(p, rounds, grid, ps, g2, w2, s); break;
    case 3: case 4: e = launch_pipelined<4, 256>(p, rounds, grid, ps, g2, w2, s); break;
    case 5: case 6: case 7: e = launch_pipelined<7, 256>(p, rounds, grid, ps, g2, w2, s); break;
    case 8: e = launch_pipelined<8, 256>(p, rounds, grid, ps, g2, w2, s); break;
    case 9: case 10: case 11: case 12: e = launch_pipelined<12, 256>(p, rounds, grid, ps, g2, w2, s); break;
    default: e = launch_pipelined<16, 256>(p, rounds, grid, ps, g2, w2, s); break;
  }
  return static_cast<int>(e);
}

}  // extern "C"
