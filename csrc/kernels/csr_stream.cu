// Persistent fused round for WIDE sparse models (rcv1 shape: d = 47 236, CSR rows, ~74 nnz/row) — sm_100a.
//
// One CTA cannot move 189 KB of parameters over NVLink fast enough, and a round of such a worker is only a few
// microseconds of work (4.2 k sampled rows, 313 k non-zeros at 8 GPUs), so launch boundaries dominate a split round
// (gate -> pull -> grad -> push kernels: ~69 us per round measured in round 1).  Here ONE resident grid runs `rounds`
// rounds with in-kernel grid barriers instead of kernel boundaries; the pull and the push are spread over every CTA
// (one float4 per thread, one NVLink round trip each):
//
//   A  CTA 0 gates (bucket / SSP / enable mask), reads the logical clock, decides `proceed`, publishes ts / seed;
//      every CTA loads its slice of w from the server (P2P) into w_local                      -> grid barrier 1
//   B  sparse gradient (warp per sampled row, dynamic 64-row chunks): gather-dot against the L2-resident w_local,
//      scatter coeff*val into g_local with L2 reductions (the reference pays an O(d) dense allocation + add per SAMPLE
//      here, mllib/linalg/BLAS.scala:74-93, 134-147); the first CTA to finish takes the logical-clock ticket
//                                                                                              -> grid barrier 2
//   C  every CTA pushes its slice: Hogwild `red.sys.add` of -eta*(g/parRecs [+ lambda w]) into the server's w
//      (SAGA: + alpha-bar), zeroes its slice of g_local, one `fence.acq_rel.sys`               -> grid barrier 3
//   D  CTA 0 publishes the result record / worker status / avail_count, resets the chunk counter, bumps `round`
//      (overlaps phase A of the next round on the other CTAs; the next barrier 1 orders it before the next phase B)
//
// The grid must be fully co-resident (barriers spin); the launcher checks occupancy.  Every spin is bounded (SpinGuard).
#include "afb_round_impl.cuh"

namespace afb {

constexpr int kCsrSThreads = 512;
constexpr int kCsrSChunkRows = 64;

struct GridBar {
  unsigned long long count;          // arrival counter, monotonically increasing within a launch (zeroed by the launcher)
  unsigned long long pad;
};

// Sense-free grid barrier on a monotonically increasing counter: barrier number `n` (1-based, counted by the caller)
// is complete when count >= n * gridDim.x.  All memory operations of the CTA before the barrier are ordered before
// those of every CTA after it (threadfence + release arrive, acquire poll + CTA barrier).  Returns false when the
// bounded wait expired (WorkerCtl.abort raised): the caller must leave the kernel.
AFB_DEVINL bool grid_barrier(GridBar* gb, unsigned long long& n, const RoundParams& p, RoundShared<kCsrSThreads>& sh) {
  __threadfence();
  __syncthreads();
  ++n;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(&gb->count), "l"(1ULL) : "memory");
    const unsigned long long target = n * static_cast<unsigned long long>(gridDim.x);
    unsigned long long v;
    SpinGuard sg;
    unsigned int ok = 1u;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&gb->count) : "memory");
      if (v < target) {
        __nanosleep(20);
        if (sg.expired(p.spin_limit_ns) || aborted(p.wctl)) { raise_abort(p.wctl); ok = 0u; break; }
      }
    } while (v < target);
    sh.flag = ok;
  }
  __syncthreads();
  const bool ok = sh.flag != 0u;
  __syncthreads();
  return ok;
}

__global__ void __launch_bounds__(kCsrSThreads, 2) csr_stream_kernel(const __grid_constant__ RoundParams p, const int rounds,
                                                                      GridBar* gb) {
  __shared__ RoundShared<kCsrSThreads> sh;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  constexpr int kWarps = kCsrSThreads / 32;
  WorkerCtl* wc = p.wctl;
  const bool saga = (p.hist != nullptr);
  const long long nvec = p.ld >> 2;
  const long long gthreads = static_cast<long long>(gridDim.x) * kCsrSThreads;
  const long long gtid = static_cast<long long>(blockIdx.x) * kCsrSThreads + tid;
  const int num_chunks = (p.n_rows + kCsrSChunkRows - 1) / kCsrSChunkRows;
  const bool gated = (p.gate_threshold > 0 || p.ssp_bound >= 0);
  unsigned long long nbar = 0;                    // barriers completed in this launch (the launcher zeroes `gb` on the stream)
  const unsigned int round0 = *reinterpret_cast<volatile unsigned int*>(&wc->round);

  for (int it = 0; it < rounds; ++it) {
    const unsigned int round = round0 + static_cast<unsigned int>(it);
    // ---------------------------------------------------------------- A: gate / clock / proceed (CTA 0) + pull (every CTA)
    if (blockIdx.x == 0 && tid < 32) {
      if (tid == 0) wc->dbg_ns[0] = globaltimer_ns();
      if (gated) gate_wait_warp(p, wc, tid);
      if (tid == 0) {
        wc->dbg_ns[1] = globaltimer_ns();
        const unsigned long long ck = ld_acquire_sys_u64(&p.sctl->clock_k);
        int proceed = PROCEED_RUN;
        if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop) || aborted(wc)) proceed = PROCEED_STOP;
        else if (p.use_mask) {
          const unsigned long long mw = ld_acquire_sys_u64(&p.sctl->enable_mask[(p.worker_id >> 6) & 3]);
          if (!((mw >> (p.worker_id & 63)) & 1ULL)) proceed = PROCEED_SKIP;
        }
        const unsigned long long now = globaltimer_ns();
        WorkerSlot* me = p.status + p.worker_id;
        if (proceed == PROCEED_RUN) {
          red_add_release_sys_s32(&p.sctl->avail_count, -1);
          if (wc->last_finish_ns) wc->wait_ns += now - wc->last_finish_ns;
          me->available = 0; me->busy_since_ns = now; me->heartbeat_ns = now; me->wait_ns = wc->wait_ns;
        } else if (proceed == PROCEED_SKIP) {
          me->heartbeat_ns = now;
        }
        wc->ts = ck & 0xffffffffULL;
        wc->k_read = ck >> 32;
        unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
        if (p.seed_mode == 0) seed_eff += wc->k_read + 1ULL;
        else if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
        wc->seed_key = mix64(seed_eff);
        wc->proceed = proceed;
        wc->t_start_ns = now;
        if (proceed == PROCEED_RUN && p.delay_mode != 0) {
          unsigned long long avg = ld_acquire_sys_u64(&p.sctl->avg_delay_ns);
          if (avg) {
            float mult = eff_delay_coeff(p, round);
            if (p.delay_mode >= 2) {
              float u = static_cast<float>(hash_u24(wc->seed_key ^ 0xD1B54A32D192ED03ULL, p.worker_id)) * (1.f / 16777216.f);
              mult = (p.delay_mode == 3) ? (2.5f + 7.5f * u) : (1.5f + u);
            }
            sleep_ns(static_cast<unsigned long long>(mult * static_cast<float>(avg)));
          }
        }
      }
    }
    // with a gate the parameters must not be read before it opens: CTA 0 gates first, everybody pulls after
    if (gated && !grid_barrier(gb, nbar, p, sh)) return;
    for (long long v = gtid; v < nvec; v += gthreads)
      *reinterpret_cast<float4*>(p.w_local + 4 * v) = ld_relaxed_sys_f32x4(p.w_server + 4 * v);          // K10, all CTAs
    if (!grid_barrier(gb, nbar, p, sh)) return;                                                            // ---- barrier 1
    if (blockIdx.x == 0 && tid == 0) wc->dbg_ns[2] = globaltimer_ns();
    const int proceed = *reinterpret_cast<volatile int*>(&wc->proceed);         // same value in every CTA after barrier 1
    if (proceed != PROCEED_RUN) {
      if (blockIdx.x == 0 && tid == 0) {
        wc->last_accepted = (proceed == PROCEED_SKIP) ? -2 : -1;
        if (proceed == PROCEED_SKIP) wc->skipped_rounds += 1u;
        __threadfence();
        st_release_gpu_u32(&wc->round, round + 1u);
      }
      if (proceed == PROCEED_STOP) return;
      continue;                                  // masked worker: the round is a no-op
    }

    // ---------------------------------------------------------------- B: sparse gradient
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
    const unsigned int thr = static_cast<unsigned int>(eff_thr24(p, round));
    const float* w = p.w_local;
    float loss_acc = 0.f;
    unsigned int cnt = 0;
    unsigned int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&wc->chunk_ctr, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < static_cast<unsigned int>(num_chunks)) {
      unsigned int nextc = 0;
      if (lane == 0) nextc = atomicAdd(&wc->chunk_ctr, 1u);
      const int base = static_cast<int>(chunk) * kCsrSChunkRows;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = base + 32 * half + lane;
        bool sel = false;
        if (r < p.n_rows) {
          unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r]) : static_cast<unsigned long long>(p.row_base + r);
          sel = hash_u24(key, g) < thr;
        }
        unsigned int m = __ballot_sync(0xffffffffu, sel);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int row = base + 32 * half + b;
          const long long s = p.indptr[row], e = p.indptr[row + 1];
          float dot = 0.f;
          for (long long k = s + lane; k < e; k += 32) dot = fmaf(p.values[k], __ldcg(w + p.indices[k]), dot);   // K1 (gather)
          const float margin = warp_sum(dot);
          float coeff, loss;
          loss_coeff_rt(p.loss_id, margin, __ldg(p.y + row), coeff, loss);                                       // K2
          float eff = coeff;
          if (saga) {                                                                                            // K4 + K7
            const float h = p.hist[row];
            eff = coeff - h;
            if (lane == 0) p.hist[row] = coeff;
          }
          for (long long k = s + lane; k < e; k += 32) red_add_gpu_f32(p.g_local + p.indices[k], eff * p.values[k]);   // K3 (scatter)
          loss_acc += loss;
          ++cnt;
        }
      }
      chunk = __shfl_sync(0xffffffffu, nextc, 0);
    }
    if (lane == 0) { sh.red[warp] = loss_acc; sh.cnt[warp] = cnt; }
    __syncthreads();
    if (tid == 0) {
      float l = 0.f; unsigned int n = 0;
#pragma unroll
      for (int w8 = 0; w8 < kWarps; ++w8) { l += sh.red[w8]; n += sh.cnt[w8]; }
      atomicAdd(&wc->loss_sum, l);
      atomicAdd(&wc->batch_count, n);
      // early logical-clock ticket by the first CTA to finish its gradient work
      if (atomicAdd(&wc->done_ctas, 1u) == 0) {
        const Ticket t = take_ticket(p, wc->ts, round);
        wc->tk_accept = t.accept; wc->tk_step = t.step; wc->tk_clock_old = t.clock_old;
        wc->tk_staleness = t.staleness; wc->tk_k_after = t.k_after;
      }
    }
    if (!grid_barrier(gb, nbar, p, sh)) return;                                                            // ---- barrier 2
    if (blockIdx.x == 0 && tid == 0) wc->dbg_ns[4] = globaltimer_ns();

    // ---------------------------------------------------------------- C: push-and-apply, every CTA its slice
    const int accept = *reinterpret_cast<volatile int*>(&wc->tk_accept);
    const float step = *reinterpret_cast<volatile float*>(&wc->tk_step);
    if (accept) hogwild_apply_range(p, p.g_local, p.w_local, step, gtid, gthreads);                       // K9 + K5/K6 (+ K8)
    for (long long v = gtid; v < nvec; v += gthreads)
      *reinterpret_cast<float4*>(p.g_local + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_sys();                                  // this CTA's reductions are performed before anybody publishes
    if (!grid_barrier(gb, nbar, p, sh)) return;                                                            // ---- barrier 3

    // ---------------------------------------------------------------- D: publication (CTA 0), overlaps the next pull
    if (blockIdx.x == 0 && tid == 0) {
      const unsigned long long now = globaltimer_ns();
      wc->dbg_ns[5] = now;
      const unsigned long long task_ns = now - wc->t_start_ns;
      const unsigned int clock_old = wc->tk_clock_old;
      const float round_loss = wc->loss_sum;
      const unsigned int round_batch = wc->batch_count;
      ResultRec* rec = p.ring + (clock_old % static_cast<unsigned int>(p.ring_size));
      rec->worker = p.worker_id; rec->accepted = accept;
      rec->ts = static_cast<long long>(wc->ts); rec->staleness = wc->tk_staleness;
      rec->k_after = wc->tk_k_after; rec->batch = round_batch;
      rec->loss_sum = round_loss; rec->step = accept ? step : 0.f;
      rec->t_start_ns = wc->t_start_ns; rec->t_end_ns = now;
      WorkerSlot* me = p.status + p.worker_id;
      const unsigned long long ntasks = wc->rounds_done + 1ULL;
      const float task_ms = static_cast<float>(task_ns) * 1e-6f;
      wc->avg_task_ms += (task_ms - wc->avg_task_ms) / static_cast<float>(ntasks);
      wc->last_finish_ns = now;
      me->staleness = wc->tk_staleness; me->num_tasks = static_cast<long long>(ntasks);
      me->last_task_ms = task_ms; me->avg_task_ms = wc->avg_task_ms;
      me->heartbeat_ns = now; me->last_finish_ns = now;
      if (accept && p.warmup_updates > 0 && wc->tk_k_after <= static_cast<long long>(p.warmup_updates)) {
        asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(&p.sctl->cul_time_ns), "l"(task_ns) : "memory");
        asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(&p.sctl->cul_count), "r"(1u) : "memory");
      }
      me->available = 1;
      fence_sys();
      *reinterpret_cast<volatile unsigned long long*>(&rec->seq) = static_cast<unsigned long long>(clock_old) + 1ULL;
      red_add_release_sys_s32(&p.sctl->avail_count, 1);
      if (accept && p.warmup_updates > 0 && wc->tk_k_after == static_cast<long long>(p.warmup_updates)) {
        unsigned long long tot = ld_acquire_sys_u64(&p.sctl->cul_time_ns);
        unsigned int c = ld_acquire_sys_u32(&p.sctl->cul_count);
        if (c) atomicExch_system(&p.sctl->avg_delay_ns, tot / c);
      }
      wc->last_accepted = accept; wc->last_batch = round_batch; wc->last_loss = round_loss;
      wc->last_step = accept ? step : 0.f; wc->last_task_ns = task_ns;
      wc->last_staleness = wc->tk_staleness; wc->last_k_after = wc->tk_k_after;
      wc->rounds_done += 1ULL;
      wc->done_ctas = 0; wc->chunk_ctr = 0; wc->loss_sum = 0.f; wc->batch_count = 0;
      wc->dbg_ns[7] = globaltimer_ns();
      __threadfence();
      st_release_gpu_u32(&wc->round, round + 1u);
    }
    // no barrier here: the next round's barrier 1 orders CTA 0's resets (chunk_ctr, done_ctas) before anybody's phase B,
    // and nobody touches chunk_ctr / done_ctas / loss_sum in phase A
  }
}

}  // namespace afb

extern "C" {

int afb_grid_bar_bytes() { return static_cast<int>(sizeof(afb::GridBar)); }

// Async algorithms only (ASGD / ASAGA without history rollback).  `grid_bar` = afb_grid_bar_bytes() bytes on the worker's GPU
// (zeroed on the stream before every launch).  The grid must be fully co-resident.
int afb_csr_stream_launch(const afb::RoundParams* params, int rounds, int grid, int share, void* grid_bar, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  if (p.ld % 4 != 0 || p.ld <= 0 || rounds <= 0 || p.indptr == nullptr) return static_cast<int>(cudaErrorInvalidValue);
  if (!(p.algo == ALGO_ASGD || p.algo == ALGO_ASAGA) || p.hist_prev != nullptr || p.seed_mode == 3 || p.pdl != 0)
    return static_cast<int>(cudaErrorInvalidValue);
  int dev = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return static_cast<int>(e);
  // resident-CTA capacity of this device: queried once (host-side driver calls sit inside the callers' timed regions)
  static int s_capacity[64] = {0};
  int capacity = (dev >= 0 && dev < 64) ? s_capacity[dev] : 0;
  if (capacity == 0) {
    int sms = 0, per_sm = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return static_cast<int>(e);
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_stream_kernel, kCsrSThreads, 0)) != cudaSuccess) return static_cast<int>(e);
    capacity = per_sm * sms;
    if (dev >= 0 && dev < 64) s_capacity[dev] = capacity;
  }
  if (share < 1) share = 1;
  if (static_cast<long long>(grid) * share > static_cast<long long>(capacity)) return static_cast<int>(cudaErrorLaunchOutOfResources);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((e = cudaMemsetAsync(grid_bar, 0, sizeof(GridBar), st)) != cudaSuccess) return static_cast<int>(e);
  csr_stream_kernel<<<grid, kCsrSThreads, 0, st>>>(p, rounds, static_cast<GridBar*>(grid_bar));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
