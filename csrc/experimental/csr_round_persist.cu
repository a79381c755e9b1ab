// EXPERIMENTAL — compiled into libafb_experimental.so, never loaded by the default path.
// Status: builds for sm_100a; NOT yet validated on hardware.
//
// Persistent fused round for WIDE sparse models (rcv1 shape: d = 47 236, CSR rows, ~74 nnz/row).  The default path
// runs such a round as four launches (gate -> multi-CTA pull -> gradient -> multi-CTA push/apply, ops/round.py
// `split`) because one CTA cannot move 189 KB of parameters over NVLink fast enough; at 8 GPUs the round is then
// ~69 us of which < 10 us is work.  Here ONE resident grid runs `rounds` rounds and uses in-kernel grid barriers
// instead of kernel boundaries, so the pull and the push are spread over every CTA (one float4 per thread, one NVLink
// round trip each):
//
//   A  every CTA loads its slice of w from the server (P2P) into w_local; CTA 0 reads the logical clock, decides
//      `proceed`, publishes ts / seed                                                      -> grid barrier 1
//   B  sparse gradient (warp per sampled row, dynamic 64-row chunks), scatter into g_local with L2 reductions;
//      the first CTA to finish takes the logical-clock ticket (staleness gate, k, step)   -> grid barrier 2
//   C  every CTA pushes its slice: Hogwild `red.sys.add` of -eta*g/parRecs into the server's w (SAGA: + alpha-bar),
//      zeroes its slice of g_local, one `fence.acq_rel.sys`                                -> grid barrier 3
//   D  CTA 0 publishes the result record / worker status / avail_count, resets the chunk counter, bumps `round`
//      (overlaps phase A of the next round on the other CTAs; the next barrier 1 orders it before the next phase B)
//
// The grid must be fully co-resident (barriers spin); the launcher checks occupancy.
#include "afb_round_impl.cuh"

namespace afb {

constexpr int kCsrPThreads = 512;
constexpr int kCsrPChunkRows = 64;

struct GridBar {
  unsigned long long count;          // arrival counter, monotonically increasing within a launch (zeroed by the launcher)
  unsigned long long pad;
};

// Sense-free grid barrier on a monotonically increasing counter: barrier number `n` (1-based, counted by the caller)
// is complete when count >= n * gridDim.x.  All memory operations of the CTA before the barrier are ordered before
// those of every CTA after it (threadfence + release arrive, acquire poll + CTA barrier).
AFB_DEVINL void grid_barrier(GridBar* gb, unsigned long long& n) {
  __threadfence();
  __syncthreads();
  ++n;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(&gb->count), "l"(1ULL) : "memory");
    const unsigned long long target = n * static_cast<unsigned long long>(gridDim.x);
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(&gb->count) : "memory");
      if (v < target) __nanosleep(20);
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCsrPThreads, 2) csr_round_persist_kernel(const __grid_constant__ RoundParams p, const int rounds,
                                                                             GridBar* gb) {
  __shared__ RoundShared<kCsrPThreads> sh;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  constexpr int kWarps = kCsrPThreads / 32;
  WorkerCtl* wc = p.wctl;
  const bool saga = (p.hist != nullptr);
  const long long nvec = p.ld >> 2;
  const long long gthreads = static_cast<long long>(gridDim.x) * kCsrPThreads;
  const long long gtid = static_cast<long long>(blockIdx.x) * kCsrPThreads + tid;
  const int num_chunks = (p.n_rows + kCsrPChunkRows - 1) / kCsrPChunkRows;
  unsigned long long nbar = 0;                    // barriers completed in this launch (the launcher zeroes `gb` on the stream)
  const unsigned int round0 = *reinterpret_cast<volatile unsigned int*>(&wc->round);

  for (int it = 0; it < rounds; ++it) {
    const unsigned int round = round0 + static_cast<unsigned int>(it);
    // ---------------------------------------------------------------- A: pull (every CTA) + clock / proceed (CTA 0)
    if (blockIdx.x == 0 && tid == 0) {
      wc->dbg_ns[0] = globaltimer_ns();
      if (p.gate_threshold > 0) {
        while (true) {
          unsigned long long ck0 = ld_acquire_sys_u64(&p.sctl->clock_k);
          if (static_cast<long long>(ck0 >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) break;
          if (ld_relaxed_sys_s32(&p.sctl->avail_count) >= p.gate_threshold) break;
          __nanosleep(200);
        }
      }
      const unsigned long long ck = ld_acquire_sys_u64(&p.sctl->clock_k);
      int proceed = 1;
      if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) proceed = 0;
      const unsigned long long now = globaltimer_ns();
      if (proceed) {
        red_add_release_sys_s32(&p.sctl->avail_count, -1);
        WorkerSlot* me = p.status + p.worker_id;
        if (wc->last_finish_ns) wc->wait_ns += now - wc->last_finish_ns;
        me->available = 0; me->busy_since_ns = now; me->heartbeat_ns = now; me->wait_ns = wc->wait_ns;
      }
      wc->ts = ck & 0xffffffffULL;
      wc->k_read = ck >> 32;
      unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
      if (p.seed_mode == 0) seed_eff += wc->k_read + 1ULL;
      else if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
      wc->seed_key = mix64(seed_eff);
      wc->proceed = proceed;
      wc->t_start_ns = now;
      if (proceed && p.delay_mode != 0) {
        unsigned long long avg = ld_acquire_sys_u64(&p.sctl->avg_delay_ns);
        if (avg) {
          float mult = p.delay_coeff;
          if (p.delay_mode >= 2) {
            float u = static_cast<float>(hash_u24(wc->seed_key ^ 0xD1B54A32D192ED03ULL, p.worker_id)) * (1.f / 16777216.f);
            mult = (p.delay_mode == 3) ? (2.5f + 7.5f * u) : (1.5f + u);
          }
          sleep_ns(static_cast<unsigned long long>(mult * static_cast<float>(avg)));
        }
      }
    }
    // with a bucket gate the parameters must not be read before the gate opens: CTA 0 gates first, everybody pulls after
    if (p.gate_threshold > 0) grid_barrier(gb, nbar);
    for (long long v = gtid; v < nvec; v += gthreads)
      *reinterpret_cast<float4*>(p.w_local + 4 * v) = ld_relaxed_sys_f32x4(p.w_server + 4 * v);          // K10, all CTAs
    grid_barrier(gb, nbar);                                                                                // ---- barrier 1
    if (blockIdx.x == 0 && tid == 0) wc->dbg_ns[2] = globaltimer_ns();
    if (!*reinterpret_cast<volatile int*>(&wc->proceed)) {
      if (blockIdx.x == 0 && tid == 0) {
        wc->last_accepted = -1;
        __threadfence();
        st_release_gpu_u32(&wc->round, round + 1u);
      }
      return;                                    // every CTA sees proceed == 0 after barrier 1
    }

    // ---------------------------------------------------------------- B: sparse gradient
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
    const float* w = p.w_local;
    float loss_acc = 0.f;
    unsigned int cnt = 0;
    unsigned int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&wc->chunk_ctr, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    while (chunk < static_cast<unsigned int>(num_chunks)) {
      unsigned int nextc = 0;
      if (lane == 0) nextc = atomicAdd(&wc->chunk_ctr, 1u);
      const int base = static_cast<int>(chunk) * kCsrPChunkRows;
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
          for (long long k = s + lane; k < e; k += 32) dot = fmaf(p.values[k], __ldcg(w + p.indices[k]), dot);
          const float margin = warp_sum(dot);
          float coeff, loss;
          loss_coeff_rt(p.loss_id, margin, __ldg(p.y + row), coeff, loss);
          float eff = coeff;
          if (saga) {
            const float h = p.hist[row];
            eff = coeff - h;
            if (lane == 0) p.hist[row] = coeff;
          }
          for (long long k = s + lane; k < e; k += 32) red_add_gpu_f32(p.g_local + p.indices[k], eff * p.values[k]);
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
        const unsigned long long ts = wc->ts;
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
          long long cmp = (p.algo == ALGO_ASAGA) ? static_cast<long long>(old >> 32) - static_cast<long long>(ts) : staleness;
          accept = (cmp <= p.taw) ? 1 : 0;
          k_old = old >> 32;
          if (accept) {
            unsigned long long o2 = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL << 32);
            k_old = o2 >> 32;
            if (static_cast<long long>(k_old) >= p.max_iters) accept = 0;
          }
        }
        wc->tk_accept = accept;
        wc->tk_step = (p.algo == ALGO_ASGD) ? p.gamma * rsqrtf(static_cast<float>(k_old / static_cast<unsigned long long>(p.num_workers)) + 1.f) : p.gamma;
        wc->tk_clock_old = static_cast<unsigned int>(old & 0xffffffffULL);
        wc->tk_staleness = staleness;
        wc->tk_k_after = static_cast<long long>(k_old) + (accept ? 1 : 0);
      }
    }
    grid_barrier(gb, nbar);                                                                                // ---- barrier 2
    if (blockIdx.x == 0 && tid == 0) wc->dbg_ns[4] = globaltimer_ns();

    // ---------------------------------------------------------------- C: push-and-apply, every CTA its slice
    const int accept = *reinterpret_cast<volatile int*>(&wc->tk_accept);
    const float step = *reinterpret_cast<volatile float*>(&wc->tk_step);
    const float a = -step * p.inv_par_recs;
    for (long long v = gtid; v < nvec; v += gthreads) {
      float4 g = __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v));
      if (accept) {
        if (p.algo == ALGO_ASGD) {
          red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x, a * g.y, a * g.z, a * g.w));
        } else {
          float4 ab = ld_relaxed_sys_f32x4(p.abar_server + 4 * v);
          red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x - step * ab.x, a * g.y - step * ab.y, a * g.z - step * ab.z, a * g.w - step * ab.w));
          red_add_sys_f32x4(p.abar_server + 4 * v, make_float4(g.x * p.inv_n, g.y * p.inv_n, g.z * p.inv_n, g.w * p.inv_n));
        }
      }
      *reinterpret_cast<float4*>(p.g_local + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_sys();                                  // this CTA's reductions are performed before anybody publishes
    grid_barrier(gb, nbar);                                                                                // ---- barrier 3

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
int afb_csr_round_persist_launch(const afb::RoundParams* params, int rounds, int grid, int share, void* grid_bar, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  if (p.ld % 4 != 0 || p.ld <= 0 || rounds <= 0 || p.indptr == nullptr) return static_cast<int>(cudaErrorInvalidValue);
  if (!(p.algo == ALGO_ASGD || p.algo == ALGO_ASAGA) || p.hist_prev != nullptr || p.seed_mode == 3 || p.pdl != 0)
    return static_cast<int>(cudaErrorInvalidValue);
  int dev = 0, sms = 0, per_sm = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_round_persist_kernel, kCsrPThreads, 0)) != cudaSuccess) return static_cast<int>(e);
  if (share < 1) share = 1;
  if (static_cast<long long>(grid) * share > static_cast<long long>(per_sm) * sms) return static_cast<int>(cudaErrorLaunchOutOfResources);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((e = cudaMemsetAsync(grid_bar, 0, sizeof(GridBar), st)) != cudaSuccess) return static_cast<int>(e);
  csr_round_persist_kernel<<<grid, kCsrPThreads, 0, st>>>(p, rounds, static_cast<GridBar*>(grid_bar));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
