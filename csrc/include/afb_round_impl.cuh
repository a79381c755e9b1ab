// Shared phases of the fused round kernels: gate + pull (phase 0) and the cross-GPU tail
// (grid reduce bookkeeping, staleness gate, Hogwild push-and-apply / BSP all-gather apply).
// Included by dense_round.cu and csr_round.cu; the gradient body in between is layout specific.
#pragma once
#include "afb_round.cuh"

namespace afb {

template <int kThreads>
struct RoundShared {
  float red[kThreads / 32];
  unsigned int cnt[kThreads / 32];
  unsigned int flag;
  float bcast[4];
};

AFB_DEVINL unsigned int fetch_chunk_ctr(unsigned int* ctr, int lane) {
  unsigned int c = 0;
  if (lane == 0) c = atomicAdd(ctr, 1u);
  return __shfl_sync(0xffffffffu, c, 0);
}

// Phase 0 is split in two so that its NVLink round trip overlaps useful work:
//   round_begin_issue()    - puller election (ticket), bucket gate, and the pull's loads are ISSUED
//                            (clock word + first chunk of w stay in registers, latency in flight);
//   ... the caller hashes its first row chunk / prefetches its first rows here ...
//   round_begin_complete() - the puller consumes the loads, publishes w_local / ts / seed and releases
//                            the CTA-visible flag; every other CTA waits for the flag.
// Returns false when the worker has been stopped (k >= max_iters or stop flag): the kernel must then
// return immediately (bookkeeping for the next launch is already done).
struct RoundPull {
  float4 wv;
  unsigned long long ck;
  bool puller;
};

template <int kThreads>
AFB_DEVINL void round_begin_issue(const RoundParams& p, RoundShared<kThreads>& sh, const unsigned int round, RoundPull& pl) {
  const int tid = threadIdx.x;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  const bool is_async = (p.algo == ALGO_ASGD || p.algo == ALGO_ASAGA);
  pl.wv = make_float4(0.f, 0.f, 0.f, 0.f);
  pl.ck = static_cast<unsigned long long>(round);               // sync: local iteration counter
  pl.puller = false;
  if (p.algo == ALGO_GRAD_ONLY) return;
  if (tid == 0) sh.flag = atomicAdd(&wc->pull_ticket, 1u);
  __syncthreads();
  pl.puller = (sh.flag == 0);
  if (!pl.puller) return;
  // This CTA is the puller: bucket gate, then ONE NVLink round trip — the logical-clock read and the
  // vectorised fetch of w are issued together (ts may be at most in-flight-updates older than w).
  if (tid == 0) {
    wc->dbg_ns[0] = globaltimer_ns();
    if (is_async && p.gate_threshold > 0) {
      // partial barrier: start only while >= floor(P*beta) workers are idle (this one included)
      while (true) {
        unsigned long long ck0 = ld_acquire_sys_u64(&p.sctl->clock_k);
        if (static_cast<long long>(ck0 >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) break;
        if (ld_relaxed_sys_s32(&p.sctl->avail_count) >= p.gate_threshold) break;
        __nanosleep(200);
      }
    }
    wc->dbg_ns[1] = globaltimer_ns();
  }
  if (is_async && p.gate_threshold > 0) __syncthreads();         // nobody reads w before the gate opens
  if (is_async) {
    if (tid < nvec) pl.wv = ld_relaxed_sys_f32x4(p.w_server + 4 * tid);          // K10 (first chunk), in flight
    if (tid == 0) pl.ck = ld_acquire_sys_u64(&p.sctl->clock_k);
  }
}

template <int kThreads>
AFB_DEVINL bool round_begin_complete(const RoundParams& p, RoundShared<kThreads>& sh, const unsigned int round, const RoundPull& pl) {
  const int tid = threadIdx.x;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  const bool is_async = (p.algo == ALGO_ASGD || p.algo == ALGO_ASAGA);
  if (p.algo != ALGO_GRAD_ONLY) {
    if (pl.puller) {
      if (tid == 0) {
        int proceed = 1;
        const unsigned long long ck = pl.ck;
        if (is_async) {
          if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) proceed = 0;
          if (proceed) {
            red_add_release_sys_s32(&p.sctl->avail_count, -1);      // busy from submit until the result lands
            WorkerSlot* me = p.status + p.worker_id;
            unsigned long long now = globaltimer_ns();
            if (wc->last_finish_ns) wc->wait_ns += now - wc->last_finish_ns;
            me->available = 0;                                      // posted stores only (no remote reads)
            me->busy_since_ns = now;
            me->heartbeat_ns = now;
            me->wait_ns = wc->wait_ns;
          }
        }
        wc->ts = ck & 0xffffffffULL;
        wc->k_read = is_async ? (ck >> 32) : round;
        unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
        if (p.seed_mode == 0) seed_eff += wc->k_read + 1ULL;
        else if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(p.pdl ? static_cast<unsigned int>(p.round_host) : round) + 1ULL;
        wc->seed_key = mix64(seed_eff);
        wc->proceed = proceed;
        wc->t_start_ns = globaltimer_ns();
      }
      if (is_async) {
        // K10: parameter pull — vectorised sys-scope loads from the server's HBM over NVLink.
        // (BSP modes keep a bit-identical local replica, so there is nothing to pull.)
        if (tid < nvec) *reinterpret_cast<float4*>(p.w_local + 4 * tid) = pl.wv;
        for (int v = tid + kThreads; v < nvec; v += kThreads)
          *reinterpret_cast<float4*>(p.w_local + 4 * v) = ld_relaxed_sys_f32x4(p.w_server + 4 * v);
      }
      __syncthreads();
      if (tid == 0) wc->dbg_ns[2] = globaltimer_ns();
      if (wc->proceed) {
        if (tid == 0 && p.delay_mode != 0) {
          // A18 straggler injector: sleep a multiple of the calibrated mean task time.
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
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_gpu_u32(&wc->pull_flag, round + 1u);
    } else {
      if (tid == 0) {
        while (ld_acquire_gpu_u32(&wc->pull_flag) != round + 1u) __nanosleep(32);
      }
      __syncthreads();
    }
    if (!*reinterpret_cast<volatile int*>(&wc->proceed)) {
      // stopped: still run the last-CTA bookkeeping so counters reset for the next launch
      __syncthreads();
      if (tid == 0) {
        unsigned int t = atomicAdd(&wc->done_ctas, 1u);
        if (t == gridDim.x - 1) {
          wc->done_ctas = 0; wc->pull_ticket = 0; wc->chunk_ctr = 0;
          wc->last_accepted = -1;
          __threadfence();
          wc->round = round + 1u;
        }
      }
      return false;
    }
  }

  return true;
}

template <int kThreads>
AFB_DEVINL bool round_begin(const RoundParams& p, RoundShared<kThreads>& sh, const unsigned int round) {
  RoundPull pl;
  round_begin_issue<kThreads>(p, sh, round, pl);
  return round_begin_complete<kThreads>(p, sh, round, pl);
}

// Phase 2b + 3.  Called by every thread after the CTA has flushed its partial gradient into
// g_local; `loss_acc` / `cnt` are the per-thread (lane-0-valid per warp) loss sum and row count.
template <int kThreads>
AFB_DEVINL void round_finish(const RoundParams& p, RoundShared<kThreads>& sh, const unsigned int round,
                             float loss_acc, unsigned int cnt) {
  constexpr int kWarps = kThreads / 32;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  const bool is_async = (p.algo == ALGO_ASGD || p.algo == ALGO_ASAGA);
  if (lane == 0) { sh.red[warp] = loss_acc; sh.cnt[warp] = cnt; }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f; unsigned int n = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { l += sh.red[w]; n += sh.cnt[w]; }
    atomicAdd(&wc->loss_sum, l);
    atomicAdd(&wc->batch_count, n);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.flag = atomicAdd(&wc->done_ctas, 1u);
  __syncthreads();
  const unsigned int my_ticket = sh.flag;
  const bool is_last_cta = (my_ticket == gridDim.x - 1);
  __syncthreads();                           // sh.flag is reused below

  // K9 (part 1): the first CTA to finish takes the logical-clock ticket — staleness gate, accepted-update
  // counter, step size — while the remaining CTAs are still draining; the last CTA only has to apply.
  if (is_async && my_ticket == 0 && tid == 0) {
    const unsigned long long ts = wc->ts;
    unsigned long long old, k_old;
    long long staleness;
    int accept;
    if (p.taw >= (1LL << 31)) {
      // never stale: one NVLink round trip ticks clock and k together
      old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, (1ULL << 32) | 1ULL);
      k_old = old >> 32;
      staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
      accept = (static_cast<long long>(k_old) < p.max_iters) ? 1 : 0;
    } else {
      old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL);
      staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
      // ASGD compares the clock-based staleness (SparkASGDThread.scala:169); ASAGA compares
      // k - ts (SparkASAGAThread.scala:184, quirk B.6) — both are available here.
      long long cmp = staleness;
      if (p.algo == ALGO_ASAGA) cmp = static_cast<long long>(old >> 32) - static_cast<long long>(ts);
      accept = (cmp <= p.taw) ? 1 : 0;
      k_old = old >> 32;
      if (accept) {
        unsigned long long o2 = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL << 32);
        k_old = o2 >> 32;
        if (static_cast<long long>(k_old) >= p.max_iters) accept = 0;
      }
    }
    float step;
    if (p.algo == ALGO_ASGD) step = p.gamma * rsqrtf(static_cast<float>(k_old / static_cast<unsigned long long>(p.num_workers)) + 1.f);
    else step = p.gamma;
    wc->tk_accept = accept; wc->tk_step = step; wc->tk_clock_old = static_cast<unsigned int>(old & 0xffffffffULL);
    wc->tk_staleness = staleness; wc->tk_k_after = static_cast<long long>(k_old) + (accept ? 1 : 0);
    __threadfence();
    st_release_gpu_u32(&wc->tk_flag, round + 1u);
  }
  if (!is_last_cta) return;

  // ------------------------------------------------------------------ phase 3 (last CTA): push-and-apply
  if (tid == 0) wc->dbg_ns[3] = globaltimer_ns();
  __threadfence();
  const float round_loss = *reinterpret_cast<volatile float*>(&wc->loss_sum);
  const unsigned int round_batch = *reinterpret_cast<volatile unsigned int*>(&wc->batch_count);

  if (p.algo == ALGO_GRAD_ONLY) {
    if (tid == 0) {
      wc->last_loss = round_loss; wc->last_batch = round_batch;
      wc->done_ctas = 0; wc->chunk_ctr = 0; wc->loss_sum = 0.f; wc->batch_count = 0;
      wc->rounds_done += 1ULL;
      __threadfence();
      wc->round = round + 1u;
    }
    return;   // g_local is the output; the caller consumes and re-zeroes it
  }

  if (is_async) {
    // K9 (part 2) + K5/K6: Hogwild reductions with the decision taken by the ticket holder.
    if (tid == 0) {
      while (ld_acquire_gpu_u32(&wc->tk_flag) != round + 1u) __nanosleep(32);
      wc->dbg_ns[4] = globaltimer_ns();
      sh.flag = static_cast<unsigned int>(*reinterpret_cast<volatile int*>(&wc->tk_accept));
      sh.bcast[0] = *reinterpret_cast<volatile float*>(&wc->tk_step);
      sh.bcast[1] = __uint_as_float(*reinterpret_cast<volatile unsigned int*>(&wc->tk_clock_old));
      wc->last_staleness = *reinterpret_cast<volatile long long*>(&wc->tk_staleness);
      wc->last_k_after = *reinterpret_cast<volatile long long*>(&wc->tk_k_after);
    }
    __syncthreads();
    const int accept = static_cast<int>(sh.flag);
    const float step = sh.bcast[0];
    if (accept) {
      if (p.algo == ALGO_ASGD) {
        const float a = -step * p.inv_par_recs;
        for (int v = tid; v < nvec; v += kThreads) {
          float4 g = __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v));
          red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x, a * g.y, a * g.z, a * g.w));
        }
      } else {
        // w <- w - gamma*(g/parRecs + abar)  (abar BEFORE its update);  abar <- abar + g/N
        const float a = -step * p.inv_par_recs;
        for (int v = tid; v < nvec; v += kThreads) {
          float4 g = __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v));
          float4 ab = ld_relaxed_sys_f32x4(p.abar_server + 4 * v);
          red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x - step * ab.x, a * g.y - step * ab.y,
                                                            a * g.z - step * ab.z, a * g.w - step * ab.w));
          red_add_sys_f32x4(p.abar_server + 4 * v, make_float4(g.x * p.inv_n, g.y * p.inv_n, g.z * p.inv_n, g.w * p.inv_n));
        }
      }
    }
    for (int v = tid; v < nvec; v += kThreads)
      *reinterpret_cast<float4*>(p.g_local + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned long long now = 0, task_ns = 0;
    unsigned int clock_old = 0;
    ResultRec* rec = nullptr;
    if (tid == 0) {
      now = globaltimer_ns();
      wc->dbg_ns[5] = now;
      task_ns = now - wc->t_start_ns;
      clock_old = __float_as_uint(sh.bcast[1]);
      // result ring (backs AsyncContext.collect_all on the fused backend): posted stores, published below
      rec = p.ring + (clock_old % static_cast<unsigned int>(p.ring_size));
      rec->worker = p.worker_id; rec->accepted = accept;
      rec->ts = static_cast<long long>(wc->ts); rec->staleness = wc->last_staleness;
      rec->k_after = wc->last_k_after; rec->batch = round_batch;
      rec->loss_sum = round_loss; rec->step = accept ? step : 0.f;
      rec->t_start_ns = wc->t_start_ns; rec->t_end_ns = now;
      // worker status (device mirror of workerState)
      WorkerSlot* me = p.status + p.worker_id;
      const unsigned long long ntasks = wc->rounds_done + 1ULL;
      const float task_ms = static_cast<float>(task_ns) * 1e-6f;
      wc->avg_task_ms += (task_ms - wc->avg_task_ms) / static_cast<float>(ntasks);
      wc->last_finish_ns = now;
      me->staleness = wc->last_staleness;
      me->num_tasks = static_cast<long long>(ntasks);
      me->last_task_ms = task_ms;
      me->avg_task_ms = wc->avg_task_ms;
      me->heartbeat_ns = now; me->last_finish_ns = now;
      // delay-injector calibration (A18): fire-and-forget accumulation during the warm-up window;
      // the worker that lands update #warmup publishes the mean.
      if (accept && p.warmup_updates > 0 && wc->last_k_after <= static_cast<long long>(p.warmup_updates)) {
        asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(&p.sctl->cul_time_ns), "l"(task_ns) : "memory");
        asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(&p.sctl->cul_count), "r"(1u) : "memory");
      }
      me->available = 1;
    }
    fence_sys();                               // ONE fence: reductions + record + status performed before publication
    __syncthreads();
    if (tid == 0) {
      wc->dbg_ns[6] = globaltimer_ns();
      *reinterpret_cast<volatile unsigned long long*>(&rec->seq) = static_cast<unsigned long long>(clock_old) + 1ULL;
      red_add_release_sys_s32(&p.sctl->avail_count, 1);
      if (accept && p.warmup_updates > 0 && wc->last_k_after == static_cast<long long>(p.warmup_updates)) {
        unsigned long long tot = ld_acquire_sys_u64(&p.sctl->cul_time_ns);
        unsigned int c = ld_acquire_sys_u32(&p.sctl->cul_count);
        if (c) atomicExch_system(&p.sctl->avg_delay_ns, tot / c);
      }
      wc->last_accepted = accept; wc->last_batch = round_batch; wc->last_loss = round_loss;
      wc->last_step = accept ? step : 0.f; wc->last_task_ns = task_ns;
      wc->rounds_done += 1ULL;
      wc->done_ctas = 0; wc->pull_ticket = 0; wc->chunk_ctr = 0; wc->loss_sum = 0.f; wc->batch_count = 0;
      wc->dbg_ns[7] = globaltimer_ns();
      __threadfence();
      wc->round = round + 1u;
    }
    return;
  }

  // ---------------------------------------------------------------- BSP: all-gather over NVLink + replicated apply
  {
    const SyncPeers* pe = p.peers;
    const int P = p.num_workers;
    const int par = static_cast<int>(round & 1u);
    const long long slot_stride = p.ld + 4;      // [ld] gradient + {loss, batch, pad, pad}
    if (tid == 0 && static_cast<int>(round) < p.warmup_updates) {
      // delay-injector calibration (sync drivers: first `warmup` iterations, SparkASGDSync.scala:157-193)
      ServerCtl* sc = p.sctl;
      sc->cul_time_ns += globaltimer_ns() - wc->t_start_ns;
      sc->cul_count += 1u;
      if (static_cast<int>(round) + 1 == p.warmup_updates) sc->avg_delay_ns = sc->cul_time_ns / sc->cul_count;
    }
    // 1. publish my gradient into slot[worker_id] of every peer's staging buffer (P2P stores)
    for (int q = 0; q < P; ++q) {
      float* dst = pe->gstage[q] + (static_cast<long long>(par) * P + p.worker_id) * slot_stride;
      for (int v = tid; v < nvec; v += kThreads) {
        float4 g = __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v));
        st_relaxed_sys_f32x4(dst + 4 * v, g);
      }
      if (tid == 0) st_relaxed_sys_f32x4(dst + p.ld, make_float4(round_loss, __uint_as_float(round_batch), 0.f, 0.f));
    }
    fence_sys();
    __syncthreads();
    if (tid < P) st_release_sys_u32(pe->flags[tid] + par * P + p.worker_id, round + 1u);
    // 2. wait for every worker's slot of this round
    if (tid < P) {
      const unsigned int* f = pe->flags[p.worker_id] + par * P + tid;
      while (ld_acquire_sys_u32(f) != round + 1u) __nanosleep(100);
    }
    __syncthreads();
    // 3. fixed-order reduction => bit-identical replicas on every rank; then the update
    const float* mine = pe->gstage[p.worker_id] + static_cast<long long>(par) * P * slot_stride;
    float tot_loss = 0.f; unsigned int tot_batch = 0;
    for (int q = 0; q < P; ++q) {
      float4 m = ld_relaxed_sys_f32x4(mine + q * slot_stride + p.ld);
      tot_loss += m.x; tot_batch += __float_as_uint(m.y);
    }
    float step = (p.algo == ALGO_SYNC_SGD) ? p.gamma * rsqrtf(static_cast<float>(round) + 1.f) : p.gamma;
    for (int v = tid; v < nvec; v += kThreads) {
      float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < P; ++q) {
        float4 g = ld_relaxed_sys_f32x4(mine + q * slot_stride + 4 * v);
        G.x += g.x; G.y += g.y; G.z += g.z; G.w += g.w;
      }
      float4 w = *reinterpret_cast<float4*>(p.w_local + 4 * v);
      const float a = step * p.inv_par_recs;
      if (p.algo == ALGO_SYNC_SGD) {
        w.x -= a * G.x; w.y -= a * G.y; w.z -= a * G.z; w.w -= a * G.w;
      } else {
        float4 ab = *reinterpret_cast<float4*>(p.abar_local + 4 * v);
        w.x -= a * G.x + step * ab.x; w.y -= a * G.y + step * ab.y;
        w.z -= a * G.z + step * ab.z; w.w -= a * G.w + step * ab.w;
        ab.x += G.x * p.inv_n; ab.y += G.y * p.inv_n; ab.z += G.z * p.inv_n; ab.w += G.w * p.inv_n;
        *reinterpret_cast<float4*>(p.abar_local + 4 * v) = ab;
      }
      *reinterpret_cast<float4*>(p.w_local + 4 * v) = w;
      if (p.worker_id == 0 && p.w_server != p.w_local) *reinterpret_cast<float4*>(p.w_server + 4 * v) = w;
      *reinterpret_cast<float4*>(p.g_local + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned long long now = globaltimer_ns();
      wc->last_accepted = 1; wc->last_batch = tot_batch; wc->last_loss = tot_loss; wc->last_step = step;
      wc->last_staleness = 0; wc->last_k_after = static_cast<long long>(round) + 1;
      wc->last_task_ns = now - wc->t_start_ns;
      wc->rounds_done += 1ULL;
      wc->done_ctas = 0; wc->pull_ticket = 0; wc->chunk_ctr = 0; wc->loss_sum = 0.f; wc->batch_count = 0;
      __threadfence();
      wc->round = round + 1u;
    }
  }
}

}  // namespace afb
