// Shared phases of the fused round kernels: gate + pull (phase 0) and the cross-GPU tail
// (grid reduce bookkeeping, staleness gate, Hogwild push-and-apply / BSP all-gather apply).
// Included by dense_stream.cu, dense_round.cu and csr_round.cu; the gradient body in between is layout specific.
//
// Reference semantics implemented here (file:line in /root/reference):
//   * submitter gating `availableWorkers >= floor(P*beta)`           SparkASGDThread.scala:230-234, 282-283
//   * ASYNCbarrier(f, STAT): arbitrary predicate / SSP via getMaxStaleness   core/rdd/RDD.scala:1050-1064, workerState.scala:74-82
//   * staleness filter, step schedule, apply                          SparkASGDThread.scala:169-197, SparkASAGAThread.scala:184-213
//   * L1 / squared-L2 updaters                                        mllib/optimization/Updater.scala:114-129, 151-158
//   * BSP rounds                                                      SparkASGDSync.scala:234-272, SparkASAGASync.scala:258-299
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

enum Proceed : int { PROCEED_STOP = 0, PROCEED_RUN = 1, PROCEED_SKIP = 2 };

AFB_DEVINL unsigned int fetch_chunk_ctr(unsigned int* ctr, int lane) {
  unsigned int c = 0;
  if (lane == 0) c = atomicAdd(ctr, 1u);
  return __shfl_sync(0xffffffffu, c, 0);
}

AFB_DEVINL bool aborted(const WorkerCtl* wc) { return *reinterpret_cast<const volatile int*>(&wc->abort) != 0; }
AFB_DEVINL void raise_abort(WorkerCtl* wc) { *reinterpret_cast<volatile int*>(&wc->abort) = 1; __threadfence(); }

// CTA-uniform view of the abort flag (call from ALL threads of the CTA; three barriers).  Per-thread reads of the flag
// could disagree while another CTA is raising it, and a divergent early return would deadlock the next CTA barrier.
template <int kThreads>
AFB_DEVINL bool cta_abort_vote(RoundShared<kThreads>& sh, const WorkerCtl* wc) {
  __syncthreads();
  if (threadIdx.x == 0) sh.flag = aborted(wc) ? 1u : 0u;
  __syncthreads();
  const bool a = sh.flag != 0u;
  __syncthreads();
  return a;
}

AFB_DEVINL float soft_threshold(float t, float shrink) {
  const float a = fmaxf(fabsf(t) - shrink, 0.f);
  return copysignf(a, t);
}

// Device-side ASYNCbarrier: the bucket rule (with the watchdog's runtime override), the SSP rule on the live status
// table, evaluated by warp 0 of the pulling CTA.  Returns once the gate is open or the run has been stopped.
AFB_DEVINL void gate_wait_warp(const RoundParams& p, WorkerCtl* wc, const int lane) {
  SpinGuard sg;
  while (true) {
    int open = 1;
    if (lane == 0) {
      const unsigned long long ck0 = ld_acquire_sys_u64(&p.sctl->clock_k);
      if (static_cast<long long>(ck0 >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) open = 2;   // stopped: let the pull decide
      else if (p.gate_threshold > 0) {
        // partial barrier with the reference's WAVE semantics (SparkASGDThread.scala:230-234: the submitter waits until
        // >= floor(P*beta) workers are free, then submits ONE job to every free worker): the first idle worker to see the
        // threshold met opens wave w+1 (CAS on ServerCtl.wave); every worker that has not run in the current wave joins it;
        // a worker that finishes mid-wave waits for the next one.  The host watchdog lowers the threshold when it
        // declares workers dead (continue with P-1).
        int thr = p.gate_threshold;
        const int dyn = ld_relaxed_sys_s32(&p.sctl->gate_dyn);
        if (dyn > 0 && dyn < thr) thr = dyn;
        const unsigned int w = ld_acquire_sys_u32(&p.sctl->wave);
        if (w != wc->wave_seen) {
          wc->wave_seen = w;                                   // a wave opened since this worker last ran: join it
        } else if (ld_relaxed_sys_s32(&p.sctl->avail_count) >= thr) {
          const unsigned int old = atom_cas_acqrel_sys_u32(&p.sctl->wave, w, w + 1u);
          wc->wave_seen = (old == w) ? w + 1u : old;           // opened it, or somebody else just did
        } else {
          open = 0;
        }
      }
    }
    open = __shfl_sync(0xffffffffu, open, 0);
    if (open == 1 && p.ssp_bound >= 0) {
      // SSP: max over the workers' last reported staleness (workerState.getMaxStaleness), one parallel sweep of the table
      long long m = 0;
      for (int q = lane; q < p.num_workers; q += 32) {
        long long sv;
        asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(sv) : "l"(&p.status[q].staleness) : "memory");
        m = sv > m ? sv : m;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { long long t = __shfl_xor_sync(0xffffffffu, m, o); m = t > m ? t : m; }
      if (m > static_cast<long long>(p.ssp_bound)) open = 0;
    }
    if (open) return;
    __nanosleep(200);
    if (sg.expired(p.spin_limit_ns)) { if (lane == 0) raise_abort(wc); return; }
    // a worker waiting at the gate is alive: keep its heartbeat fresh so the watchdog can tell it from a dead one
    if (lane == 0 && (sg.n & 1023u) == 0u) (p.status + p.worker_id)->heartbeat_ns = globaltimer_ns();
  }
}

// Phase 0 is split in two so that its NVLink round trip overlaps useful work:
//   round_begin_issue()    - puller election (ticket), gate, and the pull's loads are ISSUED
//                            (clock word + first chunk of w stay in registers, latency in flight);
//   ... the caller hashes its first row chunk / prefetches its first rows here ...
//   round_begin_complete() - the puller consumes the loads, publishes w_local / ts / seed and releases
//                            the CTA-visible flag; every other CTA waits for the flag.
// round_begin_complete returns PROCEED_RUN, PROCEED_STOP (k >= max_iters, stop flag or abort: the kernel must return,
// bookkeeping for the next launch is already done) or PROCEED_SKIP (this worker's enable-mask bit is clear: the round is
// a no-op, like a non-selected partition's empty task in the reference; a persistent kernel moves on to its next round).
struct RoundPull {
  float4 wv;
  unsigned long long ck;
  unsigned long long mask_word;
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
  pl.mask_word = ~0ULL;
  pl.puller = false;
  if (p.algo == ALGO_GRAD_ONLY) return;
  if (tid == 0) sh.flag = atomicAdd(&wc->pull_ticket, 1u);
  __syncthreads();
  pl.puller = (sh.flag == 0);
  if (!pl.puller) return;
  // This CTA is the puller: gate, then ONE NVLink round trip — the logical-clock read and the
  // vectorised fetch of w are issued together (ts may be at most in-flight-updates older than w).
  const bool gated = is_async && (p.gate_threshold > 0 || p.ssp_bound >= 0);
  if (tid == 0) wc->dbg_ns[0] = globaltimer_ns();
  if (gated) {
    if (tid < 32) gate_wait_warp(p, wc, tid);
    __syncthreads();                                             // nobody reads w before the gate opens
  }
  if (tid == 0) wc->dbg_ns[1] = globaltimer_ns();
  if (is_async) {
    if (tid < nvec) pl.wv = ld_relaxed_sys_f32x4(p.w_server + 4 * tid);          // K10 (first chunk), in flight
    if (tid == 0) {
      pl.ck = ld_acquire_sys_u64(&p.sctl->clock_k);
      if (p.use_mask) pl.mask_word = ld_acquire_sys_u64(&p.sctl->enable_mask[(p.worker_id >> 6) & 3]);
    }
  }
}

template <int kThreads>
AFB_DEVINL int round_begin_complete(const RoundParams& p, RoundShared<kThreads>& sh, const unsigned int round, const RoundPull& pl) {
  const int tid = threadIdx.x;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  const bool is_async = (p.algo == ALGO_ASGD || p.algo == ALGO_ASAGA);
  if (p.algo == ALGO_GRAD_ONLY) return PROCEED_RUN;
  if (pl.puller) {
    if (tid == 0) {
      int proceed = PROCEED_RUN;
      const unsigned long long ck = pl.ck;
      if (is_async) {
        if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop) || aborted(wc)) proceed = PROCEED_STOP;
        else if (p.use_mask && !((pl.mask_word >> (p.worker_id & 63)) & 1ULL)) proceed = PROCEED_SKIP;
        if (proceed == PROCEED_RUN) {
          red_add_release_sys_s32(&p.sctl->avail_count, -1);      // busy from submit until the result lands
          WorkerSlot* me = p.status + p.worker_id;
          unsigned long long now = globaltimer_ns();
          if (wc->last_finish_ns) wc->wait_ns += now - wc->last_finish_ns;
          me->available = 0;                                      // posted stores only (no remote reads)
          me->busy_since_ns = now;
          me->heartbeat_ns = now;
          me->wait_ns = wc->wait_ns;
        } else if (proceed == PROCEED_SKIP) {
          (p.status + p.worker_id)->heartbeat_ns = globaltimer_ns();
        }
      } else {
        (p.status + p.worker_id)->heartbeat_ns = globaltimer_ns();      // BSP workers report liveness to the server's table too
        if (*reinterpret_cast<volatile int*>(&p.sctl->stop) || aborted(wc))
          proceed = PROCEED_STOP;                                 // BSP: the host asked this rank to stop (a peer died)
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
    if (wc->proceed == PROCEED_RUN) {
      if (tid == 0 && p.delay_mode != 0) {
        // A18 straggler injector: sleep a multiple of the calibrated mean task time.
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
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_gpu_u32(&wc->pull_flag, round + 1u);
  } else {
    if (tid == 0) {
      SpinGuard sg;
      while (ld_acquire_gpu_u32(&wc->pull_flag) != round + 1u) {
        __nanosleep(32);
        if (sg.expired(p.spin_limit_ns) || aborted(wc)) { raise_abort(wc); break; }
      }
    }
    __syncthreads();
  }
  // CTA-uniform decision (the abort flag may be raised concurrently by another CTA)
  if (tid == 0) {
    int pr = *reinterpret_cast<volatile int*>(&wc->proceed);
    if (aborted(wc)) pr = PROCEED_STOP;
    sh.flag = static_cast<unsigned int>(pr);
  }
  __syncthreads();
  const int proceed = static_cast<int>(sh.flag);
  __syncthreads();
  if (proceed != PROCEED_RUN) {
    // stopped / skipped: still run the last-CTA bookkeeping so counters reset for the next round or launch
    __syncthreads();
    if (tid == 0) {
      unsigned int t = atomicAdd(&wc->done_ctas, 1u);
      if (t == gridDim.x - 1) {
        wc->done_ctas = 0; wc->pull_ticket = 0; wc->chunk_ctr = 0;
        wc->last_accepted = (proceed == PROCEED_SKIP) ? -2 : -1;
        if (proceed == PROCEED_SKIP) wc->skipped_rounds += 1u;
        __threadfence();
        st_release_gpu_u32(&wc->round, round + 1u);
      }
    }
  }
  return proceed;
}

template <int kThreads>
AFB_DEVINL int round_begin(const RoundParams& p, RoundShared<kThreads>& sh, const unsigned int round) {
  RoundPull pl;
  round_begin_issue<kThreads>(p, sh, round, pl);
  return round_begin_complete<kThreads>(p, sh, round, pl);
}

// The logical-clock ticket (K9 part 1): one or two NVLink atomics decide staleness, acceptance, k and the step size.
struct Ticket {
  int accept;
  float step;
  unsigned int clock_old;
  long long staleness;
  long long k_after;
};

AFB_DEVINL Ticket take_ticket(const RoundParams& p, const unsigned long long ts, const unsigned int round) {
  Ticket t;
  unsigned long long old, k_old;
  if (p.taw >= (1LL << 31)) {
    // never stale: one NVLink round trip ticks clock and k together
    old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, (1ULL << 32) | 1ULL);
    k_old = old >> 32;
    t.staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
    t.accept = (static_cast<long long>(k_old) < p.max_iters) ? 1 : 0;
  } else {
    old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL);
    t.staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
    // ASGD compares the clock-based staleness (SparkASGDThread.scala:169); ASAGA compares
    // k - ts (SparkASAGAThread.scala:184, quirk B.6) — both are available here.
    long long cmp = t.staleness;
    if (p.algo == ALGO_ASAGA) cmp = static_cast<long long>(old >> 32) - static_cast<long long>(ts);
    t.accept = (cmp <= p.taw) ? 1 : 0;
    k_old = old >> 32;
    if (t.accept) {
      unsigned long long o2 = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL << 32);
      k_old = o2 >> 32;
      if (static_cast<long long>(k_old) >= p.max_iters) t.accept = 0;
    }
  }
  const float gamma = eff_gamma(p, round);
  if (p.algo == ALGO_ASGD) t.step = gamma * rsqrtf(static_cast<float>(k_old / static_cast<unsigned long long>(p.num_workers)) + 1.f);
  else t.step = gamma;
  t.clock_old = static_cast<unsigned int>(old & 0xffffffffULL);
  t.k_after = static_cast<long long>(k_old) + (t.accept ? 1 : 0);
  return t;
}

// K9 (part 2) + K5/K6/K8: the Hogwild push-and-apply of one accepted gradient, executed by a whole CTA.
//   ASGD : w += -eta (g/parRecs [+ lambda w_pulled])                       SparkASGDThread.scala:185-189 (+ SquaredL2Updater)
//   ASAGA: w += -gamma (g/parRecs + abar [+ lambda w_pulled]); abar += g/N  SparkASAGAThread.scala:204-213
//   L1   : Hogwild prox — delta = soft(w_now - eta(...), eta lambda) - w_now, where w_now is re-read from the server
// `v0` / `vstride`: the float4 indices this thread covers (one CTA: tid / kThreads; a whole grid: global tid / grid threads).
AFB_DEVINL void hogwild_apply_range(const RoundParams& p, const float* g_loc, const float* w_pulled, const float step,
                                    const long long v0, const long long vstride) {
  const long long nvec = p.ld >> 2;
  const float a = -step * p.inv_par_recs;
  const float l2 = (p.reg_type == 1) ? -step * p.reg_param : 0.f;
  const float shrink = step * p.reg_param;
  for (long long v = v0; v < nvec; v += vstride) {
    const float4 g = __ldcg(reinterpret_cast<const float4*>(g_loc + 4 * v));
    float4 dl = make_float4(a * g.x, a * g.y, a * g.z, a * g.w);
    if (p.algo == ALGO_ASAGA) {
      // abar BEFORE its update
      const float4 ab = ld_relaxed_sys_f32x4(p.abar_server + 4 * v);
      dl.x -= step * ab.x; dl.y -= step * ab.y; dl.z -= step * ab.z; dl.w -= step * ab.w;
      red_add_sys_f32x4(p.abar_server + 4 * v, make_float4(g.x * p.inv_n, g.y * p.inv_n, g.z * p.inv_n, g.w * p.inv_n));
    }
    if (p.reg_type == 1) {
      const float4 wp = __ldcg(reinterpret_cast<const float4*>(w_pulled + 4 * v));
      dl.x = fmaf(l2, wp.x, dl.x); dl.y = fmaf(l2, wp.y, dl.y); dl.z = fmaf(l2, wp.z, dl.z); dl.w = fmaf(l2, wp.w, dl.w);
    } else if (p.reg_type == 2) {
      const float4 wn = ld_relaxed_sys_f32x4(p.w_server + 4 * v);
      dl.x = soft_threshold(wn.x + dl.x, shrink) - wn.x; dl.y = soft_threshold(wn.y + dl.y, shrink) - wn.y;
      dl.z = soft_threshold(wn.z + dl.z, shrink) - wn.z; dl.w = soft_threshold(wn.w + dl.w, shrink) - wn.w;
    }
    red_add_sys_f32x4(p.w_server + 4 * v, dl);
  }
}

template <int kThreads>
AFB_DEVINL void hogwild_apply(const RoundParams& p, const float* g_loc, const float* w_pulled, const float step) {
  hogwild_apply_range(p, g_loc, w_pulled, step, threadIdx.x, kThreads);
}

// Sum the `g_copies` replicas of the worker-local gradient into replica 0 and re-zero the others (stream kernels spread
// the per-CTA merges over replicas to cut same-address L2 reduction traffic).  Whole CTA; ends with a CTA barrier.
template <int kThreads>
AFB_DEVINL void merge_g_copies(const RoundParams& p, float* g) {
  if (p.g_copies <= 1) return;
  const int nvec = static_cast<int>(p.ld >> 2);
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    float4 t = __ldcg(reinterpret_cast<const float4*>(g + 4 * v));
    for (int c = 1; c < p.g_copies; ++c) {
      float* gc = g + static_cast<long long>(c) * p.ld + 4 * v;
      const float4 a = __ldcg(reinterpret_cast<const float4*>(gc));
      t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
      *reinterpret_cast<float4*>(gc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4*>(g + 4 * v) = t;
  }
  __threadfence();
  __syncthreads();
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
    const Ticket t = take_ticket(p, wc->ts, round);
    wc->tk_accept = t.accept; wc->tk_step = t.step; wc->tk_clock_old = t.clock_old;
    wc->tk_staleness = t.staleness; wc->tk_k_after = t.k_after;
    __threadfence();
    st_release_gpu_u32(&wc->tk_flag, round + 1u);
  }
  if (!is_last_cta) return;

  // ------------------------------------------------------------------ phase 3 (last CTA): push-and-apply
  if (tid == 0) wc->dbg_ns[3] = globaltimer_ns();
  __threadfence();
  merge_g_copies<kThreads>(p, p.g_local);
  const float round_loss = *reinterpret_cast<volatile float*>(&wc->loss_sum);
  const unsigned int round_batch = *reinterpret_cast<volatile unsigned int*>(&wc->batch_count);

  if (p.algo == ALGO_GRAD_ONLY) {
    if (tid == 0) {
      wc->last_loss = round_loss; wc->last_batch = round_batch;
      wc->done_ctas = 0; wc->chunk_ctr = 0; wc->loss_sum = 0.f; wc->batch_count = 0;
      wc->rounds_done += 1ULL;
      __threadfence();
      st_release_gpu_u32(&wc->round, round + 1u);
    }
    return;   // g_local is the output; the caller consumes and re-zeroes it
  }

  if (is_async) {
    // K9 (part 2) + K5/K6: Hogwild reductions with the decision taken by the ticket holder.
    if (tid == 0) {
      SpinGuard sg;
      while (ld_acquire_gpu_u32(&wc->tk_flag) != round + 1u) {
        __nanosleep(32);
        if (sg.expired(p.spin_limit_ns)) { raise_abort(wc); break; }
      }
      wc->dbg_ns[4] = globaltimer_ns();
      sh.flag = static_cast<unsigned int>(*reinterpret_cast<volatile int*>(&wc->tk_accept));
      sh.bcast[0] = *reinterpret_cast<volatile float*>(&wc->tk_step);
      sh.bcast[1] = __uint_as_float(*reinterpret_cast<volatile unsigned int*>(&wc->tk_clock_old));
      wc->last_staleness = *reinterpret_cast<volatile long long*>(&wc->tk_staleness);
      wc->last_k_after = *reinterpret_cast<volatile long long*>(&wc->tk_k_after);
    }
    __syncthreads();
    const int accept = aborted(wc) ? 0 : static_cast<int>(sh.flag);
    const float step = sh.bcast[0];
    if (accept) hogwild_apply<kThreads>(p, p.g_local, p.w_local, step);
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
      st_release_gpu_u32(&wc->round, round + 1u);
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
    // 1. publish my gradient into slot[worker_id] of every peer's staging buffer.  With an NVLS multicast alias of the
    //    symmetric block this is ONE store stream that the switch replicates into all P ranks (the TorrentBroadcast
    //    replacement of SURVEY §2.6 on the BSP path); otherwise P peer-store loops.
    const long long my_slot = (static_cast<long long>(par) * P + p.worker_id) * slot_stride;
    if (pe->mc_gstage != nullptr) {
      float* dst = pe->mc_gstage + my_slot;
      for (int v = tid; v < nvec; v += kThreads)
        multimem_st_f32x4(dst + 4 * v, __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v)));
      if (tid == 0) multimem_st_f32x4(dst + p.ld, make_float4(round_loss, __uint_as_float(round_batch), 0.f, 0.f));
      fence_sys();
      __syncthreads();
      if (tid == 0) multimem_st_release_u32(pe->mc_flags + par * P + p.worker_id, round + 1u);
    } else {
      for (int q = 0; q < P; ++q) {
        float* dst = pe->gstage[q] + my_slot;
        for (int v = tid; v < nvec; v += kThreads) {
          float4 g = __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v));
          st_relaxed_sys_f32x4(dst + 4 * v, g);
        }
        if (tid == 0) st_relaxed_sys_f32x4(dst + p.ld, make_float4(round_loss, __uint_as_float(round_batch), 0.f, 0.f));
      }
      fence_sys();
      __syncthreads();
      if (tid < P) st_release_sys_u32(pe->flags[tid] + par * P + p.worker_id, round + 1u);
    }
    // 2. wait for every worker's slot of this round (bounded; a stop request from the host ends the wait: a dead peer
    //    aborts a synchronous run cleanly instead of hanging it)
    if (tid < P) {
      const unsigned int* f = pe->flags[p.worker_id] + par * P + tid;
      SpinGuard sg;
      while (ld_acquire_sys_u32(f) != round + 1u) {
        __nanosleep(100);
        if (sg.expired(p.spin_limit_ns) || ((sg.n & 63u) == 1u && *reinterpret_cast<volatile int*>(&p.sctl->stop))) { raise_abort(wc); break; }
        if (tid == 0 && (sg.n & 1023u) == 0u) (p.status + p.worker_id)->heartbeat_ns = globaltimer_ns();    // waiting, not dead
      }
    }
    if (cta_abort_vote<kThreads>(sh, wc)) {
      if (tid == 0) {
        wc->last_accepted = -1;
        wc->done_ctas = 0; wc->pull_ticket = 0; wc->chunk_ctr = 0; wc->loss_sum = 0.f; wc->batch_count = 0;
        __threadfence();
        st_release_gpu_u32(&wc->round, round + 1u);
      }
      return;
    }
    // 3. fixed-order reduction => bit-identical replicas on every rank; then the update (K5/K6 + K8)
    const float* mine = pe->gstage[p.worker_id] + static_cast<long long>(par) * P * slot_stride;
    float tot_loss = 0.f; unsigned int tot_batch = 0;
    for (int q = 0; q < P; ++q) {
      float4 m = ld_relaxed_sys_f32x4(mine + q * slot_stride + p.ld);
      tot_loss += m.x; tot_batch += __float_as_uint(m.y);
    }
    const float gamma = eff_gamma(p, round);
    const float step = (p.algo == ALGO_SYNC_SGD) ? gamma * rsqrtf(static_cast<float>(round) + 1.f) : gamma;
    const float a = step * p.inv_par_recs;
    const float keep = (p.reg_type == 1) ? (1.f - step * p.reg_param) : 1.f;      // SquaredL2Updater: w *= (1 - eta*lambda)
    const float shrink = step * p.reg_param;
    for (int v = tid; v < nvec; v += kThreads) {
      float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < P; ++q) {
        float4 g = ld_relaxed_sys_f32x4(mine + q * slot_stride + 4 * v);
        G.x += g.x; G.y += g.y; G.z += g.z; G.w += g.w;
      }
      float4 w = *reinterpret_cast<float4*>(p.w_local + 4 * v);
      if (p.reg_type == 1) { w.x *= keep; w.y *= keep; w.z *= keep; w.w *= keep; }
      w.x -= a * G.x; w.y -= a * G.y; w.z -= a * G.z; w.w -= a * G.w;
      if (p.algo == ALGO_SYNC_SAGA) {
        float4 ab = *reinterpret_cast<float4*>(p.abar_local + 4 * v);
        w.x -= step * ab.x; w.y -= step * ab.y; w.z -= step * ab.z; w.w -= step * ab.w;
        ab.x += G.x * p.inv_n; ab.y += G.y * p.inv_n; ab.z += G.z * p.inv_n; ab.w += G.w * p.inv_n;
        *reinterpret_cast<float4*>(p.abar_local + 4 * v) = ab;
      }
      if (p.reg_type == 2) {                                                       // L1Updater: soft threshold by eta*lambda
        w.x = soft_threshold(w.x, shrink); w.y = soft_threshold(w.y, shrink);
        w.z = soft_threshold(w.z, shrink); w.w = soft_threshold(w.w, shrink);
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
      st_release_gpu_u32(&wc->round, round + 1u);
    }
  }
}

}  // namespace afb
