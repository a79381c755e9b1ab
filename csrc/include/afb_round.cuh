// Parameter blocks for the fused round kernels (mirrored 1:1 by ctypes Structures in
// asyncframework_b200/ops/_signatures.py — keep field order and types in sync).
#pragma once
#include "afb_common.cuh"

namespace afb {

enum Algo : int {
  ALGO_GRAD_ONLY = 0,   // g_local/loss/count only (used by GradientDescent, NCCL backend, tests)
  ALGO_ASGD = 1,        // async: push-and-apply  w += -gamma/sqrt(floor(k/P)+1) * g/parRecs
  ALGO_ASAGA = 2,       // async SAGA: w -= gamma*(g/parRecs + abar); abar += g/N
  ALGO_SYNC_SGD = 3,    // BSP: all-gather of g over NVLink + replicated apply
  ALGO_SYNC_SAGA = 4,
};

struct alignas(16) WorkerCtl {      // worker-local control block (device memory on the worker's GPU)
  unsigned int chunk_ctr;           // dynamic row-chunk scheduler
  unsigned int done_ctas;           // last-CTA election
  unsigned int pull_ticket;         // first-CTA election for the pull
  unsigned int pull_flag;           // == round+1 once w_local / ts / seed_key are valid
  unsigned int round;               // rounds executed by this worker (incremented by the kernel)
  int proceed;                      // set by the pulling CTA: 1 = run the round, 0 = stopped (k >= iters)
  unsigned int batch_count;         // rows sampled this round
  float loss_sum;                   // Σ loss this round
  unsigned long long ts;            // logical clock read at pull
  unsigned long long k_read;        // accepted-update counter read at pull
  unsigned long long seed_key;      // mix64(seed_eff) for this round
  unsigned long long t_start_ns;
  // results of the last finished round (read by the host without touching the server)
  int last_accepted;
  unsigned int last_batch;
  float last_loss;
  float last_step;
  long long last_staleness;
  long long last_k_after;
  unsigned long long last_task_ns;
  unsigned long long rounds_done;
  // worker-local running statistics (published to the server's status table with posted stores,
  // so the hot path never does a read-modify-write over NVLink)
  unsigned long long wait_ns;       // Σ (round start − previous finish)            (A19)
  unsigned long long last_finish_ns;
  float avg_task_ms;
  float pad_f;
  unsigned long long pad_u;
  // phase timestamps of the last round (globaltimer ns): [0] kernel entry (puller), [1] gate passed,
  // [2] w + clock landed, [3] last CTA elected (compute done), [4] clock ticket returned,
  // [5] reductions issued, [6] fence done, [7] published
  unsigned long long dbg_ns[8];
  // early logical-clock ticket: taken by the FIRST CTA that finishes its gradient work so the NVLink
  // round trip of the atomic overlaps the drain of the compute phase; consumed by the last CTA
  unsigned int tk_flag;             // == round+1 once the fields below are valid
  int tk_accept;
  float tk_step;
  unsigned int tk_clock_old;
  long long tk_staleness;
  long long tk_k_after;
  // ---- robustness -----------------------------------------------------------------------------------------------------
  int abort;                        // != 0: a bounded spin expired (SpinGuard) — results of this launch are invalid
  unsigned int skipped_rounds;      // rounds this worker sat out because its enable-mask bit was clear (ASYNCbarrier)
  unsigned int wave_seen;           // last bucket-gate wave this worker joined (ServerCtl.wave)
  unsigned int pad_tail;
};

// Per-round descriptor of the end-to-end path (FusedEngine.step): written on the host into pinned memory, copied H2D
// every step, consumed by the round kernel (slot = round % desc_slots).  `seq` = round + 1 marks the slot valid.
struct alignas(16) RoundDesc {
  float gamma;                      // step size of this round
  int thr24;                        // Bernoulli threshold (batch rate) of this round
  float delay_coeff;                // controlled-delay intensity of this round
  unsigned int seq;
};

// Double-buffered per-round state of the PIPELINED persistent rounds (dense_stream.cu): round r uses b[r & 1], so round
// r+1's pull and gradient can run while round r's NVLink tail is still in flight.
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

struct alignas(16) SyncPeers {      // BSP all-gather wiring (device array, one per worker)
  float* gstage[16];                // gstage[q] = peer q's staging buffer [2][P][ld+4] (double-buffered by round parity)
  unsigned int* flags[16];          // flags[q]  = peer q's arrival flags [2][P]
  float* mc_gstage;                 // NVLS multicast alias of the staging buffer (or null)
  unsigned int* mc_flags;           // NVLS multicast alias of the flags (or null)
};

struct RoundParams {
  // ---- local shard ---------------------------------------------------------------------
  const float* X;                   // dense: [n_rows, ld] row-major (null for CSR shards)
  const long long* indptr;          // CSR: [n_rows+1]
  const int* indices;               // CSR: [nnz] 0-based feature ids, ascending within a row
  const float* values;              // CSR: [nnz]
  const float* y;                   // [n_rows]
  const long long* gidx;            // [n_rows] global row index, or null => row_base + i
  float* hist;                      // SAGA: last residual per local row (or null)
  float* hist_prev;                 // SAGA rollback mode: previous value (or null => write-through)
  int* hist_tag;                    // SAGA rollback mode: round tag of the last write
  long long row_base;
  long long ld;
  // ---- server state (local or NVLink peer pointers) -----------------------------------------
  float* w_server;
  float* abar_server;
  ServerCtl* sctl;
  WorkerSlot* status;
  ResultRec* ring;
  // ---- worker-local scratch ------------------------------------------------------------------
  float* w_local;                   // [ld] pulled copy of w (sync modes: the local replica)
  float* abar_local;                // sync SAGA replica
  float* g_local;                   // [ld] zero on entry, zero on exit
  WorkerCtl* wctl;
  const SyncPeers* peers;           // sync modes only
  // ---- scalars ---------------------------------------------------------------------------------
  int n_rows;
  int d;
  int algo;
  int ring_size;
  int num_workers;                  // P
  int worker_id;
  int thr24;                        // Bernoulli threshold (2^24 => every row)
  int gate_threshold;               // floor(P*beta); 0 => no gating
  long long seed;
  long long taw;                    // max staleness (>= 2^31 => never stale: single-atomic fast path)
  long long max_iters;              // stop once k >= max_iters
  float gamma;
  float inv_par_recs;               // 1 / (b*N/P)  (ASGD/ASAGA) or 1 / (b*N) (sync)
  float inv_n;                      // 1 / N
  float delay_coeff;                // controlled delay: sleep coeff * avg_delay  (0 => none)
  int delay_mode;                   // 0 none, 1 controlled (this worker sleeps), 2 production normal, 3 production long tail
  int warmup_updates;               // calibration window (100*P async)
  int seed_mode;                    // 0: seed + k_read + 1 ; 1: seed + round + 1 ; 2: seed (fixed)
  int loss_id;                      // afb::LossId
  int round_host;                   // PDL: the round index counted by the host (== wctl->round at kernel start)
  int pdl;                          // 1: launched with programmatic stream serialization (see dense_round.cu)
  // ---- regularisation (K8 fused into the apply; Updater.scala:114-158) --------------------------------------------------
  float reg_param;                  // lambda
  int reg_type;                     // 0 none, 1 squared-L2 (w <- (1 - eta*lambda) w - eta g), 2 L1 (soft threshold by eta*lambda)
  // ---- gating beyond the bucket rule (ASYNCbarrier on the device; RDD.scala:1050-1064, workerState.scala:74-82) ---------
  int ssp_bound;                    // >= 0: start a round only while max_q staleness(q) <= bound (SSP); < 0: off
  int use_mask;                     // 1: honour ServerCtl.enable_mask (host-evaluated predicate): a masked worker skips the round
  // ---- end-to-end descriptor ring --------------------------------------------------------------------------------------
  const RoundDesc* desc;            // null: scalars above are used; else slot (round % desc_slots) overrides gamma / thr24 / delay_coeff
  int desc_slots;
  int pull_ahead_chunks;            // pipelined rounds: pull round r+1's parameters when this many chunks of round r remain
  unsigned long long spin_limit_ns; // bound of every in-kernel spin (0 = unbounded)
  int tma_mask;                     // stream kernel: bit s = ring stage s is filled by a TMA bulk copy (else by per-lane cp.async)
  int g_copies;                     // replicas of g_local ([g_copies][ld]; >= 1) the CTAs of a grid spread their merges over
};

// The descriptor of `round`, or null when the ring holds none for it (slot not written / already reused): the launch
// scalars apply then, so rounds enqueued without a fresh descriptor (run_rounds) simply ignore the ring.
// NB: the descriptor words are read with opaque `ld.relaxed.gpu.global` asm and the launch scalar is read FIRST into a
// register: written as `d ? d->x : p.x` nvcc merges the two loads into one `ld.global` through a selected POINTER, and the
// generic address of a __grid_constant__ parameter is not a global address (compute-sanitizer: invalid global read at
// offsetof(RoundParams, thr24)).
AFB_DEVINL unsigned int ld_desc_u32(const void* ptr) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
AFB_DEVINL const RoundDesc* round_desc(const RoundParams& p, unsigned int round) {
  const RoundDesc* base = p.desc;
  const int slots = p.desc_slots;
  if (base == nullptr || slots <= 0) return nullptr;
  const RoundDesc* d = base + (round % static_cast<unsigned int>(slots));
  return (ld_desc_u32(&d->seq) == round + 1u) ? d : nullptr;
}
AFB_DEVINL int eff_thr24(const RoundParams& p, unsigned int round) {
  int v = p.thr24;
  const RoundDesc* d = round_desc(p, round);
  if (d != nullptr) v = static_cast<int>(ld_desc_u32(&d->thr24));
  return v;
}
AFB_DEVINL float eff_gamma(const RoundParams& p, unsigned int round) {
  float v = p.gamma;
  const RoundDesc* d = round_desc(p, round);
  if (d != nullptr) v = __uint_as_float(ld_desc_u32(&d->gamma));
  return v;
}
AFB_DEVINL float eff_delay_coeff(const RoundParams& p, unsigned int round) {
  float v = p.delay_coeff;
  const RoundDesc* d = round_desc(p, round);
  if (d != nullptr) v = __uint_as_float(ld_desc_u32(&d->delay_coeff));
  return v;
}

}  // namespace afb
