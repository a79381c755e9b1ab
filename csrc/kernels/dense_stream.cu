// The default dense round kernel (sm_100a): a PERSISTENT worker loop whose row traffic moves with TMA bulk copies.
//
//   one launch  = `rounds` consecutive rounds of one logical worker (the executor task loop of the reference,
//                 core/executor/Executor.scala:290-560, without a host in it: every piece of per-round protocol state lives in
//                 device memory, rounds are separated by flags, not by kernel boundaries);
//   one round   = pull(w) -> Bernoulli sample -> x.w -> loss coefficient -> g += c.x -> (SAGA history) -> CTA/grid reduce
//                 -> logical-clock ticket -> Hogwild push-and-apply over NVLink (or the BSP all-gather + replicated apply).
//
// Data path (K1-K4, K7, K11): every warp owns a private ring of kDepth shared-memory stages and keeps kDepth sampled rows
// in flight with `cp.async.bulk.shared::cluster.global` (SASS: UBLKCP) completing on per-stage mbarriers (SYNCS); lane 0
// issues, all lanes consume with LDS.128.  The parameter vector lives in registers (the pull lands once per round), so a
// row costs one TMA write + one LDS read of shared-memory bandwidth and zero registers while it is in flight: 100 KB of rows
// are in flight per CTA at kDepth = 4, ld = 784, independent of register pressure.  The CTA-level gradient reduce reuses
// the (then idle) ring stages: warps dump their register accumulators, one pass sums the 8 slices, `red.global.add.v4`
// merges into the worker's g — no shared-memory atomics (sm_100a has no native fp32 shared atomic add: ATOMS.CAST.SPIN).
//
// Two round loops share that data path:
//   * serial rounds (template PIPE = false): the protocol of afb_round_impl.cuh, round r+1 starts when round r is
//     published; the sampling hash of the first chunk and the first kDepth row fetches are issued BEFORE that wait, so
//     they overlap the previous round's NVLink tail.  Every algorithm / gate / delay / regulariser.
//   * pipelined rounds (PIPE = true; ASGD without a bucket/SSP gate): per-round state double-buffered by round parity
//     (PipeState); the last CTA to finish round r runs its NVLink tail (ticket -> red.add -> fence -> publish) while every
//     other CTA is already streaming round r+1, and the parameters of round r+1 are pulled AHEAD — by the warp that draws
//     one of the last `pull_ahead_chunks` chunk ids of round r — so no CTA ever waits for an NVLink round trip.  The
//     worker computes round r+1 without its own round-r update (one own update in flight; the logical clock does not
//     count it), which is within the asynchronous model's semantics (staleness is measured on the server's clock).
#include <algorithm>
#include "afb_dense_rows.cuh"
#include "afb_tc.cuh"

namespace afb {

constexpr int kStreamWarps = 8;
constexpr int kStreamThreads = 32 * kStreamWarps;
constexpr int kStreamMetaBytes = 16;       // {int row; pad} behind the row

AFB_DEVINL void stream_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// packed fp32 pairs: sm_100a executes two fp32 FMAs per instruction (fma.rn.f32x2 -> FFMA2), which halves the FMA-pipe
// instructions of the dot and of the axpy (the kernel is issue-bound, not FLOP-bound: ncu, profiles/ncu_stream_*)
typedef unsigned long long f32x2;
AFB_DEVINL f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
AFB_DEVINL void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
AFB_DEVINL f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

struct WarpRing {
  uint8_t* base;            // this warp's kDepth stages (generic pointer: used by the once-per-round reduce)
  uint32_t base_s;          // ... the same as a shared-window address: the per-row path uses 32-bit shared addresses only
  uint32_t bars_s;          // this warp's kDepth mbarriers (shared-window address)
  uint32_t row_bytes;
  uint32_t stage_bytes;
};

AFB_DEVINL uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
AFB_DEVINL uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
AFB_DEVINL void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
AFB_DEVINL void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

// Two asynchronous copy engines fill a warp's ring, stage by stage (RoundParams.tma_mask):
//   * TMA stages: lane 0 arms the stage's mbarrier (expect_tx) and launches ONE `cp.async.bulk` for the row.  The TMA unit
//     serves ~64 cycles per operation per SM on top of the bytes, so 3 KB rows alone cap at ~5.2 TB/s (fp32) — measured,
//     profiles/README.md; a second 16-byte copy for the label cost as much as half a row;
//   * LSU stages: every lane issues its 16-byte `cp.async.cg` chunks of the row (LDGSTS, no register staging) and a
//     `cp.async.mbarrier.arrive.noinc` on the same mbarrier (initialised with 32 arrivals): same completion mechanism,
//     different engine.
// Alternating the two keeps both engines busy.  The row id travels with the stage (plain shared store, published to the
// warp by the __syncwarp of the consume loop), so a warp keeps NO per-stage state in registers and the consume loop is a
// single un-unrolled body.  Label / history scalars are fetched with plain loads one row AHEAD of their use (stream_rows).
template <int NV>
AFB_DEVINL void ring_issue(const WarpRing& rg, const RoundParams& p, const int row, const uint32_t st_s, const uint32_t bar_s,
                           const bool tma, const int lane, const int nchunks) {
  const uint8_t* src = reinterpret_cast<const uint8_t*>(p.X) + static_cast<long long>(row) * rg.row_bytes;
  if (lane == 0) sts_u32(st_s + rg.row_bytes, static_cast<uint32_t>(row));
  if (tma) {
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(rg.row_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(st_s), "l"(src), "r"(rg.row_bytes), "r"(bar_s) : "memory");
    }
  } else {
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int v = lane + 32 * c;
      if (c < NV - 1 || v < nchunks)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_s + 16u * v), "l"(src + 16 * v) : "memory");
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_s) : "memory");
  }
}

// Chunk ids are requested one ahead with an atomic whose RESULT IS NOT TOUCHED until the id is needed: the `__shfl_sync`
// that broadcasts lane 0's ticket is deferred to the moment the current chunk runs dry, so the ~1 us L2 round trip of the
// atomic overlaps row processing (the eager shuffle of afb_dense_rows.cuh::fetch_chunk showed up as 11 % of all stall
// samples in the first ncu capture).  `RowScheduler.next_chunk` holds lane 0's raw ticket, or 0xffffffff for "none requested".
AFB_DEVINL void stream_request_chunk(RowScheduler& s, unsigned int* ctr, const int lane) {
  unsigned int t = 0;
  if (lane == 0) t = atomicAdd(ctr, 1u);
  s.next_chunk = t;                                // valid in lane 0 only; broadcast on demand
}
// next selected row of this warp, -1 once the pool is exhausted; `cur_chunk` = id of the chunk the row comes from
AFB_DEVINL int stream_next_row(RowScheduler& s, const RoundParams& p, unsigned int* ctr, const int lane, const unsigned long long key,
                               unsigned int& cur_chunk, bool& exhausted) {
  while (true) {
    if (s.m0) { int b = __ffs(s.m0) - 1; s.m0 &= s.m0 - 1; return s.base + b; }
    if (s.m1) { int b = __ffs(s.m1) - 1; s.m1 &= s.m1 - 1; return s.base + 32 + b; }
    if (exhausted) return -1;
    const unsigned int c = __shfl_sync(0xffffffffu, s.next_chunk, 0) + s.static_chunks;
    if (c >= static_cast<unsigned int>(s.num_chunks)) { exhausted = true; return -1; }
    stream_request_chunk(s, ctr, lane);            // one ahead
    cur_chunk = c;
    load_chunk_masks(s, p, c, lane, key);
  }
}

AFB_DEVINL int pop_mask_row(RowScheduler& s) {
  if (s.m0) { int b = __ffs(s.m0) - 1; s.m0 &= s.m0 - 1; return s.base + b; }
  if (s.m1) { int b = __ffs(s.m1) - 1; s.m1 &= s.m1 - 1; return s.base + 32 + b; }
  return -1;
}

// A row is consumed in 16-byte chunks: lane l owns chunks {l + 32 c}.  fp32 rows: 4 elements per chunk; bf16 rows
// (row STORAGE only — labels, parameters, gradient and all arithmetic stay fp32): 8 elements per chunk, half the HBM bytes.
template <bool BF16> struct RowElems { static constexpr int kE = BF16 ? 8 : 4; static constexpr int kP = kE / 2; };

// one 16-byte chunk -> kP packed fp32 pairs
template <bool BF16>
AFB_DEVINL void unpack_chunk(const uint4 raw, f32x2 (&x)[RowElems<BF16>::kP]) {
  if constexpr (BF16) {
    x[0] = pack2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
    x[1] = pack2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
    x[2] = pack2(__uint_as_float(raw.z << 16), __uint_as_float(raw.z & 0xffff0000u));
    x[3] = pack2(__uint_as_float(raw.w << 16), __uint_as_float(raw.w & 0xffff0000u));
  } else {
    x[0] = (static_cast<f32x2>(raw.y) << 32) | raw.x;
    x[1] = (static_cast<f32x2>(raw.w) << 32) | raw.z;
  }
}

// Per-warp streaming state of one round: the chunk scheduler, the ring cursor and the register accumulators.
template <int NV, int kDepth, bool BF16>
struct RowStream {
  static constexpr int kP = RowElems<BF16>::kP;
  RowScheduler sch;
  int head;                 // stage holding the oldest row in flight
  int n_in;                 // rows in flight (stages head, head+1, ... in ring order)
  f32x2 acc[NV][kP];
  float loss_acc;
  unsigned int cnt;
};

template <int NV, int kDepth, bool BF16>
AFB_DEVINL void stream_reset(RowStream<NV, kDepth, BF16>& rs, const RoundParams& p, const int thr24) {
#pragma unroll
  for (int c = 0; c < NV; ++c)
#pragma unroll
    for (int j = 0; j < RowElems<BF16>::kP; ++j) rs.acc[c][j] = 0ULL;
  rs.head = 0;
  rs.n_in = 0;
  rs.loss_acc = 0.f;
  rs.cnt = 0;
  sched_init(rs.sch, p.n_rows, gridDim.x * kStreamWarps, blockIdx.x * kStreamWarps + (threadIdx.x >> 5), thr24);
}

// Prologue on immutable inputs only (X, y, gidx): hash the statically assigned first chunk, put up to kDepth row fetches
// in flight.  No atomics, no protocol state: legal before the round barrier / the pull.
template <int NV, int kDepth, bool BF16>
AFB_DEVINL void stream_prologue(RowStream<NV, kDepth, BF16>& rs, const WarpRing& rg, const RoundParams& p, const unsigned long long key,
                                const int lane) {
  const int nchunks = static_cast<int>(p.ld) / RowElems<BF16>::kE;
  const unsigned int c0 = rs.sch.next_chunk;
  rs.sch.next_chunk = 0xffffffffu;                 // "first dynamic chunk id not fetched yet"
  if (c0 < static_cast<unsigned int>(rs.sch.num_chunks)) {
    load_chunk_masks(rs.sch, p, c0, lane, key);
    while (rs.n_in < kDepth) {
      const int r = pop_mask_row(rs.sch);
      if (r < 0) break;
      ring_issue<NV>(rg, p, r, rg.base_s + rs.n_in * rg.stage_bytes, rg.bars_s + 8u * rs.n_in, (p.tma_mask >> rs.n_in) & 1, lane, nchunks);
      ++rs.n_in;
    }
  }
}

// wait for every row still in flight (a CTA must not exit, nor may its stages be reused, with copies pending)
template <int NV, int kDepth, bool BF16>
AFB_DEVINL void stream_drain(RowStream<NV, kDepth, BF16>& rs, const WarpRing& rg, unsigned int& phase) {
  while (rs.n_in > 0) {
    mbar_wait_s(rg.bars_s + 8u * rs.head, (phase >> rs.head) & 1u);
    phase ^= (1u << rs.head);
    rs.head = (rs.head + 1 == kDepth) ? 0 : rs.head + 1;
    --rs.n_in;
  }
}

// log(1 + e) for e in (0, 1]: the fast-math log of the rounded sum is accurate to ~1e-7 absolute here, which is far below
// the fp32 noise of a loss SUM over thousands of rows (the loss value is reporting only: it never feeds the gradient)
AFB_DEVINL void loss_coeff_fast(const int loss_id, const float margin, const float y, float& coeff, float& loss) {
  if (loss_id == LOSS_LOGISTIC) {
    const float e = __expf(-fabsf(margin));
    const float inv = __fdividef(1.f, 1.f + e);
    const float sig = margin >= 0.f ? inv : e * inv;
    coeff = sig - y;
    const float l1p = __logf(1.f + e) + fmaxf(-margin, 0.f);
    loss = (y > 0.f) ? l1p : l1p + margin;
  } else {
    loss_coeff_rt(loss_id, margin, y, coeff, loss);
  }
}

// The gradient phase of one round: fill the ring, then consume rows in FIFO order until the chunk pool is empty.
// `saga`: history correction + write-through history update (K4/K7).  `trigger`: the pipelined loop's pull-ahead fires
// (on_trigger()) when the warp draws a chunk id >= trigger; 0xffffffff disarms it.
template <int NV, int kDepth, bool BF16, typename Trigger>
AFB_DEVINL void stream_rows(RowStream<NV, kDepth, BF16>& rs, const WarpRing& rg, const RoundParams& p, unsigned int* chunk_ctr,
                            const unsigned long long key, const f32x2 (&wreg)[NV][RowElems<BF16>::kP], unsigned int& phase,
                            const bool prologue_done, const bool saga, const int lane, unsigned int trigger, Trigger&& on_trigger) {
  constexpr int kE = RowElems<BF16>::kE;
  constexpr int kP = RowElems<BF16>::kP;
  const int nchunks = static_cast<int>(p.ld) / kE;
  const int LOSS = p.loss_id;
  const uint32_t stage = rg.stage_bytes, row_bytes = rg.row_bytes;
  const uint32_t tma_mask = static_cast<uint32_t>(p.tma_mask);
  const bool last_partial = (lane + 32 * (NV - 1)) >= nchunks;       // this lane has no part of the last chunk column
  // the statically assigned first chunk: consumed by the prologue, or still pending (no prologue: seed known only now)
  unsigned int cur_chunk = 0;
  bool exhausted = false;
  if (!prologue_done) {
    const unsigned int c0 = rs.sch.next_chunk;
    if (c0 < static_cast<unsigned int>(rs.sch.num_chunks)) load_chunk_masks(rs.sch, p, c0, lane, key);
  }
  stream_request_chunk(rs.sch, chunk_ctr, lane);              // first dynamic chunk id: requested, not awaited
  {
    int tail = rs.head + rs.n_in;                             // first free stage (head == 0 here)
    while (rs.n_in < kDepth) {
      const int r = stream_next_row(rs.sch, p, chunk_ctr, lane, key, cur_chunk, exhausted);
      if (r < 0) break;
      ring_issue<NV>(rg, p, r, rg.base_s + tail * stage, rg.bars_s + 8u * tail, (tma_mask >> tail) & 1u, lane, nchunks);
      ++tail;
      ++rs.n_in;
    }
  }
  __syncwarp();                                               // lane 0's row-id stores are visible to the warp
  int head = rs.head;
  int n_in = rs.n_in;
  uint32_t st_s = rg.base_s + head * stage;                   // shared-window address of the head stage
  uint32_t bar_s = rg.bars_s + 8u * head;
  const uint32_t st_end = rg.base_s + kDepth * stage;
  // label / history of the row at the head of the ring: plain loads issued one row ahead of their use
  int row_nxt = 0;
  float y_nxt = 0.f, h_nxt = 0.f;
  if (n_in > 0) {
    row_nxt = static_cast<int>(lds_u32(st_s + row_bytes));
    y_nxt = __ldg(p.y + row_nxt);
    if (saga) h_nxt = __ldcg(p.hist + row_nxt);
  }
  while (n_in > 0) {
    if (cur_chunk >= trigger) { trigger = 0xffffffffu; on_trigger(); }
    const int row = row_nxt;
    const float yy = y_nxt, hh = h_nxt;
    uint32_t nst_s = st_s + stage, nbar_s = bar_s + 8u;
    int nhead = head + 1;
    if (nst_s == st_end) { nst_s = rg.base_s; nbar_s = rg.bars_s; nhead = 0; }
    if (n_in >= 2) {                                          // the next row in FIFO order sits in the next stage
      row_nxt = static_cast<int>(lds_u32(nst_s + row_bytes));
      y_nxt = __ldg(p.y + row_nxt);
      if (saga) h_nxt = __ldcg(p.hist + row_nxt);
    }
    mbar_wait_s(bar_s, (phase >> head) & 1u);
    phase ^= (1u << head);
    uint4 raw[NV];
#pragma unroll
    for (int c = 0; c < NV - 1; ++c) raw[c] = lds_v4(st_s + 16u * (lane + 32 * c));
    // only the last chunk column can fall outside the row (NV = ceil(nchunks / 32)): those lanes compute on zeros
    raw[NV - 1] = last_partial ? make_uint4(0u, 0u, 0u, 0u) : lds_v4(st_s + 16u * (lane + 32 * (NV - 1)));
    __syncwarp();                                             // every lane has read the stage: refill it
    const int nrow = stream_next_row(rs.sch, p, chunk_ctr, lane, key, cur_chunk, exhausted);
    if (nrow >= 0) ring_issue<NV>(rg, p, nrow, st_s, bar_s, (tma_mask >> head) & 1u, lane, nchunks);
    else --n_in;
    head = nhead; st_s = nst_s; bar_s = nbar_s;
    f32x2 dp0 = 0ULL, dp1 = 0ULL;
#pragma unroll
    for (int c = 0; c < NV; ++c) {                            // K1
      f32x2 x[kP];
      unpack_chunk<BF16>(raw[c], x);
      dp0 = ffma2(x[0], wreg[c][0], dp0);
      dp1 = ffma2(x[1], wreg[c][1], dp1);
      if constexpr (BF16) {
        dp0 = ffma2(x[2], wreg[c][2], dp0);
        dp1 = ffma2(x[3], wreg[c][3], dp1);
      }
    }
    float d0, d1, d2, d3;
    unpack2(dp0, d0, d1);
    unpack2(dp1, d2, d3);
    const float margin = warp_sum((d0 + d1) + (d2 + d3));
    float coeff, loss;
    loss_coeff_fast(LOSS, margin, yy, coeff, loss);           // K2
    float eff = coeff;
    if (saga) {                                               // K4 + K7 (write-through: rollback rounds use dense_round.cu)
      eff = coeff - hh;
      if (lane == 0) p.hist[row] = coeff;
    }
    const f32x2 ee = pack2(eff, eff);
#pragma unroll
    for (int c = 0; c < NV; ++c) {                            // K3 (lanes outside the row accumulate zeros)
      f32x2 x[kP];
      unpack_chunk<BF16>(raw[c], x);
#pragma unroll
      for (int j = 0; j < kP; ++j) rs.acc[c][j] = ffma2(ee, x[j], rs.acc[c][j]);
    }
    rs.loss_acc += loss;
    ++rs.cnt;
  }
  rs.head = head;
  rs.n_in = 0;
}

// CTA-level reduce of the per-warp register accumulators through the (idle) ring stages, merged into `g_dst` with vector
// L2 reductions.  Contains two CTA barriers; on return the ring may be refilled.
template <int NV, int kDepth, bool BF16>
AFB_DEVINL void stream_reduce(const RowStream<NV, kDepth, BF16>& rs, const WarpRing& rg, uint8_t* s_raw, const RoundParams& p,
                              float* g_dst, const int lane) {
  constexpr int kE = RowElems<BF16>::kE;
  constexpr int kP = RowElems<BF16>::kP;
  const int nchunks = static_cast<int>(p.ld) / kE;
  const int nvec = static_cast<int>(p.ld >> 2);
  // the 296 CTAs of a grid finish a round at about the same time: spreading their merges over `g_copies` replicas of g
  // divides the same-address L2 reduction traffic (the last CTA sums the replicas before it pushes)
  g_dst += static_cast<long long>(blockIdx.x % static_cast<unsigned int>(p.g_copies)) * p.ld;
  float* mine = reinterpret_cast<float*>(rg.base);      // ld fp32 fit: kDepth * stage_bytes >= 2 * (2 ld + 16) >= 4 ld
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int v = lane + 32 * c;
    if (v < nchunks) {
#pragma unroll
      for (int j = 0; j < kP; ++j) *reinterpret_cast<f32x2*>(mine + kE * v + 2 * j) = rs.acc[c][j];
    }
  }
  __syncthreads();
  const size_t warp_stride = static_cast<size_t>(kDepth) * rg.stage_bytes;
  for (int v = threadIdx.x; v < nvec; v += kStreamThreads) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kStreamWarps; ++w) {
      const float4 a = reinterpret_cast<const float4*>(s_raw + w * warp_stride)[v];
      t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    }
    red_add_gpu_f32x4(g_dst + 4 * v, t);
  }
  __syncthreads();
  if (lane == 0) tc::fence_proxy_async();           // generic-proxy accesses to the stages are ordered before the next bulk copies
}

template <int NV, bool BF16>
AFB_DEVINL void load_w_regs(f32x2 (&wreg)[NV][RowElems<BF16>::kP], const float* w, const int nchunks, const int lane) {
  constexpr int kE = RowElems<BF16>::kE;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int v = lane + 32 * c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (v < nchunks) {
      a = __ldcg(reinterpret_cast<const float4*>(w + kE * v));
      if constexpr (BF16) b = __ldcg(reinterpret_cast<const float4*>(w + kE * v + 4));
    }
    wreg[c][0] = pack2(a.x, a.y); wreg[c][1] = pack2(a.z, a.w);
    if constexpr (BF16) { wreg[c][2] = pack2(b.x, b.y); wreg[c][3] = pack2(b.z, b.w); }
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// serial persistent round (protocol of afb_round_impl.cuh)
// ----------------------------------------------------------------------------------------------------------------------
template <int NV, int kDepth, bool BF16>
AFB_DEVINL int stream_round_serial(const RoundParams& p, RoundShared<kStreamThreads>& sh, uint8_t* s_raw, const WarpRing& rg,
                                   unsigned int& phase, const unsigned int round, const bool wait) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nchunks = static_cast<int>(p.ld) / RowElems<BF16>::kE;
  WorkerCtl* wc = p.wctl;
  const bool saga = (p.hist != nullptr);
  RowStream<NV, kDepth, BF16> rs;
  stream_reset(rs, p, eff_thr24(p, round));
  // sampling key known without the server's k?  (seed_mode 1: seed + round + 1; 2 / grad-only: fixed seed)
  const bool key_early = (p.algo == ALGO_GRAD_ONLY) ? (p.seed_mode != 3) : (p.seed_mode == 1 || p.seed_mode == 2);
  unsigned long long key = 0;
  bool prologue = false;
  if (key_early) {
    unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
    if (p.algo != ALGO_GRAD_ONLY && p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
    key = mix64(seed_eff);
    stream_prologue(rs, rg, p, key, lane);         // immutable inputs only: the history is read later, with plain loads
    prologue = true;
  }
  if (wait) {                                       // round barrier: the previous round (tail included) is published
    if (tid == 0) {
      SpinGuard sg;
      while (ld_acquire_gpu_u32(&wc->round) != round) {
        __nanosleep(20);
        if (sg.expired(p.spin_limit_ns) || aborted(wc)) { raise_abort(wc); break; }
      }
    }
    if (cta_abort_vote<kStreamThreads>(sh, wc)) {
      stream_drain(rs, rg, phase);
      return PROCEED_STOP;
    }
  }
  RoundPull pl;
  round_begin_issue<kStreamThreads>(p, sh, round, pl);
  const int pr = round_begin_complete<kStreamThreads>(p, sh, round, pl);
  if (pr != PROCEED_RUN) {
    // rows put in flight by the prologue must land before the stages are reused (or the CTA exits)
    stream_drain(rs, rg, phase);
    return pr;
  }
  if (!key_early) key = *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
  f32x2 wreg[NV][RowElems<BF16>::kP];
  load_w_regs<NV, BF16>(wreg, p.w_local, nchunks, lane);
  stream_rows<NV, kDepth, BF16>(rs, rg, p, &wc->chunk_ctr, key, wreg, phase, prologue, saga, lane, 0xffffffffu, []() {});
  stream_reduce<NV, kDepth, BF16>(rs, rg, s_raw, p, p.g_local, lane);
  round_finish<kStreamThreads>(p, sh, round, rs.loss_acc, rs.cnt);
  return PROCEED_RUN;
}

// ----------------------------------------------------------------------------------------------------------------------
// pipelined persistent round (ASGD; per-round state double-buffered by parity)
// ----------------------------------------------------------------------------------------------------------------------
// The pull of one round, executed by ONE warp that holds ticket 0 of the round's buffer: a single NVLink round trip brings
// the logical clock and all of w; `first` marks the first round of the launch (the worker becomes busy).
// (out of line: the pull-ahead calls it from inside the row loop once per round — its 28-32 staging registers must not
// raise the register pressure of the loop itself; only the rare call site pays for saving live state)
template <int NV4>
AFB_DEVINL void pipe_pull_warp(const RoundParams& p, PipeBuf* S, float* w_loc, const unsigned int round, const bool first, const int lane) {
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  float4 wv[NV4];
  unsigned long long ck = 0;
  if (lane == 0) wc->dbg_ns[0] = globaltimer_ns();
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const int v = lane + 32 * c;
    wv[c] = (v < nvec) ? ld_relaxed_sys_f32x4(p.w_server + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);      // K10
  }
  if (lane == 0) ck = ld_acquire_sys_u64(&p.sctl->clock_k);
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const int v = lane + 32 * c;
    if (v < nvec) __stcg(reinterpret_cast<float4*>(w_loc + 4 * v), wv[c]);
  }
  if (lane == 0) {
    int proceed = 1;
    if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop) || aborted(wc)) proceed = 0;
    const unsigned long long now = globaltimer_ns();
    WorkerSlot* me = p.status + p.worker_id;
    if (proceed && first) {
      red_add_release_sys_s32(&p.sctl->avail_count, -1);        // busy for the whole persistent launch
      if (wc->last_finish_ns) wc->wait_ns += now - wc->last_finish_ns;
      me->available = 0;
      me->busy_since_ns = now;
      me->wait_ns = wc->wait_ns;
    }
    if (proceed) me->heartbeat_ns = now;
    S->ts = ck & 0xffffffffULL;
    S->k_read = ck >> 32;
    S->proceed = proceed;
    S->t_start_ns = now;
    wc->dbg_ns[2] = now;
    if (proceed && p.delay_mode != 0) {
      // A18 straggler injector: sleep a multiple of the calibrated mean task time before the round may start
      const unsigned long long avg = ld_acquire_sys_u64(&p.sctl->avg_delay_ns);
      if (avg) {
        float mult = eff_delay_coeff(p, round);
        if (p.delay_mode >= 2) {
          unsigned long long seed_eff = static_cast<unsigned long long>(p.seed) + static_cast<unsigned long long>(round) + 1ULL;
          float u = static_cast<float>(hash_u24(mix64(seed_eff) ^ 0xD1B54A32D192ED03ULL, p.worker_id)) * (1.f / 16777216.f);
          mult = (p.delay_mode == 3) ? (2.5f + 7.5f * u) : (1.5f + u);
        }
        sleep_ns(static_cast<unsigned long long>(mult * static_cast<float>(avg)));
      }
    }
  }
  __threadfence();
  __syncwarp();
  if (lane == 0) st_release_gpu_u32(&S->pull_flag, round + 1u);
}

// one warp: wait until `*flag == want` (gpu-scope acquire); false on abort / timeout
AFB_DEVINL bool warp_wait_flag(const RoundParams& p, const unsigned int* flag, const unsigned int want, const int lane) {
  int ok = 1;
  if (lane == 0) {
    SpinGuard sg;
    while (ld_acquire_gpu_u32(flag) != want) {
      __nanosleep(32);
      if (sg.expired(p.spin_limit_ns) || aborted(p.wctl)) { raise_abort(p.wctl); ok = 0; break; }
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok != 0;
}

template <int NV, int kDepth, bool BF16>
AFB_DEVINL int stream_round_pipe(const RoundParams& p, RoundShared<kStreamThreads>& sh, uint8_t* s_raw, const WarpRing& rg,
                                 unsigned int& phase, const unsigned int round, const int it, const int rounds,
                                 PipeState* ps, float* g2, float* w2) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nvec = static_cast<int>(p.ld >> 2);
  WorkerCtl* wc = p.wctl;
  PipeBuf* S = &ps->b[round & 1u];
  PipeBuf* Sn = &ps->b[(round + 1u) & 1u];
  float* g_loc = g2 + static_cast<long long>(round & 1u) * p.ld * p.g_copies;      // [parity][copy][ld]
  float* w_loc = w2 + static_cast<long long>(round & 1u) * p.ld;
  float* w_nxt = w2 + static_cast<long long>((round + 1u) & 1u) * p.ld;

  constexpr int NVW = BF16 ? 2 * NV : NV;             // float4's of w per lane (the pull is always fp32)
  const int nchunks = static_cast<int>(p.ld) / RowElems<BF16>::kE;
  RowStream<NV, kDepth, BF16> rs;
  stream_reset(rs, p, eff_thr24(p, round));
  // ---- prologue on immutable inputs: sampling key of this round, first (static) chunk, first row fetches
  unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
  if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
  const unsigned long long key = mix64(seed_eff);
  stream_prologue(rs, rg, p, key, lane);
  // ---- the pull.  Usually done AHEAD by a warp of the previous round (ticket 0 of this buffer already taken); otherwise
  //      CTA 0's warp 0 pulls right away and any other CTA's warp 0 steps in if the flag has not appeared after ~1 us
  //      (CTA 0 may be busy in a tail).  No election atomic and no extra CTA barrier for the other CTAs.
  if (warp == 0) {
    unsigned int tries = 0;
    SpinGuard sg;
    // buffer reuse: round r-2 (same parity) must be published before anything of buffer r&1 is touched
    if (it >= 2 && lane == 0) {
      while (static_cast<int>(ld_acquire_gpu_u32(&wc->round) - (round - 1u)) < 0) {
        __nanosleep(20);
        if (sg.expired(p.spin_limit_ns) || aborted(wc)) { raise_abort(wc); break; }
      }
    }
    __syncwarp();
    while (true) {
      unsigned int f = 0;
      if (lane == 0) f = ld_acquire_gpu_u32(&S->pull_flag);
      f = __shfl_sync(0xffffffffu, f, 0);
      if (f == round + 1u) break;
      if ((blockIdx.x == 0 && tries == 0) || tries == 24) {
        unsigned int tk = 1u;
        if (lane == 0) tk = atomicAdd(&S->pull_ticket, 1u);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk == 0u) { pipe_pull_warp<NVW>(p, S, w_loc, round, it == 0, lane); break; }
      }
      ++tries;
      __nanosleep(40);
      if (sg.expired(p.spin_limit_ns) || aborted(wc)) { if (lane == 0) raise_abort(wc); break; }
    }
    if (lane == 0) sh.flag = (*reinterpret_cast<volatile int*>(&S->proceed) != 0 && !aborted(wc)) ? 1u : 0u;
  }
  __syncthreads();
  const int proceed = static_cast<int>(sh.flag);
  if (!proceed) {
    // stopped (k >= max_iters, stop flag or abort): drain the prologue's copies, then the last CTA to arrive resets the
    // buffer and publishes, in round order
    stream_drain(rs, rg, phase);
    __syncthreads();
    if (tid == 0) {
      const unsigned int t = atomicAdd(&S->done_ctas, 1u);
      if (t == gridDim.x - 1) {
        SpinGuard sg;
        while (ld_acquire_gpu_u32(&wc->round) != round) {                    // the previous round's tail first
          __nanosleep(20);
          if (sg.expired(p.spin_limit_ns) || aborted(wc)) { raise_abort(wc); break; }
        }
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
    return PROCEED_STOP;
  }

  // ---- phase 1: fused gradient; the warp that draws one of the last chunk ids pulls round r+1's parameters ahead
  f32x2 wreg[NV][RowElems<BF16>::kP];
  load_w_regs<NV, BF16>(wreg, w_loc, nchunks, lane);
  // pull-ahead trigger: the chunk id from which `pull_ahead_chunks` x 64 raw rows remain in the pool (the pool ends with
  // the 32- and 16-row tiers of the guided schedule, so the id is found by walking the tiers from the end)
  unsigned int trigger;
  {
    long long rows_left = static_cast<long long>(p.pull_ahead_chunks) * kChunkRows;
    const int n_small = rs.sch.num_chunks - rs.sch.n_big - rs.sch.n_mid;
    long long from_end = 0;
    long long take = rows_left / kSmallChunkRows; if (take > n_small) take = n_small;
    from_end += take; rows_left -= take * kSmallChunkRows;
    take = rows_left / kMidChunkRows; if (take > rs.sch.n_mid) take = rs.sch.n_mid;
    from_end += take; rows_left -= take * kMidChunkRows;
    take = rows_left / kChunkRows; if (take > rs.sch.n_big) take = rs.sch.n_big;
    from_end += take;
    const long long id = static_cast<long long>(rs.sch.num_chunks) - from_end;
    trigger = static_cast<unsigned int>(id > static_cast<long long>(rs.sch.static_chunks) ? id : static_cast<long long>(rs.sch.static_chunks));
  }
  const bool ahead_armed = (p.pull_ahead_chunks > 0) && (it + 1 < rounds) && (p.delay_mode == 0);
  stream_rows<NV, kDepth, BF16>(rs, rg, p, &S->chunk_ctr, key, wreg, phase, true, false, lane, ahead_armed ? trigger : 0xffffffffu, [&]() {
    // one attempt per warp and round.  Buffer (r+1)&1 is free once round r-1 has been published; otherwise the regular
    // election at the start of round r+1 pulls
    unsigned int tk = 1u;
    if (lane == 0 && static_cast<int>(ld_acquire_gpu_u32(&wc->round) - round) >= 0) tk = atomicAdd(&Sn->pull_ticket, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk == 0u) pipe_pull_warp<NVW>(p, Sn, w_nxt, round + 1u, false, lane);
  });

  // ---- phase 2: CTA reduce -> g[parity]; loss / count; CTA ticket
  stream_reduce<NV, kDepth, BF16>(rs, rg, s_raw, p, g_loc, lane);
  if (lane == 0) { sh.red[warp] = rs.loss_acc; sh.cnt[warp] = rs.cnt; }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f; unsigned int n = 0;
#pragma unroll
    for (int w = 0; w < kStreamWarps; ++w) { l += sh.red[w]; n += sh.cnt[w]; }
    atomicAdd(&S->loss_sum, l);
    atomicAdd(&S->batch_count, n);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.flag = atomicAdd(&S->done_ctas, 1u);
  __syncthreads();
  const unsigned int my_ticket = sh.flag;
  __syncthreads();                                   // sh.flag is reused below

  // ---- early logical-clock ticket by the first CTA to finish (its NVLink round trip overlaps the drain)
  if (my_ticket == 0 && tid == 0) {
    const Ticket t = take_ticket(p, S->ts, round);
    S->tk_accept = t.accept; S->tk_step = t.step; S->tk_clock_old = t.clock_old;
    S->tk_staleness = t.staleness; S->tk_k_after = t.k_after;
    __threadfence();
    st_release_gpu_u32(&S->tk_flag, round + 1u);
  }
  if (my_ticket != gridDim.x - 1) return PROCEED_RUN;          // everybody else: on to round r+1

  // ---- tail of round `round` (this CTA only), overlapping the other CTAs' round r+1
  if (tid == 0) {
    wc->dbg_ns[3] = globaltimer_ns();
    SpinGuard sg;
    while (ld_acquire_gpu_u32(&wc->round) != round) {                        // tails publish in round order
      __nanosleep(20);
      if (sg.expired(p.spin_limit_ns) || aborted(wc)) { raise_abort(wc); break; }
    }
    while (ld_acquire_gpu_u32(&S->tk_flag) != round + 1u) {
      __nanosleep(32);
      if (sg.expired(p.spin_limit_ns) || aborted(wc)) { raise_abort(wc); break; }
    }
    wc->dbg_ns[4] = globaltimer_ns();
    sh.flag = static_cast<unsigned int>(*reinterpret_cast<volatile int*>(&S->tk_accept));
    sh.bcast[0] = *reinterpret_cast<volatile float*>(&S->tk_step);
  }
  __threadfence();
  __syncthreads();
  const int accept = aborted(wc) ? 0 : static_cast<int>(sh.flag);
  const float step = sh.bcast[0];
  merge_g_copies<kStreamThreads>(p, g_loc);                                   // replicas -> copy 0 (replicas re-zeroed)
  if (accept) hogwild_apply<kStreamThreads>(p, g_loc, w_loc, step);           // K9 + K5 (+ K8)
  for (int v = tid; v < nvec; v += kStreamThreads)
    *reinterpret_cast<float4*>(g_loc + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long now = 0, task_ns = 0;
  unsigned int clock_old = 0;
  ResultRec* rec = nullptr;
  float round_loss = 0.f;
  unsigned int round_batch = 0;
  const bool last_of_launch = (it == rounds - 1);
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
    if (accept && p.warmup_updates > 0 && S->tk_k_after <= static_cast<long long>(p.warmup_updates)) {
      asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(&p.sctl->cul_time_ns), "l"(task_ns) : "memory");
      asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(&p.sctl->cul_count), "r"(1u) : "memory");
    }
    if (last_of_launch) me->available = 1;                                   // the persistent launch ends with this round
  }
  fence_sys();                               // ONE fence: reductions + record + status performed before publication
  __syncthreads();
  if (tid == 0) {
    wc->dbg_ns[6] = globaltimer_ns();
    *reinterpret_cast<volatile unsigned long long*>(&rec->seq) = static_cast<unsigned long long>(clock_old) + 1ULL;
    if (last_of_launch) red_add_release_sys_s32(&p.sctl->avail_count, 1);
    if (accept && p.warmup_updates > 0 && S->tk_k_after == static_cast<long long>(p.warmup_updates)) {
      unsigned long long tot = ld_acquire_sys_u64(&p.sctl->cul_time_ns);
      unsigned int c = ld_acquire_sys_u32(&p.sctl->cul_count);
      if (c) atomicExch_system(&p.sctl->avg_delay_ns, tot / c);
    }
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
  return PROCEED_RUN;
}

template <int NV, int kDepth, bool PIPE, bool BF16>
__global__ void __launch_bounds__(kStreamThreads, (NV * RowElems<BF16>::kE <= 32) ? 2 : 1)
dense_stream_kernel(const __grid_constant__ RoundParams p, const int rounds, PipeState* ps, float* g2, float* w2) {
  __shared__ RoundShared<kStreamThreads> sh;
  __shared__ __align__(8) uint64_t s_bar[kStreamWarps][kDepth];
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WarpRing rg;
  rg.row_bytes = static_cast<uint32_t>(p.ld * (BF16 ? 2 : 4));
  rg.stage_bytes = rg.row_bytes + kStreamMetaBytes;
  rg.base = s_raw + static_cast<size_t>(warp) * kDepth * rg.stage_bytes;
  rg.base_s = tc::smem_u32(rg.base);
  rg.bars_s = tc::smem_u32(&s_bar[warp][0]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kDepth; ++s) tc::mbar_init(&s_bar[warp][s], ((p.tma_mask >> s) & 1) ? 1u : 32u);
    tc::fence_barrier_init();
  }
  __syncthreads();
  unsigned int phase = 0;                          // bit s = parity of the next completion of this warp's stage s
  // No CTA of this grid can have finished a round before every CTA has started (the last-CTA election needs all of
  // them), so this read is the same value in every CTA.
  const unsigned int round0 = *reinterpret_cast<volatile unsigned int*>(&p.wctl->round);
  for (int it = 0; it < rounds; ++it) {
    const unsigned int round = round0 + static_cast<unsigned int>(it);
    int r;
    if (PIPE) r = stream_round_pipe<NV, kDepth, BF16>(p, sh, s_raw, rg, phase, round, it, rounds, ps, g2, w2);
    else r = stream_round_serial<NV, kDepth, BF16>(p, sh, s_raw, rg, phase, round, it > 0);
    if (r == PROCEED_STOP) return;
    __syncthreads();                               // sh is reused by the next round
  }
}

constexpr int kMaxCachedDevices = 64;

template <int NV, int kDepth, bool PIPE, bool BF16>
static cudaError_t launch_stream(const RoundParams& p, int rounds, int grid, int share, PipeState* ps, float* g2, float* w2,
                                 cudaStream_t stream) {
  const size_t stage = static_cast<size_t>(p.ld) * (BF16 ? 2 : 4) + kStreamMetaBytes;
  const size_t smem = static_cast<size_t>(kStreamWarps) * kDepth * stage + 128;
  auto k = dense_stream_kernel<NV, kDepth, PIPE, BF16>;
  // The attribute / occupancy queries are host-side driver calls inside every caller's timed region (the GPU idles until the
  // launch arrives): do them once per (instantiation, device, shared-memory size), not once per launch.
  struct Setup { size_t smem = 0; int capacity = 0; };
  static Setup s_setup[kMaxCachedDevices];
  cudaError_t e;
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  Setup local;
  Setup& su = (dev >= 0 && dev < kMaxCachedDevices) ? s_setup[dev] : local;
  if (su.smem != smem) {
    int sms = 0, per_sm = 0;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kStreamThreads, smem)) != cudaSuccess) return e;
    su.capacity = per_sm * sms;
    su.smem = smem;
  }
  if (rounds > 1) {
    // rounds are separated by in-kernel waits: every CTA of every co-running persistent grid must be resident
    if (share < 1) share = 1;
    if (static_cast<long long>(grid) * share > static_cast<long long>(su.capacity)) return cudaErrorLaunchOutOfResources;
  }
  k<<<grid, kStreamThreads, smem, stream>>>(p, rounds, ps, g2, w2);
  return cudaGetLastError();
}

// floats per lane held for w / acc: <= 32 => 128 registers, two CTAs per SM; wider rows: one CTA per SM
static int stream_nv(long long ld, bool bf16) { const int e = bf16 ? 8 : 4; return static_cast<int>((ld / e + 31) / 32); }
static int stream_ctas(long long ld, bool bf16) { return stream_nv(ld, bf16) * (bf16 ? 8 : 4) <= 32 ? 2 : 1; }
// ring depth: as deep as 227 KB of shared memory per SM allows for the resident CTAs, at most 4 (fp32) / 8 (bf16)
static int stream_depth(long long ld, bool bf16, bool pipe = false) {
  (void)pipe;
  const long long per_cta = (stream_ctas(ld, bf16) == 2 ? 113 : 226) * 1024LL - 2048;
  const long long stage = ld * (bf16 ? 2 : 4) + kStreamMetaBytes;
  long long d = per_cta / (kStreamWarps * stage);
  const long long cap = bf16 ? 8 : 4;
  if (d > cap) d = cap;
  return static_cast<int>(d);
}

template <bool PIPE, bool BF16>
static cudaError_t dispatch_stream(const RoundParams& p, int rounds, int grid, int share, PipeState* ps, float* g2, float* w2, cudaStream_t s) {
  const int nv = stream_nv(p.ld, BF16);
  const int depth = stream_depth(p.ld, BF16, PIPE);
  if (depth < 2) return cudaErrorInvalidValue;
#define AFB_STREAM_CASE(NVX, D) return launch_stream<NVX, D, PIPE, BF16>(p, rounds, grid, share, ps, g2, w2, s)
  if constexpr (BF16) {
    // bf16 rows: depth 8 / 6 / 4 / 3 / 2
#define AFB_STREAM_DEPTH(NVX) \
    do { if (depth >= 8) AFB_STREAM_CASE(NVX, 8); if (depth >= 6) AFB_STREAM_CASE(NVX, 6); if (depth >= 4) AFB_STREAM_CASE(NVX, 4); \
         if (depth == 3) AFB_STREAM_CASE(NVX, 3); AFB_STREAM_CASE(NVX, 2); } while (0)
    switch (nv) {
      case 1: AFB_STREAM_CASE(1, 8);
      case 2: AFB_STREAM_CASE(2, 8);
      case 3: case 4: AFB_STREAM_DEPTH(4);
      case 5: case 6: AFB_STREAM_DEPTH(6);
      default: AFB_STREAM_DEPTH(8);
    }
#undef AFB_STREAM_DEPTH
  } else {
#define AFB_STREAM_DEPTH(NVX) \
    do { if (depth >= 4) AFB_STREAM_CASE(NVX, 4); if (depth == 3) AFB_STREAM_CASE(NVX, 3); AFB_STREAM_CASE(NVX, 2); } while (0)
    switch (nv) {
      case 1: AFB_STREAM_CASE(1, 4);
      case 2: AFB_STREAM_CASE(2, 4);
      case 3: case 4: AFB_STREAM_CASE(4, 4);
      case 5: case 6: case 7: AFB_STREAM_DEPTH(7);
      case 8: AFB_STREAM_DEPTH(8);
      case 9: case 10: case 11: case 12: AFB_STREAM_DEPTH(12);
      default: AFB_STREAM_DEPTH(16);
    }
#undef AFB_STREAM_DEPTH
  }
#undef AFB_STREAM_CASE
  return cudaErrorInvalidValue;
}

}  // namespace afb

extern "C" {

int afb_dense_stream_max_ld() { return 2048; }
int afb_dense_stream_ctas_per_sm(long long ld, int bf16) { return afb::stream_ctas(ld, bf16 != 0); }
int afb_dense_stream_depth(long long ld, int bf16) { return afb::stream_depth(ld, bf16 != 0); }
int afb_dense_stream_depth_pipe(long long ld, int bf16) { return afb::stream_depth(ld, bf16 != 0, true); }
int afb_pipe_state_bytes() { return static_cast<int>(sizeof(afb::PipeState)); }

static int stream_check(const afb::RoundParams& p, int rounds, int bf16) {
  if (p.ld % (bf16 ? 8 : 4) != 0 || p.ld > 2048 || p.ld <= 0 || rounds <= 0) return static_cast<int>(cudaErrorInvalidValue);
  if (p.hist_prev != nullptr || p.pdl != 0 || p.seed_mode == 3 || p.g_copies < 1) return static_cast<int>(cudaErrorInvalidValue);
  if ((reinterpret_cast<uintptr_t>(p.X) & 15) || (reinterpret_cast<uintptr_t>(p.y) & 15)) return static_cast<int>(cudaErrorMisalignedAddress);
  return 0;
}

// Serial persistent rounds: every algorithm (rollback rounds excepted: hist_prev must be null), `rounds` >= 1.
// `share` = persistent grids co-running on this GPU (logical workers per GPU); refused unless all of them fit.
// `bf16` != 0: X holds bf16 rows (ld % 8 == 0).
int afb_dense_stream_launch(const afb::RoundParams* params, int rounds, int grid, int share, int bf16, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  if (int rc = stream_check(p, rounds, bf16)) return rc;
  if (p.algo == ALGO_GRAD_ONLY && rounds != 1) return static_cast<int>(cudaErrorInvalidValue);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return static_cast<int>(bf16 ? dispatch_stream<false, true>(p, rounds, grid, share, nullptr, nullptr, nullptr, s)
                               : dispatch_stream<false, false>(p, rounds, grid, share, nullptr, nullptr, nullptr, s));
}

// Pipelined persistent rounds: ASGD only, no bucket / SSP gate, no enable mask, worker-local sampling seed (seed_mode 1 or 2);
// the grid must be fully co-resident and is expected to be the only round kernel of its GPU (one logical worker per GPU).
// `pipe_state` = afb_pipe_state_bytes() zeroed bytes, `g2` / `w2` = [2][ld] floats (g2 zeroed), all on the worker's GPU.
int afb_dense_stream_pipelined_launch(const afb::RoundParams* params, int rounds, int grid, void* pipe_state, float* g2, float* w2,
                                      int bf16, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  if (int rc = stream_check(p, rounds, bf16)) return rc;
  if (p.algo != ALGO_ASGD || p.hist != nullptr || p.gate_threshold != 0 || p.ssp_bound >= 0 || p.use_mask != 0 ||
      !(p.seed_mode == 1 || p.seed_mode == 2))
    return static_cast<int>(cudaErrorInvalidValue);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PipeState* ps = static_cast<PipeState*>(pipe_state);
  return static_cast<int>(bf16 ? dispatch_stream<true, true>(p, rounds, grid, 1, ps, g2, w2, s)
                               : dispatch_stream<true, false>(p, rounds, grid, 1, ps, g2, w2, s));
}

}  // extern "C"
