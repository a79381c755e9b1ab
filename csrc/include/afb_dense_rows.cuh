// Row streaming helpers of the dense round kernels: register-resident rows (`RowRegs`), and the warp-level
// guided self-scheduler that yields the next Bernoulli-selected local row (`RowScheduler`).  Shared by
// csrc/kernels/dense_round.cu and csrc/kernels/dense_stream.cu (scheduler + chunk masks).
#pragma once
#include "afb_round_impl.cuh"

namespace afb {

constexpr int kChunkRows = 64;

template <int NV4>
struct RowRegs {
  float4 v[NV4];
  float y;
  float h;
};

template <int NV4>
AFB_DEVINL void load_row(RowRegs<NV4>& r, const RoundParams& p, int row, int lane, int nvec, bool saga) {
  const float* xr = p.X + static_cast<long long>(row) * p.ld;
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    int v = lane + 32 * c;
    r.v[c] = (v < nvec) ? ld_stream_f32x4(xr + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  r.y = __ldg(p.y + row);
  r.h = saga ? p.hist[row] : 0.f;
}

// Warp-level scheduler state: yields the next selected local row.  Work units are row chunks taken
// from the worker's global counter (the next chunk id is requested before the current one is consumed):
//   * the first `static_chunks` chunks are assigned statically (chunk = global warp id) so a launch does
//     not start with thousands of same-address atomics;
//   * chunks are 64 rows except for the last ~2 chunks per warp, which are 16 rows (guided
//     self-scheduling) so the drain at the end of a round costs ~2 rows per warp instead of ~7.
struct RowScheduler {
  unsigned int m0, m1;      // selection masks of the current chunk (rows base+lane, base+32+lane)
  int base;                 // first row of the current chunk
  unsigned int next_chunk;  // prefetched chunk id
  int n_big;                // number of 64-row chunks
  int n_mid;                // number of 32-row chunks (between the big and the small ones)
  int num_chunks;           // total chunks (big + medium + small)
  unsigned int static_chunks;
  unsigned int thr24;       // Bernoulli threshold of this round (launch scalar or the round descriptor's)
};

constexpr int kSmallChunkRows = 16;
constexpr int kMidChunkRows = 32;

AFB_DEVINL void sched_init(RowScheduler& s, int n_rows, unsigned int total_warps, unsigned int gwarp, int thr24 = -1) {
  s.thr24 = static_cast<unsigned int>(thr24);      // callers pass eff_thr24(p, round); -1 = "use p.thr24" (set below by load_chunk_masks)
  // guided self-scheduling in three tiers: 64-row chunks, then ~2 chunks of 32 rows per warp, then ~2 chunks of 16 rows per
  // warp at the very end of the shard (never more than 1/4 of it for the two tail tiers together): a warp that draws its
  // last chunk late is at most a few sampled rows behind the others when they meet at the CTA reduce
  long long small_rows = static_cast<long long>(total_warps) * 2 * kSmallChunkRows;
  long long mid_rows = static_cast<long long>(total_warps) * 2 * kMidChunkRows;
  if (small_rows + mid_rows > n_rows / 4) { small_rows = n_rows / 12; mid_rows = n_rows / 6; }
  mid_rows = mid_rows / kMidChunkRows * kMidChunkRows;
  const int big_rows = n_rows - static_cast<int>(small_rows) - static_cast<int>(mid_rows);
  s.n_big = big_rows > 0 ? big_rows / kChunkRows : 0;
  s.n_mid = static_cast<int>(mid_rows / kMidChunkRows);
  const int rest = n_rows - s.n_big * kChunkRows - s.n_mid * kMidChunkRows;
  s.num_chunks = s.n_big + s.n_mid + (rest + kSmallChunkRows - 1) / kSmallChunkRows;
  s.static_chunks = total_warps;
  s.m0 = s.m1 = 0; s.base = 0;
  s.next_chunk = gwarp;                       // first unit: no atomic
}

AFB_DEVINL unsigned int fetch_chunk(unsigned int* ctr, int lane, unsigned int static_chunks) {
  unsigned int c = 0;
  if (lane == 0) c = atomicAdd(ctr, 1u);
  return __shfl_sync(0xffffffffu, c, 0) + static_chunks;
}

AFB_DEVINL void load_chunk_masks(RowScheduler& s, const RoundParams& p, unsigned int chunk, int lane,
                                 unsigned long long key) {
  int size;
  const int c = static_cast<int>(chunk);
  if (c < s.n_big) { s.base = c * kChunkRows; size = kChunkRows; }
  else if (c < s.n_big + s.n_mid) { s.base = s.n_big * kChunkRows + (c - s.n_big) * kMidChunkRows; size = kMidChunkRows; }
  else { s.base = s.n_big * kChunkRows + s.n_mid * kMidChunkRows + (c - s.n_big - s.n_mid) * kSmallChunkRows; size = kSmallChunkRows; }
  int r0 = s.base + lane, r1 = s.base + 32 + lane;
  bool s0 = false, s1 = false;
  const unsigned int thr = (s.thr24 == 0xffffffffu) ? static_cast<unsigned int>(p.thr24) : s.thr24;
  if (lane < size && r0 < p.n_rows) {
    unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r0]) : static_cast<unsigned long long>(p.row_base + r0);
    s0 = hash_u24(key, g) < thr;
  }
  if (32 + lane < size && r1 < p.n_rows) {
    unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r1]) : static_cast<unsigned long long>(p.row_base + r1);
    s1 = hash_u24(key, g) < thr;
  }
  s.m0 = __ballot_sync(0xffffffffu, s0);
  s.m1 = __ballot_sync(0xffffffffu, s1);
}

AFB_DEVINL int next_row(RowScheduler& s, const RoundParams& p, unsigned int* ctr, int lane,
                        unsigned long long key) {
  while (true) {
    if (s.m0) { int b = __ffs(s.m0) - 1; s.m0 &= s.m0 - 1; return s.base + b; }
    if (s.m1) { int b = __ffs(s.m1) - 1; s.m1 &= s.m1 - 1; return s.base + 32 + b; }
    unsigned int c = s.next_chunk;
    if (c >= static_cast<unsigned int>(s.num_chunks)) return -1;
    s.next_chunk = fetch_chunk(ctr, lane, s.static_chunks);     // request one ahead
    load_chunk_masks(s, p, c, lane, key);
  }
}

}  // namespace afb
