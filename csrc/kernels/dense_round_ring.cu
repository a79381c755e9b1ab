// Fused dense round kernel, per-warp TMA ring variant (sm_100a).
//
// Same phases/semantics as dense_round.cu.  Phase 1 keeps the warp-per-row mapping of the register
// variant but moves the rows-in-flight from registers into shared memory: every warp owns a private
// ring of kDepth stages and prefetches its next selected rows with `cp.async.bulk` (SASS: UBLKCP)
// completing on per-stage mbarriers.  Nothing is shared between warps (no producer warp, no hand-off),
// lane 0 issues the copies, all lanes consume with LDS.128.  With kDepth = 4 and 8 warps per CTA,
// 100 KB are in flight per CTA regardless of register pressure, so ONE CTA per SM already saturates
// HBM; two CTAs/SM (two logical workers) then overlap one worker's NVLink pull/push round trips with
// the other's streaming at full bandwidth.
#include <algorithm>
#include "afb_round_impl.cuh"
#include "afb_tc.cuh"

namespace afb {

constexpr int kRingWarps = 8;
constexpr int kRingThreads = 32 * kRingWarps;
constexpr int kRingChunkRows = 64;
constexpr int kRingMetaBytes = 32;       // {float yq[4]; float hq[4]}

__device__ __forceinline__ void ring_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

// Warp-level iterator over the selected rows of dynamically fetched 64-row chunks.
struct RingScheduler {
  unsigned int m0, m1;
  int base;
  unsigned int next_chunk;
};

AFB_DEVINL void ring_load_masks(RingScheduler& s, const RoundParams& p, unsigned int chunk, int lane, unsigned long long key) {
  s.base = static_cast<int>(chunk) * kRingChunkRows;
  const int r0 = s.base + lane, r1 = s.base + 32 + lane;
  bool s0 = false, s1 = false;
  if (r0 < p.n_rows) {
    unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r0]) : static_cast<unsigned long long>(p.row_base + r0);
    s0 = hash_u24(key, g) < static_cast<unsigned int>(p.thr24);
  }
  if (r1 < p.n_rows) {
    unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r1]) : static_cast<unsigned long long>(p.row_base + r1);
    s1 = hash_u24(key, g) < static_cast<unsigned int>(p.thr24);
  }
  s.m0 = __ballot_sync(0xffffffffu, s0);
  s.m1 = __ballot_sync(0xffffffffu, s1);
}

AFB_DEVINL int ring_next_row(RingScheduler& s, const RoundParams& p, unsigned int* ctr, int lane, unsigned long long key, int num_chunks) {
  while (true) {
    if (s.m0) { int b = __ffs(s.m0) - 1; s.m0 &= s.m0 - 1; return s.base + b; }
    if (s.m1) { int b = __ffs(s.m1) - 1; s.m1 &= s.m1 - 1; return s.base + 32 + b; }
    const unsigned int c = s.next_chunk;
    if (c >= static_cast<unsigned int>(num_chunks)) return -1;
    s.next_chunk = fetch_chunk_ctr(ctr, lane);
    ring_load_masks(s, p, c, lane, key);
  }
}

template <int NV4, int kDepth>
__global__ void __launch_bounds__(kRingThreads, 2) dense_round_ring_kernel(const RoundParams p) {
  __shared__ RoundShared<kRingThreads> sh;
  __shared__ __align__(8) uint64_t s_bar[kRingWarps][kDepth];
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = static_cast<int>(p.ld >> 2);
  const uint32_t row_bytes = static_cast<uint32_t>(p.ld * 4);
  const uint32_t stage_bytes = row_bytes + kRingMetaBytes;
  // smem: [warp][depth][row | meta] | g[ld]
  uint8_t* my_ring = s_raw + static_cast<size_t>(warp) * kDepth * stage_bytes;
  float* s_g = reinterpret_cast<float*>(s_raw + static_cast<size_t>(kRingWarps) * kDepth * stage_bytes);
  WorkerCtl* wc = p.wctl;
  const unsigned int round = wc->round;
  const bool saga = (p.hist != nullptr);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kDepth; ++s) tc::mbar_init(&s_bar[warp][s], 1);
    tc::fence_barrier_init();
  }
  if (p.algo != ALGO_GRAD_ONLY) {
    if (round_begin<kRingThreads>(p, sh, round) != PROCEED_RUN) return;
  }
  if (p.seed_mode == 3 && !*reinterpret_cast<volatile int*>(&wc->proceed)) return;
  const unsigned long long key = (p.algo == ALGO_GRAD_ONLY && p.seed_mode != 3)
                                     ? mix64(static_cast<unsigned long long>(p.seed))
                                     : *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
  for (int i = tid; i < p.ld; i += kRingThreads) s_g[i] = 0.f;
  __syncthreads();

  float4 wreg[NV4], acc[NV4];
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const int v = lane + 32 * c;
    wreg[c] = (v < nvec) ? *reinterpret_cast<const float4*>(p.w_local + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float loss_acc = 0.f;
  unsigned int cnt = 0;

  const int num_chunks = (p.n_rows + kRingChunkRows - 1) / kRingChunkRows;
  RingScheduler sch;
  sch.m0 = sch.m1 = 0; sch.base = 0;
  sch.next_chunk = fetch_chunk_ctr(&wc->chunk_ctr, lane);

  // issue(row, slot): lane 0 arms the stage barrier and launches the bulk copies
  auto issue = [&](int row, int slot) {
    if (lane == 0) {
      uint8_t* st = my_ring + static_cast<size_t>(slot) * stage_bytes;
      float* meta = reinterpret_cast<float*>(st + row_bytes);
      const bool quad_ok = (row | 3) < p.n_rows;
      uint32_t tx = row_bytes;
      if (quad_ok) tx += saga ? 32u : 16u;
      else {
        meta[row & 3] = __ldg(p.y + row);
        if (saga) meta[4 + (row & 3)] = p.hist[row];
      }
      tc::mbar_arrive_expect_tx(&s_bar[warp][slot], tx);
      ring_bulk_g2s(st, p.X + static_cast<long long>(row) * p.ld, row_bytes, &s_bar[warp][slot]);
      if (quad_ok) {
        ring_bulk_g2s(meta, p.y + (row & ~3), 16u, &s_bar[warp][slot]);
        if (saga) ring_bulk_g2s(meta + 4, p.hist + (row & ~3), 16u, &s_bar[warp][slot]);
      }
    }
  };

  // prologue: fill the ring
  int rows_q[kDepth];
#pragma unroll
  for (int s = 0; s < kDepth; ++s) {
    rows_q[s] = ring_next_row(sch, p, &wc->chunk_ctr, lane, key, num_chunks);
    if (rows_q[s] >= 0) issue(rows_q[s], s);
  }
  unsigned int it = 0;
  while (true) {
    bool any = false;
#pragma unroll
    for (int s = 0; s < kDepth; ++s) {
      const int row = rows_q[s];
      if (row < 0) continue;
      any = true;
      tc::mbar_wait(&s_bar[warp][s], it & 1u);
      const uint8_t* st = my_ring + static_cast<size_t>(s) * stage_bytes;
      const float* meta = reinterpret_cast<const float*>(st + row_bytes);
      const float yy = meta[row & 3];
      const float hh = saga ? meta[4 + (row & 3)] : 0.f;
      const float* xr = reinterpret_cast<const float*>(st);
      float4 xv[NV4];
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int c = 0; c < NV4; ++c) {
        const int v = lane + 32 * c;
        xv[c] = (v < nvec) ? *reinterpret_cast<const float4*>(xr + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
        d0 = fmaf(xv[c].x, wreg[c].x, d0);
        d1 = fmaf(xv[c].y, wreg[c].y, d1);
        d2 = fmaf(xv[c].z, wreg[c].z, d2);
        d3 = fmaf(xv[c].w, wreg[c].w, d3);
      }
      __syncwarp();                                   // every lane has read the stage: refill it
      const int nrow = ring_next_row(sch, p, &wc->chunk_ctr, lane, key, num_chunks);
      rows_q[s] = nrow;
      if (nrow >= 0) issue(nrow, s);
      const float margin = warp_sum((d0 + d1) + (d2 + d3));
      float coeff, loss;
      loss_coeff_rt(p.loss_id, margin, yy, coeff, loss);
      float eff = coeff;
      if (saga) {
        eff = coeff - hh;
        if (lane == 0) {
          if (p.hist_prev) { p.hist_prev[row] = hh; p.hist_tag[row] = static_cast<int>(round); }
          p.hist[row] = coeff;
        }
      }
#pragma unroll
      for (int c = 0; c < NV4; ++c) {
        acc[c].x = fmaf(eff, xv[c].x, acc[c].x);
        acc[c].y = fmaf(eff, xv[c].y, acc[c].y);
        acc[c].z = fmaf(eff, xv[c].z, acc[c].z);
        acc[c].w = fmaf(eff, xv[c].w, acc[c].w);
      }
      loss_acc += loss;
      ++cnt;
    }
    if (!any) break;
    ++it;
  }
#pragma unroll
  for (int c = 0; c < NV4; ++c) {
    const int v = lane + 32 * c;
    if (v < nvec) {
      atomicAdd(&s_g[4 * v + 0], acc[c].x);
      atomicAdd(&s_g[4 * v + 1], acc[c].y);
      atomicAdd(&s_g[4 * v + 2], acc[c].z);
      atomicAdd(&s_g[4 * v + 3], acc[c].w);
    }
  }
  __syncthreads();
  for (int v = tid; v < nvec; v += kRingThreads) {
    float4 gv = *reinterpret_cast<float4*>(&s_g[4 * v]);
    red_add_gpu_f32x4(p.g_local + 4 * v, gv);
  }
  round_finish<kRingThreads>(p, sh, round, loss_acc, cnt);
}

template <int NV4, int kDepth>
static cudaError_t launch_ring(const RoundParams& p, int grid, cudaStream_t stream) {
  const size_t stage = static_cast<size_t>(p.ld) * 4 + kRingMetaBytes;
  const size_t smem = static_cast<size_t>(kRingWarps) * kDepth * stage + static_cast<size_t>(p.ld) * 4 + 128;
  auto k = dense_round_ring_kernel<NV4, kDepth>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  k<<<grid, kRingThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace afb

extern "C" {

// rows wider than 896 floats would need > 113 KB per CTA at depth 4 (two CTAs must fit in 227 KB)
int afb_dense_round_ring_max_ld() { return 896; }

int afb_dense_round_ring_launch(const afb::RoundParams* params, int grid, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.ld % 4 != 0 || p.ld > 896 || p.ld <= 0) return static_cast<int>(cudaErrorInvalidValue);
  const int nv4 = static_cast<int>((p.ld / 4 + 31) / 32);
  cudaError_t e;
  switch (nv4) {
    case 1: e = launch_ring<1, 4>(p, grid, s); break;
    case 2: e = launch_ring<2, 4>(p, grid, s); break;
    case 3: case 4: e = launch_ring<4, 4>(p, grid, s); break;
    default: e = launch_ring<7, 4>(p, grid, s); break;
  }
  return static_cast<int>(e);
}

}  // extern "C"
