// Fused dense round kernel, TMA-pipelined variant (sm_100a).
//
// Same phases and semantics as dense_round.cu; phase 1 (the HBM-bound gradient) is restructured as
// a warp-specialised producer/consumer pipeline over the TMA engine:
//
//   producer warp : dynamic 64-row chunks -> counter-hash Bernoulli test -> for every selected row
//                   (lane-parallel) wait for a free stage, then `cp.async.bulk` (SASS: UBLKCP) the
//                   row (ld*4 B) + the 16-byte quads holding y[row] / hist[row] from HBM straight into
//                   the stage, completing on the stage's mbarrier.  No registers, no LSU involved.
//   consumer warps: wait on the stage's mbarrier, LDS.128 the row, dot with w (registers) -> shuffle
//                   tree -> loss coefficient (-hist for SAGA) -> axpy into register accumulators,
//                   release the stage.
//
// 32 stages of one row each (~100 KB) are in flight per CTA, independent of occupancy, so ONE CTA per
// SM saturates HBM; with 2 CTAs/SM two logical workers co-run and each worker's NVLink pull / push
// latency is hidden behind the other's streaming.
#include <algorithm>
#include "afb_round_impl.cuh"
#include "afb_tc.cuh"

namespace afb {

constexpr int kPipeConsumers = 7;                 // + 1 producer warp = 256 threads: 2 CTAs/SM at 128 regs
constexpr int kPipeThreads = 32 * (kPipeConsumers + 1);
constexpr int kPipeChunkRows = 64;
constexpr int kPipeMetaBytes = 64;        // {int row; int pad[3]; float yq[4]; float hq[4]; pad}

struct PipeMeta { int row; int pad[3]; float yq[4]; float hq[4]; int pad2[4]; };
static_assert(sizeof(PipeMeta) == kPipeMetaBytes, "meta layout");

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

template <int NV4>
__global__ void __launch_bounds__(kPipeThreads, 2) dense_round_pipe_kernel(const RoundParams p, int num_stages) {
  __shared__ RoundShared<kPipeThreads> sh;
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvec = static_cast<int>(p.ld >> 2);
  const uint32_t row_bytes = static_cast<uint32_t>(p.ld * 4);
  const uint32_t stage_bytes = row_bytes + kPipeMetaBytes;
  // smem carve-up: [stages][row|meta] | g[ld] | full[S] | empty[S]
  uint8_t* s_stage = s_raw;
  float* s_g = reinterpret_cast<float*>(s_raw + static_cast<size_t>(num_stages) * stage_bytes);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_g + p.ld);
  uint64_t* s_empty = s_full + num_stages;
  WorkerCtl* wc = p.wctl;
  const unsigned int round = wc->round;
  const bool saga = (p.hist != nullptr);

  if (tid == 0) {
    for (int s = 0; s < num_stages; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], 1); }
    tc::fence_barrier_init();
  }
  if (p.algo != ALGO_GRAD_ONLY) {
    if (round_begin<kPipeThreads>(p, sh, round) != PROCEED_RUN) return;       // contains __syncthreads(): barrier init is visible after it
  }
  if (p.seed_mode == 3 && !*reinterpret_cast<volatile int*>(&wc->proceed)) return;   // split round that was stopped at the gate
  const unsigned long long key = (p.algo == ALGO_GRAD_ONLY && p.seed_mode != 3)
                                     ? mix64(static_cast<unsigned long long>(p.seed))
                                     : *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
  for (int i = tid; i < p.ld; i += kPipeThreads) s_g[i] = 0.f;
  __syncthreads();

  float4 acc[NV4];
#pragma unroll
  for (int c = 0; c < NV4; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  float loss_acc = 0.f;
  unsigned int cnt = 0;

  if (warp == kPipeConsumers) {
    // =============================== producer warp ===============================
    const int num_chunks = (p.n_rows + kPipeChunkRows - 1) / kPipeChunkRows;
    unsigned int fill = 0;                                          // warp-uniform count of stages filled so far
    unsigned int chunk = fetch_chunk_ctr(&wc->chunk_ctr, lane);
    while (chunk < static_cast<unsigned int>(num_chunks)) {
      const unsigned int next = fetch_chunk_ctr(&wc->chunk_ctr, lane);   // requested one ahead
      const int base = static_cast<int>(chunk) * kPipeChunkRows;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = base + 32 * half + lane;
        bool sel = false;
        if (r < p.n_rows) {
          unsigned long long g = p.gidx ? static_cast<unsigned long long>(p.gidx[r]) : static_cast<unsigned long long>(p.row_base + r);
          sel = hash_u24(key, g) < static_cast<unsigned int>(p.thr24);
        }
        const unsigned int m = __ballot_sync(0xffffffffu, sel);
        const int nsel = __popc(m);
        if (lane < nsel) {
          const int row = base + 32 * half + static_cast<int>(__fns(m, 0, lane + 1));
          const unsigned int f = fill + static_cast<unsigned int>(lane);
          const int slot = static_cast<int>(f % static_cast<unsigned int>(num_stages));
          const uint32_t parity = (f / static_cast<unsigned int>(num_stages)) & 1u;
          tc::mbar_wait(&s_empty[slot], parity ^ 1u);               // stage free?
          uint8_t* st = s_stage + static_cast<size_t>(slot) * stage_bytes;
          PipeMeta* meta = reinterpret_cast<PipeMeta*>(st + row_bytes);
          meta->row = row;
          const bool quad_ok = (row | 3) < p.n_rows;                 // 16-byte quad fully inside y / hist
          uint32_t tx = row_bytes;
          if (quad_ok) tx += saga ? 32u : 16u;
          else {
            meta->yq[row & 3] = __ldg(p.y + row);
            if (saga) meta->hq[row & 3] = p.hist[row];
          }
          tc::mbar_arrive_expect_tx(&s_full[slot], tx);
          bulk_g2s(st, p.X + static_cast<long long>(row) * p.ld, row_bytes, &s_full[slot]);
          if (quad_ok) {
            bulk_g2s(meta->yq, p.y + (row & ~3), 16u, &s_full[slot]);
            if (saga) bulk_g2s(meta->hq, p.hist + (row & ~3), 16u, &s_full[slot]);
          }
        }
        fill += static_cast<unsigned int>(nsel);
      }
      chunk = next;
    }
    // one termination sentinel per consumer warp (consecutive fill indices map to distinct consumers)
    if (lane < kPipeConsumers) {
      const unsigned int f = fill + static_cast<unsigned int>(lane);
      const int slot = static_cast<int>(f % static_cast<unsigned int>(num_stages));
      const uint32_t parity = (f / static_cast<unsigned int>(num_stages)) & 1u;
      tc::mbar_wait(&s_empty[slot], parity ^ 1u);
      PipeMeta* meta = reinterpret_cast<PipeMeta*>(s_stage + static_cast<size_t>(slot) * stage_bytes + row_bytes);
      meta->row = -1;
      tc::mbar_arrive(&s_full[slot]);
    }
  } else {
    // =============================== consumer warps ===============================
    float4 wreg[NV4];
#pragma unroll
    for (int c = 0; c < NV4; ++c) {
      const int v = lane + 32 * c;
      wreg[c] = (v < nvec) ? *reinterpret_cast<const float4*>(p.w_local + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (unsigned int f = static_cast<unsigned int>(warp);; f += kPipeConsumers) {
      const int slot = static_cast<int>(f % static_cast<unsigned int>(num_stages));
      const uint32_t parity = (f / static_cast<unsigned int>(num_stages)) & 1u;
      tc::mbar_wait(&s_full[slot], parity);
      const uint8_t* st = s_stage + static_cast<size_t>(slot) * stage_bytes;
      const PipeMeta* meta = reinterpret_cast<const PipeMeta*>(st + row_bytes);
      const int row = meta->row;
      if (row < 0) break;
      const float yy = meta->yq[row & 3];
      const float hh = saga ? meta->hq[row & 3] : 0.f;
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
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&s_empty[slot]);               // row is in registers: release the stage early
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
  }
  __syncthreads();
  for (int v = tid; v < nvec; v += kPipeThreads) {
    float4 gv = *reinterpret_cast<float4*>(&s_g[4 * v]);
    red_add_gpu_f32x4(p.g_local + 4 * v, gv);
  }
  round_finish<kPipeThreads>(p, sh, round, loss_acc, cnt);
}

template <int NV4>
static cudaError_t launch_pipe(const RoundParams& p, int grid, cudaStream_t stream) {
  const size_t stage = static_cast<size_t>(p.ld) * 4 + kPipeMetaBytes;
  // two CTAs per SM must fit: ~110 KB each
  int stages = static_cast<int>((104 * 1024 - static_cast<size_t>(p.ld) * 4) / stage);
  stages = std::min(35, stages) / kPipeConsumers * kPipeConsumers;
  if (stages < kPipeConsumers) return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(stages) * stage + static_cast<size_t>(p.ld) * 4 + 2 * static_cast<size_t>(stages) * 8 + 128;
  auto k = dense_round_pipe_kernel<NV4>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  k<<<grid, kPipeThreads, smem, stream>>>(p, stages);
  return cudaGetLastError();
}

}  // namespace afb

extern "C" {

int afb_dense_round_pipe_max_ld() { return 896; }   // NV4 <= 7 validated on B200 (A/B arm, not a default)

int afb_dense_round_pipe_launch(const afb::RoundParams* params, int grid, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.ld % 4 != 0 || p.ld > 896 || p.ld <= 0) return static_cast<int>(cudaErrorInvalidValue);
  const int nv4 = static_cast<int>((p.ld / 4 + 31) / 32);
  cudaError_t e;
  switch (nv4) {
    case 1: e = launch_pipe<1>(p, grid, s); break;
    case 2: e = launch_pipe<2>(p, grid, s); break;
    case 3: case 4: e = launch_pipe<4>(p, grid, s); break;
    case 5: case 6: case 7: e = launch_pipe<7>(p, grid, s); break;
    default: e = launch_pipe<8>(p, grid, s); break;
  }
  return static_cast<int>(e);
}

}  // extern "C"
