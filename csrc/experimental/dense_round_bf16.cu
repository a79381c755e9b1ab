// EXPERIMENTAL — compiled into libafb_experimental.so, never loaded by the default path.
// Status: builds for sm_100a; NOT yet validated on hardware.
//
// Fused dense round with bf16 ROW STORAGE and fp32 math: the same kernel as csrc/kernels/dense_round.cu
// (pull -> sample -> x.w -> loss coefficient -> g += c.x -> push-and-apply; protocol code shared through
// afb_round_impl.cuh) but every row is half the bytes.  The gradient phase is HBM-bound (96.7 % of the measured copy
// bandwidth with fp32 rows), so halving the row bytes is the one remaining 2x on a single GPU.  Parameters, gradient,
// accumulators and the server state stay fp32; a row element is widened with one shift (bf16 is the upper half of an
// fp32).  For integer-pixel data (mnist8m: 0..255) bf16 storage is exact.
//
// Layout: X is [n_rows, ld] bf16 row-major with ld % 8 == 0 (a row is a whole number of 16-byte chunks); lane l of the
// warp that owns a row holds chunks {l, l+32, ...}: 8 elements each, widened to fp32 on use; w lives in shared memory as
// fp32 and is read with two LDS.128 per chunk.
#include <cuda_bf16.h>

#include "afb_dense_rows.cuh"

namespace afb {

template <int NV8>
struct RowRegsH {
  uint4 v[NV8];          // 8 bf16 each
  float y;
  float h;
};

AFB_DEVINL uint4 ld_stream_u32x4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int NV8>
AFB_DEVINL void load_row_h(RowRegsH<NV8>& r, const RoundParams& p, int row, int lane, int nchunk, bool saga) {
  const unsigned short* xr = reinterpret_cast<const unsigned short*>(p.X) + static_cast<long long>(row) * p.ld;
#pragma unroll
  for (int c = 0; c < NV8; ++c) {
    int v = lane + 32 * c;
    r.v[c] = (v < nchunk) ? ld_stream_u32x4(xr + 8 * v) : make_uint4(0u, 0u, 0u, 0u);
  }
  r.y = __ldg(p.y + row);
  r.h = saga ? p.hist[row] : 0.f;
}

AFB_DEVINL float bf_lo(unsigned int u) { return __uint_as_float(u << 16); }
AFB_DEVINL float bf_hi(unsigned int u) { return __uint_as_float(u & 0xffff0000u); }

template <int NV8, int kThreads>
__global__ void __launch_bounds__(kThreads, (NV8 <= 4) ? 2 : 1) dense_round_bf16_kernel(const RoundParams p) {
  __shared__ RoundShared<kThreads> sh;
  extern __shared__ __align__(16) float s_dyn[];
  float* s_g = s_dyn;
  float* s_w = s_dyn + p.ld;
  const bool SAGA = (p.hist != nullptr);
  const int LOSS = p.loss_id;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nchunk = static_cast<int>(p.ld >> 3);
  WorkerCtl* wc = p.wctl;
  const unsigned int round = *reinterpret_cast<volatile unsigned int*>(&wc->round);

  float acc[NV8][8];
#pragma unroll
  for (int c = 0; c < NV8; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  float loss_acc = 0.f;
  unsigned int cnt = 0;
  RowScheduler sch;
  sched_init(sch, p.n_rows, gridDim.x * (kThreads / 32), blockIdx.x * (kThreads / 32) + (tid >> 5));
  RowRegsH<NV8> cur, nxt;
  int rc = -1;
  unsigned long long key = 0;

  const bool early = (p.algo != ALGO_GRAD_ONLY) && (p.seed_mode == 1 || p.seed_mode == 2);
  if (early) {
    unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
    if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
    key = mix64(seed_eff);
    const unsigned int c0 = sch.next_chunk;
    sch.next_chunk = 0xffffffffu;
    if (c0 < static_cast<unsigned int>(sch.num_chunks)) {
      load_chunk_masks(sch, p, c0, lane, key);
      if (sch.m0) { int b = __ffs(sch.m0) - 1; sch.m0 &= sch.m0 - 1; rc = sch.base + b; }
      else if (sch.m1) { int b = __ffs(sch.m1) - 1; sch.m1 &= sch.m1 - 1; rc = sch.base + 32 + b; }
      if (rc >= 0) load_row_h<NV8>(cur, p, rc, lane, nchunk, SAGA);
    }
  }
  RoundPull pl;
  round_begin_issue<kThreads>(p, sh, round, pl);
  if (early) {
    sch.next_chunk = fetch_chunk(&wc->chunk_ctr, lane, sch.static_chunks);
    if (rc < 0) {
      rc = next_row(sch, p, &wc->chunk_ctr, lane, key);
      if (rc >= 0) load_row_h<NV8>(cur, p, rc, lane, nchunk, SAGA);
    }
  }
  if (!round_begin_complete<kThreads>(p, sh, round, pl)) return;
  if (!early) {
    key = (p.algo == ALGO_GRAD_ONLY) ? mix64(static_cast<unsigned long long>(p.seed))
                                     : *reinterpret_cast<volatile unsigned long long*>(&wc->seed_key);
  }
  for (int i = tid; i < p.ld; i += kThreads) { s_g[i] = 0.f; s_w[i] = p.w_local[i]; }
  __syncthreads();
  if (!early) {
    rc = next_row(sch, p, &wc->chunk_ctr, lane, key);
    if (rc >= 0) load_row_h<NV8>(cur, p, rc, lane, nchunk, SAGA);
  }
  while (rc >= 0) {
    int rn = next_row(sch, p, &wc->chunk_ctr, lane, key);
    if (rn >= 0) load_row_h<NV8>(nxt, p, rn, lane, nchunk, SAGA);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int c = 0; c < NV8; ++c) {
      int v = lane + 32 * c;
      if (v < nchunk) {
        const float4 wa = *reinterpret_cast<const float4*>(s_w + 8 * v);
        const float4 wb = *reinterpret_cast<const float4*>(s_w + 8 * v + 4);
        const uint4 x = cur.v[c];
        d0 = fmaf(bf_lo(x.x), wa.x, d0); d1 = fmaf(bf_hi(x.x), wa.y, d1);
        d2 = fmaf(bf_lo(x.y), wa.z, d2); d3 = fmaf(bf_hi(x.y), wa.w, d3);
        d0 = fmaf(bf_lo(x.z), wb.x, d0); d1 = fmaf(bf_hi(x.z), wb.y, d1);
        d2 = fmaf(bf_lo(x.w), wb.z, d2); d3 = fmaf(bf_hi(x.w), wb.w, d3);
      }
    }
    float margin = warp_sum((d0 + d1) + (d2 + d3));
    float coeff, loss;
    loss_coeff_rt(LOSS, margin, cur.y, coeff, loss);
    float eff = coeff;
    if (SAGA) {
      eff = coeff - cur.h;
      if (lane == 0) {
        if (p.hist_prev) { p.hist_prev[rc] = cur.h; p.hist_tag[rc] = static_cast<int>(round); }
        p.hist[rc] = coeff;
      }
    }
#pragma unroll
    for (int c = 0; c < NV8; ++c) {
      const uint4 x = cur.v[c];
      acc[c][0] = fmaf(eff, bf_lo(x.x), acc[c][0]); acc[c][1] = fmaf(eff, bf_hi(x.x), acc[c][1]);
      acc[c][2] = fmaf(eff, bf_lo(x.y), acc[c][2]); acc[c][3] = fmaf(eff, bf_hi(x.y), acc[c][3]);
      acc[c][4] = fmaf(eff, bf_lo(x.z), acc[c][4]); acc[c][5] = fmaf(eff, bf_hi(x.z), acc[c][5]);
      acc[c][6] = fmaf(eff, bf_lo(x.w), acc[c][6]); acc[c][7] = fmaf(eff, bf_hi(x.w), acc[c][7]);
    }
    loss_acc += loss;
    ++cnt;
    cur = nxt;
    rc = rn;
  }
#pragma unroll
  for (int c = 0; c < NV8; ++c) {
    int v = lane + 32 * c;
    if (v < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_g[8 * v + j], acc[c][j]);
    }
  }
  __syncthreads();
  for (int v = tid; v < (p.ld >> 2); v += kThreads) {
    float4 gv = *reinterpret_cast<float4*>(&s_g[4 * v]);
    red_add_gpu_f32x4(p.g_local + 4 * v, gv);
  }
  round_finish<kThreads>(p, sh, round, loss_acc, cnt);
}

template <int NV8, int kThreads>
static cudaError_t launch_bf16(const RoundParams& p, int grid, cudaStream_t stream) {
  size_t smem = 2 * static_cast<size_t>(p.ld) * sizeof(float);
  auto k = dense_round_bf16_kernel<NV8, kThreads>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  k<<<grid, kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace afb

extern "C" {

int afb_dense_round_bf16_max_ld() { return 2048; }
int afb_dense_round_bf16_ctas_per_sm(long long ld) { return ((ld / 8 + 31) / 32) <= 4 ? 2 : 1; }

// `params->X` points to bf16 rows ([n_rows, ld] row-major, ld % 8 == 0); everything else is as in afb_dense_round_launch.
int afb_dense_round_bf16_launch(const afb::RoundParams* params, int grid, void* stream) {
  using namespace afb;
  const RoundParams& p = *params;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.ld % 8 != 0 || p.ld > 2048 || p.ld <= 0 || p.pdl != 0 || p.seed_mode == 3) return static_cast<int>(cudaErrorInvalidValue);
  int nv8 = static_cast<int>((p.ld / 8 + 31) / 32);
  cudaError_t e;
  switch (nv8) {
    case 1: e = launch_bf16<1, 256>(p, grid, s); break;
    case 2: e = launch_bf16<2, 256>(p, grid, s); break;
    case 3: case 4: e = launch_bf16<4, 256>(p, grid, s); break;
    case 5: case 6: e = launch_bf16<6, 256>(p, grid, s); break;
    default: e = launch_bf16<8, 256>(p, grid, s); break;
  }
  return static_cast<int>(e);
}

}  // extern "C"
