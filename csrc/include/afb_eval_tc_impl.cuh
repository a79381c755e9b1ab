// tcgen05 / TMEM / TMA evaluation GEMM, templated on the loss applied in the epilogue; instantiated by csrc/kernels/eval_tc.cu
// for squared error, logistic, hinge and half-squared error (afb_eval_tc / afb_eval_loss_tc).
#pragma once
#include <algorithm>
#include "afb_common.cuh"
#include "afb_tc.cuh"

namespace afb {

constexpr int kTcStages = 4;
constexpr int kTcBlockM = 128;
constexpr int kTcBlockK = 32;               // fp32 elements = 128 bytes = one swizzle atom
constexpr int kTcThreads = 192;

template <int N>                             // N = 2 * snapshots in this pass (32 or 64)
struct TcSmem {
  alignas(1024) float a[kTcStages][kTcBlockM * kTcBlockK];
  alignas(1024) float b[kTcStages][N * kTcBlockK];
  alignas(8) uint64_t full[kTcStages];
  alignas(8) uint64_t empty[kTcStages];
  alignas(8) uint64_t tmem_full[2];
  alignas(8) uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <int N, int LOSS>
__global__ void __launch_bounds__(kTcThreads, 1)
eval_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b,
               const float* __restrict__ y, long long n_rows, int num_kb, int num_tiles, int T, double* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TcSmem<N>& sm = *reinterpret_cast<TcSmem<N>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * N <= 32) ? 32 : (2 * N <= 64 ? 64 : (2 * N <= 128 ? 128 : 256));
  constexpr uint32_t kStageBytes = (kTcBlockM + N) * kTcBlockK * sizeof(float);

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&map_x);
    tc::tma_prefetch_desc(&map_b);
    for (int s = 0; s < kTcStages; ++s) { tc::mbar_init(&sm.full[s], 1); tc::mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&sm.tmem_full[s], 1); tc::mbar_init(&sm.tmem_empty[s], 4); }
    tc::fence_barrier_init();
  } else if (warp == 1) {
    tc::tmem_alloc(&sm.tmem_base, kTmemCols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(&sm.empty[stage], phase ^ 1u);
          tc::mbar_arrive_expect_tx(&sm.full[stage], kStageBytes);
          tc::tma_load_2d(sm.a[stage], &map_x, &sm.full[stage], kb * kTcBlockK, tile * kTcBlockM);
          tc::tma_load_2d(sm.b[stage], &map_b, &sm.full[stage], kb * kTcBlockK, 0);
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(kTcBlockM, N);
      int stage = 0; uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        const uint32_t aphase = static_cast<uint32_t>(local >> 1) & 1u;
        tc::mbar_wait(&sm.tmem_empty[as], aphase ^ 1u);          // epilogue has drained this accumulator
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * N);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(&sm.full[stage], phase);                  // TMA bytes have landed
          tc::fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(sm.a[stage]);
          const uint32_t b_addr = tc::smem_u32(sm.b[stage]);
#pragma unroll
          for (int k = 0; k < kTcBlockK / 8; ++k) {               // UMMA_K = 8 for tf32 (32 bytes)
            const uint64_t da = tc::smem_desc_k_sw128(a_addr + k * 32);
            const uint64_t db = tc::smem_desc_k_sw128(b_addr + k * 32);
            tc::mma_tf32_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc::mma_commit(&sm.empty[stage]);                       // frees the smem stage when the MMAs finish
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
        tc::mma_commit(&sm.tmem_full[as]);                        // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue warps (TMEM -> registers -> squared error) =====
    const int q = warp & 3;                                       // TMEM lane quarter this warp may access
    float colsum[N / 2];
#pragma unroll
    for (int t = 0; t < N / 2; ++t) colsum[t] = 0.f;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int as = local & 1;
      const uint32_t aphase = static_cast<uint32_t>(local >> 1) & 1u;
      const long long row = static_cast<long long>(tile) * kTcBlockM + 32 * q + lane;
      const bool valid = row < n_rows;
      const float yy = valid ? __ldg(y + row) : 0.f;
      tc::mbar_wait(&sm.tmem_full[as], aphase);
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + static_cast<uint32_t>(as * N);
      // columns [0, N/2) = X·w_hi, columns [N/2, N) = X·w_lo
#pragma unroll
      for (int c = 0; c < N / 32; ++c) {
        uint32_t hi[16], lo[16];
        tc::tmem_ld_32x32b_x16(taddr + 16 * c, hi);
        tc::tmem_ld_32x32b_x16(taddr + N / 2 + 16 * c, lo);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float m = __uint_as_float(hi[j]) + __uint_as_float(lo[j]);
          if constexpr (LOSS == LOSS_LS) {
            const float dlt = m - yy;
            colsum[16 * c + j] += valid ? dlt * dlt : 0.f;
          } else {                                                  // logistic / hinge / half-squared: same K2 formulas as every
            float coeff, loss;                                      // other kernel (afb_common.cuh::loss_coeff)
            loss_coeff<LOSS>(m, yy, coeff, loss);
            colsum[16 * c + j] += valid ? loss : 0.f;
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sm.tmem_empty[as]);
    }
#pragma unroll
    for (int t = 0; t < N / 2; ++t) {
      const float s = warp_sum(colsum[t]);
      if (lane == 0 && t < T) atomicAdd(out + t, static_cast<double>(s));
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// split each snapshot into tf32 "hi" (low 13 mantissa bits cleared) and the fp32 remainder "lo"
static __global__ void split_w_kernel(const float* __restrict__ W, long long ld, int T, int Tpad, float* __restrict__ B) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long total = static_cast<long long>(Tpad) * ld;
  if (i >= total) return;
  int t = static_cast<int>(i / ld);
  long long c = i % ld;
  float w = (t < T) ? W[static_cast<long long>(t) * ld + c] : 0.f;
  float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
  B[static_cast<long long>(t) * ld + c] = hi;
  B[static_cast<long long>(Tpad + t) * ld + c] = w - hi;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map_2d(CUtensorMap* m, const float* base, long long inner, long long outer, long long row_stride_bytes,
                       int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_stride_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return static_cast<int>(r);
}

template <int N, int LOSS>
static int launch_pass(const float* X, const float* y, long long n, long long ld, const float* B, int T, double* out, cudaStream_t s) {
  CUtensorMap mx, mb;
  if (make_map_2d(&mx, X, ld, n, ld * 4, kTcBlockK, kTcBlockM) != 0) return 901;
  if (make_map_2d(&mb, B, ld, N, ld * 4, kTcBlockK, N) != 0) return 902;
  const int num_kb = static_cast<int>((ld + kTcBlockK - 1) / kTcBlockK);
  const int num_tiles = static_cast<int>((n + kTcBlockM - 1) / kTcBlockM);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(sms, num_tiles);
  const size_t smem = sizeof(TcSmem<N>) + 1024;
  auto k = eval_tc_kernel<N, LOSS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  k<<<grid, kTcThreads, smem, s>>>(mx, mb, y, n, num_kb, num_tiles, T, out);
  return static_cast<int>(cudaGetLastError());
}


// W: [T][ld] snapshots (row-major, zero padded to ld); scratch: >= 64*ld floats; out: double[T] (+=).
template <int LOSS>
static int eval_tc_all(const float* X, const float* y, long long n, long long ld, const float* W, int T, double* out, float* scratch,
                       cudaStream_t s) {
  if (ld % 4 != 0 || n <= 0) return static_cast<int>(cudaErrorInvalidValue);
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int tt = std::min(32, T - t0);
    const int tpad = tt <= 16 ? 16 : 32;
    const long long total = static_cast<long long>(tpad) * ld;
    split_w_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(W + static_cast<long long>(t0) * ld, ld, tt, tpad, scratch);
    int rc;
    if (tpad == 16) rc = launch_pass<32, LOSS>(X, y, n, ld, scratch, tt, out + t0, s);
    else rc = launch_pass<64, LOSS>(X, y, n, ld, scratch, tt, out + t0, s);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace afb
