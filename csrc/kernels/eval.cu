// Objective evaluation (K12): for T parameter snapshots W[T][ld] compute  out[t] = Σ_i loss(x_i·w_t, y_i)
// in ONE pass over the shard — the reference does `points.map(...).reduce(_+_)` producing a length-T
// vector (SparkASGDThread.scala:386-401).  This file holds the CUDA-core variants (any d, dense or
// CSR, exact fp32); eval_tc.cu holds the tcgen05/TMEM GEMM variant for dense shards.
#include <algorithm>
#include "afb_common.cuh"

namespace afb {

constexpr int kEvalThreads = 256;
constexpr int kEvalMaxNV4 = 16;    // ld <= 2048 floats

__global__ void __launch_bounds__(kEvalThreads) eval_dense_kernel(const float* __restrict__ X, const float* __restrict__ y,
                                                                  long long n, long long ld, const float* __restrict__ W,
                                                                  int T, int t_tile, int loss_id, double* out) {
  extern __shared__ __align__(16) float s_w[];      // [t_tile][ld]
  __shared__ float s_part[kEvalThreads / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = static_cast<int>(ld >> 2);
  const long long warps_total = static_cast<long long>(gridDim.x) * (kEvalThreads / 32);
  const long long gw = static_cast<long long>(blockIdx.x) * (kEvalThreads / 32) + warp;
  for (int t0 = 0; t0 < T; t0 += t_tile) {
    const int tt = min(t_tile, T - t0);
    __syncthreads();
    for (long long i = threadIdx.x; i < static_cast<long long>(tt) * ld; i += kEvalThreads) s_w[i] = W[static_cast<long long>(t0) * ld + i];
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) acc[t] = 0.f;
    for (long long row = gw; row < n; row += warps_total) {
      float4 xr[kEvalMaxNV4];
      const float* xp = X + row * ld;
#pragma unroll
      for (int c = 0; c < kEvalMaxNV4; ++c) {
        int v = lane + 32 * c;
        xr[c] = (v < nvec) ? ld_stream_f32x4(xp + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float yy = __ldg(y + row);
      for (int t = 0; t < tt; ++t) {
        const float* wt = s_w + static_cast<long long>(t) * ld;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < kEvalMaxNV4; ++c) {
          int v = lane + 32 * c;
          if (v < nvec) {
            float4 wv = *reinterpret_cast<const float4*>(wt + 4 * v);
            d = fmaf(xr[c].x, wv.x, d); d = fmaf(xr[c].y, wv.y, d); d = fmaf(xr[c].z, wv.z, d); d = fmaf(xr[c].w, wv.w, d);
          }
        }
        d = warp_sum(d);
        float coeff, loss;
        loss_coeff_rt(loss_id, d, yy, coeff, loss);
        // lane t accumulates snapshot t (t < 32 per tile)
        acc[t & 31] += loss;
      }
    }
    // every lane holds identical acc[t]; lane t publishes it
#pragma unroll
    for (int t = 0; t < 32; ++t) if (lane == t) s_part[warp][t] = (t < tt) ? acc[t] : 0.f;
    __syncthreads();
    if (warp == 0 && lane < tt) {
      float s = 0.f;
      for (int w2 = 0; w2 < kEvalThreads / 32; ++w2) s += s_part[w2][lane];
      atomicAdd(out + t0 + lane, static_cast<double>(s));
    }
  }
}

__global__ void __launch_bounds__(kEvalThreads) eval_csr_kernel(const long long* __restrict__ indptr, const int* __restrict__ indices,
                                                                const float* __restrict__ values, const float* __restrict__ y,
                                                                long long n, long long ld, const float* __restrict__ W, int T,
                                                                int loss_id, double* out) {
  __shared__ double s_acc[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warps_total = static_cast<long long>(gridDim.x) * (kEvalThreads / 32);
  const long long gw = static_cast<long long>(blockIdx.x) * (kEvalThreads / 32) + warp;
  for (int t0 = 0; t0 < T; t0 += 64) {
    const int tt = min(64, T - t0);
    if (threadIdx.x < 64) s_acc[threadIdx.x] = 0.0;
    __syncthreads();
    for (int t = 0; t < tt; ++t) {
      const float* wt = W + static_cast<long long>(t0 + t) * ld;
      float lsum = 0.f;
      for (long long row = gw; row < n; row += warps_total) {
        const long long s = indptr[row], e = indptr[row + 1];
        float d = 0.f;
        for (long long k = s + lane; k < e; k += 32) d = fmaf(values[k], __ldg(wt + indices[k]), d);
        d = warp_sum(d);
        float coeff, loss;
        loss_coeff_rt(loss_id, d, __ldg(y + row), coeff, loss);
        lsum += loss;
      }
      if (lane == 0) atomicAdd(&s_acc[t], static_cast<double>(lsum));
    }
    __syncthreads();
    if (threadIdx.x < tt) atomicAdd(out + t0 + threadIdx.x, s_acc[threadIdx.x]);
    __syncthreads();
  }
}

}  // namespace afb

extern "C" {

int afb_eval_mse_dense(const float* X, const float* y, long long n, long long ld, int d, const float* W, int T,
                       int loss_id, double* out, void* stream) {
  (void)d;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ld % 4 != 0 || ld > 2048) return static_cast<int>(cudaErrorInvalidValue);
  cudaMemsetAsync(out, 0, sizeof(double) * T, s);
  if (n == 0) return 0;
  int t_tile = static_cast<int>(std::min<long long>(32, std::max<long long>(1, (200 * 1024) / (ld * 4))));
  t_tile = std::min(t_tile, T);
  size_t smem = static_cast<size_t>(t_tile) * ld * sizeof(float);
  cudaFuncSetAttribute(afb::eval_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  int grid = static_cast<int>(std::min<long long>(148 * 2, (n + 7) / 8));
  afb::eval_dense_kernel<<<grid, afb::kEvalThreads, smem, s>>>(X, y, n, ld, W, T, t_tile, loss_id, out);
  return static_cast<int>(cudaGetLastError());
}

int afb_eval_mse_csr(const long long* indptr, const int* indices, const float* values, const float* y, long long n,
                     int ld, const float* W, int T, int loss_id, double* out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(out, 0, sizeof(double) * T, s);
  if (n == 0) return 0;
  int grid = static_cast<int>(std::min<long long>(148 * 4, (n + 7) / 8));
  afb::eval_csr_kernel<<<grid, afb::kEvalThreads, 0, s>>>(indptr, indices, values, y, n, ld, W, T, loss_id, out);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
