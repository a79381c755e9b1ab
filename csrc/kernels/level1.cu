// Level-1 BLAS (axpy / dot / scal / copy) and the Updater family (K8) as sm_100a kernels.
// These back asyncframework_b200.linalg.blas_util on CUDA tensors and optim.updater on the GPU
// paths.  d is small (<= 47 236 in every reference config) so these are latency-bound single-wave
// kernels: one launch, vectorised accesses, warp-shuffle + one atomic per block for reductions.
#include "afb_common.cuh"

namespace afb {

__global__ void axpy_kernel(float a, const float* __restrict__ x, float* __restrict__ y, long long n) {
  long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  for (; i < n; i += stride) {
    if (aligned && i + 3 < n) {
      float4 xv = *reinterpret_cast<const float4*>(x + i);
      float4 yv = *reinterpret_cast<float4*>(y + i);
      yv.x = fmaf(a, xv.x, yv.x); yv.y = fmaf(a, xv.y, yv.y); yv.z = fmaf(a, xv.z, yv.z); yv.w = fmaf(a, xv.w, yv.w);
      *reinterpret_cast<float4*>(y + i) = yv;
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) y[j] = fmaf(a, x[j], y[j]);
    }
  }
}

__global__ void scal_kernel(float a, float* __restrict__ x, long long n) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) x[i] *= a;
}

// out[0] += Σ x_i y_i  (fp32 partials, double final accumulation)
__global__ void dot_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n, double* out) {
  __shared__ float s[32];
  float acc = 0.f;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) acc = fmaf(x[i], y[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? s[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, static_cast<double>(v));
  }
}

// Updater kernels: w_new = f(w, g, eta, lambda); norm_out accumulates ||w_new||_1 or ||w_new||_2^2.
// mode 0: simple     w - eta*g
// mode 1: L1 prox    soft_threshold(w - eta*g, eta*lambda)
// mode 2: squared L2 (1 - eta*lambda)*w - eta*g
__global__ void updater_kernel(int mode, const float* __restrict__ w, const float* __restrict__ g, float* __restrict__ out,
                               long long n, float eta, float lambda, double* norm_out) {
  __shared__ float s[32];
  float acc = 0.f;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) {
    float v;
    if (mode == 0) {
      v = fmaf(-eta, g[i], w[i]);
    } else if (mode == 1) {
      float t = fmaf(-eta, g[i], w[i]);
      float sh = eta * lambda;
      v = copysignf(fmaxf(fabsf(t) - sh, 0.f), t);
      acc += fabsf(v);
    } else {
      v = fmaf(-eta, g[i], w[i] * (1.f - eta * lambda));
      acc = fmaf(v, v, acc);
    }
    out[i] = v;
  }
  if (mode != 0 && norm_out) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = (threadIdx.x < (blockDim.x >> 5)) ? s[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(norm_out, static_cast<double>(v));
    }
  }
}

}  // namespace afb

extern "C" {

static inline int grid_for(long long n, int threads, int per_thread) {
  long long g = (n + static_cast<long long>(threads) * per_thread - 1) / (static_cast<long long>(threads) * per_thread);
  if (g < 1) g = 1;
  if (g > 148 * 8) g = 148 * 8;
  return static_cast<int>(g);
}

int afb_axpy(float a, const float* x, float* y, long long n, void* stream) {
  afb::axpy_kernel<<<grid_for(n, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, x, y, n);
  return static_cast<int>(cudaGetLastError());
}
int afb_scal(float a, float* x, long long n, void* stream) {
  afb::scal_kernel<<<grid_for(n, 256, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, x, n);
  return static_cast<int>(cudaGetLastError());
}
int afb_dot(const float* x, const float* y, long long n, double* out, void* stream) {
  cudaMemsetAsync(out, 0, sizeof(double), static_cast<cudaStream_t>(stream));
  afb::dot_kernel<<<grid_for(n, 256, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, out);
  return static_cast<int>(cudaGetLastError());
}
int afb_updater(int mode, const float* w, const float* g, float* out, long long n, float eta, float lambda,
                double* norm_out, void* stream) {
  if (norm_out) cudaMemsetAsync(norm_out, 0, sizeof(double), static_cast<cudaStream_t>(stream));
  afb::updater_kernel<<<grid_for(n, 256, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(mode, w, g, out, n, eta, lambda, norm_out);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
