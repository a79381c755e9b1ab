// End-to-end step plumbing (FusedEngine.step): per step the host hands the device a round descriptor from pinned memory
// and reads the step's result record back into pinned memory.  Both copies happen EVERY step, but neither sits on the
// launching stream: the 16-byte descriptor of step i travels on an H2D stream while round i-1 still computes (the kernel of
// step i waits for it through an event), and the result record of step i is snapshotted by a one-CTA kernel behind the
// round (same stream, no engine hop) and leaves on a D2H stream while round i+1 already runs.  With the copies on the
// launching stream every step paid two copy-engine round trips (~8 us each on this box) around a ~50 us round at 8 GPUs.
//
// Reference behaviour being replaced: the driver thread that ships the task binary + broadcast id for each submission and
// takes the task result off the result queue (ASYNCsamples/.../SparkASGDThread.scala:253-299: submit, then
// `ASYNCcollectAll` on the driver).
#include <cuda_runtime.h>

#include <cstdint>
#include <new>

namespace afb {

struct E2eCtx {
  cudaStream_t h2d;
  cudaStream_t d2h;
  cudaEvent_t desc_ready[2];      // H2D of the slot's descriptor finished
  cudaEvent_t round_done[2];      // round + snapshot of the slot finished (launching stream)
  cudaEvent_t result_ready[2];    // D2H of the slot's record finished: the host may read it, the snapshot slot may be reused
  bool result_recorded[2];
};

__global__ void __launch_bounds__(64) e2e_snapshot_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int n16) {
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

}  // namespace afb

extern "C" {

void* afb_e2e_create() {
  auto* c = new (std::nothrow) afb::E2eCtx();
  if (c == nullptr) return nullptr;
  bool ok = cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < 2; ++i) {
    ok = cudaEventCreateWithFlags(&c->desc_ready[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->round_done[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->result_ready[i], cudaEventDisableTiming) == cudaSuccess;
    c->result_recorded[i] = false;
  }
  if (!ok) { delete c; return nullptr; }
  return c;
}

void afb_e2e_destroy(void* ctx) {
  auto* c = static_cast<afb::E2eCtx*>(ctx);
  if (c == nullptr) return;
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(c->desc_ready[i]);
    cudaEventDestroy(c->round_done[i]);
    cudaEventDestroy(c->result_ready[i]);
  }
  cudaStreamDestroy(c->h2d);
  cudaStreamDestroy(c->d2h);
  delete c;
}

// Before the round of this step is launched: ship its descriptor, make the launching stream wait for it.
int afb_e2e_pre(void* ctx, int slot, void* desc_dev, const void* desc_host_pinned, long long nbytes, void* main_stream) {
  auto* c = static_cast<afb::E2eCtx*>(ctx);
  slot &= 1;
  cudaError_t e = cudaMemcpyAsync(desc_dev, desc_host_pinned, static_cast<size_t>(nbytes), cudaMemcpyHostToDevice, c->h2d);
  if (e != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaEventRecord(c->desc_ready[slot], c->h2d)) != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaStreamWaitEvent(static_cast<cudaStream_t>(main_stream), c->desc_ready[slot], 0));
}

// After the round of this step was launched: snapshot the workers' result records behind it, send the snapshot home.
int afb_e2e_post(void* ctx, int slot, const void* wctl_dev, void* snap_dev, void* result_host_pinned, long long nbytes, void* main_stream) {
  auto* c = static_cast<afb::E2eCtx*>(ctx);
  slot &= 1;
  cudaStream_t ms = static_cast<cudaStream_t>(main_stream);
  cudaError_t e;
  if (nbytes <= 0 || (nbytes & 15)) return static_cast<int>(cudaErrorInvalidValue);
  if (c->result_recorded[slot]) {       // the D2H that last read this snapshot slot must be over before it is rewritten
    if ((e = cudaStreamWaitEvent(ms, c->result_ready[slot], 0)) != cudaSuccess) return static_cast<int>(e);
  }
  afb::e2e_snapshot_kernel<<<1, 64, 0, ms>>>(static_cast<const uint4*>(wctl_dev), static_cast<uint4*>(snap_dev), static_cast<int>(nbytes / 16));
  if ((e = cudaGetLastError()) != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaEventRecord(c->round_done[slot], ms)) != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaStreamWaitEvent(c->d2h, c->round_done[slot], 0)) != cudaSuccess) return static_cast<int>(e);
  if ((e = cudaMemcpyAsync(result_host_pinned, snap_dev, static_cast<size_t>(nbytes), cudaMemcpyDeviceToHost, c->d2h)) != cudaSuccess)
    return static_cast<int>(e);
  e = cudaEventRecord(c->result_ready[slot], c->d2h);
  c->result_recorded[slot] = (e == cudaSuccess);
  return static_cast<int>(e);
}

// Host side: block until the slot's record is in pinned memory.
int afb_e2e_wait(void* ctx, int slot) {
  auto* c = static_cast<afb::E2eCtx*>(ctx);
  slot &= 1;
  if (!c->result_recorded[slot]) return 0;
  return static_cast<int>(cudaEventSynchronize(c->result_ready[slot]));
}

// Drain the side streams (before tearing buffers down / synchronising the engine).
int afb_e2e_sync(void* ctx) {
  auto* c = static_cast<afb::E2eCtx*>(ctx);
  cudaError_t e = cudaStreamSynchronize(c->h2d);
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaStreamSynchronize(c->d2h));
}

}  // extern "C"
