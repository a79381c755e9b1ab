// Small utility kernels + ABI self-description + CUDA-IPC / peer-access helpers (C ABI).
#include <cstring>
#include "afb_round.cuh"

namespace afb {

// K11 as a standalone op: stream-compact the selected local row positions of a shard into `out`
// (ascending within a block, blocks in order => globally ascending).  Two passes: count per block,
// exclusive scan (single block), write.  Used by the general `ShardedDataset.sample` on HBM shards.
__global__ void sample_count_kernel(const long long* gidx, long long row_base, long long n, unsigned long long key,
                                    int thr24, int* block_counts) {
  __shared__ int s[32];
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  bool sel = false;
  if (i < n) {
    unsigned long long g = gidx ? static_cast<unsigned long long>(gidx[i]) : static_cast<unsigned long long>(row_base + i);
    sel = hash_u24(key, g) < static_cast<unsigned int>(thr24);
  }
  unsigned int m = __ballot_sync(0xffffffffu, sel);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += s[w];
    block_counts[blockIdx.x] = t;
  }
}

__global__ void sample_scan_kernel(int* block_counts, int nblocks, long long* total) {
  // single thread block exclusive scan (nblocks <= a few thousand)
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < nblocks) ? block_counts[i] : 0;
    // inclusive warp scan
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    __shared__ int wsum[32];
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int t = (threadIdx.x < (blockDim.x >> 5)) ? wsum[threadIdx.x] : 0;
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (threadIdx.x >= o) t += y; }
      wsum[threadIdx.x] = t;
    }
    __syncthreads();
    int woff = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0;
    long long excl = carry + woff + x - v;
    if (i < nblocks) block_counts[i] = static_cast<int>(excl);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void sample_write_kernel(const long long* gidx, long long row_base, long long n, unsigned long long key,
                                    int thr24, const int* block_offsets, long long* out) {
  __shared__ int s[32];
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  bool sel = false;
  if (i < n) {
    unsigned long long g = gidx ? static_cast<unsigned long long>(gidx[i]) : static_cast<unsigned long long>(row_base + i);
    sel = hash_u24(key, g) < static_cast<unsigned int>(thr24);
  }
  unsigned int m = __ballot_sync(0xffffffffu, sel);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s[warp] = __popc(m);
  __syncthreads();
  int off = block_offsets[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s[w];
  if (sel) out[off + __popc(m & ((1u << lane) - 1u))] = i;
}

// write a buffer larger than L2 (126 MB) between timed iterations
__global__ void l2_flush_kernel(float4* buf, long long n4) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

__global__ void sleep_kernel(long long ns) { if (threadIdx.x == 0) sleep_ns(static_cast<unsigned long long>(ns)); }

// Generic peer -> local copy with sys-scope loads (used for host reads of the server's control words:
// the same in-kernel NVLink load path as the round kernels, then an ordinary local D2H copy).
__global__ void peer_copy_kernel(unsigned int* dst, const unsigned int* src, long long n_words) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src + i) : "memory");
    dst[i] = v;
  }
}

// P2P litmus: ping-pong a flag between two GPUs `iters` times (release/acquire at sys scope).
__global__ void p2p_pingpong_kernel(unsigned int* mine, unsigned int* theirs, int iters, int first, unsigned long long* ns_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long t0 = globaltimer_ns();
  for (int i = 1; i <= iters; ++i) {
    if (first) {
      st_release_sys_u32(theirs, static_cast<unsigned int>(i));
      while (ld_acquire_sys_u32(mine) != static_cast<unsigned int>(i)) {}
    } else {
      while (ld_acquire_sys_u32(mine) != static_cast<unsigned int>(i)) {}
      st_release_sys_u32(theirs, static_cast<unsigned int>(i));
    }
  }
  *ns_out = globaltimer_ns() - t0;
}

}  // namespace afb

extern "C" {

// ABI self-description so Python can assert its ctypes/numpy mirrors match.
int afb_abi_sizeof(int which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(afb::RoundParams));
    case 1: return static_cast<int>(sizeof(afb::WorkerCtl));
    case 2: return static_cast<int>(sizeof(afb::ServerCtl));
    case 3: return static_cast<int>(sizeof(afb::WorkerSlot));
    case 4: return static_cast<int>(sizeof(afb::ResultRec));
    case 5: return static_cast<int>(sizeof(afb::SyncPeers));
    case 6: return static_cast<int>(sizeof(afb::RoundDesc));
    case 7: return static_cast<int>(sizeof(afb::PipeBuf));
    default: return -1;
  }
}

// Stream compaction of the Bernoulli sample: out[0..total) = selected local row positions.
int afb_sample_rows_v2(const long long* gidx, long long row_base, long long n, long long seed_eff, int thr24,
                       int* block_counts, long long* out, long long* total_dev, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n <= 0) { cudaMemsetAsync(total_dev, 0, sizeof(long long), s); return 0; }
  int nblocks = static_cast<int>((n + 255) / 256);
  unsigned long long key = 0;
  {  // host-side mix64 (same constants as the device)
    unsigned long long x = static_cast<unsigned long long>(seed_eff);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
    key = x;
  }
  afb::sample_count_kernel<<<nblocks, 256, 0, s>>>(gidx, row_base, n, key, thr24, block_counts);
  afb::sample_scan_kernel<<<1, 1024, 0, s>>>(block_counts, nblocks, total_dev);
  afb::sample_write_kernel<<<nblocks, 256, 0, s>>>(gidx, row_base, n, key, thr24, block_counts, out);
  return static_cast<int>(cudaGetLastError());
}

int afb_l2_flush(void* buf, long long nbytes, void* stream) {
  afb::l2_flush_kernel<<<148 * 4, 512, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float4*>(buf), nbytes / 16);
  return static_cast<int>(cudaGetLastError());
}

// ---- CUDA IPC / peer access (fallback provider for symmetric buffers) -------------------------
int afb_ipc_get_handle(void* dev_ptr, void* handle_out /* 64 bytes */) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, dev_ptr);
  if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
  return static_cast<int>(e);
}
int afb_ipc_open_handle(const void* handle /* 64 bytes */, void** dev_ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return static_cast<int>(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
}
int afb_ipc_close_handle(void* dev_ptr) { return static_cast<int>(cudaIpcCloseMemHandle(dev_ptr)); }

int afb_enable_peer_access(int dev, int peer) {
  int can = 0;
  cudaDeviceCanAccessPeer(&can, dev, peer);
  if (!can) return -1;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  return static_cast<int>(e);
}

int afb_sleep(long long ns, void* stream) {
  afb::sleep_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ns);
  return static_cast<int>(cudaGetLastError());
}

int afb_peer_copy(void* dst_local, const void* src_peer, long long nbytes, void* stream) {
  long long n = nbytes / 4;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid < 1) grid = 1;
  if (grid > 148) grid = 148;
  afb::peer_copy_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned int*>(dst_local),
                                                                            static_cast<const unsigned int*>(src_peer), n);
  return static_cast<int>(cudaGetLastError());
}

int afb_p2p_pingpong(void* mine, void* theirs, int iters, int first, void* ns_out, void* stream) {
  afb::p2p_pingpong_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned int*>(mine), static_cast<unsigned int*>(theirs), iters, first,
      static_cast<unsigned long long*>(ns_out));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
