// Standalone communication kernels over NVLink / NVSwitch (sm_100a):
//
//  * broadcast_mc_kernel   — server -> all workers parameter broadcast through the NVLS multicast
//    alias of a symmetric buffer: ONE `multimem.st` per 16 bytes, the switch replicates it into every
//    GPU's HBM (the TorrentBroadcast replacement).  Used to publish versioned model
//    snapshots (AsyncBroadcast ring) and for the BSP "server applies, then multicasts" variant.
//  * broadcast_p2p_kernel  — same semantics with plain peer stores (no multicast support / 1 GPU).
//  * push_apply_kernel     — standalone multi-CTA push-and-apply (K9+K5/K6) for wide models (CSR,
//    d = 47 236): CTA 0 takes the logical-clock ticket and publishes the accept/step decision, every
//    CTA streams its slice of the local gradient into the server with vector reductions, the last CTA
//    publishes the result record.  The dense round kernel has this tail fused in.
#include <algorithm>
#include "afb_round.cuh"

namespace afb {


// dst_mc: multicast address of the destination region (same offset on every rank)
__global__ void broadcast_mc_kernel(const float* __restrict__ src, float* dst_mc, long long n4, unsigned int* flag_mc,
                                    unsigned int* done_ctr) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    multimem_st_f32x4(dst_mc + 4 * i, *reinterpret_cast<const float4*>(src + 4 * i));
  fence_sys();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(done_ctr, 1u);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      if (flag_mc) multimem_red_add_u32(flag_mc, 1u);      // version counter bumps on every rank
    }
  }
}

struct PeerList { float* dst[16]; unsigned int* flag[16]; int n; };

__global__ void broadcast_p2p_kernel(const float* __restrict__ src, PeerList peers, long long n4, unsigned int* done_ctr) {
  for (int q = 0; q < peers.n; ++q)
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
      st_relaxed_sys_f32x4(peers.dst[q] + 4 * i, *reinterpret_cast<const float4*>(src + 4 * i));
  fence_sys();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(done_ctr, 1u);
    if (t == gridDim.x - 1) {
      *done_ctr = 0;
      for (int q = 0; q < peers.n; ++q)
        if (peers.flag[q]) red_add_release_sys_u32(peers.flag[q], 1u);
    }
  }
}

// ---- split round for wide models: gate -> multi-CTA pull -> (grad kernel) -> multi-CTA push ------------
// A 47 236-wide model is 189 KB: one CTA pulling / pushing it serialises ~50 NVLink round trips, so for
// wide models the round is split into small kernels whose data movement uses many CTAs.
__global__ void gate_kernel(const RoundParams p) {
  if (threadIdx.x != 0) return;
  WorkerCtl* wc = p.wctl;
  const unsigned int round = wc->round;
  int proceed = 1;
  if (p.gate_threshold > 0) {
    while (true) {
      unsigned long long ck0 = ld_acquire_sys_u64(&p.sctl->clock_k);
      if (static_cast<long long>(ck0 >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) break;
      if (ld_relaxed_sys_s32(&p.sctl->avail_count) >= p.gate_threshold) break;
      __nanosleep(200);
    }
  }
  unsigned long long ck = ld_acquire_sys_u64(&p.sctl->clock_k);
  if (static_cast<long long>(ck >> 32) >= p.max_iters || ld_relaxed_sys_s32(&p.sctl->stop)) proceed = 0;
  if (proceed) {
    red_add_release_sys_s32(&p.sctl->avail_count, -1);
    WorkerSlot* me = p.status + p.worker_id;
    unsigned long long now = globaltimer_ns();
    if (wc->last_finish_ns) wc->wait_ns += now - wc->last_finish_ns;
    me->available = 0; me->busy_since_ns = now; me->heartbeat_ns = now; me->wait_ns = wc->wait_ns;
  }
  wc->ts = ck & 0xffffffffULL;
  wc->k_read = ck >> 32;
  unsigned long long seed_eff = static_cast<unsigned long long>(p.seed);
  if (p.seed_mode == 0) seed_eff += wc->k_read + 1ULL;
  else if (p.seed_mode == 1) seed_eff += static_cast<unsigned long long>(round) + 1ULL;
  wc->seed_key = mix64(seed_eff);
  wc->proceed = proceed;
  wc->t_start_ns = globaltimer_ns();
  if (!proceed) wc->last_accepted = -1;
  if (proceed && p.delay_mode != 0) {
    unsigned long long avg = ld_acquire_sys_u64(&p.sctl->avg_delay_ns);
    if (avg) {
      float mult = p.delay_coeff;
      if (p.delay_mode >= 2) {
        float u = static_cast<float>(hash_u24(wc->seed_key ^ 0xD1B54A32D192ED03ULL, p.worker_id)) * (1.f / 16777216.f);
        mult = (p.delay_mode == 3) ? (2.5f + 7.5f * u) : (1.5f + u);
      }
      sleep_ns(static_cast<unsigned long long>(mult * static_cast<float>(avg)));
    }
  }
}

__global__ void pull_w_kernel(const RoundParams p) {
  if (!*reinterpret_cast<volatile int*>(&p.wctl->proceed)) return;
  const long long nvec = p.ld >> 2;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x)
    *reinterpret_cast<float4*>(p.w_local + 4 * v) = ld_relaxed_sys_f32x4(p.w_server + 4 * v);
}

// ---- standalone push-and-apply -------------------------------------------------------------------
struct PushCtl { unsigned int decided; int accept; float step; unsigned int clock_old; unsigned int done; unsigned int pad[3]; };

__global__ void __launch_bounds__(512) push_apply_kernel(const RoundParams p, PushCtl* pc) {
  const int tid = threadIdx.x;
  WorkerCtl* wc = p.wctl;
  if (!*reinterpret_cast<volatile int*>(&wc->proceed)) return;     // stopped round: nothing was computed
  const long long nvec = p.ld >> 2;
  __shared__ int s_accept;
  __shared__ float s_step;
  if (blockIdx.x == 0 && tid == 0) {
    unsigned long long old, k_old;
    long long staleness;
    int accept;
    const unsigned long long ts = wc->ts;
    if (p.taw >= (1LL << 31)) {
      old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, (1ULL << 32) | 1ULL);
      k_old = old >> 32;
      staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
      accept = (static_cast<long long>(k_old) < p.max_iters) ? 1 : 0;
    } else {
      old = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL);
      staleness = static_cast<long long>(old & 0xffffffffULL) - static_cast<long long>(ts);
      long long cmp = (p.algo == ALGO_ASAGA) ? static_cast<long long>(old >> 32) - static_cast<long long>(ts) : staleness;
      accept = (cmp <= p.taw) ? 1 : 0;
      k_old = old >> 32;
      if (accept) {
        unsigned long long o2 = atom_add_acqrel_sys_u64(&p.sctl->clock_k, 1ULL << 32);
        k_old = o2 >> 32;
        if (static_cast<long long>(k_old) >= p.max_iters) accept = 0;
      }
    }
    float step = (p.algo == ALGO_ASGD) ? p.gamma * rsqrtf(static_cast<float>(k_old / static_cast<unsigned long long>(p.num_workers)) + 1.f) : p.gamma;
    wc->last_staleness = staleness;
    wc->last_k_after = static_cast<long long>(k_old) + (accept ? 1 : 0);
    pc->accept = accept; pc->step = step; pc->clock_old = static_cast<unsigned int>(old & 0xffffffffULL);
    __threadfence();
    st_release_gpu_u32(&pc->decided, wc->round + 1u);
  }
  if (tid == 0) {
    while (ld_acquire_gpu_u32(&pc->decided) != wc->round + 1u) __nanosleep(64);
    s_accept = *reinterpret_cast<volatile int*>(&pc->accept);
    s_step = *reinterpret_cast<volatile float*>(&pc->step);
  }
  __syncthreads();
  const int accept = s_accept;
  const float step = s_step;
  const float a = -step * p.inv_par_recs;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + tid; v < nvec; v += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 g = __ldcg(reinterpret_cast<const float4*>(p.g_local + 4 * v));
    if (accept) {
      if (p.algo == ALGO_ASGD) {
        red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x, a * g.y, a * g.z, a * g.w));
      } else {
        float4 ab = ld_relaxed_sys_f32x4(p.abar_server + 4 * v);
        red_add_sys_f32x4(p.w_server + 4 * v, make_float4(a * g.x - step * ab.x, a * g.y - step * ab.y, a * g.z - step * ab.z, a * g.w - step * ab.w));
        red_add_sys_f32x4(p.abar_server + 4 * v, make_float4(g.x * p.inv_n, g.y * p.inv_n, g.z * p.inv_n, g.w * p.inv_n));
      }
    }
    *reinterpret_cast<float4*>(p.g_local + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_sys();
  __syncthreads();
  if (tid == 0) {
    unsigned int t = atomicAdd(&pc->done, 1u);
    if (t == gridDim.x - 1) {
      pc->done = 0;
      const unsigned long long now = globaltimer_ns();
      ResultRec* rec = p.ring + (pc->clock_old % static_cast<unsigned int>(p.ring_size));
      rec->worker = p.worker_id; rec->accepted = accept; rec->ts = static_cast<long long>(wc->ts);
      rec->staleness = wc->last_staleness; rec->k_after = wc->last_k_after; rec->batch = wc->last_batch;
      rec->loss_sum = wc->last_loss; rec->step = accept ? step : 0.f; rec->t_start_ns = wc->t_start_ns; rec->t_end_ns = now;
      WorkerSlot* me = p.status + p.worker_id;
      const float task_ms = static_cast<float>(now - wc->t_start_ns) * 1e-6f;
      wc->avg_task_ms += (task_ms - wc->avg_task_ms) / static_cast<float>(wc->rounds_done ? wc->rounds_done : 1ULL);
      me->staleness = wc->last_staleness; me->num_tasks = static_cast<long long>(wc->rounds_done);
      me->last_task_ms = task_ms; me->avg_task_ms = wc->avg_task_ms;
      me->heartbeat_ns = now; me->last_finish_ns = now; me->available = 1;
      if (accept && p.warmup_updates > 0 && wc->last_k_after <= static_cast<long long>(p.warmup_updates)) {
        asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(&p.sctl->cul_time_ns), "l"(now - wc->t_start_ns) : "memory");
        asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(&p.sctl->cul_count), "r"(1u) : "memory");
      }
      fence_sys();
      *reinterpret_cast<volatile unsigned long long*>(&rec->seq) = static_cast<unsigned long long>(pc->clock_old) + 1ULL;
      red_add_release_sys_s32(&p.sctl->avail_count, 1);
      wc->last_accepted = accept; wc->last_step = accept ? step : 0.f;
      wc->last_task_ns = now - wc->t_start_ns; wc->last_finish_ns = now;
      if (accept && p.warmup_updates > 0 && wc->last_k_after == static_cast<long long>(p.warmup_updates)) {
        unsigned long long tot = ld_acquire_sys_u64(&p.sctl->cul_time_ns);
        unsigned int c = ld_acquire_sys_u32(&p.sctl->cul_count);
        if (c) atomicExch_system(&p.sctl->avg_delay_ns, tot / c);
      }
    }
  }
}

}  // namespace afb

extern "C" {

int afb_broadcast_mc(const float* src, float* dst_mc, long long n_floats, unsigned int* flag_mc, unsigned int* done_ctr, void* stream) {
  if (n_floats % 4) return static_cast<int>(cudaErrorInvalidValue);
  long long n4 = n_floats / 4;
  int grid = static_cast<int>(std::min<long long>(148, (n4 + 255) / 256));
  if (grid < 1) grid = 1;
  afb::broadcast_mc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst_mc, n4, flag_mc, done_ctr);
  return static_cast<int>(cudaGetLastError());
}

int afb_broadcast_p2p(const float* src, void* const* dst, void* const* flags, int n_peers, long long n_floats,
                      unsigned int* done_ctr, void* stream) {
  if (n_floats % 4 || n_peers > 16) return static_cast<int>(cudaErrorInvalidValue);
  afb::PeerList pl;
  pl.n = n_peers;
  for (int q = 0; q < n_peers; ++q) { pl.dst[q] = static_cast<float*>(dst[q]); pl.flag[q] = flags ? static_cast<unsigned int*>(flags[q]) : nullptr; }
  long long n4 = n_floats / 4;
  int grid = static_cast<int>(std::min<long long>(148, (n4 + 255) / 256));
  if (grid < 1) grid = 1;
  afb::broadcast_p2p_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, pl, n4, done_ctr);
  return static_cast<int>(cudaGetLastError());
}

int afb_gate_launch(const afb::RoundParams* params, void* stream) {
  afb::gate_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*params);
  return static_cast<int>(cudaGetLastError());
}

int afb_pull_w_launch(const afb::RoundParams* params, int grid, void* stream) {
  afb::pull_w_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*params);
  return static_cast<int>(cudaGetLastError());
}

int afb_push_apply_launch(const afb::RoundParams* params, void* push_ctl, int grid, void* stream) {
  afb::push_apply_kernel<<<grid, 512, 0, static_cast<cudaStream_t>(stream)>>>(*params, static_cast<afb::PushCtl*>(push_ctl));
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
