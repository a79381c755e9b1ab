// Shared device-side definitions for the asyncframework_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AFB_DEVINL __device__ __forceinline__

namespace afb {

// ---------------------------------------------------------------------------------------------
// Counter-based Bernoulli sampler — bit-identical to asyncframework_b200/core/sampler.py.
// ---------------------------------------------------------------------------------------------
AFB_DEVINL unsigned long long mix64(unsigned long long x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}
AFB_DEVINL unsigned int hash_u24(unsigned long long key, unsigned long long idx) {
  return static_cast<unsigned int>(mix64(key + idx * 0x9E3779B97F4A7C15ULL) >> 40);
}

// ---------------------------------------------------------------------------------------------
// Memory-model primitives for cross-GPU (system-scope) communication over NVLink peer pointers.
// Writers of the Hogwild vector use relaxed sys-scope reductions; flags/counters use
// acquire/release at sys scope.
// ---------------------------------------------------------------------------------------------
AFB_DEVINL unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
AFB_DEVINL unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
AFB_DEVINL unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
AFB_DEVINL int ld_relaxed_sys_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
AFB_DEVINL void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
AFB_DEVINL void st_release_gpu_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
AFB_DEVINL void st_relaxed_sys_f32(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
AFB_DEVINL float4 ld_relaxed_sys_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
AFB_DEVINL float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
AFB_DEVINL void st_relaxed_sys_f32x4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Hogwild update primitive: fire-and-forget fp32 add into (possibly remote) HBM.
AFB_DEVINL void red_add_sys_f32(float* p, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
AFB_DEVINL void red_add_sys_f32x4(float* p, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
AFB_DEVINL void red_add_gpu_f32(float* p, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
AFB_DEVINL void red_add_gpu_f32x4(float* p, float4 v) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
AFB_DEVINL unsigned long long atom_add_acqrel_sys_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long old;
  asm volatile("atom.acq_rel.sys.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
AFB_DEVINL int atom_add_acqrel_sys_s32(int* p, int v) {
  int old;
  asm volatile("atom.acq_rel.sys.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
AFB_DEVINL unsigned int atom_cas_acqrel_sys_u32(unsigned int* p, unsigned int cmp, unsigned int val) {
  unsigned int old;
  asm volatile("atom.acq_rel.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
  return old;
}
AFB_DEVINL void red_add_release_sys_s32(int* p, int v) {
  asm volatile("red.release.sys.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
AFB_DEVINL void red_add_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// NVLS multicast (NVSwitch replicates the store into every GPU bound to the multicast object; one store, P destinations)
AFB_DEVINL void multimem_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
AFB_DEVINL void multimem_st_release_u32(unsigned int* mc, unsigned int v) {
  asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
AFB_DEVINL void multimem_red_add_u32(unsigned int* mc, unsigned int v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
AFB_DEVINL void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
AFB_DEVINL void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

AFB_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded spinning: every in-kernel wait (round barrier, pull flag, clock ticket, BSP arrival flags, bucket gate) gives
// up after `limit_ns` so that a protocol bug or a dead peer ends the kernel instead of hanging the GPU.  The caller
// raises WorkerCtl.abort; the host checks it after synchronising and treats the engine as failed.
struct SpinGuard {
  unsigned long long t0;
  unsigned int n;
  AFB_DEVINL SpinGuard() : t0(0), n(0) {}
  AFB_DEVINL bool expired(unsigned long long limit_ns) {
    if ((++n & 63u) != 0u) return false;               // look at the clock every 64 polls
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) { t0 = now; return false; }
    return limit_ns != 0 && now - t0 > limit_ns;
  }
};

AFB_DEVINL void sleep_ns(unsigned long long ns) {
  unsigned long long t0 = globaltimer_ns();
  while (globaltimer_ns() - t0 < ns) __nanosleep(ns > 100000ULL ? 20000u : 1000u);
}

// streaming 128-bit load (row data is read exactly once per round)
AFB_DEVINL float4 ld_stream_f32x4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
AFB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// Loss coefficient / loss value (K2).  IDs match asyncframework_b200/optim/gradient.py LOSS_IDS.
// ---------------------------------------------------------------------------------------------
enum LossId : int { LOSS_LS = 0, LOSS_LOGISTIC = 1, LOSS_HINGE = 2, LOSS_LS_HALF = 3 };

template <int LOSS>
AFB_DEVINL void loss_coeff(float margin, float y, float& coeff, float& loss) {
  if (LOSS == LOSS_LS) {
    coeff = margin - y; loss = coeff * coeff;
  } else if (LOSS == LOSS_LS_HALF) {
    coeff = margin - y; loss = 0.5f * coeff * coeff;
  } else if (LOSS == LOSS_LOGISTIC) {
    // mult = sigmoid(x·w) - y ; loss = log1pExp(-x·w) (+ x·w if y == 0)   (Gradient.scala:188-196)
    float e = __expf(-fabsf(margin));
    float sig = margin >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    coeff = sig - y;
    float l1p = log1pf(e) + fmaxf(-margin, 0.f);          // log(1+exp(-margin))
    loss = (y > 0.f) ? l1p : l1p + margin;
  } else {  // hinge, labels {0,1} -> ±1
    float s = 2.f * y - 1.f;
    bool active = (1.f - s * margin) > 0.f;
    coeff = active ? -s : 0.f;
    loss = active ? 1.f - s * margin : 0.f;
  }
}

AFB_DEVINL void loss_coeff_rt(int loss_id, float margin, float y, float& coeff, float& loss) {
  switch (loss_id) {   // warp-uniform branch
    case LOSS_LS: loss_coeff<LOSS_LS>(margin, y, coeff, loss); break;
    case LOSS_LOGISTIC: loss_coeff<LOSS_LOGISTIC>(margin, y, coeff, loss); break;
    case LOSS_HINGE: loss_coeff<LOSS_HINGE>(margin, y, coeff, loss); break;
    default: loss_coeff<LOSS_LS_HALF>(margin, y, coeff, loss); break;
  }
}

// ---------------------------------------------------------------------------------------------
// Server-side shared state (lives in rank-0 symmetric memory; peers hold P2P pointers to it).
// ---------------------------------------------------------------------------------------------
struct alignas(16) ResultRec {     // one entry of the result ring backing AsyncContext.collect_all()
  int worker;
  int accepted;
  long long ts;                    // logical clock at pull
  long long staleness;             // clock at merge - ts
  long long k_after;               // accepted-update counter after this result
  long long batch;                 // true number of sampled rows
  float loss_sum;                  // Σ loss_i over the batch at the (stale) w that was pulled
  float step;                      // step actually applied
  unsigned long long t_start_ns;   // globaltimer at pull
  unsigned long long t_end_ns;     // globaltimer at apply
  unsigned long long seq;          // clock value of this merge + 1 (0 = slot never written); written LAST
};

struct alignas(16) WorkerSlot {    // device mirror of workerState
  int available;
  int alive;
  long long staleness;
  long long num_tasks;
  float avg_task_ms;
  float last_task_ms;
  unsigned long long heartbeat_ns;
  unsigned long long busy_since_ns;
  unsigned long long wait_ns;      // accumulated waiting time (finish -> next start), A19
  unsigned long long last_finish_ns;
};

struct alignas(16) ServerCtl {
  unsigned long long clock_k;      // low 32: clock (merged results); high 32: k (accepted updates)
  int avail_count;                 // number of idle workers (bucket barrier)
  int stop;                        // host/device stop request
  unsigned long long cul_time_ns;  // delay-injector calibration: Σ task time over warm-up updates
  unsigned int cul_count;
  unsigned int wave;               // bucket-gate submission waves (SparkASGDThread.scala:230-234: one job per wave of free workers)
  unsigned long long avg_delay_ns; // calibrated average task time (0 until warm-up done)
  // ---- runtime-adjustable gating (written by the host: watchdog / async_barrier bridge) ------------------------------
  int gate_dyn;                    // > 0: overrides the launch-time bucket threshold when smaller (continue with P-1 workers)
  int num_alive;                   // workers not declared dead by the watchdog (informational)
  unsigned long long enable_mask[4];  // bit q = worker q may start a round (ASYNCbarrier predicate evaluated by the host); P <= 256
};

}  // namespace afb
