// Data-parallel gradient exchange over NVLink peer memory (one sum of the flat fp32 gradient per step,
// 0.84 MB at the reference shape - latency, not bandwidth).  The reference has no distributed code (train.py:55-74 is
// single device); semantics are DDP's: the mean of the rank gradients feeds AdamW (train.py:185).
//
// Every rank keeps its flat gradient in a buffer the other ranks of the box can read (CUDA peer mappings; the host
// side gets them from torch's symmetric-memory rendezvous).  One step is
//     rc_p2p_barrier          all ranks have finished writing their gradient
//     rc_p2p_adamw_step       g = sum_r grad_r[i] read straight from the peers (fixed rank order: every rank computes
//                             bit-identical sums, so the replicas stay identical), scaled by 1/world, AdamW update
//     rc_p2p_barrier          all ranks have finished reading: the gradient buffers may be overwritten
// instead of ncclAllReduce + AdamW: no reduced gradient is ever written, and the two barriers are one store + one
// polled load per peer.  Barriers are epoch based (a counter in local memory, never reset), so they are CUDA-graph
// capturable and need no host-side state.
#include "rc_misc_tile.cuh"

namespace rc {

constexpr int kMaxPeers = 16;

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_sys(int* p, int v) {
  asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Flag protocol of rc_p2p_step (rc_p2p_flag_scope).  0, the library default: st.release.sys / fence.acq_rel.sys, the pattern
// the PTX memory model asks for between two devices - measured at 7-8 us per step on this box (a system-scope fence is
// ~2-4 us, and there are two on the critical path of the exchange).  1: device-scope fences + relaxed system-scope flag
// accesses.  What the flags order is already in place when they are written: the gradients were written by EARLIER
// kernels of the publishing device (performed at its L2 at the kernel boundary; a peer reads them through that L2, over
// NVLink, with L1 bypassed), and "done reading" is published after a CTA barrier behind the loads whose values it
// reports.  The fences that remain keep each thread's own accesses in order (flag after data on the writer, data
// after flag on the reader).  tests/test_gpu_dp.py holds both protocols to bit-identical replicas against NCCL.
__device__ int g_p2p_flag_mode = 0;
__device__ __forceinline__ void flag_publish(int* p, int v) {
  if (g_p2p_flag_mode == 0) { st_release_sys(p, v); return; }
  __threadfence();
  st_relaxed_sys(p, v);
}
__device__ __forceinline__ void flag_acquire() {
  if (g_p2p_flag_mode == 0) fence_acq_rel_sys();
  else __threadfence();
}

// flags[r]: rank r's flag block (int32 [2 slots][kMaxPeers]) as mapped in this process; epochs: local int32[2].
// Thread q tells rank q that this rank has arrived (writes the epoch into slot entry [rank] of q's block) and waits
// for rank q's arrival in its own block.  >= : a peer may already be one barrier ahead on the same slot.
__global__ void __launch_bounds__(32) p2p_barrier_kernel(int* const* __restrict__ flags, int* __restrict__ epochs, int rank, int world,
                                                         int slot, int* __restrict__ timed_out) {
  __shared__ int ep;
  if (threadIdx.x == 0) ep = ++epochs[slot];
  __syncthreads();
  const int q = threadIdx.x;
  if (q < world) {
    __threadfence_system();                       // this device's earlier kernels' writes, system wide
    st_release_sys(flags[q] + slot * kMaxPeers + rank, ep);
    const int* mine = flags[rank] + slot * kMaxPeers + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < ep) {
      if (clock64() - t0 > 20000000000ll) {       // ~10 s: a peer died; do not hang the device
        atomicExch(timed_out, 1);
        break;
      }
      __nanosleep(40);
    }
  }
}

__global__ void __launch_bounds__(32) p2p_tick_kernel(const AdamTickP p) { adamw_tick_tile(p, blockIdx, gridDim); }

struct P2pAdamP {
  const float* const* grads;    // [world] peer mappings of the flat gradients
  int world;
  AdamP a;                      // a.grad unused
};

__global__ void __launch_bounds__(256) p2p_adamw_kernel(const P2pAdamP p) {
  const AdamP& a = p.a;
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)a.step[0];
    const double bc1 = 1.0 - pow((double)a.beta1, t), bc2 = 1.0 - pow((double)a.beta2, t);
    s_step_size = (float)((double)a.lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float decay = 1.0f - a.lr * a.weight_decay, one_m_b1 = 1.0f - a.beta1, one_m_b2 = 1.0f - a.beta2;
  // four elements per thread (the flat buffers are padded to multiples of four floats and 16-byte aligned)
  const long long n4 = a.n / 4;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.world; ++r) {            // fixed order: identical sums on every rank
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p.grads[r]) + i);
      g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
    }
    float4 w = reinterpret_cast<float4*>(a.param)[i], m1 = reinterpret_cast<float4*>(a.exp_avg)[i],
           v2 = reinterpret_cast<float4*>(a.exp_avg_sq)[i];
    const AdamCoef c{a.grad_scale, decay, one_m_b1, a.beta2, one_m_b2, bc2_sqrt, a.eps, step_size};
    adam_update(c, g.x, w.x, m1.x, v2.x);
    adam_update(c, g.y, w.y, m1.y, v2.y);
    adam_update(c, g.z, w.z, m1.z, v2.z);
    adam_update(c, g.w, w.w, m1.w, v2.w);
    reinterpret_cast<float4*>(a.param)[i] = w;
    reinterpret_cast<float4*>(a.exp_avg)[i] = m1;
    reinterpret_cast<float4*>(a.exp_avg_sq)[i] = v2;
  }
}

// ---------------------------------------------------------------------------------------------- fused exchange
// The whole exchange as ONE kernel on the critical path (rc_p2p_step): every CTA starts by waiting until all ranks
// have published gradient set number `epoch + 1` (CTA 0 publishes this rank's: one st.release.sys per peer), sums the
// peers' gradients in rank order, applies AdamW, and the CTA that finishes last advances the Adam step counter and the
// exchange epoch.  "Done reading" is published as soon as the rank's last CTA holds its gradients in registers, and the
// kernel does not end before every peer has published the same: the next step may overwrite the gradients at once.
// (A separate wait kernel at the start of the next step - side stream + an event the backward waited for - cost 8 us per
// step through the extra graph edges alone; rc_p2p_wait_done remains for callers that call the barriers themselves.)
// Replaces barrier -> tick -> sum + AdamW -> barrier (four dependent launches after backward).
static __device__ unsigned int g_p2p_arrivals = 0;
static long long* g_p2p_trace = nullptr;          // debug (rc_debug_p2p_trace): globaltimer (ns) at the phases of CTA 0
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

static __device__ unsigned int g_p2p_readers = 0;

__global__ void __launch_bounds__(256) p2p_step_kernel(const P2pAdamP p, int* const* __restrict__ flags, int* __restrict__ epoch,
                                                       int rank, int* __restrict__ timed_out, long long* __restrict__ trace) {
  const AdamP& a = p.a;
  const bool tr = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const long long n4 = a.n / 4;
  const long long stride = (long long)gridDim.x * 256;
  const long long first = (long long)blockIdx.x * 256 + threadIdx.x;
  // passes of this CTA (uniform over its threads: the barrier inside the loop needs that)
  const int passes = max(1, (int)((n4 - (long long)blockIdx.x * 256 + stride - 1) / stride));    // (>= 1: every CTA reports "done reading")
  // this rank's parameter and Adam moments of the first pass do not depend on the kernel before this one: fetched before
  // the programmatic-launch wait, like the weight tiles of the GEMMs
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f), m1 = w, v2 = w;
  if (first < n4) {
    w = reinterpret_cast<const float4*>(a.param)[first];
    m1 = reinterpret_cast<const float4*>(a.exp_avg)[first];
    v2 = reinterpret_cast<const float4*>(a.exp_avg_sq)[first];
  }
  pdl_entry();
  if (tr) trace[0] = globaltimer_ns();
  __shared__ float s_step_size, s_bc2_sqrt;
  __shared__ int s_flag;
  const int ep = epoch[0] + 1;                     // (advanced only by the last CTA to finish)
  if (threadIdx.x < p.world) {
    const int q = threadIdx.x;
    if (blockIdx.x == 0) {
      // release at system scope: the gradient writes of this device's earlier kernels happen-before it (kernel boundary)
      // and are covered by cumulativity - a separate fence.sys in front of it cost another ~2 us
      flag_publish(flags[q] + rank, ep);           // slot 0 of rank q's block: "rank's gradients of exchange ep are in place"
      if (tr) trace[1] = globaltimer_ns();
    }
    const int* mine = flags[rank] + q;
    const long long t0 = clock64();
    while (ld_relaxed_sys(mine) < ep) {            // relaxed polls, ONE acquire fence once the value is there
      if (clock64() - t0 > 20000000000ll) {        // ~10 s: a peer died; do not hang the device
        atomicExch(timed_out, 1);
        break;
      }
    }
    flag_acquire();
  }
  if (threadIdx.x == 0) {
    const double t = (double)(a.step[0] + 1);
    const double bc1 = 1.0 - pow((double)a.beta1, t), bc2 = 1.0 - pow((double)a.beta2, t);
    s_step_size = (float)((double)a.lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  if (tr) trace[2] = globaltimer_ns();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float decay = 1.0f - a.lr * a.weight_decay, one_m_b1 = 1.0f - a.beta1, one_m_b2 = 1.0f - a.beta2;
  auto sum_grads = [&](long long i) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4)
      for (int r = 0; r < p.world; ++r) {          // fixed order: identical sums on every rank
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p.grads[r]) + i);
        g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
      }
    return g;
  };
  float4 g = sum_grads(first);
  for (int pass = 0; pass < passes; ++pass) {
    const long long i = first + (long long)pass * stride;
    float4 gn = make_float4(0.f, 0.f, 0.f, 0.f), wn = gn, mn = gn, vn = gn;
    if (pass + 1 < passes) {
      // the next pass's gradients (one NVLink round trip) and state are in flight while this pass is updated
      const long long nx = i + stride;
      gn = sum_grads(nx);
      if (nx < n4) {
        wn = reinterpret_cast<const float4*>(a.param)[nx];
        mn = reinterpret_cast<const float4*>(a.exp_avg)[nx];
        vn = reinterpret_cast<const float4*>(a.exp_avg_sq)[nx];
      }
    } else {
      // every gradient this CTA needs is in registers.  The LAST CTA of the rank to get here tells the peers that this
      // rank is done reading their gradients (slot 1) - before the update arithmetic and stores, so that the peers'
      // matching flags have usually arrived by the time the kernel ends and waits for them.
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        s_flag = atomicAdd(&g_p2p_readers, 1u) == gridDim.x - 1;
        if (s_flag) g_p2p_readers = 0;
      }
      __syncthreads();
      if (s_flag && threadIdx.x < p.world) flag_publish(flags[threadIdx.x] + kMaxPeers + rank, ep);
    }
    if (i < n4) {
      const AdamCoef c{a.grad_scale, decay, one_m_b1, a.beta2, one_m_b2, bc2_sqrt, a.eps, step_size};
      adam_update(c, g.x, w.x, m1.x, v2.x);
      adam_update(c, g.y, w.y, m1.y, v2.y);
      adam_update(c, g.z, w.z, m1.z, v2.z);
      adam_update(c, g.w, w.w, m1.w, v2.w);
      reinterpret_cast<float4*>(a.param)[i] = w;
      reinterpret_cast<float4*>(a.exp_avg)[i] = m1;
      reinterpret_cast<float4*>(a.exp_avg_sq)[i] = v2;
    }
    g = gn; w = wn; m1 = mn; v2 = vn;
  }
  __syncthreads();
  if (tr) trace[3] = globaltimer_ns();
  if (threadIdx.x == 0) {
    __threadfence();
    s_flag = atomicAdd(&g_p2p_arrivals, 1u) == gridDim.x - 1;
    if (s_flag) {
      g_p2p_arrivals = 0;
      a.step[0] += 1;
      epoch[0] = ep;
      __threadfence();
    }
  }
  __syncthreads();
  // The kernel ends only when every peer is done reading THIS rank's gradients of exchange ep: the next step's backward
  // may then overwrite them without any further handshake (no wait kernel, no event in the next step).
  if (s_flag && threadIdx.x < p.world) {
    const int* mine = flags[rank] + kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    while (ld_relaxed_sys(mine) < ep) {
      if (clock64() - t0 > 20000000000ll) {
        atomicExch(timed_out, 1);
        break;
      }
    }
    flag_acquire();
  }
  if (tr) trace[4] = globaltimer_ns();
}

// every peer has finished reading this rank's gradients of the last completed exchange (epoch[0])
__global__ void __launch_bounds__(32) p2p_wait_done_kernel(int* const* __restrict__ flags, const int* __restrict__ epoch, int rank, int world,
                                                           int* __restrict__ timed_out, long long* __restrict__ trace) {
  const int q = threadIdx.x;
  if (trace != nullptr && q == 0) trace[5] = globaltimer_ns();
  if (q < world) {
    const int ep = epoch[0];
    const int* mine = flags[rank] + kMaxPeers + q;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < ep) {
      if (clock64() - t0 > 20000000000ll) {
        atomicExch(timed_out, 1);
        break;
      }
      __nanosleep(40);
    }
  }
  __syncwarp();
  if (trace != nullptr && q == 0) trace[6] = globaltimer_ns();
}

}  // namespace rc

using namespace rc;

extern "C" int rc_p2p_step(float* param, const float* const* peer_grads, int32_t* const* flags, int32_t* epoch, int rank, int world,
                           float* exp_avg, float* exp_avg_sq, int64_t* step, long long n, float lr, float beta1, float beta2,
                           float eps, float weight_decay, int32_t* timed_out, void* stream) {
  if (!param || !peer_grads || !flags || !epoch || !exp_avg || !exp_avg_sq || !step || !timed_out || n < 0 || world < 1 ||
      world > kMaxPeers || rank < 0 || rank >= world)
    return fail(RC_ERR_ARG, "rc_p2p_step: bad argument (world <= %d)", kMaxPeers);
  if (n % 4 || !aligned16(param) || !aligned16(exp_avg) || !aligned16(exp_avg_sq))
    return fail(RC_ERR_ARG, "rc_p2p_step: n must be a multiple of 4 and the buffers 16-byte aligned");
  long long blocks = ceil_div_ll(n / 4, 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;  // (a CTA waits for flags that PEERS set, never for another local CTA: no residency requirement)
  if (blocks < 1) blocks = 1;
  P2pAdamP p;
  p.grads = peer_grads;
  p.world = world;
  p.a = AdamP{param, nullptr, exp_avg, exp_avg_sq, reinterpret_cast<long long*>(step), n, lr, beta1, beta2, eps, weight_decay,
              1.0f / (float)world};
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, p2p_step_kernel, p, reinterpret_cast<int* const*>(flags), static_cast<int*>(epoch), rank,
                       static_cast<int*>(timed_out), g_p2p_trace);
  }
  return check_launch("p2p_step_kernel");
}

extern "C" int rc_p2p_flag_scope(int device_scope_fences) {
  const int mode = device_scope_fences ? 1 : 0;
  return cudaMemcpyToSymbol(g_p2p_flag_mode, &mode, sizeof(int)) == cudaSuccess ? RC_OK : fail(RC_ERR_CUDA, "rc_p2p_flag_scope");
}
extern "C" void rc_debug_p2p_trace(void* device_buf) { g_p2p_trace = static_cast<long long*>(device_buf); }

extern "C" int rc_p2p_wait_done(int32_t* const* flags, const int32_t* epoch, int rank, int world, int32_t* timed_out, void* stream) {
  if (!flags || !epoch || !timed_out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(RC_ERR_ARG, "rc_p2p_wait_done: bad argument");
  p2p_wait_done_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<int* const*>(flags), epoch, rank, world, timed_out, g_p2p_trace);
  return check_launch("p2p_wait_done_kernel");
}

extern "C" int rc_p2p_barrier(int32_t* const* flags, int32_t* epochs, int rank, int world, int slot, int32_t* timed_out,
                              void* stream) {
  if (!flags || !epochs || !timed_out || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || slot < 0 || slot > 1)
    return fail(RC_ERR_ARG, "rc_p2p_barrier: bad argument (world <= %d, slot 0 or 1)", kMaxPeers);
  p2p_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<int* const*>(flags), epochs, rank, world, slot,
                                                                     timed_out);
  return check_launch("p2p_barrier_kernel");
}

extern "C" int rc_p2p_adamw_step(float* param, const float* const* peer_grads, int world, float* exp_avg, float* exp_avg_sq,
                                 int64_t* step, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                 void* stream) {
  if (!param || !peer_grads || !exp_avg || !exp_avg_sq || !step || n < 0 || world < 1 || world > kMaxPeers)
    return fail(RC_ERR_ARG, "rc_p2p_adamw_step: bad argument");
  if (n % 4 || !aligned16(param) || !aligned16(exp_avg) || !aligned16(exp_avg_sq))
    return fail(RC_ERR_ARG, "rc_p2p_adamw_step: n must be a multiple of 4 and the buffers 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const AdamTickP pt{reinterpret_cast<long long*>(step)};
  p2p_tick_kernel<<<1, 32, 0, s>>>(pt);
  if (int e = check_launch("p2p_tick_kernel")) return e;
  if (n == 0) return RC_OK;
  long long blocks = ceil_div_ll(n / 4, 256);
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  if (blocks < 1) blocks = 1;
  P2pAdamP p;
  p.grads = peer_grads;
  p.world = world;
  p.a = AdamP{param, nullptr, exp_avg, exp_avg_sq, reinterpret_cast<long long*>(step), n, lr, beta1, beta2, eps, weight_decay,
              1.0f / (float)world};
  p2p_adamw_kernel<<<(int)blocks, 256, 0, s>>>(p);
  return check_launch("p2p_adamw_kernel");
}
