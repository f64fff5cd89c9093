// Shared helpers for the sm_100a kernels (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "rc_b200.h"

namespace rc {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(RC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return RC_OK;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Programmatic dependent launch.  A training step at the reference shape is a chain of ~50 dependent kernels of a few
// microseconds each; between two of them the GPU idles for the launch latency of the second.  Every kernel of the step
// therefore (a) is launched with programmaticStreamSerializationAllowed, so that its CTAs may be scheduled while the
// kernel before it in the stream is still running, and (b) starts with pdl_entry(): griddepcontrol.wait blocks until
// that kernel has completed and its writes are visible (nothing is read or written before it), and
// griddepcontrol.launch_dependents lets the NEXT kernel's launch start as soon as all CTAs of this one are resident.
// RC_PDL=0 turns the launch attribute off (the device-side instructions are then no-ops).
bool pdl_enabled();

__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename P>
inline void launch_pdl(void (*kernel)(const P), dim3 grid, dim3 block, size_t smem, cudaStream_t s, const P& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, p);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rc_gine_wide.cu: large-graph path of the aggregation kernels (H >= 128, at least kRangedMinRows rows)
constexpr int kRangedMinRows = 16384;
int gine_ranged_grid(int m);
int launch_gine_fwd_ranged(const float* x, const int* rowptr, const int* col, const float* attr, const float* w_edge,
                           const float* b_edge, const float* eps, float* h, int m, int hidden, cudaStream_t s);
int launch_gine_bwd_ranged(const float* g, const float* x, const int* t_rowptr, const int* t_dst, const float* t_attr,
                           const float* w_edge, const float* b_edge, const float* eps, const float* addend, float* dx,
                           float* partials, int m, int hidden, cudaStream_t s);

// rc_deepsets_tc.cu: tcgen05 / TMEM path of the DeepSets member Linear + ReLU + pool
bool deepsets_tc_applicable(int num_nodes, int members, int feats, int hidden);
int launch_deepsets_fwd_tc(bool bf16, const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                           int members, int feats, int hidden, cudaStream_t s);


// rc_deepsets_tc_bwd.cu: tcgen05 / TMEM path of the DeepSets pool backward (members 11 / 51)
bool deepsets_bwd_tc_applicable(int num_nodes, int members, int feats, int hidden);
int deepsets_bwd_tc_blocks(int num_nodes, int members);
void deepsets_bwd_tc_set_trace(long long* p);
int launch_deepsets_bwd_tc(const float* ens, const float* w1, const float* b1, const float* d_pooled, float* partials,
                           uint32_t* mask_out, int num_nodes, int members, int feats, int hidden, int bf16, cudaStream_t s);

// rc_gemm_tc.cu: tcgen05 3xTF32 path of rc_gemm_run (1 = activation rows, 2 = weight gradient, 0 = not applicable)
int gemm_tc_kind(const rc_gemm* g);
size_t gemm_tc_workspace(const rc_gemm* g);
int gemm_tc_wgrad_splits(const rc_gemm* g);
int gemm_tc_run(const rc_gemm* g, cudaStream_t s);
void gemm_tc_set_trace(long long* p);

}  // namespace rc
