// Debug / test instrumentation: the ReLU decisions the backward kernels take, as bit masks.
//
// Gradients of a ReLU network are discontinuous in the pre-activations: a unit within fp32 rounding of its threshold
// may sit on either side in two correct evaluations.  The parity tests therefore do not allow for such flips - they
// read the decisions the CUDA path took (these kernels evaluate the very expressions of the production kernels, on
// the same inputs) and force them into the float64 oracle, after which every gradient is held to 1e-5.
#include "rc_common.cuh"

namespace rc {

constexpr float kDbgUp = 18446744073709551616.0f;    // 2^64 (rc_gine_tiled.cu)

__device__ __forceinline__ float dbg_fma_sat(float a, float b, float c) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// one warp per transpose row j: bit c of slot q = 1[x[j,c] + t_attr[q] w[c] + b[c] > 0] as the backward evaluates it
//   tiled = 0: x + fma(a, w, b)                        (rc_gine_tile.cuh masked(), rc_gine_wide.cu masked_acc())
//   tiled = 1: sat(fma(a, w 2^64, fma(x, 2^64, b 2^64))) > 0   (rc_gine_tiled.cu masked_acc())
__global__ void __launch_bounds__(256) dbg_gine_msg_mask_kernel(const float* __restrict__ x, const int* __restrict__ t_rowptr,
                                                                const float* __restrict__ t_attr, const float* __restrict__ w,
                                                                const float* __restrict__ b, int m, int hidden, int tiled,
                                                                uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int words = (hidden + 31) / 32;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < m; row += gridDim.x * 8) {
    const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
    for (int wd = 0; wd < words; ++wd) {
      const int c = wd * 32 + lane;
      const bool ok = c < hidden;
      const float xv = ok ? __ldg(x + (size_t)row * hidden + c) : 0.f, wv = ok ? __ldg(w + c) : 0.f, bv = ok ? __ldg(b + c) : 0.f;
      const float xb = fmaf(xv, kDbgUp, bv * kDbgUp), wu = wv * kDbgUp;
      for (int q = beg; q < end; ++q) {
        const float a = __ldg(t_attr + q);
        const bool on = tiled ? dbg_fma_sat(a, wu, xb) > 0.f : xv + fmaf(a, wv, bv) > 0.f;
        const unsigned word = __ballot_sync(0xffffffffu, ok && on);
        if (lane == 0) bits[(size_t)q * words + wd] = word;
      }
    }
  }
}

// bit c of row r = 1[fma(gamma[c], (t[r,c] - mean[c]) * rstd[c], beta[c]) > 0]: the ReLU behind BatchNorm as the
// RC_EPI_BN_RELU_BWD epilogues (rc_gemm_tile.cuh, rc_gemm_tc.cu) evaluate it
__global__ void __launch_bounds__(256) dbg_bn_relu_mask_kernel(const float* __restrict__ t, int ld, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int m, int n, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int words = (n + 31) / 32;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < m; row += gridDim.x * 8) {
    for (int wd = 0; wd < words; ++wd) {
      const int c = wd * 32 + lane;
      bool on = false;
      if (c < n) {
        const float hat = (__ldg(t + (size_t)row * ld + c) - __ldg(mean + c)) * __ldg(rstd + c);
        on = fmaf(__ldg(gamma + c), hat, __ldg(beta + c)) > 0.f;
      }
      const unsigned word = __ballot_sync(0xffffffffu, on);
      if (lane == 0) bits[(size_t)row * words + wd] = word;
    }
  }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_debug_gine_msg_mask(const float* x, const int32_t* t_rowptr, const float* t_attr, const float* w_edge,
                                      const float* b_edge, int num_nodes, int hidden, int tiled, uint32_t* bits_out, void* stream) {
  if (!x || !t_rowptr || !t_attr || !w_edge || !b_edge || !bits_out || num_nodes < 0 || hidden <= 0) return fail(RC_ERR_ARG, "rc_debug_gine_msg_mask: bad argument");
  if (num_nodes == 0) return RC_OK;
  int grid = ceil_div(num_nodes, 8);
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  dbg_gine_msg_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, t_rowptr, t_attr, w_edge, b_edge, num_nodes, hidden, tiled, bits_out);
  return check_launch("dbg_gine_msg_mask_kernel");
}

extern "C" int rc_debug_bn_relu_mask(const float* t, int ld, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                     int m, int n, uint32_t* bits_out, void* stream) {
  if (!t || !mean || !rstd || !gamma || !beta || !bits_out || m < 0 || n <= 0) return fail(RC_ERR_ARG, "rc_debug_bn_relu_mask: bad argument");
  if (m == 0) return RC_OK;
  int grid = ceil_div(m, 8);
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  dbg_bn_relu_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, ld, mean, rstd, gamma, beta, m, n, bits_out);
  return check_launch("dbg_bn_relu_mask_kernel");
}

extern "C" void rc_debug_ds_trace(void* device_buf) { deepsets_bwd_tc_set_trace(static_cast<long long*>(device_buf)); }

// ---------------------------------------------------------------------------------------------- fp32 FMA peak
// MEASURED_PEAKS.json has HBM and bf16 tensor figures but no fp32 one ("the build must measure one"):
// 16 independent FMA chains per thread, 8 resident CTAs of 256 threads per SM, `iters` x 16 x 8 FFMA per thread.
namespace rc {
__global__ void __launch_bounds__(256) dbg_fma_peak_kernel(float* out, int iters, float a, float b) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;              // never true: keeps the chains alive
}
}  // namespace rc

/* launches the FMA-peak kernel; returns the number of floating-point operations it executes (2 per FMA) in *flops */
extern "C" int rc_debug_fma_peak(float* scratch, int iters, double* flops, void* stream) {
  if (!scratch || iters <= 0 || !flops) return fail(RC_ERR_ARG, "rc_debug_fma_peak: bad argument");
  const int grid = 8 * kNumSMs;
  dbg_fma_peak_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(scratch, iters, 1.000001f, 1e-7f);
  *flops = 2.0 * (double)grid * 256.0 * (double)iters * 128.0;
  return check_launch("dbg_fma_peak_kernel");
}
