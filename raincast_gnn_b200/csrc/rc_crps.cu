// Output links + closed-form CRPS value and gradient: one elementwise pass over [M, C] (HBM-bound:
// (2C+1)*4 algorithmic bytes per node).  Replaces ~2 200 ATen calls of the reference
// (models/model_utils.py:89-113, models/loss.py:203-272).
#include "rc_common.cuh"
#include "rc_crps_node.cuh"

namespace rc {

constexpr int kCrpsThreads = 256;
constexpr int kCountBlocksMax = 256;

__host__ __device__ inline int crps_count_blocks(int m) {
  int b = ceil_div(m, kCrpsThreads * 4);
  return b < 1 ? 1 : (b > kCountBlocksMax ? kCountBlocksMax : b);
}

// workspace: int32 cnt_partial[kCountBlocksMax]; double loss_partial[blocks]
struct CrpsWs {
  int* cnt;
  double* loss;
};
__host__ __device__ inline CrpsWs crps_ws(void* ws) {
  CrpsWs w;
  w.cnt = reinterpret_cast<int*>(ws);
  w.loss = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kCountBlocksMax * sizeof(int));
  return w;
}

__global__ void __launch_bounds__(kCrpsThreads) crps_count_kernel(const float* __restrict__ y, int m, int* cnt_partial) {
  int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) local += !isnan(y[i]);
  __shared__ int sh[kCrpsThreads / 32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kCrpsThreads / 32; ++w) s += sh[w];
    cnt_partial[blockIdx.x] = s;
  }
}

template <int WIDTH>
__global__ void __launch_bounds__(kCrpsThreads)
crps_main_kernel(const float* __restrict__ pred, const float* __restrict__ y, float* __restrict__ d_pred, int m,
                 int kind, int raw_input, float u_fixed, float xi, float t, const int* __restrict__ cnt_partial,
                 int n_cnt, double* __restrict__ loss_partial) {
  __shared__ int s_cnt;
  __shared__ double sh[kCrpsThreads / 32];
  if (threadIdx.x < 32) {   // every block re-derives the valid count (<= 256 ints): stateless, deterministic
    int c = 0;
    for (int i = threadIdx.x; i < n_cnt; i += 32) c += cnt_partial[i];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) s_cnt = c;
  }
  __syncthreads();
  const float inv_n = s_cnt > 0 ? 1.0f / (float)s_cnt : 0.0f;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.0f;
  if (i < m) {
    const float yi = y[i];
    float row[WIDTH], g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < WIDTH; ++j) row[j] = pred[(size_t)i * WIDTH + j];
    if (!isnan(yi)) loss = crps_node(row, yi, kind, raw_input, u_fixed, xi, t, g);
    if (d_pred != nullptr) {
#pragma unroll
      for (int j = 0; j < WIDTH; ++j) d_pred[(size_t)i * WIDTH + j] = g[j] * inv_n;
    }
  }
  double dl = warp_sum((double)loss);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kCrpsThreads / 32; ++w) s += sh[w];
    loss_partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) crps_final_kernel(const int* __restrict__ cnt_partial, int n_cnt,
                                                         const double* __restrict__ loss_partial, int n_loss,
                                                         double* loss_out, int* n_valid) {
  __shared__ double sh[8];
  __shared__ int shc[8];
  double s = 0.0;
  int c = 0;
  for (int i = threadIdx.x; i < n_loss; i += blockDim.x) s += loss_partial[i];
  for (int i = threadIdx.x; i < n_cnt; i += blockDim.x) c += cnt_partial[i];
  s = warp_sum(s);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; shc[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    int cnt = 0;
    for (int w = 0; w < 8; ++w) { tot += sh[w]; cnt += shc[w]; }
    loss_out[0] = cnt > 0 ? tot / (double)cnt : nan("");   // mean of an empty selection is NaN in torch too
    n_valid[0] = cnt;
  }
}

__global__ void __launch_bounds__(256) postprocess_fwd_kernel(const float* __restrict__ raw, float* __restrict__ post,
                                                              int m, int kind) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int w = loss_width(kind);
  float v[5];
  for (int j = 0; j < w; ++j) v[j] = raw[(size_t)i * w + j];
  apply_links(v, kind);
  for (int j = 0; j < w; ++j) post[(size_t)i * w + j] = v[j];
}

__global__ void __launch_bounds__(256) postprocess_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ d_post,
                                                              float* __restrict__ d_raw, int m, int kind) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int w = loss_width(kind);
  float r[5], g[5];
  for (int j = 0; j < w; ++j) { r[j] = raw[(size_t)i * w + j]; g[j] = d_post[(size_t)i * w + j]; }
  links_backward(r, g, kind);
  for (int j = 0; j < w; ++j) d_raw[(size_t)i * w + j] = g[j];
}

}  // namespace rc

using namespace rc;

extern "C" size_t rc_crps_workspace(int num_nodes) {
  int blocks = ceil_div(num_nodes < 1 ? 1 : num_nodes, kCrpsThreads);
  return kCountBlocksMax * sizeof(int) + (size_t)blocks * sizeof(double);
}

extern "C" int rc_crps_fwd_bwd(const float* pred, const float* y, float* d_pred, double* loss_out, int32_t* n_valid,
                               int num_nodes, int kind, int raw_input, float u_fixed, float xi, float t,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred || !y || !loss_out || !n_valid || !workspace || num_nodes < 0) return fail(RC_ERR_ARG, "rc_crps_fwd_bwd: null pointer");
  if (kind < 0 || kind > 3) return fail(RC_ERR_ARG, "rc_crps_fwd_bwd: kind %d", kind);
  if (kind >= 2 && (xi == 1.0f || xi == 2.0f || xi == 0.0f)) return fail(RC_ERR_ARG, "rc_crps_fwd_bwd: xi must not be 0, 1 or 2");
  if (workspace_bytes < rc_crps_workspace(num_nodes)) return fail(RC_ERR_WORKSPACE, "rc_crps_fwd_bwd: workspace %zu < %zu", workspace_bytes, rc_crps_workspace(num_nodes));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CrpsWs ws = crps_ws(workspace);
  const int m = num_nodes;
  const int ncnt = crps_count_blocks(m);
  const int blocks = ceil_div(m < 1 ? 1 : m, kCrpsThreads);
  crps_count_kernel<<<ncnt, kCrpsThreads, 0, s>>>(y, m, ws.cnt);
  if (int e = check_launch("crps_count_kernel")) return e;
  switch (kind) {
    case 0: crps_main_kernel<2><<<blocks, kCrpsThreads, 0, s>>>(pred, y, d_pred, m, kind, raw_input, u_fixed, xi, t, ws.cnt, ncnt, ws.loss); break;
    case 1: crps_main_kernel<3><<<blocks, kCrpsThreads, 0, s>>>(pred, y, d_pred, m, kind, raw_input, u_fixed, xi, t, ws.cnt, ncnt, ws.loss); break;
    case 2: crps_main_kernel<4><<<blocks, kCrpsThreads, 0, s>>>(pred, y, d_pred, m, kind, raw_input, u_fixed, xi, t, ws.cnt, ncnt, ws.loss); break;
    default: crps_main_kernel<5><<<blocks, kCrpsThreads, 0, s>>>(pred, y, d_pred, m, kind, raw_input, u_fixed, xi, t, ws.cnt, ncnt, ws.loss); break;
  }
  if (int e = check_launch("crps_main_kernel")) return e;
  crps_final_kernel<<<1, 256, 0, s>>>(ws.cnt, ncnt, ws.loss, blocks, loss_out, n_valid);
  return check_launch("crps_final_kernel");
}

extern "C" int rc_postprocess_fwd(const float* raw, float* post, int num_nodes, int kind, void* stream) {
  if (!raw || !post || kind < 0 || kind > 3) return fail(RC_ERR_ARG, "rc_postprocess_fwd: bad argument");
  if (num_nodes == 0) return RC_OK;
  postprocess_fwd_kernel<<<ceil_div(num_nodes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(raw, post, num_nodes, kind);
  return check_launch("postprocess_fwd_kernel");
}

extern "C" int rc_postprocess_bwd(const float* raw, const float* d_post, float* d_raw, int num_nodes, int kind, void* stream) {
  if (!raw || !d_post || !d_raw || kind < 0 || kind > 3) return fail(RC_ERR_ARG, "rc_postprocess_bwd: bad argument");
  if (num_nodes == 0) return RC_OK;
  postprocess_bwd_kernel<<<ceil_div(num_nodes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(raw, d_post, d_raw, num_nodes, kind);
  return check_launch("postprocess_bwd_kernel");
}
