// Links + CRPS tile functions shared by rc_crps.cu and the step program (rc_prog.cu).
#pragma once
#include "rc_common.cuh"
#include "rc_crps_node.cuh"

namespace rc {

struct CrpsCountP {
  const float* y;
  int m;
  int* cnt_partial;
};

struct CrpsMainP {
  const float* pred;
  const float* y;
  float* d_pred;
  int m;
  int kind;
  int raw_input;
  float u_fixed;
  float xi;
  float t;
  const int* cnt_partial;
  int n_cnt;
  double* loss_partial;
};

struct CrpsFinalP {
  const int* cnt_partial;
  int n_cnt;
  const double* loss_partial;
  int n_loss;
  double* loss_out;
  int* n_valid;
};

struct PostFwdP {
  const float* raw;
  float* post;
  int m;
  int kind;
};

struct PostBwdP {
  const float* raw;
  const float* d_post;
  float* d_raw;
  int m;
  int kind;
};



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

__device__ __forceinline__ void crps_count_tile(const CrpsCountP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ y = p.y;
  int m = p.m;
  int* cnt_partial = p.cnt_partial;
  (void)bid; (void)gdim;

  int local = 0;
  for (int i = bid.x * 256 + threadIdx.x; i < m; i += gdim.x * 256) local += !isnan(y[i]);
  __shared__ int sh[kCrpsThreads / 32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kCrpsThreads / 32; ++w) s += sh[w];
    cnt_partial[bid.x] = s;
  }
}

template <int WIDTH>
__device__ __forceinline__ void crps_main_tile(const CrpsMainP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ pred = p.pred;
  const float* __restrict__ y = p.y;
  float* __restrict__ d_pred = p.d_pred;
  int m = p.m;
  int kind = p.kind;
  int raw_input = p.raw_input;
  float u_fixed = p.u_fixed;
  float xi = p.xi;
  float t = p.t;
  const int* __restrict__ cnt_partial = p.cnt_partial;
  int n_cnt = p.n_cnt;
  double* __restrict__ loss_partial = p.loss_partial;
  (void)bid; (void)gdim;

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
  const int i = bid.x * 256 + threadIdx.x;
  float loss = 0.0f;
  if (i < m) {
    const float yi = y[i];
    float row[WIDTH], g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < WIDTH; ++j) row[j] = pred[(size_t)i * WIDTH + j];
    if (!isnan(yi)) loss = crps_node_k<WIDTH - 2>(row, yi, raw_input, u_fixed, xi, t, g);
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
    loss_partial[bid.x] = s;
  }
}

__device__ __forceinline__ void crps_final_tile(const CrpsFinalP& p, const uint3 bid, const uint3 gdim) {
  const int* __restrict__ cnt_partial = p.cnt_partial;
  int n_cnt = p.n_cnt;
  const double* __restrict__ loss_partial = p.loss_partial;
  int n_loss = p.n_loss;
  double* loss_out = p.loss_out;
  int* n_valid = p.n_valid;
  (void)bid; (void)gdim;

  __shared__ double sh[8];
  __shared__ int shc[8];
  double s = 0.0;
  int c = 0;
  for (int i = threadIdx.x; i < n_loss; i += 256) s += loss_partial[i];
  for (int i = threadIdx.x; i < n_cnt; i += 256) c += cnt_partial[i];
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

__device__ __forceinline__ void post_fwd_tile(const PostFwdP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ raw = p.raw;
  float* __restrict__ post = p.post;
  int m = p.m;
  int kind = p.kind;
  (void)bid; (void)gdim;

  const int i = bid.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int w = loss_width(kind);
  float v[5];
  for (int j = 0; j < w; ++j) v[j] = raw[(size_t)i * w + j];
  apply_links(v, kind);
  for (int j = 0; j < w; ++j) post[(size_t)i * w + j] = v[j];
}

__device__ __forceinline__ void post_bwd_tile(const PostBwdP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ raw = p.raw;
  const float* __restrict__ d_post = p.d_post;
  float* __restrict__ d_raw = p.d_raw;
  int m = p.m;
  int kind = p.kind;
  (void)bid; (void)gdim;

  const int i = bid.x * 256 + threadIdx.x;
  if (i >= m) return;
  const int w = loss_width(kind);
  float r[5], g[5];
  for (int j = 0; j < w; ++j) { r[j] = raw[(size_t)i * w + j]; g[j] = d_post[(size_t)i * w + j]; }
  links_backward(r, g, kind);
  for (int j = 0; j < w; ++j) d_raw[(size_t)i * w + j] = g[j];
}


}  // namespace rc
