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
constexpr int kCrpsMainBlocksMax = 148 * 8;      // CTAs of the main kernel (grid-stride over 256-node tiles)
__host__ __device__ inline int crps_main_blocks(int m) {
  const int b = ceil_div(m < 1 ? 1 : m, kCrpsThreads);
  return b > kCrpsMainBlocksMax ? kCrpsMainBlocksMax : b;
}

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

  // 16 targets per thread and pass (four independent 128-bit loads): with <= 256 CTAs a scalar grid-stride loop had one
  // 128-byte request in flight per warp and took 250 us for 2^24 targets - longer than the CRPS kernel itself
  int local = 0;
  const bool vec = (reinterpret_cast<uintptr_t>(y) & 15u) == 0;
  const int m16 = vec ? (m / 16) * 16 : 0;
  const float4* y4 = reinterpret_cast<const float4*>(y);
  for (int i = (bid.x * 256 + threadIdx.x) * 4; i < m16 / 4; i += gdim.x * 256 * 4) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(y4 + i + k);
#pragma unroll
    for (int k = 0; k < 4; ++k) local += !isnan(v[k].x) + !isnan(v[k].y) + !isnan(v[k].z) + !isnan(v[k].w);
  }
  for (int i = m16 + bid.x * 256 + threadIdx.x; i < m; i += gdim.x * 256) local += !isnan(y[i]);
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
  // The block's 256 rows of WIDTH floats are one contiguous span: it is moved with 128-bit coalesced accesses through
  // shared memory (a row-strided access pattern made every store instruction touch all 20 sectors of a warp's span
  // with 4-byte pieces).  A thread reads and writes only its own WIDTH slots of the tile (stride WIDTH: odd or 2, 4).
  __shared__ __align__(16) float tile_in[kCrpsThreads * WIDTH];
  __shared__ __align__(16) float tile_out[kCrpsThreads * WIDTH];
  const int tid = threadIdx.x;
  const bool in_vec = (reinterpret_cast<uintptr_t>(pred) & 15u) == 0, out_vec = (reinterpret_cast<uintptr_t>(d_pred) & 15u) == 0;
  constexpr int kF4 = kCrpsThreads * WIDTH / 4;                   // 128-bit pieces of a full tile
  constexpr int kPer = (kF4 + kCrpsThreads - 1) / kCrpsThreads;   // per thread: 1 (WIDTH <= 4) or 2
  float loss = 0.0f;
  // a CTA walks tiles bid.x, bid.x + gridDim.x, ... (<= kCrpsMainBlocksMax CTAs: the valid-count prologue and the
  // loss partial are paid once per CTA, and the final reduction reads a few hundred partials instead of M / 256).
  // The next tile's rows and targets are fetched into registers before the current tile is computed.
  float4 nx[kPer];
  float ny = 0.f;
  auto fetch = [&](int base) {
    if (base >= m) return;
    const bool full = m - base >= kCrpsThreads;
    if (full && in_vec) {
      const float4* s4 = reinterpret_cast<const float4*>(pred + (size_t)base * WIDTH);
#pragma unroll
      for (int q = 0; q < kPer; ++q)
        if (tid + q * kCrpsThreads < kF4) nx[q] = __ldcs(s4 + tid + q * kCrpsThreads);
    }
    ny = base + tid < m ? __ldcs(y + base + tid) : nanf("");
  };
  const int stride = gdim.x * kCrpsThreads;
  fetch(bid.x * kCrpsThreads);
  for (int base = bid.x * kCrpsThreads; base < m; base += stride) {
    const int n_here = min(kCrpsThreads, m - base);
    const bool full = n_here == kCrpsThreads;
    if (full && in_vec) {
#pragma unroll
      for (int q = 0; q < kPer; ++q)
        if (tid + q * kCrpsThreads < kF4) reinterpret_cast<float4*>(tile_in)[tid + q * kCrpsThreads] = nx[q];
    } else {
      const float* src = pred + (size_t)base * WIDTH;
      for (int k = tid; k < n_here * WIDTH; k += kCrpsThreads) tile_in[k] = src[k];
    }
    const int i = base + tid;
    const float yi = ny;
    __syncthreads();                 // tile_in complete (and every thread is past the stores out of tile_out of the pass before)
    fetch(base + stride);
    {
      float row[WIDTH], g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < WIDTH; ++j) row[j] = tile_in[tid * WIDTH + j];
      if (i < m && !isnan(yi)) loss += crps_node_k<WIDTH - 2>(row, yi, raw_input, u_fixed, xi, t, g);
#pragma unroll
      for (int j = 0; j < WIDTH; ++j) tile_out[tid * WIDTH + j] = g[j] * inv_n;
    }
    __syncthreads();                 // tile_out complete; tile_in may be overwritten
    if (d_pred != nullptr) {
      float* dst = d_pred + (size_t)base * WIDTH;
      if (full && out_vec) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int q = 0; q < kPer; ++q)
          if (tid + q * kCrpsThreads < kF4) __stcs(d4 + tid + q * kCrpsThreads, reinterpret_cast<const float4*>(tile_out)[tid + q * kCrpsThreads]);
      } else {
        for (int k = tid; k < n_here * WIDTH; k += kCrpsThreads) dst[k] = tile_out[k];
      }
    }
  }
  // 32 node losses in fp32 (fixed shuffle order), warps and blocks in float64
  for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = (double)loss;
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
