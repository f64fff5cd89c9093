// GINE aggregation: forward CSR gather and backward transpose gather, both atomic-free.
//
// PyG GINEConv.message/aggregate (call site models/gnn.py:27-29,39-44):
//   m_e = relu(x[src_e] + lin(edge_attr_e)),  agg_i = sum_{dst_e = i} m_e,  h_i = agg_i + (1+eps) x_i
// A sub-warp of H/4 lanes (a full warp for H >= 128) owns one destination row; every lane keeps
// float4 column slices in registers, so each gathered source row is one coalesced 128-bit-per-lane
// load and nothing of size [E, H] is ever materialised (the reference materialises four).
// HBM-bound kernel: algorithmic bytes fwd = 2*M*H*4 + E*8 + (M+1)*4 + 8*H  (SURVEY.md 8d).
#include "rc_common.cuh"

namespace rc {

constexpr int kGineThreads = 256;
constexpr int kGineWarps = kGineThreads / 32;
constexpr int kEdgeUnroll = 4;

struct GineShape {
  int lpr;   // lanes per row
  int rpw;   // rows per warp
  int ch;    // float4 chunks per lane
};

static bool gine_shape(int hidden, GineShape* s) {
  if (hidden <= 0 || hidden % 4) return false;
  if (hidden >= 128) {
    if (hidden % 128 || hidden > 512) return false;
    s->lpr = 32; s->rpw = 1; s->ch = hidden / 128;
    return true;
  }
  const int l = hidden / 4;
  if (l & (l - 1)) return false;
  s->lpr = l; s->rpw = 32 / l; s->ch = 1;
  return true;
}

__device__ __forceinline__ float4 relu_msg(float4 v, float a, float4 w, float4 b) {
  float4 r;
  r.x = fmaxf(v.x + fmaf(a, w.x, b.x), 0.f);
  r.y = fmaxf(v.y + fmaf(a, w.y, b.y), 0.f);
  r.z = fmaxf(v.z + fmaf(a, w.z, b.z), 0.f);
  r.w = fmaxf(v.w + fmaf(a, w.w, b.w), 0.f);
  return r;
}
__device__ __forceinline__ void add4(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

template <int CH>
__global__ void __launch_bounds__(kGineThreads)
gine_aggr_fwd_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                     const float* __restrict__ attr, const float* __restrict__ w_edge, const float* __restrict__ b_edge,
                     const float* __restrict__ eps_ptr, float* __restrict__ h, int m, int hidden, int lpr) {
  const int lane = threadIdx.x & 31;
  const int rpw = 32 / lpr;
  const int sub = lane / lpr, sl = lane - sub * lpr;
  const int row = (blockIdx.x * kGineWarps + (threadIdx.x >> 5)) * rpw + sub;
  if (row >= m) return;
  const float self_scale = 1.0f + __ldg(eps_ptr);
  float4 w4[CH], b4[CH], acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int cc = 4 * (sl + 32 * c);
    w4[c] = ldg4(w_edge + cc);
    b4[c] = ldg4(b_edge + cc);
    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  int s = beg;
  for (; s + kEdgeUnroll <= end; s += kEdgeUnroll) {     // 4 independent row gathers in flight
    int src[kEdgeUnroll];
    float a[kEdgeUnroll];
    float4 v[kEdgeUnroll][CH];
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) { src[k] = __ldg(col + s + k); a[k] = __ldg(attr + s + k); }
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k)
#pragma unroll
      for (int c = 0; c < CH; ++c) v[k][c] = ldg4(x + (size_t)src[k] * hidden + 4 * (sl + 32 * c));
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k)                 // accumulate in slot order (= reference edge order)
#pragma unroll
      for (int c = 0; c < CH; ++c) add4(acc[c], relu_msg(v[k][c], a[k], w4[c], b4[c]));
  }
  for (; s < end; ++s) {
    const int src = __ldg(col + s);
    const float a = __ldg(attr + s);
#pragma unroll
    for (int c = 0; c < CH; ++c) add4(acc[c], relu_msg(ldg4(x + (size_t)src * hidden + 4 * (sl + 32 * c)), a, w4[c], b4[c]));
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int cc = 4 * (sl + 32 * c);
    const float4 xi = ldg4(x + (size_t)row * hidden + cc);
    float4 o;
    o.x = acc[c].x + self_scale * xi.x;
    o.y = acc[c].y + self_scale * xi.y;
    o.z = acc[c].z + self_scale * xi.z;
    o.w = acc[c].w + self_scale * xi.w;
    st4(h + (size_t)row * hidden + cc, o);
  }
}

__device__ __forceinline__ float4 masked(float4 g, float4 xj, float a, float4 w, float4 b) {
  float4 r;
  r.x = (xj.x + fmaf(a, w.x, b.x) > 0.f) ? g.x : 0.f;
  r.y = (xj.y + fmaf(a, w.y, b.y) > 0.f) ? g.y : 0.f;
  r.z = (xj.z + fmaf(a, w.z, b.z) > 0.f) ? g.z : 0.f;
  r.w = (xj.w + fmaf(a, w.w, b.w) > 0.f) ? g.w : 0.f;
  return r;
}

// dynamic shared memory: float red[kGineWarps * rpw][2 * hidden] + float red_eps[kGineWarps * rpw]
template <int CH>
__global__ void __launch_bounds__(kGineThreads)
gine_aggr_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const int* __restrict__ t_rowptr,
                     const int* __restrict__ t_dst, const float* __restrict__ t_attr, const float* __restrict__ w_edge,
                     const float* __restrict__ b_edge, const float* __restrict__ eps_ptr, const float* __restrict__ addend,
                     float* __restrict__ dx, float* __restrict__ partials, int m, int hidden, int lpr) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / lpr;
  const int sub = lane / lpr, sl = lane - sub * lpr;
  const int rows_per_block = kGineWarps * rpw;
  const float self_scale = 1.0f + __ldg(eps_ptr);
  float4 w4[CH], b4[CH], dw[CH], db[CH];
  double deps = 0.0;   // <g, x> has heavy cancellation over M*H products: accumulate across rows in float64
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int cc = 4 * (sl + 32 * c);
    w4[c] = ldg4(w_edge + cc);
    b4[c] = ldg4(b_edge + cc);
    dw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * rows_per_block + warp * rpw + sub; row < m; row += gridDim.x * rows_per_block) {
    float4 xj[CH], gj[CH], acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int cc = 4 * (sl + 32 * c);
      xj[c] = ldg4(x + (size_t)row * hidden + cc);
      gj[c] = ldg4(g + (size_t)row * hidden + cc);
      acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      deps += (double)(gj[c].x * xj[c].x + gj[c].y * xj[c].y + gj[c].z * xj[c].z + gj[c].w * xj[c].w);
    }
    const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
    int q = beg;
    for (; q + kEdgeUnroll <= end; q += kEdgeUnroll) {
      int d[kEdgeUnroll];
      float a[kEdgeUnroll];
      float4 v[kEdgeUnroll][CH];
#pragma unroll
      for (int k = 0; k < kEdgeUnroll; ++k) { d[k] = __ldg(t_dst + q + k); a[k] = __ldg(t_attr + q + k); }
#pragma unroll
      for (int k = 0; k < kEdgeUnroll; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) v[k][c] = ldg4(g + (size_t)d[k] * hidden + 4 * (sl + 32 * c));
#pragma unroll
      for (int k = 0; k < kEdgeUnroll; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float4 gm = masked(v[k][c], xj[c], a[k], w4[c], b4[c]);
          add4(acc[c], gm);
          add4(db[c], gm);
          dw[c].x = fmaf(gm.x, a[k], dw[c].x); dw[c].y = fmaf(gm.y, a[k], dw[c].y);
          dw[c].z = fmaf(gm.z, a[k], dw[c].z); dw[c].w = fmaf(gm.w, a[k], dw[c].w);
        }
    }
    for (; q < end; ++q) {
      const int d = __ldg(t_dst + q);
      const float a = __ldg(t_attr + q);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 gm = masked(ldg4(g + (size_t)d * hidden + 4 * (sl + 32 * c)), xj[c], a, w4[c], b4[c]);
        add4(acc[c], gm);
        add4(db[c], gm);
        dw[c].x = fmaf(gm.x, a, dw[c].x); dw[c].y = fmaf(gm.y, a, dw[c].y);
        dw[c].z = fmaf(gm.z, a, dw[c].z); dw[c].w = fmaf(gm.w, a, dw[c].w);
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int cc = 4 * (sl + 32 * c);
      float4 o;
      o.x = fmaf(self_scale, gj[c].x, acc[c].x);
      o.y = fmaf(self_scale, gj[c].y, acc[c].y);
      o.z = fmaf(self_scale, gj[c].z, acc[c].z);
      o.w = fmaf(self_scale, gj[c].w, acc[c].w);
      if (addend != nullptr) add4(o, ldg4(addend + (size_t)row * hidden + cc));
      st4(dx + (size_t)row * hidden + cc, o);
    }
  }
  // ---- block reduction of d w_edge, d b_edge, d eps in a fixed order
  float* red = smem;                                   // [rows_per_block][2*hidden]
  float* red_eps = smem + rows_per_block * 2 * hidden; // [rows_per_block]
  const int slot = warp * rpw + sub;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int cc = 4 * (sl + 32 * c);
    st4(red + (size_t)slot * 2 * hidden + cc, dw[c]);
    st4(red + (size_t)slot * 2 * hidden + hidden + cc, db[c]);
  }
  // lanes of one sub-warp hold disjoint column slices of the same rows: sum their <g, x> partials
  for (int o = lpr >> 1; o > 0; o >>= 1) deps += __shfl_xor_sync(0xffffffffu, deps, o);
  if (sl == 0) red_eps[slot] = (float)deps;
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * 3 * hidden;
  for (int j = threadIdx.x; j < 2 * hidden; j += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows_per_block; ++r) s += red[(size_t)r * 2 * hidden + j];
    out[j] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < rows_per_block; ++r) s += red_eps[r];
    out[2 * hidden] = s;
  }
}

__global__ void __launch_bounds__(256)
gine_bwd_finalize_kernel(const float* __restrict__ partials, int nblocks, int hidden, float* d_w, float* d_b, float* d_eps) {
  // blockIdx.x < ceil(2H/32): 32 columns x 8 partial strides; the last block reduces d eps
  __shared__ double sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ncol_blocks = ceil_div(2 * hidden, 32);
  if ((int)blockIdx.x < ncol_blocks) {
    const int j = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (j < 2 * hidden)
      for (int b = ty; b < nblocks; b += 8) s += (double)partials[(size_t)b * 3 * hidden + j];
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && j < 2 * hidden) {
      double t = 0.0;
      for (int k = 0; k < 8; ++k) t += sh[k][tx];
      if (j < hidden) d_w[j] = (float)t; else d_b[j - hidden] = (float)t;
    }
  } else {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += (double)partials[(size_t)b * 3 * hidden + 2 * hidden];
    s = warp_sum(s);
    if (tx == 0) sh[ty][0] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < 8; ++k) t += sh[k][0];
      d_eps[0] = (float)t;
    }
  }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_gine_aggr_fwd(const float* x, const int32_t* rowptr, const int32_t* col, const float* attr,
                                const float* w_edge, const float* b_edge, const float* eps, float* h, int num_nodes,
                                int hidden, void* stream) {
  GineShape sh;
  if (!x || !rowptr || !col || !attr || !w_edge || !b_edge || !eps || !h || num_nodes < 0)
    return fail(RC_ERR_ARG, "rc_gine_aggr_fwd: null pointer");
  if (!gine_shape(hidden, &sh)) return fail(RC_ERR_ARG, "rc_gine_aggr_fwd: hidden=%d unsupported (need 4|H, H/4 a power of two below 128, 128|H up to 512)", hidden);
  if (!aligned16(x) || !aligned16(h) || !aligned16(w_edge) || !aligned16(b_edge)) return fail(RC_ERR_ARG, "rc_gine_aggr_fwd: x, h, w_edge, b_edge must be 16-byte aligned");
  if (num_nodes == 0) return RC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (sh.lpr == 32 && num_nodes >= kRangedMinRows) return launch_gine_fwd_ranged(x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, s);
  const int grid = ceil_div(num_nodes, kGineWarps * sh.rpw);
  switch (sh.ch) {
    case 1: gine_aggr_fwd_kernel<1><<<grid, kGineThreads, 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, sh.lpr); break;
    case 2: gine_aggr_fwd_kernel<2><<<grid, kGineThreads, 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, sh.lpr); break;
    case 3: gine_aggr_fwd_kernel<3><<<grid, kGineThreads, 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, sh.lpr); break;
    default: gine_aggr_fwd_kernel<4><<<grid, kGineThreads, 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, sh.lpr); break;
  }
  return check_launch("gine_aggr_fwd_kernel");
}

extern "C" int rc_gine_aggr_bwd_nblocks(int num_nodes, int hidden) {
  GineShape sh;
  if (!gine_shape(hidden, &sh) || num_nodes < 0) return -1;
  if (sh.lpr == 32 && num_nodes >= kRangedMinRows) return gine_ranged_grid(num_nodes);
  const int nb = ceil_div(num_nodes > 0 ? num_nodes : 1, kGineWarps * sh.rpw);
  const int cap = kNumSMs * 8;
  return nb < cap ? nb : cap;
}

extern "C" int rc_gine_aggr_bwd(const float* g, const float* x, const int32_t* t_rowptr, const int32_t* t_dst,
                                const float* t_attr, const float* w_edge, const float* b_edge, const float* eps,
                                const float* addend, float* dx, float* partials, int num_nodes, int hidden, void* stream) {
  GineShape sh;
  if (!g || !x || !t_rowptr || !t_dst || !t_attr || !w_edge || !b_edge || !eps || !dx || !partials || num_nodes < 0)
    return fail(RC_ERR_ARG, "rc_gine_aggr_bwd: null pointer");
  if (!gine_shape(hidden, &sh)) return fail(RC_ERR_ARG, "rc_gine_aggr_bwd: hidden=%d unsupported", hidden);
  if (!aligned16(g) || !aligned16(x) || !aligned16(dx) || !aligned16(w_edge) || !aligned16(b_edge) || !aligned16(partials) ||
      (addend && !aligned16(addend)))
    return fail(RC_ERR_ARG, "rc_gine_aggr_bwd: g, x, dx, addend, w_edge, b_edge, partials must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = rc_gine_aggr_bwd_nblocks(num_nodes, hidden);
  if (sh.lpr == 32 && num_nodes >= kRangedMinRows)
    return launch_gine_bwd_ranged(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, num_nodes, hidden, s);
  const int rpb = kGineWarps * sh.rpw;
  const size_t smem = ((size_t)rpb * 2 * hidden + rpb) * sizeof(float);
#define RC_LAUNCH_BWD(CHV)                                                                                         \
  do {                                                                                                             \
    if (smem > 48 * 1024) cudaFuncSetAttribute(gine_aggr_bwd_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    gine_aggr_bwd_kernel<CHV><<<grid, kGineThreads, smem, s>>>(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, \
                                                               dx, partials, num_nodes, hidden, sh.lpr);            \
  } while (0)
  switch (sh.ch) {
    case 1: RC_LAUNCH_BWD(1); break;
    case 2: RC_LAUNCH_BWD(2); break;
    case 3: RC_LAUNCH_BWD(3); break;
    default: RC_LAUNCH_BWD(4); break;
  }
#undef RC_LAUNCH_BWD
  return check_launch("gine_aggr_bwd_kernel");
}

extern "C" int rc_gine_aggr_bwd_finalize(const float* partials, int nblocks, int hidden, float* d_w, float* d_b,
                                         float* d_eps, void* stream) {
  if (!partials || !d_w || !d_b || !d_eps || nblocks < 0 || hidden <= 0) return fail(RC_ERR_ARG, "rc_gine_aggr_bwd_finalize: bad argument");
  const int grid = ceil_div(2 * hidden, 32) + 1;
  gine_bwd_finalize_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(partials, nblocks, hidden, d_w, d_b, d_eps);
  return check_launch("gine_bwd_finalize_kernel");
}
