// GINE aggregation tile functions (small-graph path) shared by rc_gine.cu and the step program (rc_prog.cu).
#pragma once
#include "rc_common.cuh"

namespace rc {

struct GineFwdP {
  const float* x;
  const int* rowptr;
  const int* col;
  const float* attr;
  const float* w_edge;
  const float* b_edge;
  const float* eps_ptr;
  float* h;
  int m;
  int hidden;
  int lpr;
};

struct GineBwdP {
  const float* g;
  const float* x;
  const int* t_rowptr;
  const int* t_dst;
  const float* t_attr;
  const float* w_edge;
  const float* b_edge;
  const float* eps_ptr;
  const float* addend;
  float* dx;
  float* partials;
  int m;
  int hidden;
  int lpr;
};

struct GineFinP {
  const float* partials;
  int nblocks;
  int hidden;
  float* d_w;
  float* d_b;
  float* d_eps;
};


constexpr int kGineThreads = 256;
constexpr int kGineWarps = kGineThreads / 32;
constexpr int kEdgeUnroll = 4;

struct GineShape {
  int lpr;   // lanes per row
  int rpw;   // rows per warp
  int ch;    // float4 chunks per lane
};

inline bool gine_shape(int hidden, GineShape* s) {
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
__device__ __forceinline__ void gine_fwd_tile(const GineFwdP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ x = p.x;
  const int* __restrict__ rowptr = p.rowptr;
  const int* __restrict__ col = p.col;
  const float* __restrict__ attr = p.attr;
  const float* __restrict__ w_edge = p.w_edge;
  const float* __restrict__ b_edge = p.b_edge;
  const float* __restrict__ eps_ptr = p.eps_ptr;
  float* __restrict__ h = p.h;
  int m = p.m;
  int hidden = p.hidden;
  int lpr = p.lpr;
  (void)bid; (void)gdim;

  const int lane = threadIdx.x & 31;
  const int rpw = 32 / lpr;
  const int sub = lane / lpr, sl = lane - sub * lpr;
  const int row = (bid.x * kGineWarps + (threadIdx.x >> 5)) * rpw + sub;
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
__device__ __forceinline__ void gine_bwd_tile(const GineBwdP& p, const uint3 bid, const uint3 gdim, float* smem) {
  const float* __restrict__ g = p.g;
  const float* __restrict__ x = p.x;
  const int* __restrict__ t_rowptr = p.t_rowptr;
  const int* __restrict__ t_dst = p.t_dst;
  const float* __restrict__ t_attr = p.t_attr;
  const float* __restrict__ w_edge = p.w_edge;
  const float* __restrict__ b_edge = p.b_edge;
  const float* __restrict__ eps_ptr = p.eps_ptr;
  const float* __restrict__ addend = p.addend;
  float* __restrict__ dx = p.dx;
  float* __restrict__ partials = p.partials;
  int m = p.m;
  int hidden = p.hidden;
  int lpr = p.lpr;
  (void)bid; (void)gdim;

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
  for (int row = bid.x * rows_per_block + warp * rpw + sub; row < m; row += gdim.x * rows_per_block) {
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
  float* out = partials + (size_t)bid.x * 3 * hidden;
  for (int j = threadIdx.x; j < 2 * hidden; j += kGineThreads) {
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

__device__ __forceinline__ void gine_fin_tile(const GineFinP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ partials = p.partials;
  int nblocks = p.nblocks;
  int hidden = p.hidden;
  float* d_w = p.d_w;
  float* d_b = p.d_b;
  float* d_eps = p.d_eps;
  (void)bid; (void)gdim;

  // bid.x < ceil(2H/32): 32 columns x 8 partial strides; the last block reduces d eps
  __shared__ double sh[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ncol_blocks = ceil_div(2 * hidden, 32);
  if ((int)bid.x < ncol_blocks) {
    const int j = bid.x * 32 + tx;
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
    for (int b = threadIdx.x; b < nblocks; b += kGineThreads) s += (double)partials[(size_t)b * 3 * hidden + 2 * hidden];
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
