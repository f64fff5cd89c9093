// GINE aggregation, large-graph path (H >= 128, many rows): big CTAs that walk CONTIGUOUS row ranges.
//
// Measured on B200 at BASELINE.json config 4 (100k nodes, 2.98 M edges, H = 128): with one 8-warp CTA per 8
// rows scattered round-robin over the SMs, every gathered 512-byte source row comes from L2 and the kernel sits
// on the L2 -> SM throughput cap (1.5 GB of gathers in 126 us = 12 TB/s).  When the station ids are ordered
// along a space-filling / breadth-first curve (graph.locality_order), neighbouring rows share most of their
// sources; a CTA of 32 warps marching through a contiguous id range keeps that patch of source rows in its
// SM's L1, so most gathers never leave the SM.  (Deeper per-warp gather pipelines with fewer resident warps
// were measured slower: 213 us.)
#include "rc_common.cuh"

namespace rc {

// one CTA per SM: 32 warps (64 registers each) for H = 128, 16 warps (128 registers) for wider rows.
// Measured alternatives at config 4 (forward): 24 warps with two batches of gathers in flight 143 us, 8 gathers per
// warp with shuffle-broadcast indices 213 us, this version 109 us - resident warps beat per-warp pipelining here.
__host__ __device__ constexpr int ranged_threads(int ch) { return ch == 1 ? 1024 : 512; }
// compile-time knobs kept for experiments (measured on B200 at config 4, forward): unroll/threads/CTAs-per-SM
//   4/1024/1: 109 us (default)   2/512/3: 134 us   3/512/3: 171 us   2/1024/2: 220 us   2/512/4: 225 us
#ifndef RC_FWD_UNROLL
#define RC_FWD_UNROLL 4
#endif
#ifndef RC_FWD_THREADS
#define RC_FWD_THREADS 1024
#endif
#ifndef RC_FWD_MINB
#define RC_FWD_MINB 1
#endif
constexpr int kRUnroll = 4;

// Packed fp32x2 arithmetic (fma.rn.f32x2 / add.rn.f32x2, new on sm_100): one issue slot does two lanes of the
// edge Linear, the message add and the accumulation; rounding is identical to the scalar sequence.
__device__ __forceinline__ void relu_acc(float4& acc, float4 v, float a, float4 w, float4 b) {
  const float2 a2 = make_float2(a, a);
  float2 z0 = __fadd2_rn(make_float2(v.x, v.y), __ffma2_rn(a2, make_float2(w.x, w.y), make_float2(b.x, b.y)));
  float2 z1 = __fadd2_rn(make_float2(v.z, v.w), __ffma2_rn(a2, make_float2(w.z, w.w), make_float2(b.z, b.w)));
  z0.x = fmaxf(z0.x, 0.f); z0.y = fmaxf(z0.y, 0.f);
  z1.x = fmaxf(z1.x, 0.f); z1.y = fmaxf(z1.y, 0.f);
  const float2 s0 = __fadd2_rn(make_float2(acc.x, acc.y), z0), s1 = __fadd2_rn(make_float2(acc.z, acc.w), z1);
  acc = make_float4(s0.x, s0.y, s1.x, s1.y);
}

template <int CH>
__global__ void __launch_bounds__(CH == 1 ? RC_FWD_THREADS : 512, CH == 1 ? RC_FWD_MINB : 1)
gine_aggr_fwd_ranged_kernel(const float* __restrict__ x, const int* __restrict__ rowptr, const int* __restrict__ col,
                            const float* __restrict__ attr, const float* __restrict__ w_edge, const float* __restrict__ b_edge,
                            const float* __restrict__ eps_ptr, float* __restrict__ h, int m, int hidden, int rows_per_cta) {
  constexpr int kRWarps = (CH == 1 ? RC_FWD_THREADS : 512) / 32;
  constexpr int kRUnroll = CH == 1 ? RC_FWD_UNROLL : 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float self_scale = 1.0f + __ldg(eps_ptr);
  float4 w4[CH], b4[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    w4[c] = ldg4(w_edge + 4 * (lane + 32 * c));
    b4[c] = ldg4(b_edge + 4 * (lane + 32 * c));
  }
  const float* xl = x + 4 * lane;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
  for (int row = r0 + warp; row < r1; row += kRWarps) {      // the 32 warps sweep 32 consecutive rows at a time
    float4 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    int s = beg;
    int src_n[kRUnroll];
    float a_n[kRUnroll];
    if (s + kRUnroll <= end) {
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k) { src_n[k] = __ldg(col + s + k); a_n[k] = __ldg(attr + s + k); }
    }
    for (; s + kRUnroll <= end; s += kRUnroll) {
      int src[kRUnroll];
      float a[kRUnroll];
      float4 v[kRUnroll][CH];
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k) { src[k] = src_n[k]; a[k] = a_n[k]; }
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) v[k][c] = ld4(xl + (size_t)src[k] * hidden + 128 * c);   // ld.global: allocate in L1
      if (s + 2 * kRUnroll <= end) {                          // next batch's indices travel while this batch is summed
#pragma unroll
        for (int k = 0; k < kRUnroll; ++k) { src_n[k] = __ldg(col + s + kRUnroll + k); a_n[k] = __ldg(attr + s + kRUnroll + k); }
      }
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k)                       // slot order = reference edge order
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc(acc[c], v[k][c], a[k], w4[c], b4[c]);
    }
    for (; s < end; ++s) {
      const int src = __ldg(col + s);
      const float a = __ldg(attr + s);
#pragma unroll
      for (int c = 0; c < CH; ++c) relu_acc(acc[c], ld4(xl + (size_t)src * hidden + 128 * c), a, w4[c], b4[c]);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const float4 xi = ld4(xl + (size_t)row * hidden + 128 * c);
      float4 o;
      o.x = acc[c].x + self_scale * xi.x;
      o.y = acc[c].y + self_scale * xi.y;
      o.z = acc[c].z + self_scale * xi.z;
      o.w = acc[c].w + self_scale * xi.w;
      st4(h + (size_t)row * hidden + 4 * lane + 128 * c, o);
    }
  }
}

__device__ __forceinline__ void masked_acc(float4& acc, float4& acc_a, float4 g, float4 xj, float a, float4 w, float4 b) {
  const float2 a2 = make_float2(a, a);
  const float2 z0 = __fadd2_rn(make_float2(xj.x, xj.y), __ffma2_rn(a2, make_float2(w.x, w.y), make_float2(b.x, b.y)));
  const float2 z1 = __fadd2_rn(make_float2(xj.z, xj.w), __ffma2_rn(a2, make_float2(w.z, w.w), make_float2(b.z, b.w)));
  const float2 g0 = make_float2(z0.x > 0.f ? g.x : 0.f, z0.y > 0.f ? g.y : 0.f);
  const float2 g1 = make_float2(z1.x > 0.f ? g.z : 0.f, z1.y > 0.f ? g.w : 0.f);
  const float2 s0 = __fadd2_rn(make_float2(acc.x, acc.y), g0), s1 = __fadd2_rn(make_float2(acc.z, acc.w), g1);
  const float2 t0 = __ffma2_rn(g0, a2, make_float2(acc_a.x, acc_a.y)), t1 = __ffma2_rn(g1, a2, make_float2(acc_a.z, acc_a.w));
  acc = make_float4(s0.x, s0.y, s1.x, s1.y);
  acc_a = make_float4(t0.x, t0.y, t1.x, t1.y);
}

// dynamic shared memory: float red[kRWarps][2 * hidden] + float red_eps[kRWarps]
template <int CH>
__global__ void __launch_bounds__(ranged_threads(CH), 1)
gine_aggr_bwd_ranged_kernel(const float* __restrict__ g, const float* __restrict__ x, const int* __restrict__ t_rowptr,
                            const int* __restrict__ t_dst, const float* __restrict__ t_attr, const float* __restrict__ w_edge,
                            const float* __restrict__ b_edge, const float* __restrict__ eps_ptr, const float* __restrict__ addend,
                            float* __restrict__ dx, float* __restrict__ partials, int m, int hidden, int rows_per_cta) {
  constexpr int kRWarps = ranged_threads(CH) / 32;
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float self_scale = 1.0f + __ldg(eps_ptr);
  float4 w4[CH], b4[CH], dw[CH], db[CH];
  double deps = 0.0;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    w4[c] = ldg4(w_edge + 4 * (lane + 32 * c));
    b4[c] = ldg4(b_edge + 4 * (lane + 32 * c));
    dw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* gl = g + 4 * lane;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(m, r0 + rows_per_cta);
  for (int row = r0 + warp; row < r1; row += kRWarps) {
    float4 xj[CH], gj[CH], acc[CH], acc_a[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      xj[c] = ldg4(x + (size_t)row * hidden + 4 * lane + 128 * c);
      gj[c] = ld4(gl + (size_t)row * hidden + 128 * c);
      acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc_a[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      deps += (double)(gj[c].x * xj[c].x + gj[c].y * xj[c].y + gj[c].z * xj[c].z + gj[c].w * xj[c].w);
    }
    const int beg = __ldg(t_rowptr + row), end = __ldg(t_rowptr + row + 1);
    int q = beg;
    int d_n[kRUnroll];
    float a_n[kRUnroll];
    if (q + kRUnroll <= end) {
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k) { d_n[k] = __ldg(t_dst + q + k); a_n[k] = __ldg(t_attr + q + k); }
    }
    for (; q + kRUnroll <= end; q += kRUnroll) {
      int d[kRUnroll];
      float a[kRUnroll];
      float4 v[kRUnroll][CH];
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k) { d[k] = d_n[k]; a[k] = a_n[k]; }
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) v[k][c] = ld4(gl + (size_t)d[k] * hidden + 128 * c);
      if (q + 2 * kRUnroll <= end) {
#pragma unroll
        for (int k = 0; k < kRUnroll; ++k) { d_n[k] = __ldg(t_dst + q + kRUnroll + k); a_n[k] = __ldg(t_attr + q + kRUnroll + k); }
      }
#pragma unroll
      for (int k = 0; k < kRUnroll; ++k)
#pragma unroll
        for (int c = 0; c < CH; ++c) masked_acc(acc[c], acc_a[c], v[k][c], xj[c], a[k], w4[c], b4[c]);
    }
    for (; q < end; ++q) {
      const int d = __ldg(t_dst + q);
      const float a = __ldg(t_attr + q);
#pragma unroll
      for (int c = 0; c < CH; ++c) masked_acc(acc[c], acc_a[c], ld4(gl + (size_t)d * hidden + 128 * c), xj[c], a, w4[c], b4[c]);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      // sum_e gm_e = acc and sum_e gm_e a_e = acc_a: the bias / weight gradients take them once per row
      db[c].x += acc[c].x; db[c].y += acc[c].y; db[c].z += acc[c].z; db[c].w += acc[c].w;
      dw[c].x += acc_a[c].x; dw[c].y += acc_a[c].y; dw[c].z += acc_a[c].z; dw[c].w += acc_a[c].w;
      float4 o;
      o.x = fmaf(self_scale, gj[c].x, acc[c].x);
      o.y = fmaf(self_scale, gj[c].y, acc[c].y);
      o.z = fmaf(self_scale, gj[c].z, acc[c].z);
      o.w = fmaf(self_scale, gj[c].w, acc[c].w);
      if (addend != nullptr) {
        const float4 ad = ldg4(addend + (size_t)row * hidden + 4 * lane + 128 * c);
        o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
      }
      st4(dx + (size_t)row * hidden + 4 * lane + 128 * c, o);
    }
  }
  float* red = smem;
  float* red_eps = smem + kRWarps * 2 * hidden;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    st4(red + (size_t)warp * 2 * hidden + 4 * lane + 128 * c, dw[c]);
    st4(red + (size_t)warp * 2 * hidden + hidden + 4 * lane + 128 * c, db[c]);
  }
  deps = warp_sum(deps);
  if (lane == 0) red_eps[warp] = (float)deps;
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * 3 * hidden;
  for (int j = threadIdx.x; j < 2 * hidden; j += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < kRWarps; ++r) s += red[(size_t)r * 2 * hidden + j];
    out[j] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < kRWarps; ++r) s += red_eps[r];
    out[2 * hidden] = s;
  }
}

int gine_ranged_grid(int m) { (void)m; return kNumSMs; }
static int gine_fwd_grid() { return kNumSMs * RC_FWD_MINB; }

int launch_gine_fwd_ranged(const float* x, const int* rowptr, const int* col, const float* attr, const float* w_edge,
                           const float* b_edge, const float* eps, float* h, int m, int hidden, cudaStream_t s) {
  const int grid = hidden == 128 ? gine_fwd_grid() : gine_ranged_grid(m);
  const int rpc = ceil_div(m, grid);
  switch (hidden / 128) {
    case 1: gine_aggr_fwd_ranged_kernel<1><<<grid, RC_FWD_THREADS, 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, m, hidden, rpc); break;
    case 2: gine_aggr_fwd_ranged_kernel<2><<<grid, ranged_threads(2), 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, m, hidden, rpc); break;
    case 3: gine_aggr_fwd_ranged_kernel<3><<<grid, ranged_threads(3), 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, m, hidden, rpc); break;
    default: gine_aggr_fwd_ranged_kernel<4><<<grid, ranged_threads(4), 0, s>>>(x, rowptr, col, attr, w_edge, b_edge, eps, h, m, hidden, rpc); break;
  }
  return check_launch("gine_aggr_fwd_ranged_kernel");
}

template <int CH>
static int launch_bwd_ch(const float* g, const float* x, const int* t_rowptr, const int* t_dst, const float* t_attr,
                         const float* w_edge, const float* b_edge, const float* eps, const float* addend, float* dx,
                         float* partials, int m, int hidden, cudaStream_t s) {
  const int grid = gine_ranged_grid(m);
  const int rpc = ceil_div(m, grid);
  constexpr int kRWarps = ranged_threads(CH) / 32;
  const size_t smem = ((size_t)kRWarps * 2 * hidden + kRWarps) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    cudaFuncSetAttribute(gine_aggr_bwd_ranged_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  gine_aggr_bwd_ranged_kernel<CH><<<grid, ranged_threads(CH), smem, s>>>(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx,
                                                               partials, m, hidden, rpc);
  return check_launch("gine_aggr_bwd_ranged_kernel");
}

int launch_gine_bwd_ranged(const float* g, const float* x, const int* t_rowptr, const int* t_dst, const float* t_attr,
                           const float* w_edge, const float* b_edge, const float* eps, const float* addend, float* dx,
                           float* partials, int m, int hidden, cudaStream_t s) {
  switch (hidden / 128) {
    case 1: return launch_bwd_ch<1>(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, m, hidden, s);
    case 2: return launch_bwd_ch<2>(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, m, hidden, s);
    case 3: return launch_bwd_ch<3>(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, m, hidden, s);
    default: return launch_bwd_ch<4>(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, m, hidden, s);
  }
}

}  // namespace rc
