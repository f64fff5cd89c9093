// BatchNorm finalisers, segmented reduction and AdamW tile functions shared by rc_misc.cu and rc_prog.cu.
#pragma once
#include "rc_common.cuh"

namespace rc {

struct BnStatsFinP {
  const float* stats;
  int row_tiles;
  int row_tile;
  int m;
  int n;
  float eps;
  float momentum;
  float* mean_out;
  float* rstd_out;
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
};

struct BnEvalP {
  const float* running_mean;
  const float* running_var;
  int n;
  float eps;
  float* mean;
  float* rstd;
};

struct BnBwdFinP {
  const float* stats;
  int row_tiles;
  int m;
  int n;
  int batch_stats;
  const float* gamma;
  const float* mean;
  const float* rstd;
  float* d_gamma;
  float* d_beta;
  float* c0;
  float* c1;
  float* c2;
};

struct AdamTickP {
  long long* step;
};

struct AdamP {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long* step;            // steps taken so far; this launch is step[0] + 1 and the last CTA to finish writes it back
  long long n;
  float lr;
  float beta1;
  float beta2;
  float eps;
  float weight_decay;
  float grad_scale;
};



// ---------------------------------------------------------------------------- BatchNorm forward stats
// block (32 columns, 8 tile strides); per-tile (count, mean, M2) combined in float64 in a fixed order.
__device__ __forceinline__ void bn_stats_fin_tile(const BnStatsFinP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ stats = p.stats;
  int row_tiles = p.row_tiles;
  int row_tile = p.row_tile;
  int m = p.m;
  int n = p.n;
  float eps = p.eps;
  float momentum = p.momentum;
  float* mean_out = p.mean_out;
  float* rstd_out = p.rstd_out;
  float* running_mean = p.running_mean;
  float* running_var = p.running_var;
  long long* num_batches_tracked = p.num_batches_tracked;
  (void)bid; (void)gdim;

  // Two passes instead of a chain of pairwise (Chan) merges: mean = sum_t cnt_t mean_t / M, then
  // M2 = sum_t [M2_t + cnt_t (mean_t - mean)^2] - the same quantity in real arithmetic, float64 throughout, fixed
  // summation order, and no dependent divide per tile (the merge chain made this kernel 9 us at 122 tiles, four times
  // on the forward critical path).  A thread keeps its tiles' statistics in registers across both passes.
  __shared__ double sh[8][33];
  __shared__ double sh_mean[33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = bid.x * 32 + tx;
  constexpr int kHold = 16;                        // tiles per thread kept in registers (8 * 16 = 128 tiles per pass)
  float mu[kHold], m2[kHold];
  double s = 0.0;
  const bool in_regs = row_tiles <= 8 * kHold;
  if (col < n) {
#pragma unroll
    for (int k = 0; k < kHold; ++k) {
      const int t = ty + 8 * k;
      mu[k] = 0.f; m2[k] = 0.f;
      if (t < row_tiles) {
        mu[k] = __ldg(stats + (size_t)t * 2 * n + col);
        m2[k] = __ldg(stats + (size_t)t * 2 * n + n + col);
      }
    }
#pragma unroll
    for (int k = 0; k < kHold; ++k) {
      const int t = ty + 8 * k;
      if (t < row_tiles) s += (double)min(row_tile, m - t * row_tile) * (double)mu[k];
    }
    for (int t = ty + 8 * kHold; t < row_tiles; t += 8)
      s += (double)min(row_tile, m - t * row_tile) * (double)__ldg(stats + (size_t)t * 2 * n + col);
  }
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sh[k][tx];
    sh_mean[tx] = tot / (double)m;
  }
  __syncthreads();
  const double mean = sh_mean[tx];
  double q = 0.0;
  if (col < n) {
#pragma unroll
    for (int k = 0; k < kHold; ++k) {
      const int t = ty + 8 * k;
      if (t < row_tiles) {
        const double d = (double)mu[k] - mean;
        q += (double)m2[k] + (double)min(row_tile, m - t * row_tile) * d * d;
      }
    }
    for (int t = ty + 8 * kHold; t < row_tiles; t += 8) {
      const double d = (double)__ldg(stats + (size_t)t * 2 * n + col) - mean;
      q += (double)__ldg(stats + (size_t)t * 2 * n + n + col) + (double)min(row_tile, m - t * row_tile) * d * d;
    }
  }
  (void)in_regs;
  sh[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && col < n) {
    double tot_m2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot_m2 += sh[k][tx];
    const double var_b = tot_m2 / (double)m;
    mean_out[col] = (float)mean;
    rstd_out[col] = (float)(1.0 / sqrt(var_b + (double)eps));
    if (running_mean != nullptr) running_mean[col] = (1.0f - momentum) * running_mean[col] + momentum * (float)mean;
    if (running_var != nullptr) {
      const float var_u = (float)(tot_m2 / (double)(m > 1 ? m - 1 : 1));
      running_var[col] = (1.0f - momentum) * running_var[col] + momentum * var_u;
    }
  }
  if (bid.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) num_batches_tracked[0] += 1;
}

__device__ __forceinline__ void bn_eval_tile(const BnEvalP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ running_mean = p.running_mean;
  const float* __restrict__ running_var = p.running_var;
  int n = p.n;
  float eps = p.eps;
  float* mean = p.mean;
  float* rstd = p.rstd;
  (void)bid; (void)gdim;

  const int i = bid.x * 256 + threadIdx.x;
  if (i < n) {
    mean[i] = running_mean[i];
    rstd[i] = 1.0f / sqrtf(running_var[i] + eps);
  }
}

// ---------------------------------------------------------------------------- BatchNorm backward coefficients
__device__ __forceinline__ void bn_bwd_fin_tile(const BnBwdFinP& p, const uint3 bid, const uint3 gdim) {
  const float* __restrict__ stats = p.stats;
  int row_tiles = p.row_tiles;
  int m = p.m;
  int n = p.n;
  int batch_stats = p.batch_stats;
  const float* __restrict__ gamma = p.gamma;
  const float* __restrict__ mean = p.mean;
  const float* __restrict__ rstd = p.rstd;
  float* d_gamma = p.d_gamma;
  float* d_beta = p.d_beta;
  float* c0 = p.c0;
  float* c1 = p.c1;
  float* c2 = p.c2;
  (void)bid; (void)gdim;

  __shared__ double sh[8][2][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = bid.x * 32 + tx;
  double s0 = 0.0, s1 = 0.0;
  if (col < n)
    for (int t0 = ty; t0 < row_tiles; t0 += 8 * 16) {
      float a0[16], a1[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int t = t0 + 8 * k;
        a0[k] = t < row_tiles ? __ldg(stats + (size_t)t * 2 * n + col) : 0.f;
        a1[k] = t < row_tiles ? __ldg(stats + (size_t)t * 2 * n + n + col) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) { s0 += (double)a0[k]; s1 += (double)a1[k]; }
    }
  sh[ty][0][tx] = s0; sh[ty][1][tx] = s1;
  __syncthreads();
  if (ty == 0 && col < n) {
    double db = 0.0, dg = 0.0;
    for (int k = 0; k < 8; ++k) { db += sh[k][0][tx]; dg += sh[k][1][tx]; }
    d_beta[col] = (float)db;
    d_gamma[col] = (float)dg;
    // d t = gamma*rstd * (dz - d_beta/M - xhat * d_gamma/M),  xhat = (t - mean) * rstd
    const double gr = (double)gamma[col] * (double)rstd[col];
    c0[col] = (float)gr;
    // (eval mode, running statistics: the mean/variance are constants and only c0 survives)
    c1[col] = batch_stats ? (float)(-gr * (double)rstd[col] * dg / (double)m) : 0.f;
    c2[col] = batch_stats ? (float)(-gr * db / (double)m) : 0.f;
    (void)mean;
  }
}

// ---------------------------------------------------------------------------- segmented partial reduction
struct ReduceArgs { rc_reduce_seg seg[RC_REDUCE_MAX_SEGS]; };
using ReduceP = ReduceArgs;
__device__ __forceinline__ void reduce_tile(const ReduceP& p, const uint3 bid, const uint3 gdim) {
  const ReduceArgs& args = p;
  // 32 outputs per pass; warp w sums parts w, w + 8, ... (8 loads in flight per lane), the eight partial sums meet
  // in shared memory and are added in warp order: fixed order, float64, one L2 round trip per 64 parts instead of
  // one per part (the 244-part DeepSets gradient took 26 us when every thread walked all parts alone)
  __shared__ double sh[8][33];
  const rc_reduce_seg& sg = args.seg[bid.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = bid.x * 32; j0 < sg.n; j0 += gdim.x * 32) {
    const int j = j0 + lane;
    double s = 0.0;
    if (j < sg.n) {
      for (int q0 = warp; q0 < sg.parts; q0 += 64) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int q = q0 + 8 * k;
          v[k] = q < sg.parts ? __ldg(sg.src + (size_t)q * sg.stride + j) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += (double)v[k];
      }
    }
    sh[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && j < sg.n) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += sh[w][lane];
      const float v = sg.scale * (float)tot;
      float* out = sg.dst + (sg.row_len > 0 ? (size_t)(j / sg.row_len) * sg.dst_ld + (j % sg.row_len) : (size_t)j);
      *out = sg.accumulate ? *out + v : v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------- AdamW
// torch.optim.AdamW single-tensor update (decoupled weight decay, bias-corrected moments).
__device__ __forceinline__ void adamw_tick_tile(const AdamTickP& p, const uint3 bid, const uint3 gdim) {
  long long* step = p.step;
  (void)bid; (void)gdim;
  if (threadIdx.x == 0) step[0] += 1;
}

// The step counter is read by every CTA at its start and advanced by the CTA that finishes last (a device-wide
// arrival counter that leaves itself at zero): one launch per optimiser step, replayable inside a CUDA graph.
static __device__ unsigned int g_adamw_arrivals = 0;
__device__ __forceinline__ void adamw_finish(long long* step, unsigned int n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&g_adamw_arrivals, 1u) == n_ctas - 1) {
      g_adamw_arrivals = 0;
      step[0] += 1;
      __threadfence();
    }
  }
}

// One element of torch.optim.AdamW's update (decoupled decay, lerp_, mul_ + addcmul_, bias corrections as torch computes
// them), with every rounding spelled out: rc_adamw_step and the peer-memory exchange kernel (rc_p2p_step) must produce
// the same bits from the same summed gradient - the data-parallel check compares their loss trajectories for equality -
// and that cannot be left to how the compiler contracts two differently shaped loops.
struct AdamCoef {
  float grad_scale, decay, one_m_b1, beta2, one_m_b2, bc2_sqrt, eps, step_size;
};
__device__ __forceinline__ void adam_update(const AdamCoef& c, float g, float& w, float& m, float& v) {
  g = __fmul_rn(g, c.grad_scale);
  const float q = __fmul_rn(w, c.decay);
  m = __fmaf_rn(__fsub_rn(g, m), c.one_m_b1, m);
  v = __fmaf_rn(__fmul_rn(c.one_m_b2, g), g, __fmul_rn(v, c.beta2));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
  w = __fmaf_rn(-c.step_size, __fdiv_rn(m, denom), q);
}

// 128-bit accesses when the buffers allow it (the engine's flat buffers always do).  Nothing is read before the wait for
// the previous kernel: two optimiser steps may follow each other directly, and the first one writes what the second reads.
__device__ __forceinline__ void adamw_tile(const AdamP& p, const uint3 bid, const uint3 gdim) {
  float* __restrict__ param = p.param;
  const float* __restrict__ grad = p.grad;
  float* __restrict__ exp_avg = p.exp_avg;
  float* __restrict__ exp_avg_sq = p.exp_avg_sq;
  const long long* __restrict__ step = p.step;
  const long long n = p.n;
  const float lr = p.lr, beta1 = p.beta1, beta2 = p.beta2, eps = p.eps, weight_decay = p.weight_decay, grad_scale = p.grad_scale;

  const bool vec = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                     reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0;
  const long long n4 = vec ? n / 4 : 0;
  const long long stride = (long long)gdim.x * 256;
  const long long first = (long long)bid.x * 256 + threadIdx.x;
  pdl_entry();
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)(step[0] + 1);
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const AdamCoef c{grad_scale, 1.0f - lr * weight_decay, 1.0f - beta1, beta2, 1.0f - beta2, s_bc2_sqrt, eps, s_step_size};
  auto update = [&](float g, float& pw, float& pm, float& pv) { adam_update(c, g, pw, pm, pv); };
  for (long long i = first; i < n4; i += stride) {
    float4 w = reinterpret_cast<const float4*>(param)[i], m1 = reinterpret_cast<const float4*>(exp_avg)[i],
           v2 = reinterpret_cast<const float4*>(exp_avg_sq)[i];
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    update(g.x, w.x, m1.x, v2.x);
    update(g.y, w.y, m1.y, v2.y);
    update(g.z, w.z, m1.z, v2.z);
    update(g.w, w.w, m1.w, v2.w);
    reinterpret_cast<float4*>(param)[i] = w;
    reinterpret_cast<float4*>(exp_avg)[i] = m1;
    reinterpret_cast<float4*>(exp_avg_sq)[i] = v2;
  }
  for (long long i = 4 * n4 + first; i < n; i += stride) {       // unaligned buffers / the last n % 4 elements
    float pw = param[i], pm = exp_avg[i], pv = exp_avg_sq[i];
    update(grad[i], pw, pm, pv);
    param[i] = pw; exp_avg[i] = pm; exp_avg_sq[i] = pv;
  }
  adamw_finish(p.step, gdim.x);
}


}  // namespace rc
