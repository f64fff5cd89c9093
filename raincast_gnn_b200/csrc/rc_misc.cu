// Small reductions around the GEMMs: BatchNorm statistics (forward / backward), the one-launch
// reduction of every split gradient, and the fused AdamW update on flat buffers.
#include "rc_common.cuh"

namespace rc {

// ---------------------------------------------------------------------------- BatchNorm forward stats
// block (32 columns, 8 tile strides); Chan's pairwise update in float64, merged in a fixed order.
struct Moments { double n, mean, m2; };
__device__ __forceinline__ void chan_merge(Moments& a, double nb, double mean_b, double m2_b) {
  if (nb <= 0.0) return;
  const double n = a.n + nb, delta = mean_b - a.mean;
  a.mean += delta * (nb / n);
  a.m2 += m2_b + delta * delta * (a.n * nb / n);
  a.n = n;
}

__global__ void __launch_bounds__(256)
bn_stats_finalize_kernel(const float* __restrict__ stats, int row_tiles, int row_tile, int m, int n, float eps,
                         float momentum, float* mean_out, float* rstd_out, float* running_mean, float* running_var,
                         long long* num_batches_tracked) {
  __shared__ double sh[8][3][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  Moments acc = {0.0, 0.0, 0.0};
  if (col < n)
    for (int t = ty; t < row_tiles; t += 8) {
      const int cnt = min(row_tile, m - t * row_tile);
      chan_merge(acc, (double)cnt, (double)stats[(size_t)t * 2 * n + col], (double)stats[(size_t)t * 2 * n + n + col]);
    }
  sh[ty][0][tx] = acc.n; sh[ty][1][tx] = acc.mean; sh[ty][2][tx] = acc.m2;
  __syncthreads();
  if (ty == 0 && col < n) {
    Moments tot = {0.0, 0.0, 0.0};
    for (int k = 0; k < 8; ++k) chan_merge(tot, sh[k][0][tx], sh[k][1][tx], sh[k][2][tx]);
    const double var_b = tot.m2 / (double)m;
    mean_out[col] = (float)tot.mean;
    rstd_out[col] = (float)(1.0 / sqrt(var_b + (double)eps));
    if (running_mean != nullptr) running_mean[col] = (1.0f - momentum) * running_mean[col] + momentum * (float)tot.mean;
    if (running_var != nullptr) {
      const float var_u = (float)(tot.m2 / (double)(m > 1 ? m - 1 : 1));
      running_var[col] = (1.0f - momentum) * running_var[col] + momentum * var_u;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) num_batches_tracked[0] += 1;
}

__global__ void bn_eval_prepare_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int n,
                                       float eps, float* mean, float* rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mean[i] = running_mean[i];
    rstd[i] = 1.0f / sqrtf(running_var[i] + eps);
  }
}

// ---------------------------------------------------------------------------- BatchNorm backward coefficients
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ stats, int row_tiles, int m, int n, int batch_stats, const float* __restrict__ gamma,
                       const float* __restrict__ mean, const float* __restrict__ rstd, float* d_gamma, float* d_beta,
                       float* c0, float* c1, float* c2) {
  __shared__ double sh[8][2][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  double s0 = 0.0, s1 = 0.0;
  if (col < n)
    for (int t = ty; t < row_tiles; t += 8) {
      s0 += (double)stats[(size_t)t * 2 * n + col];
      s1 += (double)stats[(size_t)t * 2 * n + n + col];
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
__global__ void __launch_bounds__(256) reduce_segments_kernel(const ReduceArgs args) {
  const rc_reduce_seg& sg = args.seg[blockIdx.y];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < sg.n; j += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < sg.parts; ++p) s += (double)__ldg(sg.src + (size_t)p * sg.stride + j);
    const float v = sg.scale * (float)s;
    float* out = sg.dst + (sg.row_len > 0 ? (size_t)(j / sg.row_len) * sg.dst_ld + (j % sg.row_len) : (size_t)j);
    *out = sg.accumulate ? *out + v : v;
  }
}

// ---------------------------------------------------------------------------- AdamW
// torch.optim.AdamW single-tensor update (decoupled weight decay, bias-corrected moments).
__global__ void adamw_tick_kernel(long long* step) { step[0] += 1; }

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
             float* __restrict__ exp_avg_sq, const long long* __restrict__ step, long long n, float lr, float beta1,
             float beta2, float eps, float weight_decay, float grad_scale) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)step[0];
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i] * grad_scale;
    float p = param[i] * (1.0f - lr * weight_decay);
    float m1 = exp_avg[i], v = exp_avg_sq[i];
    m1 = m1 + (g - m1) * (1.0f - beta1);                 // lerp_
    v = v * beta2 + (1.0f - beta2) * g * g;              // mul_ + addcmul_
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m1 / denom);
    param[i] = p; exp_avg[i] = m1; exp_avg_sq[i] = v;
  }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_bn_stats_finalize(const float* stats, int row_tiles, int row_tile, int m, int n, float eps,
                                    float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                                    int64_t* num_batches_tracked, void* stream) {
  if (!stats || !mean || !rstd || row_tiles <= 0 || row_tile <= 0 || m <= 0 || n <= 0) return fail(RC_ERR_ARG, "rc_bn_stats_finalize: bad argument");
  if (row_tiles != ceil_div(m, row_tile)) return fail(RC_ERR_ARG, "rc_bn_stats_finalize: row_tiles %d != ceil(%d/%d)", row_tiles, m, row_tile);
  bn_stats_finalize_kernel<<<ceil_div(n, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, row_tiles, row_tile, m, n, eps, momentum, mean, rstd, running_mean, running_var,
      reinterpret_cast<long long*>(num_batches_tracked));
  return check_launch("bn_stats_finalize_kernel");
}

extern "C" int rc_bn_eval_prepare(const float* running_mean, const float* running_var, int n, float eps, float* mean,
                                  float* rstd, void* stream) {
  if (!running_mean || !running_var || !mean || !rstd || n <= 0) return fail(RC_ERR_ARG, "rc_bn_eval_prepare: bad argument");
  bn_eval_prepare_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(running_mean, running_var, n, eps, mean, rstd);
  return check_launch("bn_eval_prepare_kernel");
}

extern "C" int rc_bn_bwd_finalize(const float* stats, int row_tiles, int m, int n, int batch_stats, const float* gamma, const float* mean,
                                  const float* rstd, float* d_gamma, float* d_beta, float* c0, float* c1, float* c2,
                                  void* stream) {
  if (!stats || !gamma || !mean || !rstd || !d_gamma || !d_beta || !c0 || !c1 || !c2 || row_tiles <= 0 || m <= 0 || n <= 0)
    return fail(RC_ERR_ARG, "rc_bn_bwd_finalize: bad argument");
  bn_bwd_finalize_kernel<<<ceil_div(n, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(stats, row_tiles, m, n, batch_stats, gamma, mean, rstd,
                                                                                        d_gamma, d_beta, c0, c1, c2);
  return check_launch("bn_bwd_finalize_kernel");
}

extern "C" int rc_reduce_segments(const rc_reduce_seg* segs, int n_segs, void* stream) {
  if ((!segs && n_segs > 0) || n_segs < 0) return fail(RC_ERR_ARG, "rc_reduce_segments: bad argument");
  for (int base = 0; base < n_segs; base += RC_REDUCE_MAX_SEGS) {
    const int cnt = n_segs - base < RC_REDUCE_MAX_SEGS ? n_segs - base : RC_REDUCE_MAX_SEGS;
    ReduceArgs args;
    int max_n = 0;
    for (int i = 0; i < cnt; ++i) {
      args.seg[i] = segs[base + i];
      if (!args.seg[i].src || !args.seg[i].dst || args.seg[i].n < 0 || args.seg[i].parts < 0)
        return fail(RC_ERR_ARG, "rc_reduce_segments: bad segment %d", base + i);
      if (args.seg[i].n > max_n) max_n = args.seg[i].n;
    }
    for (int i = cnt; i < RC_REDUCE_MAX_SEGS; ++i) args.seg[i] = rc_reduce_seg{nullptr, nullptr, 0, 0, 0, 0.f, 0, 0, 0};
    if (max_n == 0) continue;
    int gx = ceil_div(max_n, 256);
    if (gx > 64) gx = 64;
    reduce_segments_kernel<<<dim3(gx, cnt), 256, 0, static_cast<cudaStream_t>(stream)>>>(args);
    if (int e = check_launch("reduce_segments_kernel")) return e;
  }
  return RC_OK;
}

extern "C" int rc_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t* step, long long n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                             void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || n < 0) return fail(RC_ERR_ARG, "rc_adamw_step: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  adamw_tick_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(step));
  if (int e = check_launch("adamw_tick_kernel")) return e;
  if (n == 0) return RC_OK;
  long long blocks = ceil_div_ll(n, 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  adamw_kernel<<<(int)blocks, 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, reinterpret_cast<long long*>(step), n, lr, beta1,
                                           beta2, eps, weight_decay, grad_scale);
  return check_launch("adamw_kernel");
}
