// Small reductions around the GEMMs: BatchNorm statistics (forward / backward), the one-launch
// reduction of every split gradient, and the fused AdamW update on flat buffers.

#include "rc_misc_tile.cuh"

namespace rc {

__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(const BnStatsFinP p) {
  pdl_entry();
  bn_stats_fin_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(256) bn_eval_prepare_kernel(const BnEvalP p) {
  pdl_entry();
  bn_eval_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const BnBwdFinP p) {
  pdl_entry();
  bn_bwd_fin_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(256) reduce_segments_kernel(const ReduceP p) {
  pdl_entry();
  reduce_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(32) adamw_tick_kernel(const AdamTickP p) {
  pdl_entry();
  adamw_tick_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(256) adamw_kernel(const AdamP p) {
  adamw_tile(p, blockIdx, gridDim);        // (griddepcontrol.wait is its first statement)
}

// One batch of forecast dates out of a device-resident split: the three per-date blocks (node features, ensemble, targets)
// of `n_batch` dates are copied into the step's static inputs by one launch; blockIdx.y = position in the batch.
struct GatherDatesP {
  const float* x_all; const float* ens_all; const float* y_all;
  const long long* dates;
  long long x_len, ens_len, y_len;     // floats per date
  int n_dates;
  float* x; float* ens; float* y;
  int* bad;                            // set to 1 when a date index is out of range (the copy is skipped)
  // epoch mode (rc_gather_dates_step): `dates` is the epoch's order [n_batches][gridDim.y]; the batch taken is number
  // *step - base[0] - the optimiser's step counter minus its value at the start of the epoch; base[1] = batches in this
  // epoch (all on the device: a captured graph takes its next batch with no host write in between); outside
  // [0, min(n_batches, base[1])): *bad = 2, nothing copied
  const long long* step; const long long* base;
  int n_batches;
};

__global__ void __launch_bounds__(256) gather_dates_kernel(const GatherDatesP p) {
  pdl_entry();
  const int b = blockIdx.y;
  const long long* dates = p.dates;
  if (p.step != nullptr) {
    const long long it = *p.step - p.base[0];
    if (it < 0 || it >= p.n_batches || it >= p.base[1]) {
      if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(p.bad, 2);
      return;
    }
    dates += it * gridDim.y;
  }
  const long long d = dates[b];
  if (d < 0 || d >= p.n_dates) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(p.bad, 1);
    return;
  }
  const long long total = p.x_len + p.ens_len + p.y_len;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    if (i < p.x_len) p.x[b * p.x_len + i] = __ldg(p.x_all + d * p.x_len + i);
    else if (i < p.x_len + p.ens_len) p.ens[b * p.ens_len + (i - p.x_len)] = __ldg(p.ens_all + d * p.ens_len + (i - p.x_len));
    else p.y[b * p.y_len + (i - p.x_len - p.ens_len)] = __ldg(p.y_all + d * p.y_len + (i - p.x_len - p.ens_len));
  }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_gather_dates(const float* x_all, const float* ens_all, const float* y_all, const int64_t* dates, int n_batch,
                               int n_dates, long long x_len, long long ens_len, long long y_len, float* x, float* ens, float* y,
                               int32_t* bad, void* stream) {
  if (!x_all || !ens_all || !y_all || !dates || !x || !ens || !y || !bad || n_batch < 0 || n_dates < 0 || x_len < 0 || ens_len < 0 ||
      y_len < 0)
    return fail(RC_ERR_ARG, "rc_gather_dates: bad argument");
  if (n_batch == 0) return RC_OK;
  const long long total = x_len + ens_len + y_len;
  long long gx = ceil_div_ll(total > 0 ? total : 1, 256);
  if (gx > 2 * kNumSMs) gx = 2 * kNumSMs;
  const GatherDatesP p{x_all, ens_all, y_all, reinterpret_cast<const long long*>(dates), x_len, ens_len, y_len, n_dates, x, ens, y, bad,
                       nullptr, nullptr, 0};
  launch_pdl(gather_dates_kernel, dim3((int)gx, n_batch), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("gather_dates_kernel");
}

extern "C" int rc_gather_dates_step(const float* x_all, const float* ens_all, const float* y_all, const int64_t* order, int n_batches,
                                    const int64_t* step_count, const int64_t* epoch_base, int n_batch, int n_dates, long long x_len,
                                    long long ens_len, long long y_len, float* x, float* ens, float* y, int32_t* bad, void* stream) {
  if (!x_all || !ens_all || !y_all || !order || !step_count || !epoch_base || !x || !ens || !y || !bad || n_batches < 0 || n_batch < 0 ||
      n_dates < 0 || x_len < 0 || ens_len < 0 || y_len < 0)
    return fail(RC_ERR_ARG, "rc_gather_dates_step: bad argument");
  if (n_batch == 0) return RC_OK;
  const long long total = x_len + ens_len + y_len;
  long long gx = ceil_div_ll(total > 0 ? total : 1, 256);
  if (gx > 2 * kNumSMs) gx = 2 * kNumSMs;
  const GatherDatesP p{x_all, ens_all, y_all, reinterpret_cast<const long long*>(order), x_len, ens_len, y_len, n_dates, x, ens, y, bad,
                       reinterpret_cast<const long long*>(step_count), reinterpret_cast<const long long*>(epoch_base), n_batches};
  launch_pdl(gather_dates_kernel, dim3((int)gx, n_batch), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("gather_dates_kernel");
}

extern "C" int rc_bn_stats_finalize(const float* stats, int row_tiles, int row_tile, int m, int n, float eps,
                                    float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                                    int64_t* num_batches_tracked, void* stream) {
  if (!stats || !mean || !rstd || row_tiles <= 0 || row_tile <= 0 || m <= 0 || n <= 0) return fail(RC_ERR_ARG, "rc_bn_stats_finalize: bad argument");
  if (row_tiles != ceil_div(m, row_tile)) return fail(RC_ERR_ARG, "rc_bn_stats_finalize: row_tiles %d != ceil(%d/%d)", row_tiles, m, row_tile);
  const BnStatsFinP p{stats, row_tiles, row_tile, m, n, eps, momentum, mean, rstd, running_mean, running_var,
                      reinterpret_cast<long long*>(num_batches_tracked)};
  launch_pdl(bn_stats_finalize_kernel, dim3(ceil_div(n, 32)), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("bn_stats_finalize_kernel");
}

extern "C" int rc_bn_eval_prepare(const float* running_mean, const float* running_var, int n, float eps, float* mean,
                                  float* rstd, void* stream) {
  if (!running_mean || !running_var || !mean || !rstd || n <= 0) return fail(RC_ERR_ARG, "rc_bn_eval_prepare: bad argument");
  const BnEvalP p{running_mean, running_var, n, eps, mean, rstd};
  bn_eval_prepare_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("bn_eval_prepare_kernel");
}

extern "C" int rc_bn_bwd_finalize(const float* stats, int row_tiles, int m, int n, int batch_stats, const float* gamma, const float* mean,
                                  const float* rstd, float* d_gamma, float* d_beta, float* c0, float* c1, float* c2,
                                  void* stream) {
  if (!stats || !gamma || !mean || !rstd || !d_gamma || !d_beta || !c0 || !c1 || !c2 || row_tiles <= 0 || m <= 0 || n <= 0)
    return fail(RC_ERR_ARG, "rc_bn_bwd_finalize: bad argument");
  const BnBwdFinP p{stats, row_tiles, m, n, batch_stats, gamma, mean, rstd, d_gamma, d_beta, c0, c1, c2};
  launch_pdl(bn_bwd_finalize_kernel, dim3(ceil_div(n, 32)), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
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
    int gx = ceil_div(max_n, 32);             // a CTA finishes 32 outputs per pass
    if (gx > 2 * kNumSMs) gx = 2 * kNumSMs;
    launch_pdl(reduce_segments_kernel, dim3(gx, cnt), dim3(256), 0, static_cast<cudaStream_t>(stream), args);
    if (int e = check_launch("reduce_segments_kernel")) return e;
  }
  return RC_OK;
}

extern "C" int rc_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t* step, long long n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                             void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || n < 0) return fail(RC_ERR_ARG, "rc_adamw_step: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long blocks = ceil_div_ll(n > 0 ? n : 1, 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  const AdamP pa{param, grad, exp_avg, exp_avg_sq, reinterpret_cast<long long*>(step), n, lr, beta1, beta2, eps, weight_decay, grad_scale};
  launch_pdl(adamw_kernel, dim3((int)blocks), dim3(256), 0, s, pa);     // (n == 0: one CTA that only advances the step counter)
  return check_launch("adamw_kernel");
}
