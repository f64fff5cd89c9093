// Output links + closed-form CRPS value and gradient: one elementwise pass over [M, C] (HBM-bound:
// (2C+1)*4 algorithmic bytes per node).  Replaces ~2 200 ATen calls of the reference
// (models/model_utils.py:89-113, models/loss.py:203-272).

#include "rc_crps_tile.cuh"

namespace rc {

__global__ void __launch_bounds__(kCrpsThreads) crps_count_kernel(const CrpsCountP p) {
  pdl_entry();
  crps_count_tile(p, blockIdx, gridDim);
}

template <int WIDTH>
__global__ void __launch_bounds__(kCrpsThreads) crps_main_kernel(const CrpsMainP p) {
  pdl_entry();
  crps_main_tile<WIDTH>(p, blockIdx, gridDim);
}

// Batches of at most 1024 nodes (every reference-shape batch: 976): count, value + gradient and the mean in ONE CTA -
// the valid count and the loss sum are block reductions instead of two more kernels on the step's critical path.
// Same per-node function, same 1/n_valid scaling of the gradient; the loss sum is taken in float64 in warp order.
struct CrpsSmallP {
  CrpsMainP m;
  double* loss_out;
  int* n_valid;
};

template <int WIDTH>
__global__ void __launch_bounds__(1024) crps_small_kernel(const CrpsSmallP p) {
  pdl_entry();
  const CrpsMainP& a = p.m;
  __shared__ int shc[32];
  __shared__ double shl[32];
  __shared__ int s_cnt;
  const int i = threadIdx.x, lane = i & 31, warp = i >> 5;
  const float yi = i < a.m ? a.y[i] : nanf("");
  int c = __popc(__ballot_sync(0xffffffffu, !isnan(yi)));
  if (lane == 0) shc[warp] = c;
  __syncthreads();
  if (warp == 0) {
    int t = shc[lane];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) s_cnt = t;
  }
  __syncthreads();
  const int cnt = s_cnt;
  const float inv_n = cnt > 0 ? 1.0f / (float)cnt : 0.0f;
  float loss = 0.0f;
  if (i < a.m) {
    float row[WIDTH], g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < WIDTH; ++j) row[j] = a.pred[(size_t)i * WIDTH + j];
    if (!isnan(yi)) loss = crps_node_k<WIDTH - 2>(row, yi, a.raw_input, a.u_fixed, a.xi, a.t, g);
    if (a.d_pred != nullptr) {
#pragma unroll
      for (int j = 0; j < WIDTH; ++j) a.d_pred[(size_t)i * WIDTH + j] = g[j] * inv_n;
    }
  }
  const double dl = warp_sum((double)loss);
  if (lane == 0) shl[warp] = dl;
  __syncthreads();
  if (i == 0) {
    double tot = 0.0;
    for (int w = 0; w < 32; ++w) tot += shl[w];
    p.loss_out[0] = cnt > 0 ? tot / (double)cnt : nan("");   // mean of an empty selection is NaN in torch too
    p.n_valid[0] = cnt;
  }
}

__global__ void __launch_bounds__(256) crps_final_kernel(const CrpsFinalP p) {
  pdl_entry();
  crps_final_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(256) postprocess_fwd_kernel(const PostFwdP p) {
  pdl_entry();
  post_fwd_tile(p, blockIdx, gridDim);
}

__global__ void __launch_bounds__(256) postprocess_bwd_kernel(const PostBwdP p) {
  pdl_entry();
  post_bwd_tile(p, blockIdx, gridDim);
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
  const int blocks = crps_main_blocks(m);
  const CrpsCountP pc{y, m, ws.cnt};
  const CrpsMainP pm{pred, y, d_pred, m, kind, raw_input, u_fixed, xi, t, ws.cnt, ncnt, ws.loss};
  const CrpsFinalP pf{ws.cnt, ncnt, ws.loss, blocks, loss_out, n_valid};
  if (m > 0 && m <= 1024) {
    const CrpsSmallP ps{pm, loss_out, n_valid};
    switch (kind) {
      case 0: launch_pdl(crps_small_kernel<2>, dim3(1), dim3(1024), 0, s, ps); break;
      case 1: launch_pdl(crps_small_kernel<3>, dim3(1), dim3(1024), 0, s, ps); break;
      case 2: launch_pdl(crps_small_kernel<4>, dim3(1), dim3(1024), 0, s, ps); break;
      default: launch_pdl(crps_small_kernel<5>, dim3(1), dim3(1024), 0, s, ps); break;
    }
    return check_launch("crps_small_kernel");
  }
  launch_pdl(crps_count_kernel, dim3(ncnt), dim3(kCrpsThreads), 0, s, pc);
  if (int e = check_launch("crps_count_kernel")) return e;
  switch (kind) {
    case 0: launch_pdl(crps_main_kernel<2>, dim3(blocks), dim3(kCrpsThreads), 0, s, pm); break;
    case 1: launch_pdl(crps_main_kernel<3>, dim3(blocks), dim3(kCrpsThreads), 0, s, pm); break;
    case 2: launch_pdl(crps_main_kernel<4>, dim3(blocks), dim3(kCrpsThreads), 0, s, pm); break;
    default: launch_pdl(crps_main_kernel<5>, dim3(blocks), dim3(kCrpsThreads), 0, s, pm); break;
  }
  if (int e = check_launch("crps_main_kernel")) return e;
  launch_pdl(crps_final_kernel, dim3(1), dim3(256), 0, s, pf);
  return check_launch("crps_final_kernel");
}

extern "C" int rc_postprocess_fwd(const float* raw, float* post, int num_nodes, int kind, void* stream) {
  if (!raw || !post || kind < 0 || kind > 3) return fail(RC_ERR_ARG, "rc_postprocess_fwd: bad argument");
  if (num_nodes == 0) return RC_OK;
  postprocess_fwd_kernel<<<ceil_div(num_nodes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(PostFwdP{raw, post, num_nodes, kind});
  return check_launch("postprocess_fwd_kernel");
}

extern "C" int rc_postprocess_bwd(const float* raw, const float* d_post, float* d_raw, int num_nodes, int kind, void* stream) {
  if (!raw || !d_post || !d_raw || kind < 0 || kind > 3) return fail(RC_ERR_ARG, "rc_postprocess_bwd: bad argument");
  if (num_nodes == 0) return RC_OK;
  postprocess_bwd_kernel<<<ceil_div(num_nodes, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(PostBwdP{raw, d_post, d_raw, num_nodes, kind});
  return check_launch("postprocess_bwd_kernel");
}
