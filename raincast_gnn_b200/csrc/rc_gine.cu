// GINE aggregation: forward CSR gather and backward transpose gather, both atomic-free.
//
// PyG GINEConv.message/aggregate (call site models/gnn.py:27-29,39-44):
//   m_e = relu(x[src_e] + lin(edge_attr_e)),  agg_i = sum_{dst_e = i} m_e,  h_i = agg_i + (1+eps) x_i
// A sub-warp of H/4 lanes (a full warp for H >= 128) owns one destination row; every lane keeps
// float4 column slices in registers, so each gathered source row is one coalesced 128-bit-per-lane
// load and nothing of size [E, H] is ever materialised (the reference materialises four).
// HBM-bound kernel: algorithmic bytes fwd = 2*M*H*4 + E*8 + (M+1)*4 + 8*H  (SURVEY.md 8d).
#include "rc_gine_tile.cuh"

namespace rc {

template <int CH>
__global__ void __launch_bounds__(kGineThreads) gine_aggr_fwd_kernel(const GineFwdP p) {
  pdl_entry();
  gine_fwd_tile<CH>(p, blockIdx, gridDim);
}

template <int CH>
__global__ void __launch_bounds__(kGineThreads) gine_aggr_bwd_kernel(const GineBwdP p) {
  pdl_entry();
  extern __shared__ __align__(16) float smem[];
  gine_bwd_tile<CH>(p, blockIdx, gridDim, smem);
}

__global__ void __launch_bounds__(256) gine_bwd_finalize_kernel(const GineFinP p) {
  pdl_entry();
  gine_fin_tile(p, blockIdx, gridDim);
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
  const bool ranged = sh.lpr == 32 && num_nodes >= kRangedMinRows;
  if (ranged) return launch_gine_fwd_ranged(x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, s);
  const int grid = ceil_div(num_nodes, kGineWarps * sh.rpw);
  const GineFwdP p{x, rowptr, col, attr, w_edge, b_edge, eps, h, num_nodes, hidden, sh.lpr};
  switch (sh.ch) {
    case 1: launch_pdl(gine_aggr_fwd_kernel<1>, dim3(grid), dim3(kGineThreads), 0, s, p); break;
    case 2: launch_pdl(gine_aggr_fwd_kernel<2>, dim3(grid), dim3(kGineThreads), 0, s, p); break;
    case 3: launch_pdl(gine_aggr_fwd_kernel<3>, dim3(grid), dim3(kGineThreads), 0, s, p); break;
    default: launch_pdl(gine_aggr_fwd_kernel<4>, dim3(grid), dim3(kGineThreads), 0, s, p); break;
  }
  return check_launch("gine_aggr_fwd_kernel");
}

static int gine_small_blocks(int num_nodes, const GineShape& sh) {
  const int nb = ceil_div(num_nodes > 0 ? num_nodes : 1, kGineWarps * sh.rpw);
  const int cap = kNumSMs * 8;
  return nb < cap ? nb : cap;
}

extern "C" int rc_gine_aggr_bwd_nblocks(int num_nodes, int hidden) {
  GineShape sh;
  if (!gine_shape(hidden, &sh) || num_nodes < 0) return -1;
  if (sh.lpr == 32 && num_nodes >= kRangedMinRows) return gine_ranged_grid(num_nodes);
  return gine_small_blocks(num_nodes, sh);
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
  if (sh.lpr == 32 && num_nodes >= kRangedMinRows)
    return launch_gine_bwd_ranged(g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, num_nodes, hidden, s);
  const int grid = gine_small_blocks(num_nodes, sh);
  const int rpb = kGineWarps * sh.rpw;
  const size_t smem = ((size_t)rpb * 2 * hidden + rpb) * sizeof(float);
  const GineBwdP p{g, x, t_rowptr, t_dst, t_attr, w_edge, b_edge, eps, addend, dx, partials, num_nodes, hidden, sh.lpr};
#define RC_LAUNCH_BWD(CHV)                                                                                                  \
  do {                                                                                                                      \
    if (smem > 48 * 1024) cudaFuncSetAttribute(gine_aggr_bwd_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    launch_pdl(gine_aggr_bwd_kernel<CHV>, dim3(grid), dim3(kGineThreads), smem, s, p);                                                         \
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
  const GineFinP p{partials, nblocks, hidden, d_w, d_b, d_eps};
  launch_pdl(gine_bwd_finalize_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("gine_bwd_finalize_kernel");
}
