// DeepSets member MLP (first Linear + ReLU) fused with the SUM pool over ensemble members
// (models/gnn.py:51-56,66-67).  The second phi Linear commutes with the sum, so it is applied once
// per station afterwards (rc_gemm_run with bias_scale = members) instead of once per member: the
// member-level work drops from 2*M*Em*(F*H + H*H) to 2*M*Em*F*H FLOPs and nothing of size
// [M*Em, H] is written to HBM.  fp32 FFMA; the member rows of a station are staged in shared memory.

#include "rc_deepsets_tile.cuh"

namespace rc {

__global__ void __launch_bounds__(kDsThreads) deepsets_pool_fwd_kernel(const DsFwdP p) {
  pdl_entry();
  extern __shared__ __align__(16) float smem[];
  ds_fwd_tile(p, blockIdx, gridDim, smem);
}

template <int KQ>
__global__ void __launch_bounds__(kDsThreads) deepsets_pool_bwd_kernel(const DsBwdP p) {
  pdl_entry();
  extern __shared__ __align__(16) float smem[];
  ds_bwd_tile<KQ>(p, blockIdx, gridDim, smem);
}

}  // namespace rc

using namespace rc;

extern "C" int rc_deepsets_pool_fwd(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                                    int members, int feats, int hidden, void* stream) {
  if (!ens || !w1 || !b1 || !pooled || num_nodes < 0 || members <= 0 || feats <= 0 || hidden <= 0)
    return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd: bad argument");
  if (num_nodes == 0) return RC_OK;
  if (!aligned16(pooled) && hidden % 4 == 0) return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd: pooled must be 16-byte aligned");
  if (deepsets_tc_applicable(num_nodes, members, feats, hidden))      // large shapes: tensor cores, fp32-accurate 3xTF32
    return launch_deepsets_fwd_tc(false, ens, w1, b1, pooled, num_nodes, members, feats, hidden, static_cast<cudaStream_t>(stream));
  const int f4 = (feats + 3) & ~3;
  const size_t smem = ((size_t)f4 * kDsCols + kDsCols + (size_t)kDsWarps * kDsMemberChunk * f4) * sizeof(float);
  if (smem > 200 * 1024) return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd: feats=%d too large for the shared-memory tile", feats);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(deepsets_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_smem = smem;
  }
  int gx = ceil_div(num_nodes, kDsWarps);
  if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
  dim3 grid(gx, ceil_div(hidden, kDsCols));
  const DsFwdP p{ens, w1, b1, pooled, num_nodes, members, feats, hidden, f4};
  launch_pdl(deepsets_pool_fwd_kernel, grid, dim3(kDsThreads), smem, static_cast<cudaStream_t>(stream), p);
  return check_launch("deepsets_pool_fwd_kernel");
}

extern "C" int rc_deepsets_pool_fwd_bf16(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                                         int members, int feats, int hidden, void* stream) {
  if (!ens || !w1 || !b1 || !pooled || num_nodes < 0 || members <= 0 || feats <= 0 || hidden <= 0)
    return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd_bf16: bad argument");
  if (members > 128 || feats > 64) return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd_bf16: needs members <= 128 and feats <= 64");
  if (num_nodes == 0) return RC_OK;
  return launch_deepsets_fwd_tc(true, ens, w1, b1, pooled, num_nodes, members, feats, hidden, static_cast<cudaStream_t>(stream));
}

extern "C" int rc_deepsets_pool_bwd_nblocks(int num_nodes, int members, int feats, int hidden) {
  if (num_nodes < 0 || hidden <= 0 || members <= 0 || feats <= 0) return -1;
  if (deepsets_bwd_tc_applicable(num_nodes, members, feats, hidden)) return deepsets_bwd_tc_blocks(num_nodes, members);
  return ds_bwd_blocks(num_nodes);
}

template <int KQ>
static int ds_bwd_launch(const float* ens, const float* w1, const float* b1, const float* d_pooled, float* partials, int m,
                         int members, int feats, int hidden, int bf16_operands, uint32_t* mask_out, cudaStream_t s) {
  constexpr int FP = 8 * KQ;
  const size_t smem = ((size_t)FP * kDsCols + kDsCols + (size_t)64 * FP + (size_t)64 * kDsCols) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(deepsets_pool_bwd_kernel<KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid(ds_bwd_blocks(m), ceil_div(hidden, kDsCols));
  const DsBwdP p{ens, w1, b1, d_pooled, partials, m, members, feats, hidden, bf16_operands, mask_out};
  launch_pdl(deepsets_pool_bwd_kernel<KQ>, grid, dim3(kDsThreads), smem, s, p);
  return check_launch("deepsets_pool_bwd_kernel");
}

extern "C" int rc_deepsets_pool_bwd(const float* ens, const float* w1, const float* b1, const float* d_pooled,
                                    float* partials, int num_nodes, int members, int feats, int hidden, int bf16_operands,
                                    uint32_t* mask_bits_out, void* stream) {
  if (!ens || !w1 || !b1 || !d_pooled || !partials || num_nodes < 0 || members <= 0 || feats <= 0 || hidden <= 0)
    return fail(RC_ERR_ARG, "rc_deepsets_pool_bwd: bad argument");
  if (!aligned16(d_pooled)) return fail(RC_ERR_ARG, "rc_deepsets_pool_bwd: d_pooled must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (num_nodes > 0 && deepsets_bwd_tc_applicable(num_nodes, members, feats, hidden))
    return launch_deepsets_bwd_tc(ens, w1, b1, d_pooled, partials, mask_bits_out, num_nodes, members, feats, hidden, bf16_operands, s);

  const int kq = ceil_div(feats, 8);
  switch (kq) {
    case 1: return ds_bwd_launch<1>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 2: return ds_bwd_launch<2>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 3: return ds_bwd_launch<3>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 4: return ds_bwd_launch<4>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 5: return ds_bwd_launch<5>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 6: return ds_bwd_launch<6>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 7: return ds_bwd_launch<7>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    case 8: return ds_bwd_launch<8>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, bf16_operands, mask_bits_out, s);
    default: return fail(RC_ERR_ARG, "rc_deepsets_pool_bwd: feats=%d > 64 is not instantiated", feats);
  }
}
