// DeepSets member MLP (first Linear + ReLU) fused with the SUM pool over ensemble members
// (models/gnn.py:51-56,66-67).  The second phi Linear commutes with the sum, so it is applied once
// per station afterwards (rc_gemm_run with bias_scale = members) instead of once per member: the
// member-level work drops from 2*M*Em*(F*H + H*H) to 2*M*Em*F*H FLOPs and nothing of size
// [M*Em, H] is written to HBM.  fp32 FFMA; the member rows of a station are staged in shared memory.
#include "rc_common.cuh"

namespace rc {

constexpr int kDsThreads = 256;
constexpr int kDsWarps = 8;
constexpr int kDsCols = 128;     // hidden columns per CTA (blockIdx.y selects the chunk)
constexpr int kDsMemberChunk = 16;

// ---------------------------------------------------------------------------------------- forward
// dynamic smem: Ws[F4][128] | bias[128] | Es[8 warps][kDsMemberChunk][F4]
__global__ void __launch_bounds__(kDsThreads)
deepsets_pool_fwd_kernel(const float* __restrict__ ens, const float* __restrict__ w1, const float* __restrict__ b1,
                         float* __restrict__ pooled, int m, int members, int feats, int hidden, int f4) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* bias = Ws + f4 * kDsCols;
  float* Es_all = bias + kDsCols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * kDsCols;
  // W1 chunk, transposed to [k][col]: thread <-> column, walks its own weight row
  for (int j = tid; j < kDsCols; j += kDsThreads) {
    const int col = c0 + j;
    for (int k = 0; k < f4; ++k) Ws[k * kDsCols + j] = (col < hidden && k < feats) ? __ldg(w1 + (size_t)col * feats + k) : 0.f;
    bias[j] = col < hidden ? __ldg(b1 + col) : 0.f;
  }
  __syncthreads();
  float* Es = Es_all + warp * kDsMemberChunk * f4;
  const float4 bv = ld4(bias + 4 * lane);
  const int kq = f4 >> 2;
  for (int node = blockIdx.x * kDsWarps + warp; node < m; node += gridDim.x * kDsWarps) {
    float4 pool = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = ens + (size_t)node * members * feats;
    for (int e0 = 0; e0 < members; e0 += kDsMemberChunk) {
      const int cnt = min(kDsMemberChunk, members - e0);
      __syncwarp();
      // stage cnt member rows (contiguous in HBM) into the padded [cnt][f4] layout
      const int total = cnt * feats;
      const float* s2 = src + (size_t)e0 * feats;
      for (int idx = lane; idx < total; idx += 32) {
        const int e = idx / feats, k = idx - e * feats;
        Es[e * f4 + k] = __ldg(s2 + idx);
      }
      for (int idx = lane; idx < kDsMemberChunk * (f4 - feats); idx += 32) {     // zero the k padding
        const int e = idx / (f4 - feats), k = feats + idx - e * (f4 - feats);
        Es[e * f4 + k] = 0.f;
      }
      for (int idx = lane + cnt * feats; idx < kDsMemberChunk * feats; idx += 32) {   // and the unused rows
        const int e = idx / feats, k = idx - e * feats;
        Es[e * f4 + k] = 0.f;
      }
      __syncwarp();
      for (int e = 0; e < cnt; e += 4) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int q = 0; q < kq; ++q) {
          float4 a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = ld4(Es + (e + i) * f4 + 4 * q);
#pragma unroll
          for (int r = 0; r < 4; ++r) b[r] = ld4(Ws + (4 * q + r) * kDsCols + 4 * lane);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(a[i].x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a[i].x, b[0].y, acc[i][1]);
            acc[i][2] = fmaf(a[i].x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a[i].x, b[0].w, acc[i][3]);
            acc[i][0] = fmaf(a[i].y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a[i].y, b[1].y, acc[i][1]);
            acc[i][2] = fmaf(a[i].y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a[i].y, b[1].w, acc[i][3]);
            acc[i][0] = fmaf(a[i].z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a[i].z, b[2].y, acc[i][1]);
            acc[i][2] = fmaf(a[i].z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a[i].z, b[2].w, acc[i][3]);
            acc[i][0] = fmaf(a[i].w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a[i].w, b[3].y, acc[i][1]);
            acc[i][2] = fmaf(a[i].w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a[i].w, b[3].w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (e + i < cnt) {          // members are pooled in index order
            pool.x += fmaxf(acc[i][0] + bv.x, 0.f);
            pool.y += fmaxf(acc[i][1] + bv.y, 0.f);
            pool.z += fmaxf(acc[i][2] + bv.z, 0.f);
            pool.w += fmaxf(acc[i][3] + bv.w, 0.f);
          }
      }
    }
    const int col = c0 + 4 * lane;
    if (col + 3 < hidden) {
      st4(pooled + (size_t)node * hidden + col, pool);
    } else {
      const float o[4] = {pool.x, pool.y, pool.z, pool.w};
      for (int j = 0; j < 4; ++j)
        if (col + j < hidden) pooled[(size_t)node * hidden + col + j] = o[j];
    }
  }
}

// ---------------------------------------------------------------------------------------- backward
// Tiles of 64 consecutive member rows of the flattened [M*members, F] matrix.
//   phase 1: recompute pre = E W1^T + b1 for the tile (warp owns 8 rows, lane 4 columns), and write
//            dh[r][c] = d_pooled[node(r)][c] * 1[pre > 0] to shared memory;
//   phase 2: thread (c = tid % 128, half = tid / 128) accumulates d w1[c][k] over the tile's rows for
//            its 4*KQ features in registers; the accumulators live across all tiles of the CTA.
// dynamic smem: Ws[FP][128] | bias[128] | Es[64][FP] | dh[64][128],  FP = 8*KQ >= feats
template <int KQ>
__global__ void __launch_bounds__(kDsThreads)
deepsets_pool_bwd_kernel(const float* __restrict__ ens, const float* __restrict__ w1, const float* __restrict__ b1,
                         const float* __restrict__ d_pooled, float* __restrict__ partials, int m, int members,
                         int feats, int hidden) {
  constexpr int FP = 8 * KQ;
  constexpr int ROWS = 64;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;
  float* bias = Ws + FP * kDsCols;
  float* Es = bias + kDsCols;
  float* dh = Es + ROWS * FP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.y * kDsCols;
  for (int j = tid; j < kDsCols; j += kDsThreads) {
    const int col = c0 + j;
    for (int k = 0; k < FP; ++k) Ws[k * kDsCols + j] = (col < hidden && k < feats) ? __ldg(w1 + (size_t)col * feats + k) : 0.f;
    bias[j] = col < hidden ? __ldg(b1 + col) : 0.f;
  }
  const float4 bv_dummy = make_float4(0.f, 0.f, 0.f, 0.f);
  (void)bv_dummy;
  float dwacc[4 * KQ];
#pragma unroll
  for (int i = 0; i < 4 * KQ; ++i) dwacc[i] = 0.f;
  float dbacc = 0.f;
  const long long total_rows = (long long)m * members;
  const int c_own = tid & (kDsCols - 1), half = tid >> 7;
  __syncthreads();
  const float4 bv = ld4(bias + 4 * lane);
  for (long long row0 = (long long)blockIdx.x * ROWS; row0 < total_rows; row0 += (long long)gridDim.x * ROWS) {
    const int nrows = (int)min((long long)ROWS, total_rows - row0);
    // stage the tile (contiguous floats) into [64][FP], zero padded
    const float* src = ens + (size_t)row0 * feats;
    for (int idx = tid; idx < ROWS * FP; idx += kDsThreads) {
      const int r = idx / FP, k = idx - r * FP;
      Es[idx] = (r < nrows && k < feats) ? __ldg(src + (size_t)r * feats + k) : 0.f;
    }
    __syncthreads();
    // phase 1
    {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int q = 0; q < 2 * KQ; ++q) {
        float4 b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) b[r] = ld4(Ws + (4 * q + r) * kDsCols + 4 * lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 a = ld4(Es + (warp * 8 + i) * FP + 4 * q);
          acc[i][0] = fmaf(a.x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a.x, b[0].y, acc[i][1]);
          acc[i][2] = fmaf(a.x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a.x, b[0].w, acc[i][3]);
          acc[i][0] = fmaf(a.y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a.y, b[1].y, acc[i][1]);
          acc[i][2] = fmaf(a.y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a.y, b[1].w, acc[i][3]);
          acc[i][0] = fmaf(a.z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a.z, b[2].y, acc[i][1]);
          acc[i][2] = fmaf(a.z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a.z, b[2].w, acc[i][3]);
          acc[i][0] = fmaf(a.w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a.w, b[3].y, acc[i][1]);
          acc[i][2] = fmaf(a.w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a.w, b[3].w, acc[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 8 + i;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) {
          const long long node = (row0 + r) / members;
          const int col = c0 + 4 * lane;
          float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col + 3 < hidden) {
            d = ldg4(d_pooled + (size_t)node * hidden + col);
          } else {
            if (col < hidden) d.x = __ldg(d_pooled + (size_t)node * hidden + col);
            if (col + 1 < hidden) d.y = __ldg(d_pooled + (size_t)node * hidden + col + 1);
            if (col + 2 < hidden) d.z = __ldg(d_pooled + (size_t)node * hidden + col + 2);
          }
          o.x = (acc[i][0] + bv.x > 0.f) ? d.x : 0.f;
          o.y = (acc[i][1] + bv.y > 0.f) ? d.y : 0.f;
          o.z = (acc[i][2] + bv.z > 0.f) ? d.z : 0.f;
          o.w = (acc[i][3] + bv.w > 0.f) ? d.w : 0.f;
        }
        st4(dh + r * kDsCols + 4 * lane, o);
      }
    }
    __syncthreads();
    // phase 2
    for (int r = 0; r < nrows; ++r) {
      const float d = dh[r * kDsCols + c_own];
      if (half == 0) dbacc += d;
#pragma unroll
      for (int q = 0; q < KQ; ++q) {
        const float4 e = ld4(Es + r * FP + half * 4 * KQ + 4 * q);
        dwacc[4 * q + 0] = fmaf(d, e.x, dwacc[4 * q + 0]);
        dwacc[4 * q + 1] = fmaf(d, e.y, dwacc[4 * q + 1]);
        dwacc[4 * q + 2] = fmaf(d, e.z, dwacc[4 * q + 2]);
        dwacc[4 * q + 3] = fmaf(d, e.w, dwacc[4 * q + 3]);
      }
    }
    __syncthreads();
  }
  // partials[blockIdx.x][hidden*feats + hidden]: every CTA of column chunk blockIdx.y writes its slice
  float* out = partials + (size_t)blockIdx.x * ((size_t)hidden * feats + hidden);
  const int col = c0 + c_own;
  if (col < hidden) {
#pragma unroll
    for (int i = 0; i < 4 * KQ; ++i) {
      const int k = half * 4 * KQ + i;
      if (k < feats) out[(size_t)col * feats + k] = dwacc[i];
    }
    if (half == 0) out[(size_t)hidden * feats + col] = dbacc;
  }
}

static int ds_bwd_blocks(int m) {
  // the member count is not known here; one CTA per 64 member rows is the natural upper bound, and
  // the persistent loop makes any smaller grid correct
  int nb = ceil_div(m > 0 ? m : 1, 4);
  const int cap = 2 * kNumSMs;
  return nb < cap ? nb : cap;
}

}  // namespace rc

using namespace rc;

extern "C" int rc_deepsets_pool_fwd(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                                    int members, int feats, int hidden, void* stream) {
  if (!ens || !w1 || !b1 || !pooled || num_nodes < 0 || members <= 0 || feats <= 0 || hidden <= 0)
    return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd: bad argument");
  if (num_nodes == 0) return RC_OK;
  if (!aligned16(pooled) && hidden % 4 == 0) return fail(RC_ERR_ARG, "rc_deepsets_pool_fwd: pooled must be 16-byte aligned");
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
  deepsets_pool_fwd_kernel<<<grid, kDsThreads, smem, static_cast<cudaStream_t>(stream)>>>(ens, w1, b1, pooled, num_nodes, members,
                                                                                         feats, hidden, f4);
  return check_launch("deepsets_pool_fwd_kernel");
}

extern "C" int rc_deepsets_pool_bwd_nblocks(int num_nodes, int hidden) {
  if (num_nodes < 0 || hidden <= 0) return -1;
  return ds_bwd_blocks(num_nodes);
}

template <int KQ>
static int ds_bwd_launch(const float* ens, const float* w1, const float* b1, const float* d_pooled, float* partials, int m,
                         int members, int feats, int hidden, cudaStream_t s) {
  constexpr int FP = 8 * KQ;
  const size_t smem = ((size_t)FP * kDsCols + kDsCols + (size_t)64 * FP + (size_t)64 * kDsCols) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(deepsets_pool_bwd_kernel<KQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid(ds_bwd_blocks(m), ceil_div(hidden, kDsCols));
  deepsets_pool_bwd_kernel<KQ><<<grid, kDsThreads, smem, s>>>(ens, w1, b1, d_pooled, partials, m, members, feats, hidden);
  return check_launch("deepsets_pool_bwd_kernel");
}

extern "C" int rc_deepsets_pool_bwd(const float* ens, const float* w1, const float* b1, const float* d_pooled,
                                    float* partials, int num_nodes, int members, int feats, int hidden, void* stream) {
  if (!ens || !w1 || !b1 || !d_pooled || !partials || num_nodes < 0 || members <= 0 || feats <= 0 || hidden <= 0)
    return fail(RC_ERR_ARG, "rc_deepsets_pool_bwd: bad argument");
  if (!aligned16(d_pooled)) return fail(RC_ERR_ARG, "rc_deepsets_pool_bwd: d_pooled must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int kq = ceil_div(feats, 8);
  switch (kq) {
    case 1: return ds_bwd_launch<1>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 2: return ds_bwd_launch<2>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 3: return ds_bwd_launch<3>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 4: return ds_bwd_launch<4>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 5: return ds_bwd_launch<5>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 6: return ds_bwd_launch<6>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 7: return ds_bwd_launch<7>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    case 8: return ds_bwd_launch<8>(ens, w1, b1, d_pooled, partials, num_nodes, members, feats, hidden, s);
    default: return fail(RC_ERR_ARG, "rc_deepsets_pool_bwd: feats=%d > 64 is not instantiated", feats);
  }
}
