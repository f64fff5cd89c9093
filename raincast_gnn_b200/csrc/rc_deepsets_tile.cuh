// DeepSets pool tile functions shared by rc_deepsets.cu and the step program (rc_prog.cu).
#pragma once
#include <cuda_bf16.h>

#include "rc_common.cuh"

namespace rc {

struct DsFwdP {
  const float* ens;
  const float* w1;
  const float* b1;
  float* pooled;
  int m;
  int members;
  int feats;
  int hidden;
  int f4;
};

struct DsBwdP {
  const float* ens;
  const float* w1;
  const float* b1;
  const float* d_pooled;
  float* partials;
  int m;
  int members;
  int feats;
  int hidden;
  int bf16_operands;   // forward ran with bf16 operands: recompute the ReLU mask (and use the inputs) as the tensor cores saw them
  uint32_t* mask_out;  // test instrumentation (NULL in production): the ReLU mask used, bits [m*members][ceil(hidden/32)]
};

__device__ __forceinline__ float round_operand(float v, int bf16) {
  return bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}



constexpr int kDsThreads = 256;
constexpr int kDsWarps = 8;
constexpr int kDsCols = 128;     // hidden columns per CTA (bid.y selects the chunk)
constexpr int kDsMemberChunk = 16;

// ---------------------------------------------------------------------------------------- forward
// dynamic smem: Ws[F4][128] | bias[128] | Es[8 warps][kDsMemberChunk][F4]
// idx / d for 0 <= idx < 2^20, 0 < d < 2^10 through one float multiply (inv = 1.0f / d): (idx + 0.5) / d is at least
// 0.5 / d away from an integer, far more than the rounding error of the product.
__device__ __forceinline__ int fast_div(int idx, float inv) { return __float2int_rd(((float)idx + 0.5f) * inv); }

// W1 chunk [ncols][feats] (contiguous in HBM) -> shared [k][col], zero padded to fp x kDsCols.  Coalesced, eight loads
// in flight per thread before the first store: a thread walking its own weight row k by k paid one L2 round trip per
// element (this staging was 10 us of the 28 us pool backward at the reference shape).
__device__ __forceinline__ void stage_w1(float* Ws, const float* __restrict__ w1, int c0, int hidden, int feats, int fp, int bf16) {
  const int ncols = max(0, min(kDsCols, hidden - c0));
  const int total = ncols * feats;
  const float inv_feats = 1.0f / (float)feats;
  const float* __restrict__ src = w1 + (size_t)c0 * feats;
  for (int i0 = threadIdx.x; i0 < total; i0 += 8 * kDsThreads) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = i0 + u * kDsThreads;
      v[u] = idx < total ? __ldg(src + idx) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = i0 + u * kDsThreads;
      if (idx < total) {
        const int col = fast_div(idx, inv_feats), k = idx - col * feats;
        Ws[k * kDsCols + col] = round_operand(v[u], bf16);
      }
    }
  }
  for (int idx = threadIdx.x; idx < (fp - feats) * kDsCols; idx += kDsThreads) Ws[feats * kDsCols + idx] = 0.f;      // k padding
  const int pad = kDsCols - ncols;
  for (int idx = threadIdx.x; idx < feats * pad; idx += kDsThreads) {                                                // column padding
    const int k = idx / pad;
    Ws[k * kDsCols + ncols + (idx - k * pad)] = 0.f;
  }
}

__device__ __forceinline__ void ds_fwd_tile(const DsFwdP& p, const uint3 bid, const uint3 gdim, float* smem) {
  const float* __restrict__ ens = p.ens;
  const float* __restrict__ w1 = p.w1;
  const float* __restrict__ b1 = p.b1;
  float* __restrict__ pooled = p.pooled;
  int m = p.m;
  int members = p.members;
  int feats = p.feats;
  int hidden = p.hidden;
  int f4 = p.f4;
  (void)bid; (void)gdim;

  float* Ws = smem;
  float* bias = Ws + f4 * kDsCols;
  float* Es_all = bias + kDsCols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = bid.y * kDsCols;
  // W1 chunk, transposed to [k][col]
  stage_w1(Ws, w1, c0, hidden, feats, f4, 0);
  for (int j = tid; j < kDsCols; j += kDsThreads) bias[j] = c0 + j < hidden ? __ldg(b1 + c0 + j) : 0.f;
  __syncthreads();
  float* Es = Es_all + warp * kDsMemberChunk * f4;
  const float inv_feats = 1.0f / (float)feats;
  const float4 bv = ld4(bias + 4 * lane);
  const int kq = f4 >> 2;
  for (int node = bid.x * kDsWarps + warp; node < m; node += gdim.x * kDsWarps) {
    float4 pool = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = ens + (size_t)node * members * feats;
    for (int e0 = 0; e0 < members; e0 += kDsMemberChunk) {
      const int cnt = min(kDsMemberChunk, members - e0);
      __syncwarp();
      // stage cnt member rows (contiguous in HBM) into the padded [cnt][f4] layout
      const int total = cnt * feats;
      const float* s2 = src + (size_t)e0 * feats;
      for (int i0 = lane; i0 < total; i0 += 4 * 32) {          // four loads in flight before the first store
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = i0 + 32 * u < total ? __ldg(s2 + i0 + 32 * u) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = i0 + 32 * u;
          if (idx < total) {
            const int e = fast_div(idx, inv_feats), k = idx - e * feats;
            Es[e * f4 + k] = v[u];
          }
        }
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
//   phase 2: thread (cp = tid % 64, fh = tid / 64 % 2, rh = tid / 128) accumulates d w1[c][k] for its two columns
//            2cp, 2cp+1 and its 4*KQ features over its half of the tile's rows (32 * rh ...) in registers; the accumulators
//            live across all tiles of the CTA and the two row halves are added through shared memory at the end.
//            (At scale the kernel is bound by the shared-memory pipe - ncu: l1tex 73 %, FMA pipe 46 % - and two columns per
//            thread read every broadcast feature value once per two columns: 7 wavefronts per 40 FFMA instead of 12.)
// dynamic smem: Ws[FP][128] | bias[128] | Es[64][FP] | dh[64][128],  FP = 8*KQ >= feats
template <int KQ>
__device__ __forceinline__ void ds_bwd_tile(const DsBwdP& p, const uint3 bid, const uint3 gdim, float* smem) {
  const float* __restrict__ ens = p.ens;
  const float* __restrict__ w1 = p.w1;
  const float* __restrict__ b1 = p.b1;
  const float* __restrict__ d_pooled = p.d_pooled;
  float* __restrict__ partials = p.partials;
  int m = p.m;
  int members = p.members;
  int feats = p.feats;
  int hidden = p.hidden;
  const int bf16_operands = p.bf16_operands;
  (void)bid; (void)gdim;

  constexpr int FP = 8 * KQ;
  constexpr int ROWS = 64;
  float* Ws = smem;
  float* bias = Ws + FP * kDsCols;
  float* Es = bias + kDsCols;
  float* dh = Es + ROWS * FP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = bid.y * kDsCols;
  stage_w1(Ws, w1, c0, hidden, feats, FP, bf16_operands);
  for (int j = tid; j < kDsCols; j += kDsThreads) bias[j] = c0 + j < hidden ? __ldg(b1 + c0 + j) : 0.f;
  const float4 bv_dummy = make_float4(0.f, 0.f, 0.f, 0.f);
  (void)bv_dummy;
  float dwacc[2][4 * KQ];
#pragma unroll
  for (int i = 0; i < 4 * KQ; ++i) dwacc[0][i] = dwacc[1][i] = 0.f;
  float dbacc[2] = {0.f, 0.f};
  const long long total_rows = (long long)m * members;
  const int cp = tid & 63, fh = (tid >> 6) & 1, rh = tid >> 7;
  __syncthreads();
  const float4 bv = ld4(bias + 4 * lane);
  for (long long row0 = (long long)bid.x * ROWS; row0 < total_rows; row0 += (long long)gdim.x * ROWS) {
    const int nrows = (int)min((long long)ROWS, total_rows - row0);
    // stage the tile (contiguous floats) into [64][FP], zero padded
    const float* src = ens + (size_t)row0 * feats;
    {
      constexpr int kIt = ROWS * FP / kDsThreads;            // all loads of the tile in flight before the first store
      float v[kIt];
#pragma unroll
      for (int u = 0; u < kIt; ++u) {
        const int idx = tid + u * kDsThreads;
        const int r = idx / FP, k = idx - r * FP;
        v[u] = (r < nrows && k < feats) ? __ldg(src + (size_t)r * feats + k) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kIt; ++u) Es[tid + u * kDsThreads] = round_operand(v[u], bf16_operands);
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
          if (p.mask_out != nullptr) {
            // lane l holds columns 4l..4l+3 of the chunk: its nibble goes to bits 4*(l%8).. of word l/8
            unsigned v = ((acc[i][0] + bv.x > 0.f) ? 1u : 0u) | ((acc[i][1] + bv.y > 0.f) ? 2u : 0u) |
                         ((acc[i][2] + bv.z > 0.f) ? 4u : 0u) | ((acc[i][3] + bv.w > 0.f) ? 8u : 0u);
            v <<= 4 * (lane & 7);
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            v |= __shfl_xor_sync(0xffffffffu, v, 4);
            if ((lane & 7) == 0 && col < hidden) p.mask_out[(size_t)(row0 + r) * ((hidden + 31) / 32) + (col >> 5)] = v;
          }
        }
        st4(dh + r * kDsCols + 4 * lane, o);
      }
    }
    __syncthreads();
    // phase 2
    {
      const int r_end = min(nrows, 32 * rh + 32);
#pragma unroll 4
      for (int r = 32 * rh; r < r_end; ++r) {
        const float2 d = *reinterpret_cast<const float2*>(dh + r * kDsCols + 2 * cp);
        if (fh == 0) { dbacc[0] += d.x; dbacc[1] += d.y; }
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
          const float4 e = ld4(Es + r * FP + fh * 4 * KQ + 4 * q);
          dwacc[0][4 * q + 0] = fmaf(d.x, e.x, dwacc[0][4 * q + 0]); dwacc[1][4 * q + 0] = fmaf(d.y, e.x, dwacc[1][4 * q + 0]);
          dwacc[0][4 * q + 1] = fmaf(d.x, e.y, dwacc[0][4 * q + 1]); dwacc[1][4 * q + 1] = fmaf(d.y, e.y, dwacc[1][4 * q + 1]);
          dwacc[0][4 * q + 2] = fmaf(d.x, e.z, dwacc[0][4 * q + 2]); dwacc[1][4 * q + 2] = fmaf(d.y, e.z, dwacc[1][4 * q + 2]);
          dwacc[0][4 * q + 3] = fmaf(d.x, e.w, dwacc[0][4 * q + 3]); dwacc[1][4 * q + 3] = fmaf(d.y, e.w, dwacc[1][4 * q + 3]);
        }
      }
    }
    __syncthreads();
  }
  // the upper row half hands its sums to the lower one through shared memory (Es | dh are free and contiguous:
  // [8*KQ + 2][128] floats <= [64][8*KQ + 128])
  {
    float* comb = Es + (tid & 127);
    if (rh == 1) {
#pragma unroll
      for (int i = 0; i < 4 * KQ; ++i) { comb[(2 * i) * 128] = dwacc[0][i]; comb[(2 * i + 1) * 128] = dwacc[1][i]; }
      comb[(8 * KQ) * 128] = dbacc[0];
      comb[(8 * KQ + 1) * 128] = dbacc[1];
    }
    __syncthreads();
    if (rh == 1) return;
#pragma unroll
    for (int i = 0; i < 4 * KQ; ++i) { dwacc[0][i] += comb[(2 * i) * 128]; dwacc[1][i] += comb[(2 * i + 1) * 128]; }
    dbacc[0] += comb[(8 * KQ) * 128];
    dbacc[1] += comb[(8 * KQ + 1) * 128];
  }
  // partials[bid.x][hidden*feats + hidden]: every CTA of column chunk bid.y writes its slice
  float* out = partials + (size_t)bid.x * ((size_t)hidden * feats + hidden);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int col = c0 + 2 * cp + j;
    if (col < hidden) {
#pragma unroll
      for (int i = 0; i < 4 * KQ; ++i) {
        const int k = fh * 4 * KQ + i;
        if (k < feats) out[(size_t)col * feats + k] = dwacc[j][i];
      }
      if (fh == 0) out[(size_t)hidden * feats + col] = dbacc[j];
    }
  }
}

inline int ds_bwd_blocks(int m) {
  // the member count is not known here; one CTA per 64 member rows is the natural upper bound, and
  // the persistent loop makes any smaller grid correct
  int nb = ceil_div(m > 0 ? m : 1, 4);
  const int cap = 2 * kNumSMs;
  return nb < cap ? nb : cap;
}


}  // namespace rc
