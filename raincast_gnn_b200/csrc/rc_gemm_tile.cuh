// GEMM tile function shared by the standalone kernel (rc_gemm.cu) and the step program (rc_prog.cu).
#pragma once
#include "rc_common.cuh"
#include "rc_gine_tile.cuh"

namespace rc {

constexpr int kGemmThreads = 256;
constexpr int kBN = 128;   // output columns per CTA
// reduction slice RK: 32 (throughput regime) or 128 (small, latency-bound problems: one memory round trip
// covers K = 128 instead of four)

struct GemmP {
  rc_gemm g;
  int a_vec, b_vec, a2_vec, b2_vec, d_vec;   // 128-bit access allowed (ld % 4 == 0 and 16-byte aligned base)
  int tiles1, tiles2;                        // reduction slices of segment 1 / 2
  int b_early;                               // first B tile fetched before griddepcontrol.wait (rc_gemm.b_static, plain B)
};

__device__ __forceinline__ float apply_op(const rc_operand& o, float v, int row, int col) {
  switch (o.op) {
    case RC_OP_BN_RELU:
      return fmaxf((v - __ldg(o.p0 + col)) * __ldg(o.p1 + col) * __ldg(o.p2 + col) + __ldg(o.p3 + col), 0.f);
    case RC_OP_BITMASK:
      return ((__ldg(o.bits + (size_t)row * o.ld_bits + (col >> 5)) >> (col & 31)) & 1u) ? v : 0.f;
    case RC_OP_AFFINE2:
      return fmaf(__ldg(o.p0 + col), v, fmaf(__ldg(o.p1 + col), __ldg(o.aux + (size_t)row * o.ld_aux + col) - __ldg(o.p3 + col), __ldg(o.p2 + col)));
    default:      // (RC_OP_GINE_AGGR never comes here: gemm_tile builds that operand with gine_aggr4)
      return v;
  }
}

// four consecutive stored columns of one stored row, zero outside [0,nrows) x [0,ncols)
__device__ __forceinline__ float4 load_op4(const rc_operand& o, const float* base, int ld, int row, int col, int nrows,
                                           int ncols, bool vec, bool use_op) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows || col >= ncols) return v;
  const float* p = base + (size_t)row * ld + col;
  if (vec && col + 3 < ncols) {
    v = ldg4(p);
  } else {
    v.x = __ldg(p);
    if (col + 1 < ncols) v.y = __ldg(p + 1);
    if (col + 2 < ncols) v.z = __ldg(p + 2);
    if (col + 3 < ncols) v.w = __ldg(p + 3);
  }
  if (use_op && o.op != RC_OP_NONE) {
    v.x = apply_op(o, v.x, row, col);
    if (col + 1 < ncols) v.y = apply_op(o, v.y, row, col + 1);
    if (col + 2 < ncols) v.z = apply_op(o, v.z, row, col + 2);
    if (col + 3 < ncols) v.w = apply_op(o, v.w, row, col + 3);
  }
  return v;
}

// RC_OP_GINE_AGGR: four consecutive columns of the aggregated row (PyG GINEConv message + 'add' aggregation + self term),
// same arithmetic and slot order as gine_fwd_tile (rc_gine_tile.cuh).  The lanes that share a row walk its edge list
// together: with 128-long reduction slices that is a whole warp gathering one 512-byte source row per edge.
__device__ __forceinline__ float4 gine_aggr4(const rc_operand& o, int row, int col, int nrows, int ncols) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows || col >= ncols) return acc;
  const float* __restrict__ x = o.ptr;
  const float4 w4 = ldg4(o.p0 + col), b4 = ldg4(o.p1 + col);
  const int beg = __ldg(o.idx0 + row), end = __ldg(o.idx0 + row + 1);
  int s = beg;
  for (; s + kEdgeUnroll <= end; s += kEdgeUnroll) {     // 4 independent row gathers in flight
    int src[kEdgeUnroll];
    float a[kEdgeUnroll];
    float4 v[kEdgeUnroll];
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) { src[k] = __ldg(o.idx1 + s + k); a[k] = __ldg(o.aux + s + k); }
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) v[k] = ldg4(x + (size_t)src[k] * o.ld + col);
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) add4(acc, relu_msg(v[k], a[k], w4, b4));
  }
  for (; s < end; ++s) {
    const int src = __ldg(o.idx1 + s);
    const float a = __ldg(o.aux + s);
    add4(acc, relu_msg(ldg4(x + (size_t)src * o.ld + col), a, w4, b4));
  }
  const float self_scale = 1.0f + __ldg(o.p2);
  const float4 xi = ldg4(x + (size_t)row * o.ld + col);
  return make_float4(acc.x + self_scale * xi.x, acc.y + self_scale * xi.y, acc.z + self_scale * xi.z, acc.w + self_scale * xi.w);
}

// One CTA tile of the GEMM: `bid` plays blockIdx (x: row tile, y: column tile, z: reduction split), `smem` is the
// dynamic shared memory base.  Called by gemm_kernel (one launch per GEMM) and by the step program (rc_prog.cu).
template <int RM, int AL, int BL, int kRK>
__device__ __forceinline__ void gemm_tile(const GemmP& p, const uint3 bid, float* smem) {
  constexpr int kPadK = kRK + 4;
  constexpr int BM = 8 * RM;
  constexpr int SA = (AL == RC_A_ROW) ? kPadK : (BM + 4);            // A tile row stride (floats)
  constexpr int A_ROWS = (AL == RC_A_ROW) ? BM : kRK;
  constexpr int A_F4_PER_ROW = (AL == RC_A_ROW) ? (kRK / 4) : (BM / 4);
  constexpr int A_SLOTS = A_ROWS * A_F4_PER_ROW;                      // = 8*BM
  constexpr int A_IT = (A_SLOTS + kGemmThreads - 1) / kGemmThreads;
  constexpr int SB = (BL == RC_B_COL) ? kPadK : kBN;
  constexpr int B_ROWS = (BL == RC_B_COL) ? kBN : kRK;
  constexpr int B_F4_PER_ROW = (BL == RC_B_COL) ? (kRK / 4) : (kBN / 4);
  constexpr int B_IT = (B_ROWS * B_F4_PER_ROW) / kGemmThreads;        // = 4
  constexpr int A_STAGE = A_ROWS * SA;
  constexpr int B_STAGE = B_ROWS * SB;

  float* As = smem;                       // 2 stages
  float* Bs = smem + 2 * A_STAGE;         // 2 stages

  const rc_gemm& g = p.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = bid.x * BM, n0 = bid.y * kBN;
  const int z = bid.z;

  // reduction slices handled by this CTA
  const int tiles = p.tiles1 + p.tiles2;
  int t_beg = 0, t_end = tiles;
  if (g.splits > 1) {
    const int per = ceil_div(tiles, g.splits);
    t_beg = z * per;
    t_end = min(tiles, t_beg + per);
  }

  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[A_IT], rb[B_IT];
  float4 csum[A_IT];
#pragma unroll
  for (int i = 0; i < A_IT; ++i) csum[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_a = [&](int t) {
    const bool seg2 = t >= p.tiles1;
    const int k0 = (seg2 ? t - p.tiles1 : t) * kRK;
    const int kk = seg2 ? g.k2 : g.k;
    const float* abase = seg2 ? g.a2 : g.a.ptr;
    const int lda = seg2 ? g.lda2 : g.a.ld;
    const bool avec = seg2 ? p.a2_vec : p.a_vec;
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int s = tid + it * kGemmThreads;
      if (A_SLOTS >= kGemmThreads || s < A_SLOTS) {
        const int row = s / A_F4_PER_ROW, c4 = (s % A_F4_PER_ROW) * 4;
        if (AL == RC_A_ROW && g.a.op == RC_OP_GINE_AGGR && !seg2) {
          ra[it] = gine_aggr4(g.a, m0 + row, k0 + c4, g.m, kk);
          if (g.a_out != nullptr && bid.y == 0 && m0 + row < g.m && k0 + c4 < kk) st4(g.a_out + (size_t)(m0 + row) * g.ld_a_out + k0 + c4, ra[it]);
        } else if (AL == RC_A_ROW) ra[it] = load_op4(g.a, abase, lda, m0 + row, k0 + c4, g.m, kk, avec, !seg2);
        else                ra[it] = load_op4(g.a, abase, lda, k0 + row, m0 + c4, kk, g.m, avec, !seg2);
      }
    }
  };
  auto load_b = [&](int t) {
    const bool seg2 = t >= p.tiles1;
    const int k0 = (seg2 ? t - p.tiles1 : t) * kRK;
    const int kk = seg2 ? g.k2 : g.k;
    const float* bbase = seg2 ? g.b2 : g.b.ptr;
    const int ldb = seg2 ? g.ldb2 : g.b.ld;
    const bool bvec = seg2 ? p.b2_vec : p.b_vec;
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
      const int s = tid + it * kGemmThreads;
      const int row = s / B_F4_PER_ROW, c4 = (s % B_F4_PER_ROW) * 4;
      // (a B prologue only exists for weight-gradient GEMMs, A stored [r][i]: compiled out of the activation GEMMs)
      if (BL == RC_B_COL) rb[it] = load_op4(g.b, bbase, ldb, n0 + row, k0 + c4, g.n, kk, bvec, AL == RC_A_RED && !seg2);
      else                rb[it] = load_op4(g.b, bbase, ldb, k0 + row, n0 + c4, kk, g.n, bvec, AL == RC_A_RED && !seg2);
    }
  };
  auto load_tile = [&](int t) { load_a(t); load_b(t); };
  auto store_tile = [&](int stage) {
    float* as = As + stage * A_STAGE;
    float* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int s = tid + it * kGemmThreads;
      if (A_SLOTS >= kGemmThreads || s < A_SLOTS) {
        const int row = s / A_F4_PER_ROW, c4 = (s % A_F4_PER_ROW) * 4;
        st4(as + row * SA + c4, ra[it]);
        if (AL == RC_A_RED) { csum[it].x += ra[it].x; csum[it].y += ra[it].y; csum[it].z += ra[it].z; csum[it].w += ra[it].w; }
      }
    }
#pragma unroll
    for (int it = 0; it < B_IT; ++it) {
      const int s = tid + it * kGemmThreads;
      const int row = s / B_F4_PER_ROW, c4 = (s % B_F4_PER_ROW) * 4;
      st4(bs + row * SB + c4, rb[it]);
    }
  };

  // Programmatic dependent launch (rc_common.cuh): this CTA may be resident while the kernel before it still runs.
  // A parameter operand (b_static) is fetched before waiting for that kernel - a 64 KB weight tile per CTA at the
  // reference shape, the longest memory round trip of these few-microsecond kernels.
  // (same promise, same place: the residual / auxiliary operand of the epilogue - a saved forward activation or the
  // layer input - is fetched up front instead of behind the reduction, where its L2 round trip cannot overlap anything)
  float e_pre[4] = {0.f, 0.f, 0.f, 0.f};
  const bool e_early = RM == 1 && p.b_early &&
                       (g.epi == RC_EPI_RELU_RES || g.epi == RC_EPI_ADD_RES || g.epi == RC_EPI_MASK_POS || g.epi == RC_EPI_BN_RELU_BWD);
  if (e_early) {
    const int row = m0 + warp;
    const bool use_res = g.epi == RC_EPI_RELU_RES || g.epi == RC_EPI_ADD_RES;
    const float* src = use_res ? g.res : g.e_aux;
    const int ld = use_res ? g.ld_res : g.ld_e_aux;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + ((BL == RC_B_COL) ? (lane + 32 * j) : (4 * lane + j));
      if (row < g.m && col < g.n) e_pre[j] = __ldg(src + (size_t)row * ld + col);
    }
  }
  constexpr bool kWarpSplit = (RM == 1 && kRK == 128 && AL == RC_A_ROW);
  if constexpr (kWarpSplit) {
    if (p.b_early && t_beg < t_end) {
      load_b(t_beg);
      pdl_entry();
      load_a(t_beg);
      store_tile(0);
    } else {
      pdl_entry();
      if (t_beg < t_end) {
        load_tile(t_beg);
        store_tile(0);
      }
    }
    __syncthreads();
  }
  // Warp-split reduction for the latency-bound shape (8-row tile, 128-long slices - every Linear of the reference
  // step at B = 8 graphs): warp w takes reduction indices [16 w, 16 w + 16) of every slice for ALL eight rows instead
  // of the whole reduction for one row, so that the CTA reads the 64 KB weight tile from shared memory
  // once instead of eight times (the shared-memory pipe, not FFMA, bounded this shape); the eight partial
  // [8 x 128] tiles meet in shared memory and warp w sums row w in a fixed order.
  if constexpr (kWarpSplit) {
    float acc8[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc8[i][j] = 0.f;
    int cur = 0;
    for (int t = t_beg; t < t_end; ++t) {
      const bool more = t + 1 < t_end;
      if (more) load_tile(t + 1);
      const float* as = As + cur * A_STAGE;
      const float* bs = Bs + cur * B_STAGE;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r4 = warp * 4 + q;
        float a[8][4], b[4][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = ld4(as + i * SA + r4 * 4);
          a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
        }
        if (BL == RC_B_COL) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = ld4(bs + (lane + 32 * j) * SB + r4 * 4);
            b[0][j] = v.x; b[1][j] = v.y; b[2][j] = v.z; b[3][j] = v.w;
          }
        } else {
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const float4 v = ld4(bs + (r4 * 4 + rr) * SB + 4 * lane);
            b[rr][0] = v.x; b[rr][1] = v.y; b[rr][2] = v.z; b[rr][3] = v.w;
          }
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc8[i][j] = fmaf(a[i][rr], b[rr][j], acc8[i][j]);
      }
      if (more) store_tile(cur ^ 1);
      __syncthreads();                // (last slice: every warp is done with the operand tiles, they become the partial buffer)
      cur ^= 1;
    }
    float* part = smem;               // [8 warps][8 rows][kBN]
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* pr = part + (warp * 8 + i) * kBN;
      if (BL == RC_B_COL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[lane + 32 * j] = acc8[i][j];
      } else {
        st4(pr + 4 * lane, make_float4(acc8[i][0], acc8[i][1], acc8[i][2], acc8[i][3]));
      }
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float* pr = part + (w * 8 + warp) * kBN;
      if (BL == RC_B_COL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][j] += pr[lane + 32 * j];
      } else {
        const float4 v = ld4(pr + 4 * lane);
        acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
      }
    }
    __syncthreads();                  // the epilogue reuses the front of shared memory
  }
  int cur = 0;
  // (the warp-split instantiation never runs this loop: it is kept out of its code image.)  ONE site for the operand
  // loads: the pass t = t_beg - 1 only fetches - with the wait for the previous kernel behind an early B fetch - so that
  // the loaders (sixteen-fold unrolled, with their prologues) exist once in the kernel instead of twice: these kernels
  // start with a cold instruction cache and their time follows their code size.
  if constexpr (!kWarpSplit)
#pragma unroll 1
  for (int t = t_beg - 1; t < t_end; ++t) {
    const bool more = t + 1 < t_end;
    const bool first = t == t_beg - 1;
    const bool early = p.b_early && more;
    if (first && !early) pdl_entry();
    if (more) load_b(t + 1);
    if (first && early) pdl_entry();
    if (more) load_a(t + 1);
    const float* as = As + cur * A_STAGE;
    const float* bs = Bs + cur * B_STAGE;
    // 8 reduction quads per unrolled body: the long-slice variant (32 quads) stays small enough for the instruction
    // cache, which is cold at every launch of a step made of ~75 different small kernels
    if (!first)
#pragma unroll 8
    for (int r4 = 0; r4 < kRK / 4; ++r4) {
      float a[RM][4], b[4][4];   // a[i][rr], b[rr][j]
      if (AL == RC_A_ROW) {
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const float4 v = ld4(as + (warp * RM + i) * SA + r4 * 4);
          a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float* ap = as + (r4 * 4 + rr) * SA + warp * RM;
          if (RM >= 4) {
#pragma unroll
            for (int i4 = 0; i4 < RM / 4; ++i4) {
              const float4 v = ld4(ap + 4 * i4);
              a[4 * i4 + 0][rr] = v.x; a[4 * i4 + 1][rr] = v.y; a[4 * i4 + 2][rr] = v.z; a[4 * i4 + 3][rr] = v.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i][rr] = ap[i];
          }
        }
      }
      if (BL == RC_B_COL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = ld4(bs + (lane + 32 * j) * SB + r4 * 4);
          b[0][j] = v.x; b[1][j] = v.y; b[2][j] = v.z; b[3][j] = v.w;
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float4 v = ld4(bs + (r4 * 4 + rr) * SB + 4 * lane);
          b[rr][0] = v.x; b[rr][1] = v.y; b[rr][2] = v.z; b[rr][3] = v.w;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][rr], b[rr][j], acc[i][j]);
    }
    if (more) store_tile(first ? 0 : (cur ^ 1));
    __syncthreads();
    if (!first) cur ^= 1;
  }

  // ------------------------------------------------------------------------------------------ epilogue
  float* red = smem;   // the tile buffers are free now (all warps passed the last barrier)
  float* dout = g.d + (g.splits > 1 ? (size_t)z * g.split_stride : 0);
  const int valid_rows = min(BM, g.m - m0);

  auto tile_col = [&](int j) { return (BL == RC_B_COL) ? (lane + 32 * j) : (4 * lane + j); };

  // bias
  float bias_v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + tile_col(j);
    bias_v[j] = (g.bias != nullptr && col < g.n) ? g.bias_scale * __ldg(g.bias + col) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += bias_v[j];

  // (weight-gradient GEMMs, A stored [r][i], have no epilogue beyond the store: the variants below are compiled out of them)
  if (AL != RC_A_RED && (g.epi == RC_EPI_BN_STATS || g.epi == RC_EPI_BN_RELU_BWD)) {
    // per-column reductions over the rows of this tile: warp partials -> smem -> 128 column threads
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.epi == RC_EPI_BN_RELU_BWD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tile_col(j);
        const bool cok = col < g.n;
        const float mean = cok ? __ldg(g.e_p0 + col) : 0.f, rstd = cok ? __ldg(g.e_p1 + col) : 0.f;
        const float gamma = cok ? __ldg(g.e_p2 + col) : 0.f, beta = cok ? __ldg(g.e_p3 + col) : 0.f;
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const int row = m0 + warp * RM + i;
          float hat = 0.f, dz = 0.f;
          if (cok && row < g.m) {
            hat = ((e_early ? e_pre[j] : __ldg(g.e_aux + (size_t)row * g.ld_e_aux + col)) - mean) * rstd;
            dz = (fmaf(gamma, hat, beta) > 0.f) ? acc[i][j] : 0.f;
          }
          acc[i][j] = dz;
          s0[j] += dz;
          s1[j] = fmaf(dz, hat, s1[j]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < RM; ++i)
          if (m0 + warp * RM + i < g.m) s0[j] += acc[i][j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[warp * kBN + tile_col(j)] = s0[j];
      red[(8 + warp) * kBN + tile_col(j)] = s1[j];
    }
    __syncthreads();
    float* stats = g.stats + (size_t)bid.x * 2 * g.n;
    if (g.epi == RC_EPI_BN_RELU_BWD) {
      if (tid < kBN && n0 + tid < g.n) {
        float t0 = 0.f, t1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { t0 += red[w * kBN + tid]; t1 += red[(8 + w) * kBN + tid]; }
        stats[n0 + tid] = t0;
        stats[g.n + n0 + tid] = t1;
      }
    } else {
      // tile mean, then centred second moment (two passes over registers: robust to |mean| >> std)
      if (tid < kBN) {
        float t0 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t0 += red[w * kBN + tid];
        red[16 * kBN + tid] = t0 / (float)valid_rows;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float mean = red[16 * kBN + tile_col(j)];
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < RM; ++i)
          if (m0 + warp * RM + i < g.m) { const float d = acc[i][j] - mean; q = fmaf(d, d, q); }
        red[warp * kBN + tile_col(j)] = q;
      }
      __syncthreads();
      if (tid < kBN && n0 + tid < g.n) {
        float q = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) q += red[w * kBN + tid];
        stats[n0 + tid] = red[16 * kBN + tid];
        stats[g.n + n0 + tid] = q;
      }
    }
  }

  // elementwise epilogue + store
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int row = m0 + warp * RM + i;
    const bool rok = row < g.m;
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tile_col(j);
      const bool ok = rok && col < g.n;
      float v = acc[i][j];
      bool pos = v > 0.f;
      if (AL == RC_A_RED) {
      } else if (g.epi == RC_EPI_RELU) {
        v = fmaxf(v, 0.f);
      } else if (g.epi == RC_EPI_RELU_RES) {
        v = (ok ? (e_early ? e_pre[j] : __ldg(g.res + (size_t)row * g.ld_res + col)) : 0.f) + fmaxf(v, 0.f);
      } else if (g.epi == RC_EPI_ADD_RES) {
        v += ok ? (e_early ? e_pre[j] : __ldg(g.res + (size_t)row * g.ld_res + col)) : 0.f;
      } else if (g.epi == RC_EPI_MASK_POS) {
        v = (ok && (e_early ? e_pre[j] : __ldg(g.e_aux + (size_t)row * g.ld_e_aux + col)) > 0.f) ? v : 0.f;
      }
      out[j] = v;
      if (AL != RC_A_RED && BL == RC_B_COL && g.bits_out != nullptr) {
        const unsigned word = __ballot_sync(0xffffffffu, ok && pos);   // lane l <-> column 32*j + l of the tile
        if (lane == 0 && rok && n0 + 32 * j < g.n) g.bits_out[(size_t)row * g.ld_bits_out + (n0 >> 5) + j] = word;
      }
    }
    if (!rok) continue;
    float* dp = dout + (size_t)row * g.ldd;
    if (BL == RC_B_RED) {
      const int col = n0 + 4 * lane;
      if (p.d_vec && col + 3 < g.n) {
        st4(dp + col, make_float4(out[0], out[1], out[2], out[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < g.n) dp[col + j] = out[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + lane + 32 * j;
        if (col < g.n) dp[col] = out[j];
      }
    }
  }

  // column sums of the stored A operand over this CTA's reduction slice (bias gradient)
  if (AL == RC_A_RED && g.colsum_a != nullptr && bid.y == 0) {
    __syncthreads();
    float* cs = smem;   // [kRK][BM]
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
      const int s = tid + it * kGemmThreads;
      if (A_SLOTS >= kGemmThreads || s < A_SLOTS) {
        const int row = s / A_F4_PER_ROW, c4 = (s % A_F4_PER_ROW) * 4;
        st4(cs + row * BM + c4, csum[it]);
      }
    }
    __syncthreads();
    if (tid < BM && m0 + tid < g.m) {
      float t = 0.f;
#pragma unroll 8
      for (int r = 0; r < kRK; ++r) t += cs[r * BM + tid];
      g.colsum_a[(size_t)z * g.m + m0 + tid] = t;
    }
  }
}

}  // namespace rc
