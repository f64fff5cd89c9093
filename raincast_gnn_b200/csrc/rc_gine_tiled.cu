// GINE aggregation over station tiles (rc_tiles.cu): the large-graph path.
//
// PyG GINEConv message + 'add' aggregation (call site models/gnn.py:27-29,39-44) gathers one 4*H-byte source row per
// edge.  At ~30 edges per station that is 30x the compulsory traffic, and the plain warp-per-row kernels of
// rc_gine.cu / rc_gine_wide.cu sit at the L2 -> SM throughput cap (profiles/r01_ncu_gine_aggr.txt).  Here a CTA
// stages the rows a tile gathers in shared memory once (own rows + halo, ~2.3 rows per owned row on the config-4
// graph, exactly 1.0 on batched reference graphs) and every edge reads its source row from shared memory.
//   - staging is asynchronous (cp.async, global -> shared without registers): the tile's block of row + edge records
//     and one 16-byte-per-lane copy per gathered row, issued by all warps; it overlaps the other CTA of the SM
//   - one warp per row, one lane per 4 columns: conflict-free 128-bit LDS, coalesced 128-bit row stores
//   - edge records {byte offset of the staged row, attr} are read from shared memory two at a time (one broadcast
//     128-bit LDS per two edges): no global load inside the edge loop
//   - sums run in CSR slot order, with the same expressions as the untiled kernels: results are bitwise equal
// Backward: the transpose tiles stage g rows; the ReLU mask is recomputed from the row's own x and the edge attr
// (nothing per-edge saved); d w_edge / d b_edge / d eps partials per CTA, same format as rc_gine_aggr_bwd.
#include "rc_common.cuh"
#include "rc_prog.h"

namespace rc {

constexpr int kTiledThreads = 512;
constexpr int kTiledWarps = kTiledThreads / 32;
constexpr int kSmemPerSM = 227 * 1024;

struct TilesP {
  int n_tiles;
  int row_bytes;     // 4 * hidden
  int rows_bytes;    // shared-memory bytes of the staged-row region (max_staged * row_bytes)
  int blk_bytes;     // shared-memory bytes of the block region (largest tile block)
  const int* __restrict__ tile_stage_ptr;
  const int* __restrict__ tile_blk_ptr;
  const int* __restrict__ stage_id;
  const int* __restrict__ blocks;
};

__host__ __device__ constexpr int tiled_ctas_per_sm(int ch) { return ch == 1 ? 2 : 1; }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// All warps: start the asynchronous copies of one tile (cp.async, 16 bytes per lane: one warp instruction moves one
// 4*128-byte row slice global -> shared without passing through registers).  The caller commits / waits.
// (One cp.async.bulk per row was measured first: ~60 cycles of serialised issue per 512-byte copy, 11k cycles per tile.)
template <int CH>
__device__ __forceinline__ void issue_tile(const float* __restrict__ src, const TilesP& t, int tile, unsigned char* rows,
                                           unsigned char* blk, int lane, int warp) {
  const int s0 = __ldg(t.tile_stage_ptr + tile), nst = __ldg(t.tile_stage_ptr + tile + 1) - s0;
  const int b0 = __ldg(t.tile_blk_ptr + tile), bunits = __ldg(t.tile_blk_ptr + tile + 1) - b0;
  // the tile's block of row / edge records: contiguous, 16 bytes per thread per pass
  const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(t.blocks) + (size_t)b0 * 16;
  for (int u = threadIdx.x; u < bunits; u += kTiledThreads)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(blk + 16 * u)), "l"(bsrc + 16 * (size_t)u) : "memory");
  // gathered rows: lane k of a warp fetches the id of the warp's k-th row, then the ids are broadcast row by row
  const unsigned char* base = reinterpret_cast<const unsigned char*>(src) + 16 * lane;
  for (int l0 = warp; l0 < nst; l0 += 32 * kTiledWarps) {
    const int mine = l0 + lane * kTiledWarps;
    const int my_id = mine < nst ? __ldg(t.stage_id + s0 + mine) : 0;
    const int cnt = min(32, (nst - l0 + kTiledWarps - 1) / kTiledWarps);
    for (int k = 0; k < cnt; ++k) {
      const int id = __shfl_sync(0xffffffffu, my_id, k);
      const uint32_t dst = smem_addr(rows + (size_t)(l0 + k * kTiledWarps) * t.row_bytes + 16 * lane);
      const unsigned char* sp = base + (size_t)id * t.row_bytes;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512 * c), "l"(sp + 512 * c) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void wait_tile(int* next_row) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (threadIdx.x == 0) *next_row = 0;     // nobody claims before the barrier below; the barrier at the tile's end orders the reset
  __syncthreads();
}

__device__ __forceinline__ int claim_row(int* next_row, int lane) {
  int r = 0;
  if (lane == 0) r = atomicAdd(next_row, 1);
  return __shfl_sync(0xffffffffu, r, 0);
}

__device__ __forceinline__ void relu_acc2(float4& acc, float4 v, float a, float4 w, float4 b) {
  const float2 a2 = make_float2(a, a);
  float2 z0 = __fadd2_rn(make_float2(v.x, v.y), __ffma2_rn(a2, make_float2(w.x, w.y), make_float2(b.x, b.y)));
  float2 z1 = __fadd2_rn(make_float2(v.z, v.w), __ffma2_rn(a2, make_float2(w.z, w.w), make_float2(b.z, b.w)));
  z0.x = fmaxf(z0.x, 0.f); z0.y = fmaxf(z0.y, 0.f);
  z1.x = fmaxf(z1.x, 0.f); z1.y = fmaxf(z1.y, 0.f);
  const float2 s0 = __fadd2_rn(make_float2(acc.x, acc.y), z0), s1 = __fadd2_rn(make_float2(acc.z, acc.w), z1);
  acc = make_float4(s0.x, s0.y, s1.x, s1.y);
}

__device__ __forceinline__ float4 lds4(const unsigned char* p) { return *reinterpret_cast<const float4*>(p); }

template <int CH>
__global__ void __launch_bounds__(kTiledThreads, tiled_ctas_per_sm(CH))
gine_aggr_fwd_tiled_kernel(const float* __restrict__ x, const TilesP t, const float* __restrict__ w_edge,
                           const float* __restrict__ b_edge, const float* __restrict__ eps_ptr, float* __restrict__ h, int hidden) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* rows = smem;
  unsigned char* blk = smem + t.rows_bytes;
  int* next_row = reinterpret_cast<int*>(blk + t.blk_bytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float self_scale = 1.0f + __ldg(eps_ptr);
  float4 w4[CH], b4[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    w4[c] = ldg4(w_edge + 4 * (lane + 32 * c));
    b4[c] = ldg4(b_edge + 4 * (lane + 32 * c));
  }
  const unsigned char* rl = rows + 16 * lane;        // this lane's 4 columns inside a staged row
#ifdef RC_TILED_PROFILE
  long long prof_issue = 0, prof_wait = 0, prof_rows = 0, prof_bar = 0, prof_tiles = 0;
#endif
  for (int tile = blockIdx.x; tile < t.n_tiles; tile += gridDim.x) {
#ifdef RC_TILED_PROFILE
    const long long c0 = clock64();
#endif
    issue_tile<CH>(x, t, tile, rows, blk, lane, warp);
#ifdef RC_TILED_PROFILE
    const long long c1 = clock64();
#endif
    wait_tile(next_row);
#ifdef RC_TILED_PROFILE
    const long long c2 = clock64();
    if (threadIdx.x == 0) { prof_issue += c1 - c0; prof_wait += c2 - c1; }
#endif
    const int nrows = *reinterpret_cast<const int*>(blk);
    // rows are claimed in order (longest first) from a shared counter: warps finish a tile within one short row of
    // each other whatever the degree distribution
    for (int r = claim_row(next_row, lane); r < nrows; r = claim_row(next_row, lane)) {
      const int4 rec = *reinterpret_cast<const int4*>(blk + 16 + 16 * r);      // {node, edge byte offset, degree, 0}
      const unsigned char* ep = blk + rec.y;
      const int deg = rec.z;
      float4 acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      int k = 0;
      for (; k + 4 <= deg; k += 4) {                   // slot order = reference edge order
        const int4 m01 = *reinterpret_cast<const int4*>(ep + 8 * k);
        const int4 m23 = *reinterpret_cast<const int4*>(ep + 8 * k + 16);
        float4 v[4][CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          v[0][c] = lds4(rl + m01.x + 512 * c);
          v[1][c] = lds4(rl + m01.z + 512 * c);
          v[2][c] = lds4(rl + m23.x + 512 * c);
          v[3][c] = lds4(rl + m23.z + 512 * c);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], v[0][c], __int_as_float(m01.y), w4[c], b4[c]);
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], v[1][c], __int_as_float(m01.w), w4[c], b4[c]);
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], v[2][c], __int_as_float(m23.y), w4[c], b4[c]);
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], v[3][c], __int_as_float(m23.w), w4[c], b4[c]);
      }
      if (k + 2 <= deg) {
        const int4 m01 = *reinterpret_cast<const int4*>(ep + 8 * k);
        float4 v[2][CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          v[0][c] = lds4(rl + m01.x + 512 * c);
          v[1][c] = lds4(rl + m01.z + 512 * c);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], v[0][c], __int_as_float(m01.y), w4[c], b4[c]);
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], v[1][c], __int_as_float(m01.w), w4[c], b4[c]);
        k += 2;
      }
      if (k < deg) {
        const int2 m = *reinterpret_cast<const int2*>(ep + 8 * k);
#pragma unroll
        for (int c = 0; c < CH; ++c) relu_acc2(acc[c], lds4(rl + m.x + 512 * c), __int_as_float(m.y), w4[c], b4[c]);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 xi = lds4(rl + (size_t)r * t.row_bytes + 512 * c);
        float4 o;
        o.x = acc[c].x + self_scale * xi.x;
        o.y = acc[c].y + self_scale * xi.y;
        o.z = acc[c].z + self_scale * xi.z;
        o.w = acc[c].w + self_scale * xi.w;
        st4(h + (size_t)rec.x * hidden + 4 * lane + 128 * c, o);
      }
    }
#ifdef RC_TILED_PROFILE
    const long long c3 = clock64();
#endif
    __syncthreads();       // every read of this tile is done before the next tile's copies land
#ifdef RC_TILED_PROFILE
    if (threadIdx.x == 0) { prof_rows += c3 - c2; prof_bar += clock64() - c3; ++prof_tiles; }
#endif
  }
#ifdef RC_TILED_PROFILE
  if (threadIdx.x == 0) {
    long long* o = reinterpret_cast<long long*>(h) ;
    (void)o;
    printf("cta %d tiles %lld issue %lld wait %lld rows(warp0) %lld bar %lld\n", blockIdx.x, prof_tiles, prof_issue, prof_wait, prof_rows, prof_bar);
  }
#endif
}

__device__ __forceinline__ void masked_acc2(float4& acc, float4& acc_a, float4 g, float4 xj, float a, float4 w, float4 b) {
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

// dynamic shared memory: staged rows | tile block | mbarrier; the final reduction reuses the front of it
// (red[kTiledWarps][2*hidden] + red_eps[kTiledWarps], never larger than what the launch reserves)
template <int CH>
__global__ void __launch_bounds__(kTiledThreads, tiled_ctas_per_sm(CH))
gine_aggr_bwd_tiled_kernel(const float* __restrict__ g, const float* __restrict__ x, const TilesP t,
                           const float* __restrict__ w_edge, const float* __restrict__ b_edge, const float* __restrict__ eps_ptr,
                           const float* __restrict__ addend, float* __restrict__ dx, float* __restrict__ partials, int hidden) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* rows = smem;
  unsigned char* blk = smem + t.rows_bytes;
  int* next_row = reinterpret_cast<int*>(blk + t.blk_bytes);
  constexpr int kU = CH <= 2 ? 2 : 1;      // edge pairs in flight per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float self_scale = 1.0f + __ldg(eps_ptr);
  float4 w4[CH], b4[CH], dw[CH], db[CH];
  double deps = 0.0;   // <g, x> cancels heavily over M*H products: float64 across rows
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    w4[c] = ldg4(w_edge + 4 * (lane + 32 * c));
    b4[c] = ldg4(b_edge + 4 * (lane + 32 * c));
    dw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const unsigned char* rl = rows + 16 * lane;
  for (int tile = blockIdx.x; tile < t.n_tiles; tile += gridDim.x) {
    issue_tile<CH>(g, t, tile, rows, blk, lane, warp);
    wait_tile(next_row);
    const int nrows = *reinterpret_cast<const int*>(blk);
    for (int r = claim_row(next_row, lane); r < nrows; r = claim_row(next_row, lane)) {
      const int4 rec = *reinterpret_cast<const int4*>(blk + 16 + 16 * r);
      const unsigned char* ep = blk + rec.y;
      const int deg = rec.z;
      float4 xj[CH], acc[CH], acc_a[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        xj[c] = ldg4(x + (size_t)rec.x * hidden + 4 * lane + 128 * c);
        acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        acc_a[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      int k = 0;
      for (; k + 2 * kU <= deg; k += 2 * kU) {
        int4 m[kU];
        float4 v[2 * kU][CH];
#pragma unroll
        for (int u = 0; u < kU; ++u) m[u] = *reinterpret_cast<const int4*>(ep + 8 * k + 16 * u);
#pragma unroll
        for (int u = 0; u < kU; ++u)
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            v[2 * u][c] = lds4(rl + m[u].x + 512 * c);
            v[2 * u + 1][c] = lds4(rl + m[u].z + 512 * c);
          }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
#pragma unroll
          for (int c = 0; c < CH; ++c) masked_acc2(acc[c], acc_a[c], v[2 * u][c], xj[c], __int_as_float(m[u].y), w4[c], b4[c]);
#pragma unroll
          for (int c = 0; c < CH; ++c) masked_acc2(acc[c], acc_a[c], v[2 * u + 1][c], xj[c], __int_as_float(m[u].w), w4[c], b4[c]);
        }
      }
      for (; k < deg; ++k) {
        const int2 m = *reinterpret_cast<const int2*>(ep + 8 * k);
#pragma unroll
        for (int c = 0; c < CH; ++c) masked_acc2(acc[c], acc_a[c], lds4(rl + m.x + 512 * c), xj[c], __int_as_float(m.y), w4[c], b4[c]);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float4 gj = lds4(rl + (size_t)r * t.row_bytes + 512 * c);
        deps += (double)(gj.x * xj[c].x + gj.y * xj[c].y + gj.z * xj[c].z + gj.w * xj[c].w);
        // sum_e gm_e = acc and sum_e gm_e a_e = acc_a: the bias / weight gradients take them once per row
        db[c].x += acc[c].x; db[c].y += acc[c].y; db[c].z += acc[c].z; db[c].w += acc[c].w;
        dw[c].x += acc_a[c].x; dw[c].y += acc_a[c].y; dw[c].z += acc_a[c].z; dw[c].w += acc_a[c].w;
        float4 o;
        o.x = fmaf(self_scale, gj.x, acc[c].x);
        o.y = fmaf(self_scale, gj.y, acc[c].y);
        o.z = fmaf(self_scale, gj.z, acc[c].z);
        o.w = fmaf(self_scale, gj.w, acc[c].w);
        if (addend != nullptr) {
          const float4 ad = ldg4(addend + (size_t)rec.x * hidden + 4 * lane + 128 * c);
          o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
        }
        st4(dx + (size_t)rec.x * hidden + 4 * lane + 128 * c, o);
      }
    }
    __syncthreads();
  }
  float* red = reinterpret_cast<float*>(smem);
  float* red_eps = red + kTiledWarps * 2 * hidden;
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
#pragma unroll
    for (int r = 0; r < kTiledWarps; ++r) s += red[(size_t)r * 2 * hidden + j];
    out[j] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < kTiledWarps; ++r) s += red_eps[r];
    out[2 * hidden] = s;
  }
}

static int tiled_grid(int n_tiles, int ch) {
  const int cap = kNumSMs * tiled_ctas_per_sm(ch);
  return n_tiles < cap ? n_tiles : cap;
}

static int smem_budget(int hidden) { return kSmemPerSM / tiled_ctas_per_sm(hidden / 128) - 1024; }   // 1 KiB per CTA is system-reserved

static size_t tiled_smem(const TilesP& t, int hidden, bool bwd) {
  size_t b = (size_t)t.rows_bytes + t.blk_bytes + 16;
  const size_t red = ((size_t)kTiledWarps * 2 * hidden + kTiledWarps) * sizeof(float);
  if (bwd && red > b) b = red;
  return b;
}

static int check_tiles(const rc_gine_tiles* t, int hidden, const char* who) {
  if (!t || t->n_tiles < 0 || !t->tile_stage_ptr || !t->tile_blk_ptr || !t->stage_id || !t->blocks)
    return fail(RC_ERR_ARG, "%s: null tile array", who);
  if (hidden < 128 || hidden % 128 || hidden > 512) return fail(RC_ERR_ARG, "%s: hidden=%d unsupported (128 | H, H <= 512)", who, hidden);
  if (t->row_bytes != hidden * 4) return fail(RC_ERR_ARG, "%s: tiles were built for rows of %d bytes, hidden=%d needs %d", who, t->row_bytes, hidden, hidden * 4);
  if (t->max_staged < 0 || t->max_block_bytes < 0 || t->max_block_bytes % 16 ||
      (long long)t->max_staged * t->row_bytes + t->max_block_bytes + 16 > smem_budget(hidden))
    return fail(RC_ERR_ARG, "%s: a tile of %d rows + %d block bytes exceeds the %d bytes of shared memory per CTA", who, t->max_staged,
                t->max_block_bytes, smem_budget(hidden));
  if (!aligned16(t->blocks)) return fail(RC_ERR_ARG, "%s: blocks must be 16-byte aligned", who);
  return RC_OK;
}

static TilesP tiles_param(const rc_gine_tiles* t) {
  return TilesP{t->n_tiles, t->row_bytes, t->max_staged * t->row_bytes, t->max_block_bytes, t->tile_stage_ptr, t->tile_blk_ptr,
                t->stage_id, t->blocks};
}

}  // namespace rc

using namespace rc;

extern "C" int rc_gine_tiles_limits(int hidden, int* max_src, int* max_block_bytes) {
  if (!max_src || !max_block_bytes || hidden < 128 || hidden % 128 || hidden > 512) return fail(RC_ERR_ARG, "rc_gine_tiles_limits: hidden=%d unsupported", hidden);
  // split of the per-CTA shared memory: ~3/4 staged rows, the rest for the tile's row + edge records
  const int budget = smem_budget(hidden) - 16;
  const int rows = (budget * 25 / 32) / (hidden * 4);
  *max_src = rows;
  *max_block_bytes = (budget - rows * hidden * 4) / 16 * 16;
  return RC_OK;
}

extern "C" int rc_gine_aggr_fwd_tiled(const float* x, const rc_gine_tiles* tiles, const float* w_edge, const float* b_edge,
                                      const float* eps, float* h, int num_nodes, int hidden, void* stream) {
  if (!x || !w_edge || !b_edge || !eps || !h || num_nodes < 0) return fail(RC_ERR_ARG, "rc_gine_aggr_fwd_tiled: null pointer");
  if (int rc = check_tiles(tiles, hidden, "rc_gine_aggr_fwd_tiled")) return rc;
  if (recording()) return fail(RC_ERR_ARG, "rc_gine_aggr_fwd_tiled: not available inside a step program (large-graph path)");
  if (!aligned16(x) || !aligned16(h) || !aligned16(w_edge) || !aligned16(b_edge))
    return fail(RC_ERR_ARG, "rc_gine_aggr_fwd_tiled: x, h, w_edge, b_edge must be 16-byte aligned");
  if (num_nodes == 0 || tiles->n_tiles == 0) return RC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ch = hidden / 128;
  const int grid = tiled_grid(tiles->n_tiles, ch);
  const TilesP t = tiles_param(tiles);
  const size_t smem = tiled_smem(t, hidden, false);
#define RC_LAUNCH(CHV)                                                                                                       \
  do {                                                                                                                       \
    cudaFuncSetAttribute(gine_aggr_fwd_tiled_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    gine_aggr_fwd_tiled_kernel<CHV><<<grid, kTiledThreads, smem, s>>>(x, t, w_edge, b_edge, eps, h, hidden);                 \
  } while (0)
  switch (ch) {
    case 1: RC_LAUNCH(1); break;
    case 2: RC_LAUNCH(2); break;
    case 3: RC_LAUNCH(3); break;
    default: RC_LAUNCH(4); break;
  }
#undef RC_LAUNCH
  return check_launch("gine_aggr_fwd_tiled_kernel");
}

extern "C" int rc_gine_aggr_bwd_tiled_nblocks(const rc_gine_tiles* tiles, int hidden) {
  if (!tiles || hidden < 128 || hidden % 128 || hidden > 512) return -1;
  const int nb = tiled_grid(tiles->n_tiles, hidden / 128);
  return nb > 0 ? nb : 1;
}

extern "C" int rc_gine_aggr_bwd_tiled(const float* g, const float* x, const rc_gine_tiles* tiles, const float* w_edge,
                                      const float* b_edge, const float* eps, const float* addend, float* dx, float* partials,
                                      int num_nodes, int hidden, void* stream) {
  if (!g || !x || !w_edge || !b_edge || !eps || !dx || !partials || num_nodes < 0) return fail(RC_ERR_ARG, "rc_gine_aggr_bwd_tiled: null pointer");
  if (int rc = check_tiles(tiles, hidden, "rc_gine_aggr_bwd_tiled")) return rc;
  if (recording()) return fail(RC_ERR_ARG, "rc_gine_aggr_bwd_tiled: not available inside a step program (large-graph path)");
  if (!aligned16(g) || !aligned16(x) || !aligned16(dx) || !aligned16(w_edge) || !aligned16(b_edge) || !aligned16(partials) ||
      (addend && !aligned16(addend)))
    return fail(RC_ERR_ARG, "rc_gine_aggr_bwd_tiled: g, x, dx, addend, w_edge, b_edge, partials must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ch = hidden / 128;
  const int grid = rc_gine_aggr_bwd_tiled_nblocks(tiles, hidden);   // >= 1: an empty graph still zeroes its partials
  const TilesP t = tiles_param(tiles);
  const size_t smem = tiled_smem(t, hidden, true);
#define RC_LAUNCH(CHV)                                                                                                       \
  do {                                                                                                                       \
    cudaFuncSetAttribute(gine_aggr_bwd_tiled_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    gine_aggr_bwd_tiled_kernel<CHV><<<grid, kTiledThreads, smem, s>>>(g, x, t, w_edge, b_edge, eps, addend, dx, partials, hidden); \
  } while (0)
  switch (ch) {
    case 1: RC_LAUNCH(1); break;
    case 2: RC_LAUNCH(2); break;
    case 3: RC_LAUNCH(3); break;
    default: RC_LAUNCH(4); break;
  }
#undef RC_LAUNCH
  return check_launch("gine_aggr_bwd_tiled_kernel");
}
