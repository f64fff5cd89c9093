// GINE aggregation over station tiles (rc_tiles.cu): the large-graph path.
//
// PyG GINEConv message + 'add' aggregation (call site models/gnn.py:27-29,39-44) gathers one 4*H-byte source row per
// edge.  At ~30 edges per station that is 30x the compulsory traffic, and the plain warp-per-row kernels of
// rc_gine.cu / rc_gine_wide.cu sit at the L2 -> SM throughput cap (profiles/r01_ncu_gine_aggr.txt).  Here one
// persistent CTA per SM walks its share of the tiles through a two-buffer shared-memory pipeline:
//   - producer warps (one per SM sub-partition) stage the rows a tile gathers (own rows + halo, ~2.5 rows per owned
//     row on the config-4 graph, exactly 1.0 on batched reference graphs) and the tile's block of group / entry
//     records with cp.async (global -> shared without registers, 16 bytes per lane) and signal an mbarrier when
//     the copies have landed; the next tile is in flight while the consumers work on the current one
//   - consumer warps claim GROUPS of three neighbouring rows from a shared counter (groups with most edges first).
//     Each distinct source row of the group is read from shared memory once and used by every row of the group that
//     has an edge from it (2.2 edges per 512-byte read on config 4) - the shared-memory data pipe was the limiter of
//     the row-at-a-time version (profiles/r01_ncu_gine_tiled.txt).  A warp that finds a tile's groups all taken
//     releases the buffer (mbarrier) and moves on to the next tile: no CTA-wide barrier in the loop
//   - one lane per 4 columns of a 128-column chunk: conflict-free 128-bit LDS, coalesced 128-bit row stores.  Wider
//     rows are walked chunk by chunk (a CTA stays on one chunk, so tiles do not depend on the width)
//   - a group's entries are sorted by class (= which of its rows use the source); one straight-line loop per class,
//     one broadcast 16-byte record {staged offset, attr per row} per entry: no per-edge test, no global load
//   - per (edge, 4 columns): 2 packed FMA (a*w + x_j), 4 max, 2 packed adds.  The bias is taken out of the loop:
//     relu(x_j + a*w + b) = max(x_j + a*w, -b) + b, so a row adds degree * b once at the end
//   - deterministic: the summation order is fixed by the tiles (class by class, not the CSR slot order of the
//     untiled kernels - results agree to rounding, not bit for bit)
// Backward: the transpose tiles stage g rows; the ReLU mask is recomputed from the row's own x and the edge attr
// (nothing per-edge saved); d w_edge / d b_edge / d eps partials per CTA, same format as rc_gine_aggr_bwd.
#include "rc_common.cuh"
#include "rc_prog.h"

namespace rc {

constexpr int kSmemPerCTA = 227 * 1024;   // opt-in maximum of dynamic shared memory
constexpr int kTileRowBytes = 512;        // one 128-column chunk of a row
constexpr int kProdWarps = 4;             // producer warps: warp w runs on sub-partition w
constexpr int kFwdThreads = 1024;
constexpr int kBwdThreads = 768;
constexpr int kCtrlBytes = 64;            // full[2], empty[2] mbarriers + two item counters

struct TilesP {
  int n_tiles;
  int rows_bytes;    // shared-memory bytes of one buffer's staged-row region (max_staged * 512)
  int buf_bytes;     // one buffer: staged rows + the largest tile block
  const int* __restrict__ tile_stage_ptr;
  const int* __restrict__ tile_blk_ptr;
  const int* __restrict__ stage_id;
  const int* __restrict__ blocks;
};

struct Ctrl {
  unsigned long long full[2], empty[2];
  int counter[2];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar)) : "memory");
}
// arrive (without incrementing the pending count) once all cp.async of this thread issued so far have landed
__device__ __forceinline__ void mbar_arrive_on_copies(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// Producer warps: stage this CTA's tiles, two in flight.  `src` already points at the CTA's column chunk; rows are
// `row_stride` bytes apart in global memory and 512 bytes apart in shared memory.
__device__ __forceinline__ void produce_tiles(const float* __restrict__ src, size_t row_stride, const TilesP& t, int first, int step,
                                              unsigned char* smem, Ctrl* ctrl, int lane, int warp) {
  const unsigned char* base = reinterpret_cast<const unsigned char*>(src) + 16 * lane;
  int it = 0;
  for (int tile = first; tile < t.n_tiles; tile += step, ++it) {
    const int b = it & 1;
    unsigned char* rows = smem + (size_t)b * t.buf_bytes;
    unsigned char* blk = rows + t.rows_bytes;
    if (it >= 2) mbar_wait(&ctrl->empty[b], ((it >> 1) & 1) ^ 1);      // every consumer warp is done with this buffer
    const int s0 = __ldg(t.tile_stage_ptr + tile), nst = __ldg(t.tile_stage_ptr + tile + 1) - s0;
    const int b0 = __ldg(t.tile_blk_ptr + tile), bunits = __ldg(t.tile_blk_ptr + tile + 1) - b0;
    // the tile's block of group / entry records: contiguous, 16 bytes per thread per pass
    const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(t.blocks) + (size_t)b0 * 16;
    for (int u = warp * 32 + lane; u < bunits; u += 32 * kProdWarps)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(blk + 16 * u)), "l"(bsrc + 16 * (size_t)u) : "memory");
    // gathered rows: lane k of a warp fetches the id of the warp's k-th row, then the ids are broadcast row by row
    for (int l0 = warp; l0 < nst; l0 += 32 * kProdWarps) {
      const int mine = l0 + lane * kProdWarps;
      const int my_id = mine < nst ? __ldg(t.stage_id + s0 + mine) : 0;
      const int cnt = min(32, (nst - l0 + kProdWarps - 1) / kProdWarps);
      for (int k = 0; k < cnt; ++k) {
        const int id = __shfl_sync(0xffffffffu, my_id, k);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(rows + (size_t)(l0 + k * kProdWarps) * kTileRowBytes + 16 * lane)),
                     "l"(base + (size_t)id * row_stride) : "memory");
      }
    }
    if (warp == 0 && lane == 0) {
      ctrl->counter[b] = 0;                 // nobody claims from this buffer between its release and the arrival below
      mbar_arrive(&ctrl->full[b]);          // release: orders the counter store before the consumers' claims
    }
    mbar_arrive_on_copies(&ctrl->full[b]);
  }
}

__device__ __forceinline__ void init_ctrl(Ctrl* ctrl, int consumer_warps) {
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctrl->full[b], 32 * kProdWarps + 1);
      mbar_init(&ctrl->empty[b], consumer_warps);
      ctrl->counter[b] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

__device__ __forceinline__ int claim_item(int* next_item, int lane) {
  int r = 0;
  if (lane == 0) r = atomicAdd(next_item, 1);
  return __shfl_sync(0xffffffffu, r, 0);
}

__device__ __forceinline__ float4 lds4(const unsigned char* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ int4 ldsi4(const unsigned char* p) { return *reinterpret_cast<const int4*>(p); }

// acc += max(a * w + v, nb)   (nb = -b_edge; the row adds degree * b_edge at the end)
__device__ __forceinline__ void relu_acc(float4& acc, float4 v, float a, float4 w, float4 nb) {
  const float2 a2 = make_float2(a, a);
  float2 z0 = __ffma2_rn(a2, make_float2(w.x, w.y), make_float2(v.x, v.y));
  float2 z1 = __ffma2_rn(a2, make_float2(w.z, w.w), make_float2(v.z, v.w));
  z0.x = fmaxf(z0.x, nb.x); z0.y = fmaxf(z0.y, nb.y);
  z1.x = fmaxf(z1.x, nb.z); z1.y = fmaxf(z1.y, nb.w);
  const float2 s0 = __fadd2_rn(make_float2(acc.x, acc.y), z0), s1 = __fadd2_rn(make_float2(acc.z, acc.w), z1);
  acc = make_float4(s0.x, s0.y, s1.x, s1.y);
}

template <int MASK>
__device__ __forceinline__ void fwd_apply(float4 (&acc)[3], float4 v, int4 rec, float4 w, float4 nb) {
  if (MASK & 1) relu_acc(acc[0], v, __int_as_float(rec.y), w, nb);
  if (MASK & 2) relu_acc(acc[1], v, __int_as_float(rec.z), w, nb);
  if (MASK & 4) relu_acc(acc[2], v, __int_as_float(rec.w), w, nb);
}

// the n entries of one class, two in flight; returns the next class's first entry
template <int MASK>
__device__ __forceinline__ const unsigned char* fwd_class(float4 (&acc)[3], const unsigned char* p, int n, const unsigned char* rl,
                                                          float4 w, float4 nb) {
  for (; n >= 2; n -= 2, p += 32) {
    const int4 r0 = ldsi4(p), r1 = ldsi4(p + 16);
    const float4 v0 = lds4(rl + r0.x), v1 = lds4(rl + r1.x);
    fwd_apply<MASK>(acc, v0, r0, w, nb);
    fwd_apply<MASK>(acc, v1, r1, w, nb);
  }
  if (n) {
    const int4 r0 = ldsi4(p);
    fwd_apply<MASK>(acc, lds4(rl + r0.x), r0, w, nb);
    p += 16;
  }
  return p;
}

// grid = chunks * ctas_per_chunk; CTA b works on column chunk b % chunks and on tiles b / chunks + k * ctas_per_chunk
__global__ void __launch_bounds__(kFwdThreads, 1)
gine_aggr_fwd_tiled_kernel(const float* __restrict__ x, const TilesP t, const float* __restrict__ w_edge,
                           const float* __restrict__ b_edge, const float* __restrict__ eps_ptr, float* __restrict__ h, int hidden,
                           int chunks) {
  extern __shared__ __align__(128) unsigned char smem[];
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + 2 * (size_t)t.buf_bytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x % chunks, first = blockIdx.x / chunks, step = gridDim.x / chunks;
  init_ctrl(ctrl, kFwdThreads / 32 - kProdWarps);
  if (warp < kProdWarps) {
    produce_tiles(x + 128 * c, (size_t)hidden * 4, t, first, step, smem, ctrl, lane, warp);
    return;
  }
  const float self_scale = 1.0f + __ldg(eps_ptr);
  const float4 w = ldg4(w_edge + 4 * lane + 128 * c), b = ldg4(b_edge + 4 * lane + 128 * c);
  const float4 nb = make_float4(-b.x, -b.y, -b.z, -b.w);
  float* hc = h + 128 * c + 4 * lane;
  int it = 0;
  for (int tile = first; tile < t.n_tiles; tile += step, ++it) {
    const int bf = it & 1;
    const unsigned char* rows = smem + (size_t)bf * t.buf_bytes;
    const unsigned char* blk = rows + t.rows_bytes;
    const unsigned char* rl = rows + 16 * lane;        // this lane's 4 columns inside a staged row
    mbar_wait(&ctrl->full[bf], (it >> 1) & 1);
    const int n_items = reinterpret_cast<const int*>(blk)[2];
    // groups are claimed in order (most edges first) from a shared counter: warps leave a tile within one short group
    // of each other whatever the degree distribution
    for (int grp = claim_item(&ctrl->counter[bf], lane); grp < n_items; grp = claim_item(&ctrl->counter[bf], lane)) {
      const unsigned char* gp = blk + 16 + 48 * grp;
      const int4 u0 = ldsi4(gp), u1 = ldsi4(gp + 16), u2 = ldsi4(gp + 32);
      float4 acc[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const unsigned char* p = blk + u0.w;
      p = fwd_class<1>(acc, p, u1.x & 0xffff, rl, w, nb);
      p = fwd_class<2>(acc, p, (unsigned)u1.x >> 16, rl, w, nb);
      p = fwd_class<3>(acc, p, u1.y & 0xffff, rl, w, nb);
      p = fwd_class<4>(acc, p, (unsigned)u1.y >> 16, rl, w, nb);
      p = fwd_class<5>(acc, p, u1.z & 0xffff, rl, w, nb);
      p = fwd_class<6>(acc, p, (unsigned)u1.z >> 16, rl, w, nb);
      p = fwd_class<7>(acc, p, u1.w, rl, w, nb);
      const int node[3] = {u0.x, u0.y, u0.z};
      const float deg[3] = {__int_as_float(u2.x), __int_as_float(u2.y), __int_as_float(u2.z)};
      const unsigned char* self = rl + u2.w;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (node[k] < 0) break;                      // rows of a group are packed to the front
        const float4 xi = lds4(self + k * kTileRowBytes);
        float4 o;
        o.x = fmaf(self_scale, xi.x, fmaf(deg[k], b.x, acc[k].x));
        o.y = fmaf(self_scale, xi.y, fmaf(deg[k], b.y, acc[k].y));
        o.z = fmaf(self_scale, xi.z, fmaf(deg[k], b.z, acc[k].z));
        o.w = fmaf(self_scale, xi.w, fmaf(deg[k], b.w, acc[k].w));
        st4(hc + (size_t)node[k] * hidden, o);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctrl->empty[bf]);      // this warp reads nothing more from the buffer
  }
}

// One (entry, row) pair of the backward: m = 1[x_s + a*w + b > 0] as 1.0 / 0.0, acc += g * m, s += a * m
// (s: the entry's sum_k a_k m_k; the weight gradient takes g * s once per entry)
__device__ __forceinline__ void masked_acc(float4& acc, float4& s, float4 g, float4 xs, float a, float4 w, float4 nb) {
  const float2 a2 = make_float2(a, a);
  const float2 z0 = __ffma2_rn(a2, make_float2(w.x, w.y), make_float2(xs.x, xs.y));
  const float2 z1 = __ffma2_rn(a2, make_float2(w.z, w.w), make_float2(xs.z, xs.w));
  const float2 m0 = make_float2(z0.x > nb.x ? 1.f : 0.f, z0.y > nb.y ? 1.f : 0.f);
  const float2 m1 = make_float2(z1.x > nb.z ? 1.f : 0.f, z1.y > nb.w ? 1.f : 0.f);
  const float2 c0 = __ffma2_rn(make_float2(g.x, g.y), m0, make_float2(acc.x, acc.y));
  const float2 c1 = __ffma2_rn(make_float2(g.z, g.w), m1, make_float2(acc.z, acc.w));
  const float2 t0 = __ffma2_rn(a2, m0, make_float2(s.x, s.y)), t1 = __ffma2_rn(a2, m1, make_float2(s.z, s.w));
  acc = make_float4(c0.x, c0.y, c1.x, c1.y);
  s = make_float4(t0.x, t0.y, t1.x, t1.y);
}

template <int MASK>
__device__ __forceinline__ void bwd_apply(float4 (&acc)[3], float4& dw, const float4 (&xs)[3], float4 g, int4 rec, float4 w, float4 nb) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MASK & 1) masked_acc(acc[0], s, g, xs[0], __int_as_float(rec.y), w, nb);
  if (MASK & 2) masked_acc(acc[1], s, g, xs[1], __int_as_float(rec.z), w, nb);
  if (MASK & 4) masked_acc(acc[2], s, g, xs[2], __int_as_float(rec.w), w, nb);
  const float2 d0 = __ffma2_rn(make_float2(g.x, g.y), make_float2(s.x, s.y), make_float2(dw.x, dw.y));
  const float2 d1 = __ffma2_rn(make_float2(g.z, g.w), make_float2(s.z, s.w), make_float2(dw.z, dw.w));
  dw = make_float4(d0.x, d0.y, d1.x, d1.y);
}

template <int MASK>
__device__ __forceinline__ const unsigned char* bwd_class(float4 (&acc)[3], float4& dw, const float4 (&xs)[3], const unsigned char* p,
                                                          int n, const unsigned char* rl, float4 w, float4 nb) {
  for (; n >= 2; n -= 2, p += 32) {
    const int4 r0 = ldsi4(p), r1 = ldsi4(p + 16);
    const float4 v0 = lds4(rl + r0.x), v1 = lds4(rl + r1.x);
    bwd_apply<MASK>(acc, dw, xs, v0, r0, w, nb);
    bwd_apply<MASK>(acc, dw, xs, v1, r1, w, nb);
  }
  if (n) {
    const int4 r0 = ldsi4(p);
    bwd_apply<MASK>(acc, dw, xs, lds4(rl + r0.x), r0, w, nb);
    p += 16;
  }
  return p;
}

// dynamic shared memory: two tile buffers | control block; the final reduction reuses the front of it
// (red[warps][2*128] + red_eps[warps]).  partials[blockIdx.x]: [3][H], zero outside the CTA's column chunk.
__global__ void __launch_bounds__(kBwdThreads, 1)
gine_aggr_bwd_tiled_kernel(const float* __restrict__ g, const float* __restrict__ x, const TilesP t,
                           const float* __restrict__ w_edge, const float* __restrict__ b_edge, const float* __restrict__ eps_ptr,
                           const float* __restrict__ addend, float* __restrict__ dx, float* __restrict__ partials, int hidden,
                           int chunks) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NW = kBwdThreads / 32;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + 2 * (size_t)t.buf_bytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x % chunks, first = blockIdx.x / chunks, step = gridDim.x / chunks;
  init_ctrl(ctrl, NW - kProdWarps);
  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f), db = make_float4(0.f, 0.f, 0.f, 0.f);
  double deps = 0.0;   // <g, x> cancels heavily over M*H products: float64 across rows
  if (warp < kProdWarps) {
    produce_tiles(g + 128 * c, (size_t)hidden * 4, t, first, step, smem, ctrl, lane, warp);
  } else {
    const float self_scale = 1.0f + __ldg(eps_ptr);
    const float4 w = ldg4(w_edge + 4 * lane + 128 * c), b = ldg4(b_edge + 4 * lane + 128 * c);
    const float4 nb = make_float4(-b.x, -b.y, -b.z, -b.w);
    const size_t col0 = 128 * c + 4 * lane;
    int it = 0;
    for (int tile = first; tile < t.n_tiles; tile += step, ++it) {
      const int bf = it & 1;
      const unsigned char* rows = smem + (size_t)bf * t.buf_bytes;
      const unsigned char* blk = rows + t.rows_bytes;
      const unsigned char* rl = rows + 16 * lane;
      mbar_wait(&ctrl->full[bf], (it >> 1) & 1);
      const int n_items = reinterpret_cast<const int*>(blk)[2];
      for (int grp = claim_item(&ctrl->counter[bf], lane); grp < n_items; grp = claim_item(&ctrl->counter[bf], lane)) {
        const unsigned char* gp = blk + 16 + 48 * grp;
        const int4 u0 = ldsi4(gp), u1 = ldsi4(gp + 16), u2 = ldsi4(gp + 32);
        const int node[3] = {u0.x, u0.y, u0.z};
        float4 xs[3], acc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          xs[k] = node[k] >= 0 ? ldg4(x + (size_t)node[k] * hidden + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
          acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const unsigned char* p = blk + u0.w;
        p = bwd_class<1>(acc, dw, xs, p, u1.x & 0xffff, rl, w, nb);
        p = bwd_class<2>(acc, dw, xs, p, (unsigned)u1.x >> 16, rl, w, nb);
        p = bwd_class<3>(acc, dw, xs, p, u1.y & 0xffff, rl, w, nb);
        p = bwd_class<4>(acc, dw, xs, p, (unsigned)u1.y >> 16, rl, w, nb);
        p = bwd_class<5>(acc, dw, xs, p, u1.z & 0xffff, rl, w, nb);
        p = bwd_class<6>(acc, dw, xs, p, (unsigned)u1.z >> 16, rl, w, nb);
        p = bwd_class<7>(acc, dw, xs, p, u1.w, rl, w, nb);
        const unsigned char* self = rl + u2.w;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (node[k] < 0) break;
          const float4 gj = lds4(self + k * kTileRowBytes);
          deps += (double)(gj.x * xs[k].x + gj.y * xs[k].y + gj.z * xs[k].z + gj.w * xs[k].w);
          // sum_e gm_e = acc: the bias gradient takes it once per row
          db.x += acc[k].x; db.y += acc[k].y; db.z += acc[k].z; db.w += acc[k].w;
          float4 o;
          o.x = fmaf(self_scale, gj.x, acc[k].x);
          o.y = fmaf(self_scale, gj.y, acc[k].y);
          o.z = fmaf(self_scale, gj.z, acc[k].z);
          o.w = fmaf(self_scale, gj.w, acc[k].w);
          if (addend != nullptr) {
            const float4 ad = ldg4(addend + (size_t)node[k] * hidden + col0);
            o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
          }
          st4(dx + (size_t)node[k] * hidden + col0, o);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->empty[bf]);
    }
  }
  __syncthreads();     // every tile has been read: the buffers become the reduction scratch
  float* red = reinterpret_cast<float*>(smem);
  float* red_eps = red + NW * 256;
  st4(red + warp * 256 + 4 * lane, dw);
  st4(red + warp * 256 + 128 + 4 * lane, db);
  deps = warp_sum(deps);
  if (lane == 0) red_eps[warp] = (float)deps;
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * 3 * hidden;
  for (int j = threadIdx.x; j < 2 * hidden; j += blockDim.x) {
    const int which = j / hidden, col = j - which * hidden;      // 0: d w_edge, 1: d b_edge
    float s = 0.f;
    if (col / 128 == c) {
#pragma unroll
      for (int r = kProdWarps; r < NW; ++r) s += red[r * 256 + which * 128 + (col & 127)];
    }
    out[j] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = kProdWarps; r < NW; ++r) s += red_eps[r];
    out[2 * hidden] = s;
  }
}

static int tiled_chunk_ctas(int n_tiles, int chunks) {
  const int cap = kNumSMs / chunks;
  return n_tiles < cap ? (n_tiles > 0 ? n_tiles : 1) : cap;
}

static int buffer_budget() { return ((kSmemPerCTA - kCtrlBytes) / 2) / 16 * 16; }

static size_t tiled_smem(const TilesP& t) { return 2 * (size_t)t.buf_bytes + kCtrlBytes; }

static int check_tiles(const rc_gine_tiles* t, int hidden, const char* who) {
  if (!t || t->n_tiles < 0 || !t->tile_stage_ptr || !t->tile_blk_ptr || !t->stage_id || !t->blocks)
    return fail(RC_ERR_ARG, "%s: null tile array", who);
  if (hidden < 128 || hidden % 128 || hidden > 512) return fail(RC_ERR_ARG, "%s: hidden=%d unsupported (128 | H, H <= 512)", who, hidden);
  if (t->row_bytes != kTileRowBytes) return fail(RC_ERR_ARG, "%s: tiles must be built with row_bytes = %d (one 128-column chunk), not %d", who, kTileRowBytes, t->row_bytes);
  if (t->max_staged < 0 || t->max_block_bytes < 0 || t->max_block_bytes % 16 ||
      (long long)t->max_staged * kTileRowBytes + t->max_block_bytes > buffer_budget())
    return fail(RC_ERR_ARG, "%s: a tile of %d rows + %d block bytes exceeds the %d bytes of one shared-memory buffer", who, t->max_staged,
                t->max_block_bytes, buffer_budget());
  if (!aligned16(t->blocks)) return fail(RC_ERR_ARG, "%s: blocks must be 16-byte aligned", who);
  return RC_OK;
}

static TilesP tiles_param(const rc_gine_tiles* t) {
  const int rows_bytes = t->max_staged * kTileRowBytes;
  return TilesP{t->n_tiles, rows_bytes, rows_bytes + t->max_block_bytes, t->tile_stage_ptr, t->tile_blk_ptr, t->stage_id, t->blocks};
}

}  // namespace rc

using namespace rc;

extern "C" int rc_gine_tiles_limits(int hidden, int* max_src, int* max_block_bytes) {
  if (!max_src || !max_block_bytes || hidden < 128 || hidden % 128 || hidden > 512) return fail(RC_ERR_ARG, "rc_gine_tiles_limits: hidden=%d unsupported", hidden);
  // tiles hold 128-column chunks of rows whatever the width; split of one buffer: ~25/32 staged rows, the rest for
  // the tile's group + entry records
  const int budget = buffer_budget();
  const int rows = (budget * 25 / 32) / kTileRowBytes;
  *max_src = rows;
  *max_block_bytes = (budget - rows * kTileRowBytes) / 16 * 16;
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
  const int chunks = hidden / 128;
  const int grid = tiled_chunk_ctas(tiles->n_tiles, chunks) * chunks;
  const TilesP t = tiles_param(tiles);
  const size_t smem = tiled_smem(t);
  cudaFuncSetAttribute(gine_aggr_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gine_aggr_fwd_tiled_kernel<<<grid, kFwdThreads, smem, s>>>(x, t, w_edge, b_edge, eps, h, hidden, chunks);
  return check_launch("gine_aggr_fwd_tiled_kernel");
}

extern "C" int rc_gine_aggr_bwd_tiled_nblocks(const rc_gine_tiles* tiles, int hidden) {
  if (!tiles || hidden < 128 || hidden % 128 || hidden > 512) return -1;
  const int chunks = hidden / 128;
  return tiled_chunk_ctas(tiles->n_tiles, chunks) * chunks;   // >= chunks: an empty graph still zeroes its partials
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
  const int chunks = hidden / 128;
  const int grid = rc_gine_aggr_bwd_tiled_nblocks(tiles, hidden);
  const TilesP t = tiles_param(tiles);
  size_t smem = tiled_smem(t);
  const size_t red = ((size_t)(kBwdThreads / 32) * 256 + kBwdThreads / 32) * sizeof(float);
  if (red > smem) smem = red;
  cudaFuncSetAttribute(gine_aggr_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gine_aggr_bwd_tiled_kernel<<<grid, kBwdThreads, smem, s>>>(g, x, t, w_edge, b_edge, eps, addend, dx, partials, hidden, chunks);
  return check_launch("gine_aggr_bwd_tiled_kernel");
}
