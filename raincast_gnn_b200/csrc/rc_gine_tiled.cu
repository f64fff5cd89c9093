// GINE aggregation over station tiles (rc_tiles.cu): the large-graph path.
//
// PyG GINEConv message + 'add' aggregation (call site models/gnn.py:27-29,39-44) gathers one 4*H-byte source row per
// edge.  At ~30 edges per station that is 30x the compulsory traffic, and the plain warp-per-row kernels of
// rc_gine.cu / rc_gine_wide.cu sit at the L2 -> SM throughput cap (profiles/r01_ncu_gine_aggr.txt).  Here one
// persistent CTA per SM walks its share of the tiles through a two-buffer shared-memory pipeline:
//   - eight producer warps (two per SM sub-partition) stage the rows a tile gathers (own rows + halo, ~2.5 rows per owned
//     row on the config-4 graph, exactly 1.0 on batched reference graphs) and the tile's block of group / entry
//     records with cp.async (global -> shared without registers, 16 bytes per lane) and signal an mbarrier when
//     the copies have landed; the next tile is in flight while the consumers work on the current one.  Forward tiles
//     are handed out from a global counter (dynamic schedule), backward tiles round robin (reproducible partial sums)
//   - consumer warps claim GROUPS of three neighbouring rows from a shared counter (groups with most edges first).
//     Each distinct source row of the group is read from shared memory once and used by every row of the group that
//     has an edge from it (2.2 edges per 512-byte read on config 4) - the shared-memory data pipe was the limiter of
//     the row-at-a-time version (profiles/r01_ncu_gine_tiled.txt).  A warp that finds a tile's groups all taken
//     releases the buffer (mbarrier) and moves on to the next tile: no CTA-wide barrier in the loop
//   - one lane per 4 columns of a 128-column chunk: conflict-free 128-bit LDS, coalesced 128-bit row stores.  Wider
//     rows are walked chunk by chunk (a CTA stays on one chunk, so tiles do not depend on the width)
//   - a group's entries are sorted by class (= which of its rows use the source); one straight-line loop per class,
//     one broadcast 16-byte record {staged offset, attr per row} per entry: no per-edge test, no global load
//   - the ReLU rides on the FMA: with operands scaled by 2^-64 (exact), fma.rn.sat clamps a*w + x_j + b to [0, 1],
//     i.e. computes relu for every message below 2^64.  Per (edge, 4 columns): 4 FFMA.SAT + 2 packed adds = 9 issue
//     cycles per sub-partition; FFMA2 + 4 FMNMX + 2 FADD2 measured 14.6 (tools/ubench/pipes.cu: FMNMX issues every
//     second cycle and does not overlap the FMA pipe on B200)
//   - deterministic: the summation order is fixed by the tiles (class by class, not the CSR slot order of the
//     untiled kernels - results agree to rounding, not bit for bit)
// Backward: the transpose tiles stage g rows; the ReLU mask is recomputed from the row's own x and the edge attr
// (nothing per-edge saved); d w_edge / d b_edge / d eps partials per CTA, same format as rc_gine_aggr_bwd.
#include "rc_common.cuh"

namespace rc {

constexpr int kSmemPerCTA = 227 * 1024;   // opt-in maximum of dynamic shared memory
constexpr int kTileRowBytes = 512;        // one 128-column chunk of a row
#ifndef RC_PROD_WARPS
#define RC_PROD_WARPS 8
#endif
#ifndef RC_FWD_THREADS
#define RC_FWD_THREADS 1024
#endif
#ifndef RC_BWD_THREADS
#define RC_BWD_THREADS 768
#endif
constexpr int kProdWarps = RC_PROD_WARPS;  // producer warps: warp w runs on sub-partition w % 4
static_assert(32 * kProdWarps * 512 >= 227 * 1024 / 2, "one staged row per producer lane");
constexpr int kFwdThreads = RC_FWD_THREADS;
constexpr int kBwdThreads = RC_BWD_THREADS;
constexpr int kCtrlBytes = 64;            // full[2], empty[2] mbarriers, two item counters, two claimed tile indices

// scale factors of the saturating-FMA ReLU (see relu_acc / masked_acc)
constexpr float kDown = 5.421010862427522e-20f;   // 2^-64
constexpr float kUp = 18446744073709551616.0f;    // 2^64

struct TilesP {
  int n_tiles;
  int rows_bytes;    // shared-memory bytes of one buffer's staged-row region (max_staged * 512)
  int buf_bytes;     // one buffer: staged rows + the largest tile block
  const int* __restrict__ tile_stage_ptr;
  const int* __restrict__ tile_blk_ptr;
  const int* __restrict__ stage_id;
  const int* __restrict__ blocks;
  int* sched;        // [8] zero between launches: next tile per column chunk [0..3], CTAs out of tiles per chunk [4..7]
};

struct Ctrl {
  unsigned long long full[2], empty[2];
  int counter[2];
  int claimed[2];    // dynamic schedule: the tile index the producers claimed for a buffer (>= n_tiles: no tile left)
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
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, int parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// waiting warps back off: a spinning warp takes issue slots from the warps it is waiting for
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
  while (!mbar_test(bar, parity)) __nanosleep(64);
}

// Producer warps: stage this CTA's tiles, two in flight.  `src` already points at the CTA's column chunk; rows are
// `row_stride` bytes apart in global memory and 512 bytes apart in shared memory.  Lane l of warp w owns staged row
// w + l * kProdWarps of every tile (32 * kProdWarps >= the most rows a buffer can hold): the tile's ranges and the
// lane's row id are fetched BEFORE waiting for the buffer, so that a released buffer is refilled without a trip to
// global memory first.
// DYN: tiles are claimed from a global counter (tile order = build order; every CTA keeps claiming until the counter
// runs past the last tile, the last CTA to get there zeroes the counters for the next launch) instead of the static
// round robin `first + k * step`.  Which CTA runs which tile then varies from launch to launch; the forward results do
// not depend on it, the backward keeps the static schedule so that its per-CTA partial sums stay reproducible.
template <bool DYN>
__device__ __forceinline__ void produce_tiles(const float* __restrict__ src, size_t row_stride, const TilesP& t, int first, int step,
                                              unsigned char* smem, Ctrl* ctrl, int lane, int warp, int chunk = 0) {
  const unsigned char* base = reinterpret_cast<const unsigned char*>(src);
  const int mine = warp + lane * kProdWarps;
  auto claim = [&](int it, int prev) -> int {
    if (!DYN) return it == 0 ? first : prev + step;
    if (warp == 0 && lane == 0) ctrl->claimed[it & 1] = atomicAdd(t.sched + chunk, 1);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kProdWarps) : "memory");      // producer warps only
    return ctrl->claimed[it & 1];
  };
  // ranges and row id of a tile are fetched while the copies of the tile before it are in flight
  int s0 = 0, nst = 0, b0 = 0, bunits = 0, my_id = 0;
  auto fetch = [&](int tile) {
    if (tile < t.n_tiles) {
      s0 = __ldg(t.tile_stage_ptr + tile); nst = __ldg(t.tile_stage_ptr + tile + 1) - s0;
      b0 = __ldg(t.tile_blk_ptr + tile); bunits = __ldg(t.tile_blk_ptr + tile + 1) - b0;
      my_id = mine < nst ? __ldg(t.stage_id + s0 + mine) : 0;
    }
  };
  int tile = claim(0, 0);
  fetch(tile);
  for (int it = 0;; ++it) {
    const int b = it & 1;
    if (!DYN && tile >= t.n_tiles) break;
    unsigned char* rows = smem + (size_t)b * t.buf_bytes;
    unsigned char* blk = rows + t.rows_bytes;
    const unsigned char* my_src = base + (size_t)my_id * row_stride;   // start of this lane's row
    if (it >= 2) mbar_wait(&ctrl->empty[b], ((it >> 1) & 1) ^ 1);      // every consumer warp is done with this buffer
    if (tile < t.n_tiles) {
      // the tile's block of group / entry records: contiguous, 16 bytes per thread per pass
      const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(t.blocks) + (size_t)b0 * 16;
      for (int u = warp * 32 + lane; u < bunits; u += 32 * kProdWarps)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(blk + 16 * u)), "l"(bsrc + 16 * (size_t)u) : "memory");
      // gathered rows: the row addresses are broadcast lane by lane; every lane moves 16 bytes of every row of the warp
      const int cnt = nst > warp ? (nst - warp + kProdWarps - 1) / kProdWarps : 0;
      unsigned char* dst0 = rows + (size_t)warp * kTileRowBytes + 16 * lane;
#ifndef RC_TILED_NO_GATHER      // (tools/build_variant.sh nogather -DRC_TILED_NO_GATHER: everything but the row gather, to time the SM-side floor)
      for (int k = 0; k < cnt; ++k) {
        const unsigned long long sp = __shfl_sync(0xffffffffu, (unsigned long long)my_src, k) + 16 * lane;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst0 + k * (kProdWarps * kTileRowBytes))), "l"(sp) : "memory");
      }
#else
      (void)cnt; (void)dst0; (void)my_src;
#endif
    }
    const bool last = tile >= t.n_tiles;      // (dynamic schedule only: tells the consumers to stop)
    if (warp == 0 && lane == 0) {
      ctrl->counter[b] = last ? (1 << 30) : 0;   // nobody claims from this buffer between its release and the arrivals below
      mbar_arrive(&ctrl->full[b]);               // release: orders the counter store before the consumers' claims
    }
    mbar_arrive_on_copies(&ctrl->full[b]);
    if (last) {
      if (warp == 0 && lane == 0 && atomicAdd(t.sched + 4 + chunk, 1) == step - 1) {
        t.sched[chunk] = 0;                      // every CTA of this chunk has made its last claim
        t.sched[4 + chunk] = 0;
      }
      break;
    }
    tile = claim(it + 1, tile);
    fetch(tile);
  }
}

__device__ __forceinline__ void init_ctrl(Ctrl* ctrl, int consumer_warps) {
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ctrl->full[b], 32 * kProdWarps + 1);     // every producer lane's copies + the counter reset
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

// The ReLU rides on the FMA: fma.rn.sat clamps its result to [0, 1], so with the operands scaled by kDown = 2^-64
// (exact: a power of two) sat(a * (w kDown) + (x_j + b) kDown) = kDown * relu(x_j + a w + b) for every message below
// 2^64.  Measured on B200 (tools/ubench/pipes.cu): 2 FFMA.SAT + FADD2 per two columns cost 4.7 issue cycles per
// sub-partition against 7.3 for FFMA2 + 2 FMNMX + FADD2 - FMNMX runs at half rate and does not overlap the FMA pipe.
// (A NaN input gives 0 here, NaN in the reference.)

__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float d;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// acc += sat(a * ws + vs)   (ws = w kDown, vs = (x_j + b) kDown)
__device__ __forceinline__ void relu_acc(float4& acc, float4 vs, float a, float4 ws) {
  const float2 z0 = make_float2(fma_sat(a, ws.x, vs.x), fma_sat(a, ws.y, vs.y));
  const float2 z1 = make_float2(fma_sat(a, ws.z, vs.z), fma_sat(a, ws.w, vs.w));
  const float2 s0 = __fadd2_rn(make_float2(acc.x, acc.y), z0), s1 = __fadd2_rn(make_float2(acc.z, acc.w), z1);
  acc = make_float4(s0.x, s0.y, s1.x, s1.y);
}

// The staged rows are raw x_j; every entry folds bias and scale in with four FMAs, shared by the rows of the group:
// vs = x_j * 2^-64 + b * 2^-64.  (Letting the producer warps rewrite the staged rows once per tile instead measured
// the same on config 4 and 3 % slower on batched reference graphs: two more shared-memory wavefronts per 128 bytes
// and a producer that has to wait for its own copies.)
__device__ __forceinline__ float4 prebias(float4 v, float4 bs) {
  return make_float4(fmaf(v.x, kDown, bs.x), fmaf(v.y, kDown, bs.y), fmaf(v.z, kDown, bs.z), fmaf(v.w, kDown, bs.w));
}

template <int MASK>
__device__ __forceinline__ void fwd_apply(float4 (&acc)[3], float4 v, int4 rec, float4 w) {
  if (MASK & 1) relu_acc(acc[0], v, __int_as_float(rec.y), w);
  if (MASK & 2) relu_acc(acc[1], v, __int_as_float(rec.z), w);
  if (MASK & 4) relu_acc(acc[2], v, __int_as_float(rec.w), w);
}

// the n entries of one class, two in flight; returns the next class's first entry
template <int MASK>
__device__ __forceinline__ const unsigned char* fwd_class(float4 (&acc)[3], const unsigned char* p, int n, const unsigned char* rl,
                                                          float4 w, float4 bs) {
#pragma unroll 1
  for (; n >= 2; n -= 2, p += 32) {
    const int4 r0 = ldsi4(p), r1 = ldsi4(p + 16);
    float4 v0 = lds4(rl + r0.x), v1 = lds4(rl + r1.x);
    v0 = prebias(v0, bs);
    v1 = prebias(v1, bs);
    fwd_apply<MASK>(acc, v0, r0, w);
    fwd_apply<MASK>(acc, v1, r1, w);
  }
  if (n) {
    const int4 r0 = ldsi4(p);
    float4 v0 = lds4(rl + r0.x);
    v0 = prebias(v0, bs);
    fwd_apply<MASK>(acc, v0, r0, w);
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
  const float4 w = ldg4(w_edge + 4 * lane + 128 * c), b = ldg4(b_edge + 4 * lane + 128 * c);
  const float4 bs = make_float4(b.x * kDown, b.y * kDown, b.z * kDown, b.w * kDown);
  if (warp < kProdWarps) {
    produce_tiles<true>(x + 128 * c, (size_t)hidden * 4, t, first, step, smem, ctrl, lane, warp, c);
    return;
  }
  const float self_scale = 1.0f + __ldg(eps_ptr);
  const float4 ws = make_float4(w.x * kDown, w.y * kDown, w.z * kDown, w.w * kDown);
  float* hc = h + 128 * c + 4 * lane;
  for (int it = 0;; ++it) {
    const int bf = it & 1;
    const unsigned char* rows = smem + (size_t)bf * t.buf_bytes;
    const unsigned char* blk = rows + t.rows_bytes;
    const unsigned char* rl = rows + 16 * lane;        // this lane's 4 columns inside a staged row
    mbar_wait(&ctrl->full[bf], (it >> 1) & 1);
    if (*reinterpret_cast<volatile int*>(&ctrl->counter[bf]) >= (1 << 30)) break;     // the producers found no tile left
    const int n_items = reinterpret_cast<const int*>(blk)[2];
    // groups are claimed in order (most edges first) from a shared counter: warps leave a tile within one short group
    // of each other whatever the degree distribution
    for (int grp = claim_item(&ctrl->counter[bf], lane); grp < n_items; grp = claim_item(&ctrl->counter[bf], lane)) {
      const unsigned char* gp = blk + 16 + 48 * grp;
      const int4 u0 = ldsi4(gp), u1 = ldsi4(gp + 16);
      const int u2w = *reinterpret_cast<const int*>(gp + 44);
      float4 acc[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const unsigned char* p = blk + u0.w;
      p = fwd_class<1>(acc, p, u1.x & 0xffff, rl, ws, bs);
      p = fwd_class<2>(acc, p, (unsigned)u1.x >> 16, rl, ws, bs);
      p = fwd_class<3>(acc, p, u1.y & 0xffff, rl, ws, bs);
      p = fwd_class<4>(acc, p, (unsigned)u1.y >> 16, rl, ws, bs);
      p = fwd_class<5>(acc, p, u1.z & 0xffff, rl, ws, bs);
      p = fwd_class<6>(acc, p, (unsigned)u1.z >> 16, rl, ws, bs);
      p = fwd_class<7>(acc, p, u1.w, rl, ws, bs);
      const int node[3] = {u0.x, u0.y, u0.z};
      const unsigned char* self = rl + u2w;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (node[k] < 0) break;                      // rows of a group are packed to the front
        const float4 xs = lds4(self + k * kTileRowBytes);       // x_i is staged too (own rows lead the tile)
        float4 o;
        o.x = fmaf(self_scale, xs.x, acc[k].x * kUp);
        o.y = fmaf(self_scale, xs.y, acc[k].y * kUp);
        o.z = fmaf(self_scale, xs.z, acc[k].z * kUp);
        o.w = fmaf(self_scale, xs.w, acc[k].w * kUp);
        st4(hc + (size_t)node[k] * hidden, o);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&ctrl->empty[bf]);      // this warp reads nothing more from the buffer
  }
}

// One (entry, row) pair of the backward.  The mask 1[x_s + a w + b > 0] comes out of a saturating FMA as 1.0 / 0.0:
// m = sat((a w + x_s + b) * 2^64) (operands pre-scaled; 0 for every z <= 0, 1 for every z >= 2^-64).  Then
// acc += g * m and s += a * m (s: the entry's sum_k a_k m_k; the weight gradient takes g * s once per entry).
// Scalar FMAs on purpose: FFMA2 issues at 2.6 cycles per sub-partition against 1.2 for FFMA (tools/ubench/pipes.cu).
__device__ __forceinline__ void masked_acc(float4& acc, float4& s, float4 g, float4 xb, float a, float4 wu) {
  const float m0 = fma_sat(a, wu.x, xb.x), m1 = fma_sat(a, wu.y, xb.y), m2 = fma_sat(a, wu.z, xb.z), m3 = fma_sat(a, wu.w, xb.w);
  acc.x = fmaf(g.x, m0, acc.x); acc.y = fmaf(g.y, m1, acc.y); acc.z = fmaf(g.z, m2, acc.z); acc.w = fmaf(g.w, m3, acc.w);
  s.x = fmaf(a, m0, s.x); s.y = fmaf(a, m1, s.y); s.z = fmaf(a, m2, s.z); s.w = fmaf(a, m3, s.w);
}

template <int MASK>
__device__ __forceinline__ void bwd_apply(float4 (&acc)[3], float4& dw, const float4 (&xb)[3], float4 g, int4 rec, float4 wu) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MASK & 1) masked_acc(acc[0], s, g, xb[0], __int_as_float(rec.y), wu);
  if (MASK & 2) masked_acc(acc[1], s, g, xb[1], __int_as_float(rec.z), wu);
  if (MASK & 4) masked_acc(acc[2], s, g, xb[2], __int_as_float(rec.w), wu);
  dw.x = fmaf(g.x, s.x, dw.x); dw.y = fmaf(g.y, s.y, dw.y); dw.z = fmaf(g.z, s.z, dw.z); dw.w = fmaf(g.w, s.w, dw.w);
}

template <int MASK>
__device__ __forceinline__ const unsigned char* bwd_class(float4 (&acc)[3], float4& dw, const float4 (&xs)[3], const unsigned char* p,
                                                          int n, const unsigned char* rl, float4 w) {
#pragma unroll 1
  for (; n >= 2; n -= 2, p += 32) {
    const int4 r0 = ldsi4(p), r1 = ldsi4(p + 16);
    const float4 v0 = lds4(rl + r0.x), v1 = lds4(rl + r1.x);
    bwd_apply<MASK>(acc, dw, xs, v0, r0, w);
    bwd_apply<MASK>(acc, dw, xs, v1, r1, w);
  }
  if (n) {
    const int4 r0 = ldsi4(p);
    bwd_apply<MASK>(acc, dw, xs, lds4(rl + r0.x), r0, w);
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
    produce_tiles<false>(g + 128 * c, (size_t)hidden * 4, t, first, step, smem, ctrl, lane, warp);
  } else {
    const float self_scale = 1.0f + __ldg(eps_ptr);
    const float4 w = ldg4(w_edge + 4 * lane + 128 * c), b = ldg4(b_edge + 4 * lane + 128 * c);
    const float4 wu = make_float4(w.x * kUp, w.y * kUp, w.z * kUp, w.w * kUp);
    const float4 bu = make_float4(b.x * kUp, b.y * kUp, b.z * kUp, b.w * kUp);
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
        const unsigned char* self = rl + u2.w;
        float4 xs[3], acc[3];       // xs: the row's own (x_s + b) * 2^64 (after its <g_s, x_s> went into d eps)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          xs[k] = node[k] >= 0 ? ldg4(x + (size_t)node[k] * hidden + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
          acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (node[k] >= 0) {
            const float4 gj = lds4(self + k * kTileRowBytes);
            deps += (double)(gj.x * xs[k].x + gj.y * xs[k].y + gj.z * xs[k].z + gj.w * xs[k].w);
          }
          xs[k] = make_float4(fmaf(xs[k].x, kUp, bu.x), fmaf(xs[k].y, kUp, bu.y), fmaf(xs[k].z, kUp, bu.z), fmaf(xs[k].w, kUp, bu.w));
        }
        const unsigned char* p = blk + u0.w;
        p = bwd_class<1>(acc, dw, xs, p, u1.x & 0xffff, rl, wu);
        p = bwd_class<2>(acc, dw, xs, p, (unsigned)u1.x >> 16, rl, wu);
        p = bwd_class<3>(acc, dw, xs, p, u1.y & 0xffff, rl, wu);
        p = bwd_class<4>(acc, dw, xs, p, (unsigned)u1.y >> 16, rl, wu);
        p = bwd_class<5>(acc, dw, xs, p, u1.z & 0xffff, rl, wu);
        p = bwd_class<6>(acc, dw, xs, p, (unsigned)u1.z >> 16, rl, wu);
        p = bwd_class<7>(acc, dw, xs, p, u1.w, rl, wu);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (node[k] < 0) break;
          const float4 gj = lds4(self + k * kTileRowBytes);
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
  if (!t || t->n_tiles < 0 || !t->tile_stage_ptr || !t->tile_blk_ptr || !t->stage_id || !t->blocks || !t->sched)
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
  return TilesP{t->n_tiles, rows_bytes, rows_bytes + t->max_block_bytes, t->tile_stage_ptr, t->tile_blk_ptr, t->stage_id, t->blocks,
                t->sched};
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
