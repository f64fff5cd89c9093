// Tensor-core (tcgen05 + TMEM) path of rc_gemm_run for the large Linear layers: every contraction of the node MLPs,
// dim_red, rho and the head with >= 16384 rows (BASELINE.json configs 4 and 5; models/gnn.py:21-26,51-62,113,123).
//
// fp32 parity (1e-5) on TF32 tensor cores: 3xTF32 - every operand is split v = hi + lo (hi = cvt.rna.tf32) and
//   D += a_hi b_hi + a_hi b_lo + a_lo b_hi          (fp32 accumulation in TMEM, error ~2^-21 per product).
//
// The problem is transposed so that the epilogue is free of cross-thread traffic:
//   rows kernel   (forward / backward-data):  D^T[out channel, row] = Wp[out channel, k] . X'[row, k]
//       (statistics tiles - BatchNorm forward / backward partial sums - are the 64-row halves of a 128-row tile)
//       A = packed weight block (128 channels x 32 k, hi | lo), pre-split once per call by tc_pack_kernel and brought
//           into shared memory by ONE bulk copy (TMA engine, cp.async.bulk) per stage
//       B = 128 activation rows x 32 k, loaded by the producer warps (128-bit loads), operand prologue applied
//           (BatchNorm+ReLU, bit mask, BatchNorm backward), split and stored K-major
//       D in TMEM: lane = output channel, column = row.  The thread that owns a channel walks the rows: bias, ReLU,
//       residual, BatchNorm tile statistics and the BatchNorm backward column sums are per-thread serial loops, the
//       ReLU bit mask is one ballot per row, and every global access is coalesced across the 32 channels of a warp.
//   wgrad kernel  (weight gradient):  D[i, j] = sum_r A'[r, i] . B'[r, j]  over >= 16384 samples r
//       both operands are stored sample-major; the producers transpose on the fly (coalesced 32-bit loads of 4
//       consecutive samples -> one 16-byte K-major chunk), the reduction is split over the CTAs and finished by
//       rc_reduce_segments in float64; the bias gradient (column sums of A') falls out of the producer registers.
// Operand tiles: K-major, no swizzle, 8-row groups of 1024 bytes (LBO = 128, SBO = 1024), see rc_umma.cuh.
// Pipeline: 3 stages of 64 KB (A hi|lo, B hi|lo), mbarriers full/empty, one MMA thread, two TMEM accumulator sets in the
// rows kernel so that the epilogue of tile t overlaps the MMAs of tile t+1; persistent CTAs (one per SM).
//
// Accumulation chains.  The tensor core adds every MMA's 8-term dot product into the fp32 accumulator with truncation
// (round toward zero): a chain of L accumulating MMAs shrinks the result by ~L * 3e-8 relative (measured: 1.2e-5 on a
// weight gradient over 1900 samples per CTA; 1.0 - 1.5e-5 on the first layers' gradients of the config-4 model, whose
// backward chains ten such GEMMs, with chains of 48).  So no accumulator takes more than 2 stages (24 MMAs, ~4e-7):
// the rows kernel gives every group of 2 k-blocks its own accumulator and the epilogue adds the groups (K <= 256; wider
// layers - config 5, a bf16 / 1e-2 configuration - use groups of 4), the wgrad kernel flushes its accumulator into
// round-to-nearest register sums every 2 sample blocks (two accumulators, so the flush of one overlaps the MMAs into the
// other).
#include <stdlib.h>

#include "rc_gemm_tile.cuh"
#include "rc_umma.cuh"

namespace rc {

// rows kernel: 8 epilogue warps (two per TMEM lane quarter, 64 of the tile's 128 rows each), 8 producer warps, the MMA warp
// (17 warps: 96 registers per thread)
constexpr int kTcEpiWarps = 8, kTcRowsProdWarps = 8;
constexpr int kTcThreads = 32 * (kTcEpiWarps + kTcRowsProdWarps + 1);
constexpr int kTcProdWarps = 8;                                      // wgrad kernel
constexpr int kTcStages = 3;
constexpr int kTcStagesAtm = 6;                                      // rows kernel with the weight tile in TMEM: a stage holds B only
constexpr int kTcBlkFloats = 128 * 32;                               // one hi (or lo) block: 128 rows x 32 k
constexpr int kTcBlkBytes = kTcBlkFloats * 4;
constexpr int kTcStageBytes = 4 * kTcBlkBytes;                       // A hi | A lo | B hi | B lo
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 256 + 1024 + 128;  // + barriers, + column sums, + alignment slack
constexpr int kTcChain = 2;                                          // stages accumulated into one TMEM accumulator (wgrad; rows: TcRowsP::chain)
constexpr int kTcWgEpiWarps = 8;                                     // wgrad: two epilogue warps per TMEM lane quarter
constexpr int kTcWgThreads = 32 * (kTcWgEpiWarps + kTcProdWarps + 1);
constexpr uint32_t kTcLBO = 128, kTcSBO = 1024;

// float index of element (row, kk) inside a 128 x 32 block (K-major, 8-row groups)
__host__ __device__ inline int tc_blk_index(int row, int kk) { return (row >> 3) * 256 + (kk >> 2) * 32 + (row & 7) * 4 + (kk & 3); }

struct TcPackP {
  const float* b; int ldb; int b_layout;
  const float* b2; int ldb2;
  int n, k, k2, kb1, kblocks, n_tiles;
  float* out;                               // [n_tiles][kblocks][hi | lo][128 x 32]
};

// weight (B operand of rc_gemm: stored [j][r] or [r][j]) -> packed, pre-split K-major blocks, zero padded
__global__ void __launch_bounds__(256) tc_pack_kernel(const TcPackP p) {
  pdl_entry();
  const long long total = (long long)p.n_tiles * p.kblocks * kTcBlkFloats;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int blk = (int)(e / kTcBlkFloats), w = (int)(e % kTcBlkFloats);
    const int nt = blk / p.kblocks, kb = blk % p.kblocks;
    const int row = w >> 5, kk = w & 31;                  // consecutive threads: consecutive k of one output channel
    const int j = nt * 128 + row;
    float v = 0.f;
    if (j < p.n) {
      if (kb < p.kb1) {
        const int r = kb * 32 + kk;
        if (r < p.k) v = __ldg(p.b_layout == RC_B_COL ? p.b + (size_t)j * p.ldb + r : p.b + (size_t)r * p.ldb + j);
      } else {
        const int r = (kb - p.kb1) * 32 + kk;
        if (r < p.k2) v = __ldg(p.b_layout == RC_B_COL ? p.b2 + (size_t)j * p.ldb2 + r : p.b2 + (size_t)r * p.ldb2 + j);
      }
    }
    const float hi = to_tf32(v);
    float* blk_out = p.out + (size_t)blk * 2 * kTcBlkFloats;
    const int idx = tc_blk_index(row, kk);
    blk_out[idx] = hi;
    blk_out[kTcBlkFloats + idx] = v - hi;
  }
}

struct TcSmem {
  unsigned char* stages;
  uint32_t full, empty, tmem_full, tmem_empty, w_ready, tmem_slot;   // shared-memory addresses of the barrier arrays / slot
  uint32_t* tmem_slot_ptr;
};

// the operand stages take 3 x 64 KB in both pipelines (A|B stages, or 6 stages of B alone); barriers follow
__device__ __forceinline__ TcSmem tc_carve(unsigned char* raw) {
  TcSmem s;
  unsigned char* base = raw + ((128 - (smem_u32(raw) & 127)) & 127);
  s.stages = base;
  unsigned char* bars = base + kTcStages * kTcStageBytes;
  s.full = smem_u32(bars);                  // up to 6 x 8
  s.empty = s.full + 48;                    // up to 6 x 8
  s.tmem_full = s.empty + 48;               // 2 x 8
  s.tmem_empty = s.tmem_full + 16;          // 2 x 8
  s.w_ready = s.tmem_empty + 16;            // 8
  s.tmem_slot = s.w_ready + 8;
  s.tmem_slot_ptr = reinterpret_cast<uint32_t*>(bars + 136);
  return s;
}

// 12 MMAs of one stage: 4 k-steps of 8, three TF32 products each
__device__ __forceinline__ void tc_issue_stage(uint32_t stage_addr, uint32_t tmem_d, uint32_t idesc, bool first) {
  const UmmaDesc ah = umma_desc2(stage_addr, kTcLBO, kTcSBO);           // A hi; A lo, B hi, B lo follow at 16 KB steps
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t off = ks * 2 * kTcLBO;
    umma_tf32(tmem_d, ah.at(kTcBlkBytes + off), ah.at(2 * kTcBlkBytes + off), idesc, (first && ks == 0) ? 0u : 1u);
    umma_tf32(tmem_d, ah.at(off), ah.at(3 * kTcBlkBytes + off), idesc, 1u);
    umma_tf32(tmem_d, ah.at(off), ah.at(2 * kTcBlkBytes + off), idesc, 1u);
  }
}

// the same with the A operand (the weight tile, hi at a_tmem, lo 128 columns further) read from TMEM: the tensor core
// fetches only B from shared memory, which halves the shared-memory traffic of a stage
__device__ __forceinline__ void tc_issue_stage_atm(uint32_t b_addr, uint32_t a_tmem, uint32_t tmem_d, uint32_t idesc, bool first) {
  const UmmaDesc bh = umma_desc2(b_addr, kTcLBO, kTcSBO);               // B hi; B lo follows 16 KB later
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t off = ks * 2 * kTcLBO;
    umma_tf32_ts(tmem_d, a_tmem + 128 + ks * 8, bh.at(off), idesc, (first && ks == 0) ? 0u : 1u);
    umma_tf32_ts(tmem_d, a_tmem + ks * 8, bh.at(kTcBlkBytes + off), idesc, 1u);
    umma_tf32_ts(tmem_d, a_tmem + ks * 8, bh.at(off), idesc, 1u);
  }
}

// ------------------------------------------------------------------------------------------------ rows kernel
struct TcRowsP {
  rc_gemm g;
  const float* wpack;
  int kb1, kblocks, n_tiles, row_tiles;
  int a_vec, a2_vec, a_out_vec;
  int chain;                 // k-blocks per accumulator: 2, or 4 when K > 256 (at most four accumulators per tile)
  long long* trace;          // debug (rc_debug_tc_trace): CTA 0 records [role][tile][begin, end] clocks; NULL in production
  int dbg;                   // debug (RC_TC_DBG): 2 = one accumulator set (no MMA / epilogue overlap)
};
constexpr int kTcTraceTiles = 16;
__device__ __forceinline__ void tc_trace(long long* trace, int role, uint32_t tile_no, int which) {
  if (trace != nullptr && blockIdx.x == 0 && tile_no < kTcTraceTiles && (threadIdx.x & 31) == 0)
    trace[(role * kTcTraceTiles + tile_no) * 2 + which] = clock64();
}

// Producer role of the rows kernel.  A thread converts eight 16-byte chunks per stage (8 rows x 4 chunks per warp and
// load instruction: 64 contiguous bytes per row, conflict-free 16-byte shared-memory stores).  The raw global loads of
// stage n+2 are issued before stage n is converted, so two stages of loads are in flight per SM while it converts.
// The common case - a full 128-row tile, 16-byte aligned rows, a k-block inside the matrix - is straight-line code
// (the role is issue-bound: ~100 instructions per stage and thread there, several hundred on the guarded path).
struct TcChunk {
  float4 x, aux;        // activation values / BatchNorm-backward second operand
  uint32_t bits;        // ReLU mask word of (row, k-block)
  uint32_t valid;       // bit e: element e is inside the matrix (padding stays zero through the operand prologue); 16: no prologue
};

// walks this CTA's (tile, k-block) sequence
struct TcCursor {
  int tile, kb, row0, nt, tl;
  __device__ __forceinline__ void start(const TcRowsP& p) { tile = blockIdx.x; kb = 0; tl = 0; set(p); }
  __device__ __forceinline__ void set(const TcRowsP& p) {
    const int rt = tile / p.n_tiles;
    row0 = rt * 128;
    nt = tile - rt * p.n_tiles;
  }
  __device__ __forceinline__ void next(const TcRowsP& p) {
    if (++kb == p.kblocks) { kb = 0; tile += gridDim.x; ++tl; set(p); }
  }
};

template <int OP, bool ATM>
__device__ __forceinline__ void tc_rows_producer(const TcRowsP& p, const TcSmem& sm, int pw, int lane, int n_tiles_total) {
  constexpr uint32_t kStages = ATM ? kTcStagesAtm : kTcStages;
  constexpr int kStageBytes = ATM ? 2 * kTcBlkBytes : kTcStageBytes;
  const rc_gemm& g = p.g;
  const int i8 = lane & 7, j4 = lane >> 3;                // row inside its 8-row group, chunk inside its half
  const int my_tiles = (int)blockIdx.x < n_tiles_total ? (n_tiles_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_tiles * p.kblocks;
  constexpr int kCh = 32 / kTcRowsProdWarps;             // chunks per thread and stage (32 units of 8 rows x 4 chunks per stage)
  // chunk `it` of this thread: unit u = it * 8 + pw -> row (u >> 1) * 8 + i8 of the tile, k offset 4 * ((u & 1) * 4 + j4) in the block
  const int r0 = (pw >> 1) * 8 + i8, kofs_t = 4 * ((pw & 1) * 4 + j4);
  const int s0 = (pw >> 1) * 256 + (kofs_t >> 2) * 32 + i8 * 4;
  auto rofs_f = [&](int it) { return r0 + it * (kTcRowsProdWarps / 2) * 8; };
  auto sidx_f = [&](int it) { return s0 + it * (kTcRowsProdWarps / 2) * 256; };

  auto load = [&](const TcCursor& cu, TcChunk (&ch)[kCh]) {
    const bool seg2 = cu.kb >= p.kb1;
    const int k0 = (seg2 ? cu.kb - p.kb1 : cu.kb) * 32;
    const float* abase = seg2 ? g.a2 : g.a.ptr;
    const int lda = seg2 ? g.lda2 : g.a.ld;
    const int klen = seg2 ? g.k2 : g.k;
    const bool vec = seg2 ? p.a2_vec : p.a_vec;
    if (vec && cu.row0 + 128 <= g.m && k0 + 32 <= klen) {          // whole block inside the matrix, 128-bit loads
      const float* tile_base = abase + (size_t)cu.row0 * lda + k0;
      const float* aux_base = (OP == RC_OP_AFFINE2 && !seg2) ? g.a.aux + (size_t)cu.row0 * g.a.ld_aux + k0 : nullptr;
      const uint32_t* bits_base = (OP == RC_OP_BITMASK && !seg2) ? g.a.bits + (size_t)cu.row0 * g.a.ld_bits + cu.kb : nullptr;
#pragma unroll
      for (int it = 0; it < kCh; ++it) {
        ch[it].x = ldg4(tile_base + rofs_f(it) * lda + kofs_t);
        ch[it].valid = seg2 ? 31u : 15u;
        if (OP == RC_OP_AFFINE2 && !seg2) ch[it].aux = ldg4(aux_base + rofs_f(it) * g.a.ld_aux + kofs_t);
        if (OP == RC_OP_BITMASK && !seg2) ch[it].bits = __ldg(bits_base + rofs_f(it) * g.a.ld_bits);
      }
      return;
    }
#pragma unroll
    for (int it = 0; it < kCh; ++it) {
      const int row = cu.row0 + rofs_f(it), k = k0 + kofs_t;
      TcChunk& c = ch[it];
      c.x = make_float4(0.f, 0.f, 0.f, 0.f);
      c.aux = c.x;
      c.bits = 0u;
      c.valid = 0u;
      if (row < g.m && k < klen) {
        const float* src = abase + (size_t)row * lda + k;
        float t[4] = {0.f, 0.f, 0.f, 0.f}, a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (k + e < klen) {
            t[e] = __ldg(src + e);
            c.valid |= 1u << e;
            if (OP == RC_OP_AFFINE2 && !seg2) a[e] = __ldg(g.a.aux + (size_t)row * g.a.ld_aux + k + e);
          }
        c.x = make_float4(t[0], t[1], t[2], t[3]);
        c.aux = make_float4(a[0], a[1], a[2], a[3]);
        if (OP == RC_OP_BITMASK && !seg2) c.bits = __ldg(g.a.bits + (size_t)row * g.a.ld_bits + (k >> 5));
        if (seg2) c.valid |= 16u;
      }
    }
  };

  auto convert = [&](const TcCursor& cu, uint32_t n, const TcChunk (&ch)[kCh]) {
    const uint32_t s = n % kStages, ph = (n / kStages) & 1;
    const int k0 = cu.kb * 32;                            // (the prologue only applies to segment 1, where this is the column)
    if (pw == 0 && cu.kb == 0) tc_trace(p.trace, 0, cu.tl, 0);
    if (ATM && n == 4) mbar_wait(sm.w_ready, 0);          // stages 4 and 5 stage the weight tile on its way to TMEM
    mbar_wait(sm.empty + 8 * s, ph ^ 1);
    unsigned char* st = sm.stages + (size_t)s * kStageBytes;
    if (!ATM && pw == 0 && lane == 0) {
      mbar_arrive_expect_tx(sm.full + 8 * s, 2 * kTcBlkBytes);
      bulk_g2s(smem_u32(st), p.wpack + ((size_t)cu.nt * p.kblocks + cu.kb) * 2 * kTcBlkFloats, 2 * kTcBlkBytes, sm.full + 8 * s);
    }
    float* bh = reinterpret_cast<float*>(st + (ATM ? 0 : 2 * kTcBlkBytes));
    float* bl = bh + kTcBlkFloats;
#pragma unroll
    for (int it = 0; it < kCh; ++it) {
      const TcChunk& q = ch[it];
      float v[4] = {q.x.x, q.x.y, q.x.z, q.x.w};
      if (OP != RC_OP_NONE && (q.valid & 31u) == 15u) {   // whole chunk inside the matrix: per-column vectors as 128-bit loads
        const int col = k0 + kofs_t;
        if (OP == RC_OP_BN_RELU) {
          const float4 m4 = ldg4(g.a.p0 + col), r4 = ldg4(g.a.p1 + col), g4 = ldg4(g.a.p2 + col), b4 = ldg4(g.a.p3 + col);
          v[0] = fmaxf((v[0] - m4.x) * r4.x * g4.x + b4.x, 0.f);
          v[1] = fmaxf((v[1] - m4.y) * r4.y * g4.y + b4.y, 0.f);
          v[2] = fmaxf((v[2] - m4.z) * r4.z * g4.z + b4.z, 0.f);
          v[3] = fmaxf((v[3] - m4.w) * r4.w * g4.w + b4.w, 0.f);
        }
        if (OP == RC_OP_BITMASK) {
          const uint32_t w = q.bits >> (col & 31);
          v[0] = (w & 1u) ? v[0] : 0.f; v[1] = (w & 2u) ? v[1] : 0.f; v[2] = (w & 4u) ? v[2] : 0.f; v[3] = (w & 8u) ? v[3] : 0.f;
        }
        if (OP == RC_OP_AFFINE2) {
          const float4 c0 = ldg4(g.a.p0 + col), c1 = ldg4(g.a.p1 + col), c2 = ldg4(g.a.p2 + col), m4 = ldg4(g.a.p3 + col);
          v[0] = fmaf(c0.x, v[0], fmaf(c1.x, q.aux.x - m4.x, c2.x));
          v[1] = fmaf(c0.y, v[1], fmaf(c1.y, q.aux.y - m4.y, c2.y));
          v[2] = fmaf(c0.z, v[2], fmaf(c1.z, q.aux.z - m4.z, c2.z));
          v[3] = fmaf(c0.w, v[3], fmaf(c1.w, q.aux.w - m4.w, c2.w));
        }
      } else if (OP != RC_OP_NONE && q.valid != 0u && !(q.valid & 16u)) {
        const float ax[4] = {q.aux.x, q.aux.y, q.aux.z, q.aux.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (q.valid & (1u << e)) {
            const int col = k0 + kofs_t + e;
            if (OP == RC_OP_BN_RELU)
              v[e] = fmaxf((v[e] - __ldg(g.a.p0 + col)) * __ldg(g.a.p1 + col) * __ldg(g.a.p2 + col) + __ldg(g.a.p3 + col), 0.f);
            if (OP == RC_OP_BITMASK) v[e] = ((q.bits >> (col & 31)) & 1u) ? v[e] : 0.f;
            if (OP == RC_OP_AFFINE2)
              v[e] = fmaf(__ldg(g.a.p0 + col), v[e], fmaf(__ldg(g.a.p1 + col), ax[e] - __ldg(g.a.p3 + col), __ldg(g.a.p2 + col)));
          }
        }
      }
      float4 hi, lo;
      split_tf32_rn(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
      st4(bh + sidx_f(it), hi);
      st4(bl + sidx_f(it), lo);
      // the operand after its prologue, once (first column tile): the weight-gradient GEMM of the same layer then reads it
      // as a plain operand instead of re-deriving it element by element (bit mask, BatchNorm backward, ...)
      if (OP != RC_OP_NONE && g.a_out != nullptr && cu.nt == 0 && cu.kb < p.kb1) {
        const int row = cu.row0 + rofs_f(it), col = k0 + kofs_t;
        if (row < g.m) {
          float* dst = g.a_out + (size_t)row * g.ld_a_out + col;
          if (col + 3 < g.k && p.a_out_vec) {
            st4(dst, make_float4(v[0], v[1], v[2], v[3]));
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < g.k) dst[e] = v[e];
          }
        }
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.full + 8 * s);
    if (pw == 0 && cu.kb == p.kblocks - 1) tc_trace(p.trace, 0, cu.tl, 1);
  };

  TcCursor lc, cc;                                        // load cursor (kDepth stages ahead), convert cursor
  lc.start(p);
  cc.start(p);
  constexpr int kDepth = OP == RC_OP_AFFINE2 ? 2 : 3;     // stages of raw loads in flight (registers: 8 floats per chunk with AFFINE2)
  TcChunk c0[kCh], c1[kCh], c2[kCh];
  if (total > 0) { load(lc, c0); lc.next(p); }
  if (total > 1) { load(lc, c1); lc.next(p); }
  if (kDepth > 2 && total > 2) { load(lc, c2); lc.next(p); }
  for (int n = 0; n < total; n += kDepth) {
    convert(cc, n, c0);
    cc.next(p);
    if (n + kDepth < total) { load(lc, c0); lc.next(p); }
    if (n + 1 < total) {
      convert(cc, n + 1, c1);
      cc.next(p);
      if (n + 1 + kDepth < total) { load(lc, c1); lc.next(p); }
    }
    if (kDepth > 2 && n + 2 < total) {
      convert(cc, n + 2, c2);
      cc.next(p);
      if (n + 2 + kDepth < total) { load(lc, c2); lc.next(p); }
    }
  }
}

// Epilogue of one 128 x 128 tile of the rows kernel.  The thread owns output channel `col` (its TMEM lane) and walks the
// tile's rows, 32 at a time; EPI / BITS are compile-time so that the per-row work is a handful of instructions.
template <int EPI, bool BITS>
__device__ __forceinline__ void tc_rows_epilogue(const rc_gemm& g, uint32_t taddr, int groups, int rt, int nt, int warp, int lane) {
  const int half = warp >> 2, quarter = warp & 3;         // which 64 rows of the tile / which 32 channels
  const int row0 = rt * 128 + half * 64, col = nt * 128 + quarter * 32 + lane;
  const bool cok = col < g.n;
  const int valid_rows = min(64, g.m - row0);              // <= 0: this half tile lies outside the matrix
  taddr += half * 64;
  const float bias = (g.bias != nullptr && cok) ? g.bias_scale * __ldg(g.bias + col) : 0.f;
  // 16 rows of this channel: the sum of the tile's accumulators (round to nearest)
  auto load_cols = [&](int cb, uint32_t (&r)[16]) {
    tmem_ld16(taddr + cb * 16, r);
    tmem_ld_wait();
    for (int gi = 1; gi < groups; ++gi) {
      uint32_t q[16];
      tmem_ld16(taddr + gi * 128 + cb * 16, q);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(q[i]));
    }
  };
  float s0 = 0.f, s1 = 0.f;
  float e_mean = 0.f, e_rstd = 0.f, e_gamma = 0.f, e_beta = 0.f;
  if (EPI == RC_EPI_BN_RELU_BWD && cok) {
    e_mean = __ldg(g.e_p0 + col); e_rstd = __ldg(g.e_p1 + col); e_gamma = __ldg(g.e_p2 + col); e_beta = __ldg(g.e_p3 + col);
  }
  if (EPI == RC_EPI_BN_STATS && valid_rows > 0) {
    // tile mean first (the centred second moment needs it); the accumulator is read again below
    float sum = 0.f;
#pragma unroll 1
    for (int cb = 0; cb * 16 < valid_rows; ++cb) {
      uint32_t r[16];
      load_cols(cb, r);
      const int nrows = valid_rows - cb * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nrows) sum += __uint_as_float(r[i]) + bias;
    }
    s0 = sum / (float)valid_rows;
  }
  constexpr bool kRes = EPI == RC_EPI_RELU_RES || EPI == RC_EPI_ADD_RES;
  constexpr bool kAux = EPI == RC_EPI_MASK_POS || EPI == RC_EPI_BN_RELU_BWD;
  const int ld_x = kRes ? g.ld_res : g.ld_e_aux;
  const float* xp = (kRes ? g.res : g.e_aux) + (size_t)row0 * ld_x + col;      // only dereferenced when kRes || kAux
  float* dp = g.d + (size_t)row0 * g.ldd + col;
  uint32_t* bp = g.bits_out + (size_t)row0 * g.ld_bits_out + nt * 4 + quarter;
  const bool wok = nt * 128 + 32 * quarter < g.n;
  const size_t ldd = g.ldd, ldx = ld_x, ldb = g.ld_bits_out;
  auto row_op = [&](float acc, float xi, uint32_t* bpi) -> float {
    float v = acc + bias;
    const bool pos = v > 0.f;
    if (EPI == RC_EPI_RELU) v = fmaxf(v, 0.f);
    if (EPI == RC_EPI_RELU_RES) v = xi + fmaxf(v, 0.f);
    if (EPI == RC_EPI_ADD_RES) v += xi;
    if (EPI == RC_EPI_MASK_POS) v = xi > 0.f ? v : 0.f;
    if (EPI == RC_EPI_BN_STATS) { const float d = v - s0; s1 = fmaf(d, d, s1); }
    if (EPI == RC_EPI_BN_RELU_BWD) {
      const float hat = (xi - e_mean) * e_rstd;
      v = (cok && fmaf(e_gamma, hat, e_beta) > 0.f) ? v : 0.f;
      s0 += v;
      s1 = fmaf(v, hat, s1);
    }
    if (BITS) {
      const unsigned word = __ballot_sync(0xffffffffu, cok && pos);      // lane <-> channel 32*quarter + lane of the tile
      if (lane == 0 && wok) *bpi = word;
    }
    return v;
  };
  // (Tried: passing each 16-row block through the warp's shared memory so that global accesses are 16-byte row pieces,
  //  four 512-byte instructions instead of sixteen 128-byte ones - no gain, 45 vs 41 us at config 4: the epilogue's
  //  stores are bound by the bytes in flight towards L2 / HBM, not by their instruction count.)
#pragma unroll 1
  for (int cb = 0; cb * 16 < valid_rows; ++cb) {
    const int nrows = valid_rows - cb * 16;               // warp-uniform
    const float* xpi = xp + (size_t)cb * 16 * ldx;
    float* dpi = dp + (size_t)cb * 16 * ldd;
    uint32_t* bpi = bp + (size_t)cb * 16 * ldb;
    float x[16];
    uint32_t r[16];
    if (nrows >= 16) {                                    // whole block of 16 rows: straight-line code
      if (kRes || kAux) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = cok ? __ldg(xpi + i * ldx) : 0.f;
      }
      load_cols(cb, r);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = row_op(__uint_as_float(r[i]), (kRes || kAux) ? x[i] : 0.f, bpi + i * ldb);
        if (cok) dpi[i * ldd] = v;
      }
    } else {                                              // the matrix ends inside this block
      if (kRes || kAux) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = (cok && i < nrows) ? __ldg(xpi + i * ldx) : 0.f;
      }
      load_cols(cb, r);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < nrows) {
          const float v = row_op(__uint_as_float(r[i]), (kRes || kAux) ? x[i] : 0.f, bpi + i * ldb);
          if (cok) dpi[i * ldd] = v;
        }
    }
  }
  if ((EPI == RC_EPI_BN_STATS || EPI == RC_EPI_BN_RELU_BWD) && cok && valid_rows > 0) {
    float* stats = g.stats + (size_t)(rt * 2 + half) * 2 * g.n;
    stats[col] = s0;
    stats[g.n + col] = s1;
  }
}

template <bool ATM>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_rows_kernel(const TcRowsP p) {
  pdl_entry();
  extern __shared__ unsigned char smem_raw[];
  const TcSmem sm = tc_carve(smem_raw);
  const rc_gemm& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles_total = p.row_tiles * p.n_tiles;
  const int groups = ceil_div(p.kblocks, p.chain);         // accumulators per tile (<= 4; 1 with the weight tile in TMEM)
  const uint32_t nbuf = (groups <= 2 && !(p.dbg & 2)) ? 2u : 1u;   // accumulator sets: tile t+1's MMAs overlap tile t's epilogue
  constexpr uint32_t kStages = ATM ? kTcStagesAtm : kTcStages;
  constexpr int kStageBytes = ATM ? 2 * kTcBlkBytes : kTcStageBytes;

  if (warp == 0) tmem_alloc(sm.tmem_slot, 512);
  if (tid == 32) {
    for (uint32_t s = 0; s < kStages; ++s) {
      mbar_init(sm.full + 8 * s, ATM ? kTcRowsProdWarps : kTcRowsProdWarps + 1);   // producer warps (+ the bulk copy's expect_tx arrival)
      mbar_init(sm.empty + 8 * s, 1);                                      // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(sm.tmem_full + 8 * a, 1);
      mbar_init(sm.tmem_empty + 8 * a, kTcEpiWarps);
    }
    mbar_init(sm.w_ready, kTcEpiWarps);
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot_ptr;
  const uint32_t w_tmem = tmem_base + 256;                 // ATM: weight tile hi in columns [256, 384), lo in [384, 512)

  if (warp >= kTcEpiWarps && warp < kTcEpiWarps + kTcRowsProdWarps) {
    // ================================================================= producers: activation rows -> B blocks
    const int pw = warp - kTcEpiWarps;                    // 0..3
    switch (g.a.op) {
      case RC_OP_BN_RELU: tc_rows_producer<RC_OP_BN_RELU, ATM>(p, sm, pw, lane, n_tiles_total); break;
      case RC_OP_BITMASK: tc_rows_producer<RC_OP_BITMASK, ATM>(p, sm, pw, lane, n_tiles_total); break;
      case RC_OP_AFFINE2: tc_rows_producer<RC_OP_AFFINE2, ATM>(p, sm, pw, lane, n_tiles_total); break;
      default: tc_rows_producer<RC_OP_NONE, ATM>(p, sm, pw, lane, n_tiles_total); break;
    }
  } else if (warp == kTcEpiWarps + kTcRowsProdWarps) {
    // ================================================================= MMA issuer (the whole warp waits, one elected lane issues)
    const uint32_t idesc = umma_idesc(2u, 128);
    uint32_t cnt = 0, tcount = 0;
    if (ATM) {
      mbar_wait(sm.w_ready, 0);                            // the weight tile is in TMEM
      tc_fence_after();
    }
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount % nbuf, aph = (tcount / nbuf) & 1;
      mbar_wait(sm.tmem_empty + 8 * acc, aph ^ 1);
      tc_fence_after();
      if (lane == 0) tc_trace(p.trace, 1, tcount, 0);
      for (int kb = 0; kb < p.kblocks; ++kb, ++cnt) {
        const uint32_t s = cnt % kStages, ph = (cnt / kStages) & 1;
        const uint32_t tmem_d = tmem_base + (acc * groups + kb / p.chain) * 128;
        mbar_wait(sm.full + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = smem_u32(sm.stages + (size_t)s * kStageBytes);
        if (elect_one()) {
          if (ATM) tc_issue_stage_atm(st, w_tmem + kb * 32, tmem_d, idesc, kb % p.chain == 0);
          else tc_issue_stage(st, tmem_d, idesc, kb % p.chain == 0);
          umma_commit(sm.empty + 8 * s);
          if (kb == p.kblocks - 1) umma_commit(sm.tmem_full + 8 * acc);
        }
        __syncwarp();
      }
      if (lane == 0) tc_trace(p.trace, 1, tcount, 1);
    }
  } else {
    // ================================================================= epilogue: thread = output channel
    if (ATM) {
      // the weight tile (this thread's output channel = its TMEM lane; k = column), split hi | lo, loaded once per CTA.
      // Stored [k][channel] (backward-data) the global loads are coalesced as they are; stored [channel][k] (forward) half
      // a tile at a time passes through shared memory: coalesced 128-bit loads in, one row per thread out (16-byte chunk
      // q of row n at n*256 + ((q ^ (n & 7)) * 16): conflict-free both ways).
      const int j = (warp & 3) * 32 + lane, half = warp >> 2;      // the two warps of a lane quarter take 64 of the 128 columns each
      const uint32_t taddr = w_tmem + ((uint32_t)((warp & 3) * 32) << 16);
      // staging: the last two operand stages (64 KB), which the producers do not touch before w_ready
      float* wst = reinterpret_cast<float*>(sm.stages + 4 * kStageBytes);
      const bool staged = g.b_layout == RC_B_COL;
      const bool wvec = (g.b.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.b.ptr) & 15u) == 0);
      const int kpad = p.kblocks * 32;
      if (warp == 0) tc_trace(p.trace, 2, 8, 0);
      if (staged) {
        // row n, 16-byte chunk q (of 32) at n*512 + ((q ^ (n & 7)) * 16) bytes
#pragma unroll 4
        for (int f = tid; f < 128 * 32; f += 32 * kTcEpiWarps) {
          const int row = f >> 5, q = f & 31, r = 4 * q;
          float* dst = wst + row * 128 + ((q ^ (row & 7)) << 2);
          if (row < g.n && r + 3 < g.k && wvec) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(g.b.ptr + (size_t)row * g.b.ld + r) : "memory");
          } else {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < g.n && r < g.k) {
              const float* src = g.b.ptr + (size_t)row * g.b.ld + r;
              v.x = __ldg(src);
              if (r + 1 < g.k) v.y = __ldg(src + 1);
              if (r + 2 < g.k) v.z = __ldg(src + 2);
              if (r + 3 < g.k) v.w = __ldg(src + 3);
            }
            st4(dst, v);
          }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("bar.sync 2, %0;" :: "r"(32 * kTcEpiWarps) : "memory");
      }
      if (warp == 0) tc_trace(p.trace, 2, 9, 0);
#pragma unroll 1
      for (int c = 64 * half; c < min(kpad, 64 * half + 64); c += 16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) {
          float v[4];
          if (staged) {
            const float4 t = ld4(wst + j * 128 + (((c >> 2) + e4) ^ (j & 7)) * 4);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = c + 4 * e4 + e;
              v[e] = (j < g.n && r < g.k) ? __ldg(g.b.ptr + (size_t)r * g.b.ld + j) : 0.f;
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float hh = to_tf32(v[e]);
            hi[4 * e4 + e] = __float_as_uint(hh);
            lo[4 * e4 + e] = __float_as_uint(v[e] - hh);
          }
        }
        tmem_st16(taddr + c, hi);
        tmem_st16(taddr + 128 + c, lo);
      }
      if (warp == 0) tc_trace(p.trace, 2, 9, 1);
      if (warp == 0) tc_trace(p.trace, 2, 8, 1);
      tmem_st_wait();
      if (warp == 0) tc_trace(p.trace, 2, 11, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.w_ready);
    }
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount % nbuf, aph = (tcount / nbuf) & 1;
      mbar_wait(sm.tmem_full + 8 * acc, aph);
      tc_fence_after();
      if (warp == 0) tc_trace(p.trace, 2, tcount, 0);
      const uint32_t taddr = tmem_base + acc * groups * 128 + ((uint32_t)((warp & 3) * 32) << 16);
      const int rt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const bool bits = g.bits_out != nullptr;
      switch (g.epi) {
        case RC_EPI_RELU:
          if (bits) tc_rows_epilogue<RC_EPI_RELU, true>(g, taddr, groups, rt, nt, warp, lane);
          else tc_rows_epilogue<RC_EPI_RELU, false>(g, taddr, groups, rt, nt, warp, lane);
          break;
        case RC_EPI_RELU_RES:
          if (bits) tc_rows_epilogue<RC_EPI_RELU_RES, true>(g, taddr, groups, rt, nt, warp, lane);
          else tc_rows_epilogue<RC_EPI_RELU_RES, false>(g, taddr, groups, rt, nt, warp, lane);
          break;
        case RC_EPI_ADD_RES: tc_rows_epilogue<RC_EPI_ADD_RES, false>(g, taddr, groups, rt, nt, warp, lane); break;
        case RC_EPI_MASK_POS: tc_rows_epilogue<RC_EPI_MASK_POS, false>(g, taddr, groups, rt, nt, warp, lane); break;
        case RC_EPI_BN_STATS: tc_rows_epilogue<RC_EPI_BN_STATS, false>(g, taddr, groups, rt, nt, warp, lane); break;
        case RC_EPI_BN_RELU_BWD: tc_rows_epilogue<RC_EPI_BN_RELU_BWD, false>(g, taddr, groups, rt, nt, warp, lane); break;
        default: tc_rows_epilogue<RC_EPI_NONE, false>(g, taddr, groups, rt, nt, warp, lane); break;
      }
      // the accumulator is free once every thread's TMEM loads have completed (they all complete inside the call)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.tmem_empty + 8 * acc);
      if (warp == 0) tc_trace(p.trace, 2, tcount, 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ wgrad kernel
struct TcWgradP {
  rc_gemm g;                     // D[i, j] (g.m x g.n), reduction over g.k samples; a stored [r][i], b stored [r][j]
  int i_tiles, j_tiles;
  int mblocks, per_split;        // 32-sample blocks in total / per reduction split
};

// OPA / OPB: the operand prologues as compile-time constants for the combinations a training step uses (-1: read them
// from the descriptor at run time)
template <int OPA, int OPB>
__global__ void __launch_bounds__(kTcWgThreads, 1) gemm_tc_wgrad_kernel(const TcWgradP p) {
  pdl_entry();
  extern __shared__ unsigned char smem_raw[];
  const TcSmem sm = tc_carve(smem_raw);
  const rc_gemm& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int z = blockIdx.x;
  const int it_ = blockIdx.y / p.j_tiles, jt = blockIdx.y % p.j_tiles;
  const int mb_beg = z * p.per_split, mb_end = min(p.mblocks, mb_beg + p.per_split);
  const int n_groups = mb_beg < mb_end ? ceil_div(mb_end - mb_beg, kTcChain) : 0;
  const int i0 = it_ * 128, j0 = jt * 128;

  if (warp == 0) tmem_alloc(sm.tmem_slot, 256);
  if (tid == 32) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(sm.full + 8 * s, kTcProdWarps);
      mbar_init(sm.empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(sm.tmem_full + 8 * a, 1);
      mbar_init(sm.tmem_empty + 8 * a, kTcWgEpiWarps);
    }
    mbar_init_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_slot_ptr;
  float* cs_smem = reinterpret_cast<float*>(sm.stages + kTcStages * kTcStageBytes + 256);     // [2][128] column sums

  if (warp >= kTcWgEpiWarps && warp < kTcWgEpiWarps + kTcProdWarps) {
    // ================================================================= producers: transpose 4 samples -> one K-major chunk
    const int pw = warp - kTcWgEpiWarps;
    const int il = (pw & 3) * 32 + lane;                  // this thread's row of both operand blocks (fixed)
    const int ia = i0 + il, jb = j0 + il;
    const bool a_ok = ia < g.m, b_ok = jb < g.n;
    float colsum = 0.f;
    const int opa = OPA >= 0 ? OPA : g.a.op, opb = OPB >= 0 ? OPB : g.b.op;
    // the operand prologues' per-column vectors: this thread's column never changes
    float pa[4] = {0.f, 0.f, 0.f, 0.f}, pb[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok && (opa == RC_OP_BN_RELU || opa == RC_OP_AFFINE2)) {
      pa[0] = __ldg(g.a.p0 + ia); pa[1] = __ldg(g.a.p1 + ia); pa[2] = __ldg(g.a.p2 + ia); pa[3] = __ldg(g.a.p3 + ia);
    }
    if (b_ok && (opb == RC_OP_BN_RELU || opb == RC_OP_AFFINE2)) {
      pb[0] = __ldg(g.b.p0 + jb); pb[1] = __ldg(g.b.p1 + jb); pb[2] = __ldg(g.b.p2 + jb); pb[3] = __ldg(g.b.p3 + jb);
    }
    // stored element -> operand value (same arithmetic as apply_op in rc_gemm_tile.cuh)
    auto prologue = [](int op, float v, float aux, uint32_t bits, int col, const float (&pv)[4]) -> float {
      switch (op) {
        case RC_OP_BN_RELU: return fmaxf((v - pv[0]) * pv[1] * pv[2] + pv[3], 0.f);
        case RC_OP_BITMASK: return ((bits >> (col & 31)) & 1u) ? v : 0.f;
        case RC_OP_AFFINE2: return fmaf(pv[0], v, fmaf(pv[1], aux - pv[3], pv[2]));
        default: return v;
      }
    };
    uint32_t cnt = 0;
    for (int mb = mb_beg; mb < mb_end; ++mb, ++cnt) {
      const uint32_t s = cnt % kTcStages, ph = (cnt / kTcStages) & 1;
      float4 va[4], vb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = q * 2 + (pw >> 2);                  // chunk of 4 samples
        const int r = mb * 32 + 4 * c;
        float t[4], u[4], xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t wa[4] = {0u, 0u, 0u, 0u}, wb[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int e = 0; e < 4; ++e) {                     // every load of the chunk is issued before the first use
          const bool ra = a_ok && r + e < g.k, rb = b_ok && r + e < g.k;
          t[e] = ra ? __ldg(g.a.ptr + (size_t)(r + e) * g.a.ld + ia) : 0.f;
          u[e] = rb ? __ldg(g.b.ptr + (size_t)(r + e) * g.b.ld + jb) : 0.f;
          if (opa == RC_OP_AFFINE2 && ra) xa[e] = __ldg(g.a.aux + (size_t)(r + e) * g.a.ld_aux + ia);
          if (opb == RC_OP_AFFINE2 && rb) xb[e] = __ldg(g.b.aux + (size_t)(r + e) * g.b.ld_aux + jb);
          if (opa == RC_OP_BITMASK && ra) wa[e] = __ldg(g.a.bits + (size_t)(r + e) * g.a.ld_bits + (ia >> 5));
          if (opb == RC_OP_BITMASK && rb) wb[e] = __ldg(g.b.bits + (size_t)(r + e) * g.b.ld_bits + (jb >> 5));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (opa != RC_OP_NONE && a_ok && r + e < g.k) t[e] = prologue(opa, t[e], xa[e], wa[e], ia, pa);
          if (opb != RC_OP_NONE && b_ok && r + e < g.k) u[e] = prologue(opb, u[e], xb[e], wb[e], jb, pb);
        }
        va[q] = make_float4(t[0], t[1], t[2], t[3]);
        vb[q] = make_float4(u[0], u[1], u[2], u[3]);
        colsum += (t[0] + t[1]) + (t[2] + t[3]);
      }
      mbar_wait(sm.empty + 8 * s, ph ^ 1);
      float* ah = reinterpret_cast<float*>(sm.stages + (size_t)s * kTcStageBytes);
      float* al = ah + kTcBlkFloats, *bh = al + kTcBlkFloats, *bl = bh + kTcBlkFloats;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = q * 2 + (pw >> 2);
        const int idx = (il >> 3) * 256 + c * 32 + (il & 7) * 4;
        float4 hi, lo;
        split_tf32_rn(va[q], hi, lo);
        st4(ah + idx, hi);
        st4(al + idx, lo);
        split_tf32_rn(vb[q], hi, lo);
        st4(bh + idx, hi);
        st4(bl + idx, lo);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.full + 8 * s);
    }
    // bias gradient: two warps hold partial column sums of every operand row
    if (g.colsum_a != nullptr && jt == 0) {
      cs_smem[(pw >> 2) * 128 + il] = colsum;
      asm volatile("bar.sync 1, %0;" :: "r"(32 * kTcProdWarps) : "memory");
      if (pw < 4 && a_ok) g.colsum_a[(size_t)z * g.m + ia] = cs_smem[il] + cs_smem[128 + il];
    }
  } else if (warp == kTcWgEpiWarps + kTcProdWarps) {
    const uint32_t idesc = umma_idesc(2u, 128);
    uint32_t cnt = 0;
    for (int gi = 0; gi < n_groups; ++gi) {
      const uint32_t acc = gi & 1, aph = (gi >> 1) & 1;
      mbar_wait(sm.tmem_empty + 8 * acc, aph ^ 1);
      tc_fence_after();
      const int g_beg = mb_beg + gi * kTcChain, g_end = min(mb_end, g_beg + kTcChain);
      for (int mb = g_beg; mb < g_end; ++mb, ++cnt) {
        const uint32_t s = cnt % kTcStages, ph = (cnt / kTcStages) & 1;
        mbar_wait(sm.full + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          tc_issue_stage(smem_u32(sm.stages + (size_t)s * kTcStageBytes), tmem_base + acc * 128, idesc, mb == g_beg);
          umma_commit(sm.empty + 8 * s);
          if (mb == g_end - 1) umma_commit(sm.tmem_full + 8 * acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= epilogue: thread = row i of D, 64 of its 128 columns
    const int i = i0 + (warp & 3) * 32 + lane;
    const int half = warp >> 2;
    float sums[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) sums[e] = 0.f;
    for (int gi = 0; gi < n_groups; ++gi) {
      const uint32_t acc = gi & 1, aph = (gi >> 1) & 1;
      mbar_wait(sm.tmem_full + 8 * acc, aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 128 + half * 64 + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t r[16];
        tmem_ld16(taddr + q * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) sums[q * 16 + e] += __uint_as_float(r[e]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.tmem_empty + 8 * acc);
    }
    float* dout = g.d + (size_t)z * g.split_stride;
    const bool vec = (g.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(dout) & 15u) == 0);
    const int jbase = j0 + half * 64;
    if (i < g.m && jbase < g.n) {
      float* dp = dout + (size_t)i * g.ldd + jbase;
      if (vec && jbase + 63 < g.n) {
#pragma unroll
        for (int e = 0; e < 64; e += 4) st4(dp + e, make_float4(sums[e], sums[e + 1], sums[e + 2], sums[e + 3]));
      } else {
#pragma unroll
        for (int e = 0; e < 64; ++e)
          if (jbase + e < g.n) dp[e] = sums[e];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int kTcMinRows = 16384;

static bool tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RC_GEMM_TC");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on == 1;
}

static long long* g_tc_trace = nullptr;
void gemm_tc_set_trace(long long* p) { g_tc_trace = p; }

static bool tc_atm_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RC_GEMM_TC_ATMEM");
    on = (e && atoi(e) != 0) ? 1 : 0;
  }
  return on == 1;
}

// 1 = rows kernel, 2 = wgrad kernel, 0 = not applicable (the SIMT kernels of rc_gemm.cu take it)
int gemm_tc_kind(const rc_gemm* g) {
  if (!tc_enabled() || g->m <= 0 || g->n <= 0) return 0;
  if (g->a_layout == RC_A_ROW) {
    if (g->m < kTcMinRows || g->splits > 1 || g->b.op != RC_OP_NONE || g->colsum_a || g->a.op == RC_OP_GINE_AGGR) return 0;
    if (g->k + g->k2 < 32) return 0;
    if (ceil_div(g->k, 32) + (g->k2 > 0 ? ceil_div(g->k2, 32) : 0) > 16) return 0;   // four accumulators of four k-blocks per tile at most
    return 1;
  }
  if (g->a_layout == RC_A_RED && g->b_layout == RC_B_RED) {
    if (g->k < kTcMinRows || g->k2 > 0 || g->epi != RC_EPI_NONE || g->bias || g->bits_out) return 0;
    if (g->m < 32 || g->n < 32) return 0;               // the head's 5-row gradient stays on the SIMT kernel
    return 2;
  }
  return 0;
}

size_t gemm_tc_workspace(const rc_gemm* g) {
  if (gemm_tc_kind(g) != 1) return 0;
  const size_t kblocks = ceil_div(g->k, 32) + (g->k2 > 0 ? ceil_div(g->k2, 32) : 0);
  return (size_t)ceil_div(g->n, 128) * kblocks * 2 * kTcBlkBytes;
}

int gemm_tc_wgrad_splits(const rc_gemm* g) {
  const int tiles = ceil_div(g->m, 128) * ceil_div(g->n, 128);
  const int mblocks = ceil_div(g->k, 32);
  int splits = kNumSMs / tiles;
  if (splits < 1) splits = 1;
  if (splits > mblocks) splits = mblocks;
  return splits;
}

int gemm_tc_run(const rc_gemm* g, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(gemm_tc_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    cudaError_t e2 = cudaFuncSetAttribute(gemm_tc_wgrad_kernel<-1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(gemm_tc_wgrad_kernel<RC_OP_NONE, RC_OP_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(gemm_tc_wgrad_kernel<RC_OP_AFFINE2, RC_OP_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(gemm_tc_wgrad_kernel<RC_OP_BITMASK, RC_OP_BN_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    cudaError_t e3 = cudaFuncSetAttribute(gemm_tc_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e1 == cudaSuccess) e1 = e3;
    if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(RC_ERR_CUDA, "rc_gemm_run (tensor cores): %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    attr_set = true;
  }
  auto vec_ok = [](const float* ptr, int ld) { return ptr != nullptr && (ld % 4 == 0) && aligned16(ptr); };
  const int kind = gemm_tc_kind(g);
  if (kind == 1) {
    TcRowsP p;
    p.g = *g;
    p.kb1 = ceil_div(g->k, 32);
    p.kblocks = p.kb1 + (g->k2 > 0 ? ceil_div(g->k2, 32) : 0);
    p.n_tiles = ceil_div(g->n, 128);
    p.row_tiles = ceil_div(g->m, 128);
    p.a_vec = vec_ok(g->a.ptr, g->a.ld);
    if (g->a.op == RC_OP_AFFINE2) p.a_vec = p.a_vec && vec_ok(g->a.aux, g->a.ld_aux);
    p.a2_vec = g->k2 > 0 ? vec_ok(g->a2, g->lda2) : 0;
    p.a_out_vec = g->a_out ? vec_ok(g->a_out, g->ld_a_out) : 0;
    p.trace = g_tc_trace;
    p.dbg = getenv("RC_TC_DBG") ? atoi(getenv("RC_TC_DBG")) : 0;
    p.wpack = nullptr;
    int grid = p.row_tiles * p.n_tiles;
    if (grid > kNumSMs) grid = kNumSMs;
    p.chain = p.kblocks <= 8 ? 2 : 4;
    // one output tile wide, K <= 128: the weight tile can live in TMEM for the whole kernel (no packing pass); it leaves
    // room for ONE accumulator per tile only, i.e. chains of 48 MMAs: opt-in (RC_GEMM_TC_ATMEM=1), measured 39 us against
    // 44 us for a 100k x 128 x 128 Linear, not worth 1e-6 of systematic error per layer
    if (p.n_tiles == 1 && p.kblocks <= 4 && g->k2 == 0 && tc_atm_enabled()) {
      p.chain = 4;
      launch_pdl(gemm_tc_rows_kernel<true>, dim3(grid), dim3(kTcThreads), (size_t)kTcSmemBytes, s, p);
      return check_launch("gemm_tc_rows_kernel");
    }
    if (!g->tc_ws || g->tc_ws_bytes < gemm_tc_workspace(g)) return fail(RC_ERR_WORKSPACE, "rc_gemm_run: tensor-core workspace missing or too small");
    if (!aligned16(g->tc_ws)) return fail(RC_ERR_ARG, "rc_gemm_run: tc_ws must be 16-byte aligned");
    TcPackP pp;
    pp.b = g->b.ptr; pp.ldb = g->b.ld; pp.b_layout = g->b_layout; pp.b2 = g->b2; pp.ldb2 = g->ldb2;
    pp.n = g->n; pp.k = g->k; pp.k2 = g->k2;
    pp.kb1 = p.kb1;
    pp.kblocks = p.kblocks;
    pp.n_tiles = p.n_tiles;
    pp.out = static_cast<float*>(g->tc_ws);
    const long long total = (long long)pp.n_tiles * pp.kblocks * kTcBlkFloats;
    int pgrid = (int)ceil_div_ll(total, 256);
    if (pgrid > 2 * kNumSMs) pgrid = 2 * kNumSMs;
    launch_pdl(tc_pack_kernel, dim3(pgrid), dim3(256), 0, s, pp);
    if (int e = check_launch("tc_pack_kernel")) return e;
    p.wpack = pp.out;
    launch_pdl(gemm_tc_rows_kernel<false>, dim3(grid), dim3(kTcThreads), (size_t)kTcSmemBytes, s, p);
    return check_launch("gemm_tc_rows_kernel");
  }
  if (kind == 2) {
    TcWgradP p;
    p.g = *g;
    p.i_tiles = ceil_div(g->m, 128);
    p.j_tiles = ceil_div(g->n, 128);
    p.mblocks = ceil_div(g->k, 32);
    const int splits = g->splits > 1 ? g->splits : 1;
    p.per_split = ceil_div(p.mblocks, splits);
    const dim3 wgrid(splits, p.i_tiles * p.j_tiles);
    if (g->a.op == RC_OP_NONE && g->b.op == RC_OP_NONE)
      launch_pdl(gemm_tc_wgrad_kernel<RC_OP_NONE, RC_OP_NONE>, wgrid, dim3(kTcWgThreads), (size_t)kTcSmemBytes, s, p);
    else if (g->a.op == RC_OP_AFFINE2 && g->b.op == RC_OP_NONE)
      launch_pdl(gemm_tc_wgrad_kernel<RC_OP_AFFINE2, RC_OP_NONE>, wgrid, dim3(kTcWgThreads), (size_t)kTcSmemBytes, s, p);
    else if (g->a.op == RC_OP_BITMASK && g->b.op == RC_OP_BN_RELU)
      launch_pdl(gemm_tc_wgrad_kernel<RC_OP_BITMASK, RC_OP_BN_RELU>, wgrid, dim3(kTcWgThreads), (size_t)kTcSmemBytes, s, p);
    else
      launch_pdl(gemm_tc_wgrad_kernel<-1, -1>, wgrid, dim3(kTcWgThreads), (size_t)kTcSmemBytes, s, p);
    return check_launch("gemm_tc_wgrad_kernel");
  }
  return fail(RC_ERR_ARG, "rc_gemm_run: tensor-core path not applicable");
}

}  // namespace rc
