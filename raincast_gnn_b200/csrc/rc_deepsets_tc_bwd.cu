// DeepSets member Linear + ReLU + SUM pool, BACKWARD, on the 5th-generation tensor cores (models/gnn.py:51-56,66-67;
// BASELINE.json configs 4 and 5).  Nothing was saved in forward: per stage of whole stations (<= 64 member rows)
//   MMA1   D1[channel, row] = W1 . E^T                       (as forward: A = W1 tile in shared memory, B = member rows)
//   E1     thread = channel: mask = D1 + b1 > 0,  dh = mask ? d_pooled[station(row), channel] : 0,  d b1 += dh,
//          dh split hi | lo and written with tcgen05.st into TMEM as the A operand of
//   MMA2   D2[channel, feature] += dh . E                    (A = dh from TMEM, B = E^T in shared memory, K = member rows)
// so neither the pre-activations nor dh [M*members, H] ever exist outside TMEM, and the two contractions that were
// 3.0 ms of FFMA at config 4 run as 3xTF32 tcgen05 MMAs (fp32 parity; bf16-rounded operands after a bf16 forward).
//
// Roles (14 warps): 8 x E1 (two per TMEM lane quarter, 32 of the stage's 64 columns each), 4 x converter (raw member
// rows -> B1 = E as [row][feature] K-major and B2 = E^T as [feature][row] K-major, hi | lo), the MMA issuer, the loader
// (one bulk copy of the stage's contiguous member rows per stage - TMA engine, cp.async.bulk).
// Pipelines: raw ring (3) -> B1|B2 ring (3) -> D1 (2 in TMEM) -> A2 (2 in TMEM) -> D2 (2 in TMEM).  A D2 accumulator
// takes one stage (21 accumulating MMAs: the tensor core truncates every accumulation, rc_gemm_tc.cu) and is then
// flushed by the E1 warps into round-to-nearest register sums while the next stage accumulates into the other one.
// The member count is compile-time (the reference's ensembles: 11 reforecast / 51 forecast members), so the station of
// a column is static; other member counts take the SIMT kernel (rc_deepsets.cu).
#include <cuda_bf16.h>
#include <stdlib.h>

#include <type_traits>

#include "rc_umma.cuh"

namespace rc {

constexpr int kDbE1Warps = 8, kDbCvtWarps = 4;
constexpr int kDbThreads = 32 * (kDbE1Warps + kDbCvtWarps + 2);
constexpr int kDbRows = 64;                 // MMA1 N / MMA2 K: member rows per stage (whole stations, zero padded)
constexpr int kDbRaw = 3, kDbBuf = 3;

struct DsBwdTcP {
  const float* ens; const float* w1; const float* b1; const float* d_pooled;
  float* partials;                          // [gridDim.x][hidden*feats + hidden]
  uint32_t* mask_out;                       // debug: ReLU mask bits [m*members][hidden/32] (NULL in production)
  int m, members, feats, hidden;
  int kp, np;                               // feats padded to 8 (MMA1 K) / to 16 (MMA2 N)
  int bf16;                                 // operands rounded to bf16 first (the forward ran on bf16 tensor cores)
  int n_stages;                             // ceil(m / stations per stage)
  long long* trace;                         // debug (rc_debug_ds_trace): CTA 0, [event 0..7][stage < 32] clock
};
__device__ __forceinline__ void db_trace(long long* trace, int ev, int i) {
  if (trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && i < 32 && (threadIdx.x & 31) == 0) trace[ev * 32 + i] = clock64();
}

// barrier indices (8 bytes each)
enum { DB_RAW_FULL = 0, DB_RAW_EMPTY = 3, DB_B_FULL = 6, DB_B_EMPTY = 9, DB_D1_FULL = 12, DB_D1_EMPTY = 14, DB_A2_FULL = 16,
       DB_A2_EMPTY = 18, DB_D2_FULL = 20, DB_D2_EMPTY = 22, DB_NBARS = 24 };

__host__ __device__ inline size_t db_smem_bytes(int feats, int kp, int np) {
  const size_t raw = ((size_t)kDbRows * feats + 8 + 3) / 4 * 4;                 // + misalignment slack, 16-byte multiple
  return 128 + 4 * (kDbRaw * raw + 2 * (size_t)128 * kp + kDbBuf * 2 * ((size_t)kDbRows * kp + (size_t)np * kDbRows)) + 8 * DB_NBARS + 16;
}

// FIXED: 33..40 features (the reference has 35): KP = 40, NP = 48 are compile-time and the converter loops unroll
template <int MEMBERS, bool DUMP, bool FIXED>
__global__ void __launch_bounds__(kDbThreads, 1) deepsets_pool_bwd_tc_kernel(const DsBwdTcP p) {
  pdl_entry();
  constexpr int NPT = kDbRows / MEMBERS;            // stations per stage
  constexpr int USED = NPT * MEMBERS;               // member rows per full stage
  constexpr int KS2 = (USED + 7) / 8;               // MMA2 k-steps
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  const int F = p.feats, KP = FIXED ? 40 : p.kp, NP = FIXED ? 48 : p.np;
  const int raw_floats = (kDbRows * F + 8 + 3) / 4 * 4;
  float* raw = reinterpret_cast<float*>(base);
  float* a1 = raw + kDbRaw * raw_floats;
  const int a1_floats = 128 * KP, b1_floats = kDbRows * KP, b2_floats = NP * kDbRows;
  float* b1buf = a1 + 2 * a1_floats;
  float* b2buf = b1buf + kDbBuf * 2 * b1_floats;
  const uint32_t bars = smem_u32(b2buf + kDbBuf * 2 * b2_floats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b2buf + kDbBuf * 2 * b2_floats) + 2 * DB_NBARS;
  auto bar = [&](int idx) { return bars + 8u * idx; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * 128;                  // hidden chunk of this CTA
  const long long total_f = (long long)p.m * MEMBERS * F;

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  if (tid == 32) {
    for (int i = 0; i < kDbRaw; ++i) { mbar_init(bar(DB_RAW_FULL + i), 1); mbar_init(bar(DB_RAW_EMPTY + i), kDbCvtWarps); }
    for (int i = 0; i < kDbBuf; ++i) { mbar_init(bar(DB_B_FULL + i), kDbCvtWarps); mbar_init(bar(DB_B_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(DB_D1_FULL + i), 1); mbar_init(bar(DB_D1_EMPTY + i), kDbE1Warps);
      mbar_init(bar(DB_A2_FULL + i), kDbE1Warps); mbar_init(bar(DB_A2_EMPTY + i), 1);
      mbar_init(bar(DB_D2_FULL + i), 1); mbar_init(bar(DB_D2_EMPTY + i), kDbE1Warps);
    }
    mbar_init_fence();
  }
  // W1 tile -> A1 (K-major, 8-row groups: LBO = 128, SBO = (KP/4)*128), split hi | lo; coalesced reads of the [128][F] block
  {
    const int ncols = max(0, min(128, p.hidden - c0));
    for (int idx = tid; idx < 128 * KP; idx += kDbThreads) {
      const int c = idx / KP, k = idx - c * KP;
      float v = (c < ncols && k < F) ? __ldg(p.w1 + (size_t)(c0 + c) * F + k) : 0.f;
      if (p.bf16) v = __bfloat162float(__float2bfloat16_rn(v));
      const float hi = to_tf32(v);
      const int o = (c >> 3) * (KP / 4) * 32 + (k >> 2) * 32 + (c & 7) * 4 + (k & 3);
      a1[o] = hi;
      a1[a1_floats + o] = v - hi;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_d1 = tmem, t_a2 = tmem + 128, t_d2 = tmem + 384;

  const int n_my = (int)blockIdx.x < p.n_stages ? (p.n_stages - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // floats between the 16-byte boundary below the stage's first element and that element
  auto stage_first = [&](int st) -> long long { return (long long)st * NPT * MEMBERS * F; };
  auto misalign = [&](int st) -> int { return (int)(((reinterpret_cast<uintptr_t>(p.ens) >> 2) + stage_first(st)) & 3); };

  if (warp == kDbE1Warps + kDbCvtWarps + 1) {
    // ================================================================= loader: raw member rows of a stage, one bulk copy
    if (lane == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int st = blockIdx.x + i * gridDim.x;
        const uint32_t s = i % kDbRaw, ph = (i / kDbRaw) & 1;
        const int n_nodes = min(NPT, p.m - st * NPT);
        const int mis = misalign(st);
        const long long g0 = stage_first(st) - mis;                          // float index of the first 16-byte chunk
        long long g1 = stage_first(st) + (long long)n_nodes * MEMBERS * F;    // one past the last float
        long long g1a = (g1 + 3) & ~3LL;
        if (g1a > total_f) g1a = g1 & ~3LL;                                  // never read past the tensor: the converters fetch the tail
        mbar_wait(bar(DB_RAW_EMPTY + s), ph ^ 1);
        const uint32_t bytes = (uint32_t)((g1a - g0) * 4);
        mbar_arrive_expect_tx(bar(DB_RAW_FULL + s), bytes);
        if (bytes) bulk_g2s(smem_u32(raw + s * raw_floats), p.ens + g0, bytes, bar(DB_RAW_FULL + s));
      }
    }
  } else if (warp == kDbE1Warps + kDbCvtWarps) {
    // ================================================================= MMA issuer (the whole warp waits, one elected lane issues)
    const uint32_t idesc1 = umma_idesc(2u, kDbRows), idesc2 = umma_idesc(2u, NP);
    const uint32_t lbo = 128, sbo1a = (uint32_t)(KP / 4) * 128, sbo2 = (kDbRows / 4) * 128;
    const int ks1 = KP / 8;
    const UmmaDesc a1h = umma_desc2(smem_u32(a1), lbo, sbo1a);
    const uint32_t a1_lo_off = a1_floats * 4;
    auto mma2 = [&](int i) {                                                // D2 += dh(stage i) . E(stage i)
      const uint32_t ab = i & 1, aph = (i >> 1) & 1, bb = i % kDbBuf;
      const uint32_t pair = i & 1, pph = (i >> 1) & 1;                       // one D2 accumulator per stage, alternating
      mbar_wait(bar(DB_A2_FULL + ab), aph);
      mbar_wait(bar(DB_D2_EMPTY + pair), pph ^ 1);                           // the accumulator was flushed
      tc_fence_after();
      if (lane == 0) db_trace(p.trace, 3, i);
      const uint32_t a_hi = t_a2 + ab * 128, a_lo = a_hi + 64, d2 = t_d2 + pair * 64;
      const UmmaDesc bh = umma_desc2(smem_u32(b2buf + bb * 2 * b2_floats), lbo, sbo2);
      const uint32_t b_lo_off = b2_floats * 4;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KS2; ++ks) {
          const uint32_t off = ks * 2 * lbo;
          umma_tf32_ts(d2, a_lo + ks * 8, bh.at(off), idesc2, ks == 0 ? 0u : 1u);
          if (!p.bf16) umma_tf32_ts(d2, a_hi + ks * 8, bh.at(b_lo_off + off), idesc2, 1u);
          umma_tf32_ts(d2, a_hi + ks * 8, bh.at(off), idesc2, 1u);
        }
        umma_commit(bar(DB_B_EMPTY + bb));
        umma_commit(bar(DB_A2_EMPTY + ab));
        umma_commit(bar(DB_D2_FULL + pair));
      }
      __syncwarp();
    };
    for (int i = 0; i < n_my; ++i) {
      const uint32_t db = i & 1, dph = (i >> 1) & 1, bb = i % kDbBuf, bph = (i / kDbBuf) & 1;
      mbar_wait(bar(DB_B_FULL + bb), bph);
      mbar_wait(bar(DB_D1_EMPTY + db), dph ^ 1);
      tc_fence_after();
      if (lane == 0) db_trace(p.trace, 2, i);
      const UmmaDesc bh = umma_desc2(smem_u32(b1buf + bb * 2 * b1_floats), lbo, sbo1a);
      const uint32_t b_lo_off = b1_floats * 4;
      const uint32_t d1 = t_d1 + db * 64;
      if (elect_one()) {
        for (int ks = 0; ks < ks1; ++ks) {
          const uint32_t off = ks * 2 * lbo;
          if (!p.bf16) {
            umma_tf32(d1, a1h.at(a1_lo_off + off), bh.at(off), idesc1, ks == 0 ? 0u : 1u);
            umma_tf32(d1, a1h.at(off), bh.at(b_lo_off + off), idesc1, 1u);
            umma_tf32(d1, a1h.at(off), bh.at(off), idesc1, 1u);
          } else {
            umma_tf32(d1, a1h.at(off), bh.at(off), idesc1, ks == 0 ? 0u : 1u);
          }
        }
        umma_commit(bar(DB_D1_FULL + db));
      }
      __syncwarp();
      if (i >= 1) mma2(i - 1);
    }
    if (n_my >= 1) mma2(n_my - 1);
  } else if (warp >= kDbE1Warps) {
    // ================================================================= converters: raw rows -> B1 (E) and B2 (E^T), hi | lo
    const int ct = tid - 32 * kDbE1Warps;                                     // 0..127
    const float inv_np = 1.0f / (float)NP;
    for (int i = 0; i < n_my; ++i) {
      const int st = blockIdx.x + i * gridDim.x;
      const uint32_t s = i % kDbRaw, ph = (i / kDbRaw) & 1, bb = i % kDbBuf, bph = (i / kDbBuf) & 1;
      const int n_nodes = min(NPT, p.m - st * NPT);
      const int rows = n_nodes * MEMBERS;
      const int mis = misalign(st);
      float* rw = raw + s * raw_floats;
      mbar_wait(bar(DB_RAW_FULL + s), ph);
      {  // tail floats the aligned bulk copy could not take without reading past the tensor
        const long long g1 = stage_first(st) + (long long)rows * F;
        if (((g1 + 3) & ~3LL) > total_f) {
          const int done = (int)((g1 & ~3LL) - (stage_first(st) - mis));
          if (ct < (int)(g1 & 3)) rw[done + ct] = __ldg(p.ens + (g1 & ~3LL) + ct);
          asm volatile("bar.sync 3, %0;" :: "r"(32 * kDbCvtWarps) : "memory");
        }
      }
      mbar_wait(bar(DB_B_EMPTY + bb), bph ^ 1);
      if (ct == 0) db_trace(p.trace, 0, i);
      const float* e = rw + mis;
      float* b1h = b1buf + bb * 2 * b1_floats, *b1l = b1h + b1_floats;
      float* b2h = b2buf + bb * 2 * b2_floats, *b2l = b2h + b2_floats;
      auto round_op = [&](float v) { return p.bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v; };
      if constexpr (FIXED) {
        // every shared-memory read of the stage is issued before the first conversion (one converter warp per
        // sub-partition: without instruction-level parallelism the role is bound by the 29-cycle LDS latency)
        constexpr int N1 = kDbRows * (40 / 4) / (32 * kDbCvtWarps);           // 5 chunks of B1 per thread
        constexpr int N2 = 48 * (kDbRows / 4) / (32 * kDbCvtWarps);           // 6 chunks of B2 per thread
        float v1[N1][4], v2[N2][4];
#pragma unroll
        for (int it = 0; it < N1; ++it) {
          const int idx = ct + it * 32 * kDbCvtWarps;
          const int q = idx >> 6, r = idx & 63;                               // (4-feature chunk, row): consecutive threads = consecutive rows
#pragma unroll
          for (int t = 0; t < 4; ++t) v1[it][t] = (r < rows && 4 * q + t < F) ? e[r * F + 4 * q + t] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < N2; ++it) {
          const int idx = ct + it * 32 * kDbCvtWarps;
          const int q = idx / 48, f = idx - q * 48;                           // (4-row chunk, feature): consecutive threads = consecutive features
#pragma unroll
          for (int t = 0; t < 4; ++t) v2[it][t] = (4 * q + t < rows && f < F) ? e[(4 * q + t) * F + f] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < N1; ++it) {
          const int idx = ct + it * 32 * kDbCvtWarps;
          const int q = idx >> 6, r = idx & 63;
          float4 hi, lo;
          split_tf32_trunc(make_float4(round_op(v1[it][0]), round_op(v1[it][1]), round_op(v1[it][2]), round_op(v1[it][3])), hi, lo);
          const int o = (r >> 3) * (40 / 4) * 32 + q * 32 + (r & 7) * 4;
          st4(b1h + o, hi);
          st4(b1l + o, lo);
        }
#pragma unroll
        for (int it = 0; it < N2; ++it) {
          const int idx = ct + it * 32 * kDbCvtWarps;
          const int q = idx / 48, f = idx - q * 48;
          float4 hi, lo;
          split_tf32_trunc(make_float4(round_op(v2[it][0]), round_op(v2[it][1]), round_op(v2[it][2]), round_op(v2[it][3])), hi, lo);
          const int o = (f >> 3) * (kDbRows / 4) * 32 + q * 32 + (f & 7) * 4;
          st4(b2h + o, hi);
          st4(b2l + o, lo);
        }
      } else {
      // B1: [64 rows][KP] K-major over features; thread <-> (row, 4-feature chunk), consecutive threads = consecutive rows
      for (int idx = ct; idx < kDbRows * (KP / 4); idx += 32 * kDbCvtWarps) {
        const int q = idx / kDbRows, r = idx - q * kDbRows;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int k = 4 * q + t;
          v[t] = round_op((r < rows && k < F) ? e[r * F + k] : 0.f);
        }
        float4 hi, lo;
        split_tf32_trunc(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
        const int o = (r >> 3) * (KP / 4) * 32 + q * 32 + (r & 7) * 4;
        st4(b1h + o, hi);
        st4(b1l + o, lo);
      }
      // B2: [NP features][64 rows] K-major over rows; thread <-> (feature, 4-row chunk), consecutive threads = consecutive features
      for (int idx = ct; idx < NP * (kDbRows / 4); idx += 32 * kDbCvtWarps) {
        const int q = __float2int_rd(((float)idx + 0.5f) * inv_np), f = idx - q * NP;     // idx / NP without an integer division
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = 4 * q + t;
          v[t] = round_op((r < rows && f < F) ? e[r * F + f] : 0.f);
        }
        float4 hi, lo;
        split_tf32_trunc(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
        const int o = (f >> 3) * (kDbRows / 4) * 32 + q * 32 + (f & 7) * 4;
        st4(b2h + o, hi);
        st4(b2l + o, lo);
      }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(DB_B_FULL + bb));
        mbar_arrive(bar(DB_RAW_EMPTY + s));
      }
      if (ct == 0) db_trace(p.trace, 1, i);
    }
  } else {
    // ================================================================= E1: thread = channel (TMEM lane), 32 of the stage's columns
    const int quarter = warp & 3, half = warp >> 2;
    const int c = quarter * 32 + lane, col = c0 + c;
    const bool cok = col < p.hidden;
    const float bias = cok ? __ldg(p.b1 + col) : 0.f;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    float dbsum = 0.f;
    float sums[32];                                   // D2 columns [32*half, 32*half + 32) of this channel (half 1: 16 used)
#pragma unroll
    for (int e = 0; e < 32; ++e) sums[e] = 0.f;
    auto flush = [&](int pair_idx) {                  // add the D2 accumulator of stage `pair_idx` into the register sums
      const uint32_t pair = pair_idx & 1, pph = (pair_idx >> 1) & 1;
      mbar_wait(bar(DB_D2_FULL + pair), pph);
      tc_fence_after();
      const uint32_t ta = t_d2 + pair * 64 + half * 32 + lane_base;
      if (half * 32 < NP) {
        uint32_t r[16];
        tmem_ld16(ta, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) sums[e] += __uint_as_float(r[e]);
        if (half * 32 + 16 < NP) {
          tmem_ld16(ta + 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) sums[16 + e] += __uint_as_float(r[e]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(DB_D2_EMPTY + pair));
    };
    for (int i = 0; i < n_my; ++i) {
      const int st = blockIdx.x + i * gridDim.x;
      const int n0 = st * NPT;
      const uint32_t db = i & 1, dph = (i >> 1) & 1;
      // d_pooled of this channel for the stage's stations (static column -> station map)
      float dpv[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) dpv[j] = (cok && n0 + j < p.m) ? __ldg(p.d_pooled + (size_t)(n0 + j) * p.hidden + col) : 0.f;
      mbar_wait(bar(DB_D1_FULL + db), dph);
      mbar_wait(bar(DB_A2_EMPTY + db), dph ^ 1);
      tc_fence_after();
      if (warp == 0) db_trace(p.trace, 4, i);
      const uint32_t td = t_d1 + db * 64 + half * 32 + lane_base;
      const uint32_t ta = t_a2 + db * 128 + half * 32 + lane_base;
      auto e1_body = [&](auto half_c) {                   // `half` as a compile-time constant: the column -> station map is static
        constexpr int H = decltype(half_c)::value;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t r[16], hi[16], lo[16];
          tmem_ld16(td + q * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            constexpr int dummy = 0; (void)dummy;
            const int icol = H * 32 + q * 16 + e;          // column of the stage = member row
            float dh = 0.f;
            bool on = false;
            if (icol < USED) {
              on = __uint_as_float(r[e]) + bias > 0.f;
              dh = on ? dpv[icol / MEMBERS < NPT ? icol / MEMBERS : 0] : 0.f;
            }
            if (DUMP) {                                     // (compile-time: the production kernel carries no trace of it)
              const unsigned word = __ballot_sync(0xffffffffu, on && cok);
              const long long row = (long long)n0 * MEMBERS + icol;
              if (lane == 0 && icol < USED && row < (long long)p.m * MEMBERS && c0 + quarter * 32 < p.hidden)
                p.mask_out[row * ((p.hidden + 31) / 32) + (c0 >> 5) + quarter] = word;
            }
            dbsum += dh;
            const float h = __uint_as_float(__float_as_uint(dh) & 0xffffe000u);
            hi[e] = __float_as_uint(h);
            lo[e] = __float_as_uint(dh - h);
          }
          tmem_st16(ta + q * 16, hi);
          tmem_st16(ta + 64 + q * 16, lo);
        }
      };
      if (half == 0) e1_body(std::integral_constant<int, 0>{});
      else e1_body(std::integral_constant<int, 1>{});
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(DB_D1_EMPTY + db));
        mbar_arrive(bar(DB_A2_FULL + db));
      }
      if (warp == 0) db_trace(p.trace, 5, i);
      if (i >= 1) flush(i - 1);                         // MMA2 of the previous stage was issued when this stage's D1 was: long completed
      if (warp == 0) db_trace(p.trace, 6, i);
    }
    if (n_my >= 1) flush(n_my - 1);
    // partials[blockIdx.x][hidden*feats + hidden]: d W1 [channel][feature] and d b1 [channel] of this CTA's stages
    float* out = p.partials + (size_t)blockIdx.x * ((size_t)p.hidden * F + p.hidden);
    if (cok) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int f = half * 32 + e;
        if (f < F) out[(size_t)col * F + f] = sums[e];
      }
    }
    // d b1: the two warps of a lane quarter hold the sums of their column halves
    float* dbs = reinterpret_cast<float*>(raw);         // raw ring is idle now (all stages consumed)
    asm volatile("bar.sync 2, %0;" :: "r"(32 * kDbE1Warps) : "memory");
    if (half == 1) dbs[c] = dbsum;
    asm volatile("bar.sync 2, %0;" :: "r"(32 * kDbE1Warps) : "memory");
    if (half == 0 && cok) out[(size_t)p.hidden * F + col] = dbsum + dbs[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ host side
static long long* g_ds_trace = nullptr;
void deepsets_bwd_tc_set_trace(long long* p) { g_ds_trace = p; }

bool deepsets_bwd_tc_applicable(int num_nodes, int members, int feats, int hidden) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("RC_DEEPSETS_TC");
    forced = e ? (atoi(e) ? 1 : 2) : 0;                 // 1 = always when legal, 2 = never, 0 = by size
  }
  const bool legal = (members == 11 || members == 51) && feats >= 1 && feats <= 64 && hidden >= 1;
  if (!legal || forced == 2) return false;
  if (forced == 1) return true;
  return (long long)num_nodes * members >= 8192;
}

int deepsets_bwd_tc_blocks(int num_nodes, int members) {
  const int npt = kDbRows / members;
  const int n_stages = ceil_div(num_nodes > 0 ? num_nodes : 1, npt);
  return n_stages < kNumSMs ? n_stages : kNumSMs;
}

int launch_deepsets_bwd_tc(const float* ens, const float* w1, const float* b1, const float* d_pooled, float* partials,
                           uint32_t* mask_out, int num_nodes, int members, int feats, int hidden, int bf16, cudaStream_t s) {
  DsBwdTcP p;
  p.ens = ens; p.w1 = w1; p.b1 = b1; p.d_pooled = d_pooled; p.partials = partials; p.mask_out = mask_out;
  p.m = num_nodes; p.members = members; p.feats = feats; p.hidden = hidden;
  p.kp = ceil_div(feats, 8) * 8;
  p.np = ceil_div(feats, 16) * 16;
  p.bf16 = bf16;
  const int npt = kDbRows / members;
  p.n_stages = ceil_div(num_nodes, npt);
  p.trace = g_ds_trace;
  const size_t smem = db_smem_bytes(feats, p.kp, p.np);
  if (smem > 227 * 1024) return fail(RC_ERR_ARG, "deepsets tensor-core backward: feats=%d needs %zu bytes of shared memory", feats, smem);
  if (!aligned16(ens)) return fail(RC_ERR_ARG, "deepsets tensor-core backward: ens must be 16-byte aligned");
  dim3 grid(deepsets_bwd_tc_blocks(num_nodes, members), ceil_div(hidden, 128));
  auto go = [&](auto kern, size_t& attr) -> int {
    if (smem > attr) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(RC_ERR_CUDA, "deepsets tensor-core backward: %s", cudaGetErrorString(e));
      attr = smem;
    }
    launch_pdl(kern, grid, dim3(kDbThreads), smem, s, p);
    return RC_OK;
  };
  static size_t attr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool fixed = p.kp == 40 && p.np == 48;
  int rc;
  if (members == 11) {
    if (fixed) rc = mask_out ? go(deepsets_pool_bwd_tc_kernel<11, true, true>, attr[0]) : go(deepsets_pool_bwd_tc_kernel<11, false, true>, attr[1]);
    else rc = mask_out ? go(deepsets_pool_bwd_tc_kernel<11, true, false>, attr[2]) : go(deepsets_pool_bwd_tc_kernel<11, false, false>, attr[3]);
  } else {
    if (fixed) rc = mask_out ? go(deepsets_pool_bwd_tc_kernel<51, true, true>, attr[4]) : go(deepsets_pool_bwd_tc_kernel<51, false, true>, attr[5]);
    else rc = mask_out ? go(deepsets_pool_bwd_tc_kernel<51, true, false>, attr[6]) : go(deepsets_pool_bwd_tc_kernel<51, false, false>, attr[7]);
  }
  if (rc) return rc;
  return check_launch("deepsets_pool_bwd_tc_kernel");
}

}  // namespace rc
