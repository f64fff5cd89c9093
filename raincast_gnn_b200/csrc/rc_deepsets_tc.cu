// DeepSets member Linear + ReLU + SUM pool on the 5th-generation tensor cores (tcgen05 + TMEM), for the shapes
// where members x stations x hidden is a real dense contraction (BASELINE.json configs 4 and 5).
//
// The problem is transposed so that the pooling epilogue is free:   D[channel, member row] = W1 . E^T
//   A = W1 tile      (M = 128 hidden channels, K = features)            shared memory, K-major, loaded once
//   B = member rows  (N = 128 rows = whole stations, K = features)      shared memory, K-major, per tile
//   D in TMEM: lane = channel, column = member row  ->  the thread that owns a lane sums ReLU(D + b1) over the
//   columns of each station in registers and writes pooled[station, channel] (coalesced across the 128 lanes).
// Nothing of size [M*members, H] exists anywhere (the reference materialises it twice, models/gnn.py:66-67).
//
// Precision: fp32 parity (1e-5) needs more than one TF32 product, so the fp32 mode issues the 3xTF32 split
//   a = a_hi + a_lo (a_hi = cvt.rna.tf32), D += a_hi b_hi + a_hi b_lo + a_lo b_hi        (error ~2^-21 per product)
// and the bf16 mode (config 5) one kind::f16 product with fp32 accumulation.
// Operand tiles use the canonical no-swizzle K-major layout: 16-byte chunk c of row r at  c*2048 + r*16
// (8 rows x 16 B core matrices; SBO = 128 B between 8-row groups, LBO = 2048 B between K-adjacent chunks).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rc_common.cuh"

namespace rc {

// threads per member row / hidden channel (they split the K chunks and the accumulator columns): measured at config 4 / 5,
// fp32 mode 470 us with one, 438 us with two; bf16 mode (H = 512) 908 us with one, 1145 us with two
__host__ __device__ constexpr int tc_threads_per_row(bool bf16) { return bf16 ? 1 : 2; }
constexpr int kTcRows = 128;          // MMA N: member rows per tile
constexpr int kTcChunkBytes = kTcRows * 16;   // one 16-byte K chunk for 128 rows

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  // start address (>>4) | LBO (>>4) << 16 | SBO (>>4) << 32 | version 1 << 46 | SWIZZLE_NONE
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kTcChunkBytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ bool tc_elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 32 lanes x 32 columns of the accumulator -> one column per register; completes at tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// dynamic shared memory (bytes), all 16-byte aligned:
//   ny x { A_hi [chunks][128][16] | A_lo (fp32 mode) } | B_hi | B_lo (fp32 mode) | staging [128*feats + 8] fp32 | bias [ny*128] | part [128] | mbar | tmem ptr
// ny = hidden-channel chunks (of 128) handled inside one CTA: a tile's member rows are staged and converted once and
// multiplied with ny W1 tiles one after the other (bf16 mode, where ny W1 tiles fit beside two CTAs per SM); with ny = 1 the
// chunks are blockIdx.y.
// A tile's member rows are contiguous in HBM; they travel as 16-byte cp.async chunks of the aligned span that covers them
// (the tile starts `mis` floats into the first chunk), issued for tile i+1 right after tile i has been converted, so the
// HBM latency of the next tile hides behind the MMAs, the TMEM read-back and the pooling of the current one.
// MEMBERS > 0 / KQ > 0: the member count / ceil(feats / 8) are compile-time (instantiated for the reference's ensembles of
// 11 and 51 members and its 35 features): station boundaries in the pooling epilogue and the operand conversion are then
// static - no per-column boundary test or branch, all shared-memory loads of a row in flight at once.  0 = any value.
template <bool BF16, int MEMBERS, int KQ>
__global__ void __launch_bounds__(128 * tc_threads_per_row(BF16))
deepsets_pool_fwd_tc_kernel(const float* __restrict__ ens, const float* __restrict__ w1, const float* __restrict__ b1,
                            float* __restrict__ pooled, int m, int members, int feats, int hidden, int nodes_per_tile,
                            int chunks, int ny) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kParts = BF16 ? 1 : 2;                 // hi (+ lo)
  constexpr int NH = tc_threads_per_row(BF16);
  constexpr int kTcThreads = 128 * NH;
  constexpr int kElemsPerChunk = BF16 ? 8 : 4;
  const int op_bytes = chunks * kTcChunkBytes;
  const int a_bytes = kParts * op_bytes;               // one W1 tile: hi (+ lo)
  unsigned char* b_hi = smem_raw + ny * a_bytes;
  unsigned char* b_lo = b_hi + op_bytes;
  float* staging = reinterpret_cast<float*>(b_hi + kParts * op_bytes);
  float* bias = staging + kTcRows * feats + 8;
  float* part = bias + ny * 128;                           // upper-half partial sum of the station that straddles column 64
  uint64_t* mbar = reinterpret_cast<uint64_t*>(part + 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int lane_row = tid & 127, half = tid >> 7;      // row of the operand tile / TMEM lane, and which half of its work
  const int c0 = blockIdx.y * 128 * ny;

  const int n_tiles = ceil_div(m, nodes_per_tile);
  const long long total_f = (long long)m * members * feats;
  // floats between the 16-byte boundary below the tile's first element and that element
  auto misalign = [&](int tile) -> int {
    const float* first = ens + (size_t)tile * nodes_per_tile * members * feats;
    return (int)((reinterpret_cast<uintptr_t>(first) & 15) >> 2);
  };
  auto prefetch = [&](int tile) {
    const int n0 = tile * nodes_per_tile;
    const int mis = misalign(tile);
    const long long g0 = (long long)n0 * members * feats - mis;        // float index of the first chunk (may be < 0)
    const int nch = (mis + min(nodes_per_tile, m - n0) * members * feats + 3) >> 2;
    for (int c = tid; c < nch; c += kTcThreads) {
      const long long g = g0 + 4 * c;
      if (g >= 0 && g + 4 <= total_f) {
        cp_async16(smem_u32(staging + 4 * c), ens + g);
      } else {                                                           // the chunk straddles an end of the buffer
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (g + e >= 0 && g + e < total_f) staging[4 * c + e] = __ldg(ens + g + e);
      }
    }
  };
  if ((int)blockIdx.x < n_tiles) prefetch(blockIdx.x);

  // ---- one-time setup: TMEM (128 fp32 accumulator columns), mbarrier, W1 tile, bias
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int y = 0; y < ny; ++y) {
    const int col = c0 + y * 128 + lane_row;             // thread pair <-> hidden channel (row of the A tile)
    unsigned char* a_hi = smem_raw + y * a_bytes;
    unsigned char* a_lo = a_hi + op_bytes;
    if (half == 0) bias[y * 128 + lane_row] = col < hidden ? __ldg(b1 + col) : 0.f;
    for (int c = half; c < chunks; c += NH) {
      float v[8];
#pragma unroll
      for (int e = 0; e < kElemsPerChunk; ++e) {
        const int k = c * kElemsPerChunk + e;
        v[e] = (col < hidden && k < feats) ? __ldg(w1 + (size_t)col * feats + k) : 0.f;
      }
      if (BF16) {
        __nv_bfloat162 p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) p[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(a_hi + c * kTcChunkBytes + lane_row * 16) = *reinterpret_cast<uint4*>(p);
      } else {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { hi[e] = to_tf32(v[e]); lo[e] = v[e] - hi[e]; }
        *reinterpret_cast<float4*>(a_hi + c * kTcChunkBytes + lane_row * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(a_lo + c * kTcChunkBytes + lane_row * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  cp_async_wait_all();
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();                                        // the first tile has landed, the W1 tile is in place
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = *tmem_slot;
  // instruction descriptor: D fp32, A/B tf32 (2) or bf16 (1), both K-major, N = 128 (>>3), M = 128 (>>4)
  const uint32_t fmt = BF16 ? 1u : 2u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTcRows >> 3) << 17) | ((128u >> 4) << 24);
  const int ksteps = chunks / 2;                         // one MMA covers 32 bytes of K = 2 chunks
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n0 = tile * nodes_per_tile;
    const int n_nodes = min(nodes_per_tile, m - n0);
    const int rows = n_nodes * members;
    // ---- thread pair <-> member row: convert the staged rows to the operand format, canonical K-major layout
    {
      const float* row = staging + misalign(tile) + lane_row * feats;
      const bool live = lane_row < rows;
      auto put_chunk = [&](int c, const float* v) {                   // kElemsPerChunk values -> 16 bytes of B_hi (and B_lo)
        if (BF16) {
          __nv_bfloat162 p[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) p[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          *reinterpret_cast<uint4*>(b_hi + c * kTcChunkBytes + lane_row * 16) = *reinterpret_cast<uint4*>(p);
        } else {
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { hi[e] = to_tf32(v[e]); lo[e] = v[e] - hi[e]; }
          *reinterpret_cast<float4*>(b_hi + c * kTcChunkBytes + lane_row * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(b_lo + c * kTcChunkBytes + lane_row * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      };
      if constexpr (KQ > 0) {
        constexpr int KP = BF16 ? (8 * KQ + 15) / 16 * 16 : 8 * KQ;
        constexpr int CH = KP / kElemsPerChunk;                       // == chunks (checked at launch)
        static_assert(CH % NH == 0, "the threads of a row take the same number of chunks each");
        constexpr int HC = CH / NH, HK = HC * kElemsPerChunk;
        const int k0 = half * HK;
        float v[HK];
#pragma unroll
        for (int k = 0; k < HK; ++k) v[k] = (live && k0 + k < feats) ? row[k0 + k] : 0.f;
#pragma unroll
        for (int cc = 0; cc < HC; ++cc) put_chunk(half * HC + cc, v + cc * kElemsPerChunk);
      } else {
        for (int c = half; c < chunks; c += NH) {
          float v[8];
#pragma unroll
          for (int e = 0; e < kElemsPerChunk; ++e) {
            const int k = c * kElemsPerChunk + e;
            v[e] = (live && k < feats) ? row[k] : 0.f;
          }
          put_chunk(c, v);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();                                                // B is complete, nobody reads the staged rows any more
    if (tile + (int)gridDim.x < n_tiles) prefetch(tile + gridDim.x);
    for (int y = 0; y < ny; ++y) {
      // ---- one thread issues the MMAs of this hidden chunk; completion arrives on the mbarrier
      if (warp == 0) {                                    // converged here (just past the barrier): one elected lane issues
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t ah = smem_u32(smem_raw) + y * a_bytes, al = ah + op_bytes, bh = smem_u32(b_hi), bl = smem_u32(b_lo);
        if (tc_elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t off = ks * 2 * kTcChunkBytes;
            if (BF16) {
              umma_bf16(tmem_base, umma_desc(ah + off), umma_desc(bh + off), idesc, ks > 0);
            } else {
              umma_tf32(tmem_base, umma_desc(ah + off), umma_desc(bh + off), idesc, ks > 0);
              umma_tf32(tmem_base, umma_desc(ah + off), umma_desc(bl + off), idesc, 1);
              umma_tf32(tmem_base, umma_desc(al + off), umma_desc(bh + off), idesc, 1);
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(mbar), phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;");
      // ---- epilogue: lane = channel, column = member row; pool per station (members summed in index order)
      const int col = c0 + y * 128 + lane_row;
      const float my_bias = bias[y * 128 + lane_row];
      const uint32_t taddr0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      if constexpr (MEMBERS > 0) {
        // thread `half` of a channel takes columns [W*half, W*half + W); with two threads the station that straddles column 64
        // is finished by the lower one from the upper one's partial sum
        constexpr int NPT = kTcRows / MEMBERS;               // == nodes_per_tile (checked at launch)
        constexpr int USED = NPT * MEMBERS;
        constexpr int W = kTcRows / NH, NQ = W / 32;
        constexpr int JS = 63 / MEMBERS;                     // station that owns column 63
        constexpr bool STRADDLE = NH == 2 && (JS + 1) * MEMBERS > 64 && USED > 64;
        float acc[NPT];
#pragma unroll
        for (int j = 0; j < NPT; ++j) acc[j] = 0.f;
        uint32_t r[NQ][32];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
          if (W * half + 32 * q < USED) tmem_ld32(taddr0 + (uint32_t)(W * half + 32 * q), r[q]);
        tmem_ld_wait();
        if (half == 0) {
#pragma unroll
          for (int c = 0; c < (USED < W ? USED : W); ++c) acc[c / MEMBERS] += fmaxf(__uint_as_float(r[c >> 5][c & 31]) + my_bias, 0.f);
        } else {
#pragma unroll
          for (int c = W; c < USED; ++c) acc[c / MEMBERS] += fmaxf(__uint_as_float(r[(c - W) >> 5][c & 31]) + my_bias, 0.f);
          if (STRADDLE) part[lane_row] = acc[JS];
        }
        if (STRADDLE) __syncthreads();
        if (col < hidden) {
#pragma unroll
          for (int j = 0; j < NPT; ++j) {
            const bool lower = NH == 1 || (j + 1) * MEMBERS <= 64 || (STRADDLE && j == JS);     // who writes station j
            if (j < n_nodes && (half == 0) == lower) {
              float v = acc[j];
              if (STRADDLE && j == JS) v += part[lane_row];
              pooled[(size_t)(n0 + j) * hidden + col] = v;
            }
          }
        }
      } else if (half == 0) {
        float sum = 0.f;
        int cnt = 0, node = n0;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          if (q * 32 >= rows) break;
          uint32_t r[32];
          tmem_ld32(taddr0 + (uint32_t)(q * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (q * 32 + i < rows) {
              sum += fmaxf(__uint_as_float(r[i]) + my_bias, 0.f);
              if (++cnt == members) {
                if (col < hidden) pooled[(size_t)node * hidden + col] = sum;
                sum = 0.f; cnt = 0; ++node;
              }
            }
          }
        }
      }
      if (y + 1 < ny) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();                                    // the accumulator is free for the next chunk's MMAs
      }
    }  // hidden chunks
    cp_async_wait_all();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();                                      // TMEM and the B tile are free again, the next tile has landed
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128));
}

static size_t tc_smem_bytes(bool bf16, int feats, int chunks, int ny) {
  const size_t op = (size_t)chunks * kTcChunkBytes;
  return (bf16 ? 1 : 2) * op * (ny + 1) + ((size_t)kTcRows * feats + 8) * sizeof(float) + (ny + 1) * 128 * sizeof(float) + 64;
}

// Is the tensor-core path applicable (and worth it) for this shape?
bool deepsets_tc_applicable(int num_nodes, int members, int feats, int hidden) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("RC_DEEPSETS_TC");
    forced = e ? (atoi(e) ? 1 : 2) : 0;                 // 1 = always when legal, 2 = never, 0 = by size
  }
  const bool legal = members >= 1 && members <= kTcRows && feats >= 1 && feats <= 64 && hidden >= 1;
  if (!legal || forced == 2) return false;
  if (forced == 1) return true;
  // from ~8k member rows the tensor-core kernel also wins at the reference shape (8 graphs x 122 stations x 11 members =
  // 10 736 rows: the training step went from 0.376 to 0.362 ms with both DeepSets kernels on tcgen05)
  return (long long)num_nodes * members >= 8192 && hidden % 128 == 0;
}

template <bool BF16, int MEMBERS, int KQ>
static int launch_tc_inst(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes, int members,
                          int feats, int hidden, int npt, int chunks, int ny, dim3 grid, size_t smem, cudaStream_t s) {
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(deepsets_pool_fwd_tc_kernel<BF16, MEMBERS, KQ>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "deepsets tensor-core path: %s", cudaGetErrorString(e));
    attr = smem;
  }
  deepsets_pool_fwd_tc_kernel<BF16, MEMBERS, KQ><<<grid, 128 * tc_threads_per_row(BF16), smem, s>>>(ens, w1, b1, pooled, num_nodes, members, feats,
                                                                              hidden, npt, chunks, ny);
  return check_launch("deepsets_pool_fwd_tc_kernel");
}

template <bool BF16, int MEMBERS>
static int launch_tc_members(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes, int members,
                             int feats, int hidden, int npt, int chunks, int ny, dim3 grid, size_t smem, cudaStream_t s) {
  if (ceil_div(feats, 8) == 5)           // 33..40 features (the reference has 35)
    return launch_tc_inst<BF16, MEMBERS, 5>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s);
  return launch_tc_inst<BF16, MEMBERS, 0>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s);
}

template <bool BF16>
static int launch_tc(const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes, int members, int feats,
                     int hidden, int npt, int chunks, int ny, dim3 grid, size_t smem, cudaStream_t s) {
  switch (members) {                     // the reference's ensembles: 11 reforecast members, 51 forecast members
    case 11: return launch_tc_members<BF16, 11>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s);
    case 51: return launch_tc_members<BF16, 51>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s);
    default: return launch_tc_members<BF16, 0>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s);
  }
}

int launch_deepsets_fwd_tc(bool bf16, const float* ens, const float* w1, const float* b1, float* pooled, int num_nodes,
                           int members, int feats, int hidden, cudaStream_t s) {
  const int kp = bf16 ? ((feats + 15) / 16) * 16 : ((feats + 7) / 8) * 8;
  const int chunks = bf16 ? kp / 8 : kp / 4;
  // hidden chunks inside the CTA (member rows converted once per tile) where all the W1 tiles fit beside two CTAs per SM
  const int hchunks = ceil_div(hidden, 128);
  const int ny = (bf16 && tc_smem_bytes(bf16, feats, chunks, hchunks) <= 110 * 1024) ? hchunks : 1;
  const size_t smem = tc_smem_bytes(bf16, feats, chunks, ny);
  const int npt = kTcRows / members;
  const int n_tiles = ceil_div(num_nodes, npt);
  int gx = 2 * kNumSMs;
  if (gx > n_tiles) gx = n_tiles;
  dim3 grid(gx, ceil_div(hchunks, ny));
  return bf16 ? launch_tc<true>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s)
              : launch_tc<false>(ens, w1, b1, pooled, num_nodes, members, feats, hidden, npt, chunks, ny, grid, smem, s);
}

}  // namespace rc
