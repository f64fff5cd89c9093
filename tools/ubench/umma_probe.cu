// UMMA layout probe: one 128 x N x 32 TF32 tcgen05.mma per variant, checked against the CPU, to pin down
// the shared-memory descriptor / instruction-descriptor encodings the kernels in raincast_gnn_b200/csrc rely on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench/umma_probe tools/ubench/umma_probe.cu
// Variants (A is 128 x 32, B is N x 32, D = A B^T, small integers so that TF32 is exact):
//   0  K-major, no swizzle, chunk-major   (LBO = rows*16, SBO = 128)          - the layout rc_deepsets_tc.cu uses
//   1  K-major, no swizzle, group-major   (LBO = 128, SBO = chunks*128)
//   2  K-major, 128-byte swizzle          (SBO = 1024)
//   3  B MN-major, 128-byte swizzle
//   4  A MN-major, 128-byte swizzle
//   5  A from TMEM (tcgen05.st), B as variant 0
//   6  B MN-major, no swizzle             (SBO = 128 between MN groups of 4, LBO between K groups of 8)
//   7  A MN-major, no swizzle
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

constexpr int M = 128, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void umma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
               "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

// byte offset of element (row, k) of an operand tile with `rows` rows
__device__ __host__ inline int off_k_chunk(int row, int k, int rows) { return (k / 4) * rows * 16 + row * 16 + (k % 4) * 4; }
__device__ __host__ inline int off_k_group(int row, int k) { return (row / 8) * (K / 4) * 128 + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4; }
__device__ __host__ inline int off_k_sw128(int row, int k) { return (row / 8) * 1024 + (row % 8) * 128 + (((k / 4) ^ (row % 8)) * 16) + (k % 4) * 4; }
__device__ __host__ inline int off_mn_sw128(int row, int k) {   // row = M/N index, stored transposed: 128-byte lines of 32 MN elements
  return (row / 32) * (K / 8) * 1024 + (k / 8) * 1024 + (k % 8) * 128 + ((((row % 32) / 4) ^ (k % 8)) * 16) + (row % 4) * 4;
}
__device__ __host__ inline int off_mn_none(int row, int k, int rows) {   // core matrix: 8 k x 4 MN elements
  return (row / 4) * 128 + (k / 8) * (rows / 4) * 128 + (k % 8) * 16 + (row % 4) * 4;
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int N, int variant) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);   // swizzle atoms want 1024-byte alignment
  unsigned char* sa = smem;                 // 16 KB
  unsigned char* sb = smem + 16384;         // 32 KB (N <= 256)
  uint64_t& mbar = *reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(smem + 49152 + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  const bool a_mn = variant == 4 || variant == 7, b_mn = variant == 3 || variant == 6;
  for (int i = tid; i < M * K; i += 128) {
    const int r = i / K, k = i % K;
    int o;
    switch (variant) {
      case 1: o = off_k_group(r, k); break;
      case 2: case 3: o = off_k_sw128(r, k); break;
      case 4: o = off_mn_sw128(r, k); break;
      case 7: o = off_mn_none(r, k, M); break;
      default: o = off_k_chunk(r, k, M); break;
    }
    *reinterpret_cast<float*>(sa + o) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    int o;
    switch (variant) {
      case 1: o = off_k_group(r, k); break;
      case 2: case 4: o = off_k_sw128(r, k); break;
      case 3: o = off_mn_sw128(r, k); break;
      case 6: o = off_mn_none(r, k, N); break;
      default: o = off_k_chunk(r, k, N); break;
    }
    *reinterpret_cast<float*>(sb + o) = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  const uint32_t a_tmem = tmem + 256;        // columns 256.. hold A for variant 5
  if (variant == 5) {
    // lane = row of A, column = k: each thread stores its row's 32 values (4 x 8 registers)
    const uint32_t taddr = a_tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < K; c += 8) {
      uint32_t v[8];
      for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(A[tid * K + c + e]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr + c), "r"(v[0]), "r"(v[1]),
                   "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (a_mn) idesc |= 1u << 15;
  if (b_mn) idesc |= 1u << 16;
  if (tid == 0) {
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da, db;
      const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sb);
      switch (variant) {
        case 1:
          da = desc(a0 + ks * 256, 128, (K / 4) * 128, 0); db = desc(b0 + ks * 256, 128, (K / 4) * 128, 0); break;
        case 2:
          da = desc(a0 + ks * 32, 16, 1024, 2); db = desc(b0 + ks * 32, 16, 1024, 2); break;
        case 3:
          da = desc(a0 + ks * 32, 16, 1024, 2); db = desc(b0 + ks * 1024, (K / 8) * 1024, 1024, 2); break;
        case 4:
          da = desc(a0 + ks * 1024, (K / 8) * 1024, 1024, 2); db = desc(b0 + ks * 32, 16, 1024, 2); break;
        case 6:
          da = desc(a0 + ks * 2 * M * 16, M * 16, 128, 0); db = desc(b0 + ks * (N / 4) * 128, (N / 4) * 128, 128, 0); break;
        case 7:
          da = desc(a0 + ks * (M / 4) * 128, (M / 4) * 128, 128, 0); db = desc(b0 + ks * 2 * N * 16, N * 16, 128, 0); break;
        default:
          da = desc(a0 + ks * 2 * M * 16, M * 16, 128, 0); db = desc(b0 + ks * 2 * N * 16, N * 16, 128, 0); break;
      }
      if (variant == 5) umma_tf32_ts(tmem, a_tmem + ks * 8, db, idesc, ks > 0);
      else umma_tf32_ss(tmem, da, db, idesc, ks > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 8; ++e) D[tid * N + c + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  const int Ns[2] = {128, 48};
  float *A, *B, *D, *dA, *dB, *dD;
  A = (float*)malloc(M * K * 4); B = (float*)malloc(256 * K * 4); D = (float*)malloc(M * 256 * 4);
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, 256 * K * 4); cudaMalloc(&dD, M * 256 * 4);
  srand(1);
  for (int i = 0; i < M * K; ++i) A[i] = (float)(rand() % 17 - 8);
  for (int i = 0; i < 256 * K; ++i) B[i] = (float)(rand() % 13 - 6);
  cudaMemcpy(dA, A, M * K * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B, 256 * K * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 2048);
  for (int ni = 0; ni < 2; ++ni) {
    const int N = Ns[ni];
    for (int v = 0; v < 8; ++v) {
      if ((v == 3 || v == 6) && N % 32 != 0 && v == 3) continue;   // swizzled MN-major B wants whole 32-element blocks
      cudaMemset(dD, 0xff, M * 256 * 4);
      probe<<<1, 128, 49152 + 2048, 0>>>(dA, dB, dD, N, v);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d N=%d: CUDA error %s\n", v, N, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D, dD, M * N * 4, cudaMemcpyDeviceToHost);
      double worst = 0; int bad = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
          const double d = fabs(ref - D[m * N + n]);
          if (!(d <= 1e-3)) ++bad;
          if (d > worst || d != d) worst = d;
        }
      printf("variant %d N=%3d: %s (max |err| %.3g, %d bad of %d)\n", v, N, bad ? "MISMATCH" : "ok", worst, bad, M * N);
      if (bad) {
        for (int m = 0; m < 2; ++m) {
          printf("   D[%d][0..7] =", m);
          for (int n = 0; n < 8; ++n) printf(" %g", D[m * N + n]);
          printf("   ref =");
          for (int n = 0; n < 8; ++n) { double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k]; printf(" %g", ref); }
          printf("\n");
        }
      }
    }
  }
  return 0;
}
