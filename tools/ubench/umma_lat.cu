// How long does a chain of dependent tcgen05.mma (same accumulator) take, against the same MMAs spread over 2 / 4
// independent accumulators?  (M = 128, kind::tf32, K = 8; A from shared memory or TMEM; N = 48 / 64 / 128)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ubench/umma_lat tools/ubench/umma_lat.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128) lat(int N, int ts, int chains, int count, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = desc(smem_u32(smem), 128, 1024), db = desc(smem_u32(smem) + 16384, 128, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t d = tmem + (i % chains) * 128;
      if (ts) mma_ts(d, tmem + 504, db, idesc, i >= chains);
      else mma_ss(d, da, db, idesc, i >= chains);
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const int Ns[3] = {48, 64, 128};
  for (int ts = 0; ts < 2; ++ts)
    for (int ni = 0; ni < 3; ++ni)
      for (int chains = 1; chains <= 3; ++chains) {
        if (Ns[ni] == 128 && chains == 3 && ts) continue;
        for (int rep = 0; rep < 2; ++rep) {
          lat<<<1, 128, 49152>>>(Ns[ni], ts, chains, 48, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("A from %s, N=%3d, %d chain(s): 48 MMAs issued in %5lld cycles, complete after %5lld  (%.1f cycles per MMA)\n", ts ? "TMEM" : "smem", Ns[ni], chains,
               h[0], h[1], h[1] / 48.0);
      }
  return 0;
}
