// Micro-benchmark: issue cost of the instruction mixes of the GINE aggregation inner loop on one SM sub-partition.
// Each variant runs ITER iterations of an unrolled body on 8 warps per sub-partition (32 warps per SM, 148 CTAs) and
// reports cycles per warp-instruction per sub-partition.   nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int V>
__global__ void __launch_bounds__(1024, 1) bench(float* out, float seed, long long* cycles) {
  float2 a[8];
  float2 w = make_float2(seed, seed * 0.5f), nb = make_float2(-seed, -0.25f * seed);
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed * (i + 1) + threadIdx.x, seed * (i + 2));
  float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (V == 0) {            // FMNMX reg, reg  x2
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(nb.x));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].y) : "f"(nb.y));
      } else if (V == 1) {     // FMNMX reg, 0  x2
        asm volatile("max.f32 %0, %0, 0f00000000;" : "+f"(a[i].x));
        asm volatile("max.f32 %0, %0, 0f00000000;" : "+f"(a[i].y));
      } else if (V == 2) {     // FFMA2 x2
        a[i] = __ffma2_rn(a[i], w, nb);
        a[(i + 4) & 7] = __ffma2_rn(a[(i + 4) & 7], w, nb);
      } else if (V == 3) {     // FADD2 x2
        a[i] = __fadd2_rn(a[i], w);
        a[(i + 4) & 7] = __fadd2_rn(a[(i + 4) & 7], nb);
      } else if (V == 4) {     // the forward pair body as shipped: FFMA2 + 2 FMNMX + FADD2 (per 2 columns)
        float2 z = __ffma2_rn(a[i], w, acc[i & 3]);
        asm volatile("max.f32 %0, %0, %1;" : "+f"(z.x) : "f"(nb.x));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(z.y) : "f"(nb.y));
        acc[i & 3] = __fadd2_rn(acc[i & 3], z);
      } else if (V == 5) {     // scalar FFMA x2 (3 registers)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(w.x), "f"(nb.x));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(w.y), "f"(nb.y));
      } else if (V == 6) {     // saturating scalar FFMA x2 + FADD2
        float2 z;
        asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(z.x) : "f"(a[i].x), "f"(w.x), "f"(acc[i & 3].x));
        asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(z.y) : "f"(a[i].y), "f"(w.y), "f"(acc[i & 3].y));
        acc[i & 3] = __fadd2_rn(acc[i & 3], z);
      } else if (V == 12) {    // saturating scalar FFMA x2 + scalar FADD x2
        float2 z;
        asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(z.x) : "f"(a[i].x), "f"(w.x), "f"(acc[i & 3].x));
        asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(z.y) : "f"(a[i].y), "f"(w.y), "f"(acc[i & 3].y));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(acc[i & 3].x) : "f"(z.x));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(acc[i & 3].y) : "f"(z.y));
      } else if (V == 13) {    // saturating scalar FFMA x2
        asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(w.x), "f"(nb.x));
        asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(w.y), "f"(nb.y));
      } else if (V == 14) {    // the backward pair body as shipped: FFMA2 z, 2 FSET, FFMA2 acc, FFMA2 s (per 2 columns)
        float2 z = __ffma2_rn(a[i], w, acc[i & 3]);
        float2 m = make_float2(z.x > nb.x ? 1.f : 0.f, z.y > nb.y ? 1.f : 0.f);
        acc[i & 3] = __ffma2_rn(a[(i + 1) & 7], m, acc[i & 3]);
        acc[(i + 1) & 3] = __ffma2_rn(w, m, acc[(i + 1) & 3]);
      } else if (V == 15) {    // backward with the mask from a saturating FFMA: 2 FFMA.SAT + 2 FFMA + 2 FFMA (per 2 columns)
        float2 m;
        asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(m.x) : "f"(a[i].x), "f"(w.x), "f"(acc[i & 3].x));
        asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(m.y) : "f"(a[i].y), "f"(w.y), "f"(acc[i & 3].y));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i & 3].x) : "f"(a[(i + 1) & 7].x), "f"(m.x));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i & 3].y) : "f"(a[(i + 1) & 7].y), "f"(m.y));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[(i + 1) & 3].x) : "f"(w.x), "f"(m.x));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[(i + 1) & 3].y) : "f"(w.y), "f"(m.y));
      } else if (V == 16) {    // scalar forward: 2 FFMA + 2 FMNMX + 2 FADD
        float2 z;
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(z.x) : "f"(a[i].x), "f"(w.x), "f"(acc[i & 3].x));
        asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(z.y) : "f"(a[i].y), "f"(w.y), "f"(acc[i & 3].y));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(z.x) : "f"(nb.x));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(z.y) : "f"(nb.y));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(acc[i & 3].x) : "f"(z.x));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(acc[i & 3].y) : "f"(z.y));
      } else if (V == 7) {     // scalar FADD with |.| x2
        asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+f"(acc[i & 3].x) : "f"(a[i].x));
        asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+f"(acc[i & 3].y) : "f"(a[i].y));
      } else if (V == 8) {     // FFMA2 + 2 FADD|.| + FADD2: relu as (z + |z|) / 2, all on the FMA pipe
        float2 z = __ffma2_rn(a[i], w, acc[(i + 1) & 3]);
        asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+f"(acc[i & 3].x) : "f"(z.x));
        asm volatile("{.reg .f32 t; abs.f32 t, %1; add.f32 %0, %0, t;}" : "+f"(acc[i & 3].y) : "f"(z.y));
        acc[(i + 1) & 3] = __fadd2_rn(acc[(i + 1) & 3], z);
      } else if (V == 9) {     // 2 FMNMX alternating with 2 scalar FFMA
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(nb.x));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i & 3].x) : "f"(w.x), "f"(nb.x));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].y) : "f"(nb.y));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i & 3].y) : "f"(w.y), "f"(nb.y));
      } else if (V == 10) {    // FMNMX alternating with FFMA2 (1 : 1)
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(nb.x));
        acc[i & 3] = __ffma2_rn(acc[i & 3], w, nb);
      } else if (V == 11) {    // relu through the integer pipe: 2 x (shift, and-not) + FFMA2 + FADD2
        float2 z = __ffma2_rn(a[i], w, acc[i & 3]);
        int zx = __float_as_int(z.x), zy = __float_as_int(z.y);
        asm volatile("{.reg .s32 t; shr.s32 t, %0, 31; not.b32 t, t; and.b32 %0, %0, t;}" : "+r"(zx));
        asm volatile("{.reg .s32 t; shr.s32 t, %0, 31; not.b32 t, t; and.b32 %0, %0, t;}" : "+r"(zy));
        acc[i & 3] = __fadd2_rn(acc[i & 3], make_float2(__int_as_float(zx), __int_as_float(zy)));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  for (int i = 0; i < 4; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int instr_per_body, float* out, long long* cyc) {
  bench<V><<<148, 1024>>>(out, 1.25f, cyc);
  bench<V><<<148, 1024>>>(out, 1.25f, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  // 8 warps per sub-partition, 8 bodies per iteration
  const double winst = 8.0 * 8.0 * instr_per_body * ITER;
  printf("%-58s %6.2f cycles per warp-instruction per sub-partition (%d instr per body, %.0f cycles)\n", name, avg / winst, instr_per_body, avg);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<0>("FMNMX reg,reg", 2, out, cyc);
  run<1>("FMNMX reg,0", 2, out, cyc);
  run<2>("FFMA2", 2, out, cyc);
  run<3>("FADD2", 2, out, cyc);
  run<5>("FFMA scalar 3-reg", 2, out, cyc);
  run<7>("FADD scalar |.|", 2, out, cyc);
  run<9>("FMNMX : FFMA scalar 1:1", 4, out, cyc);
  run<10>("FMNMX : FFMA2 1:1", 2, out, cyc);
  run<4>("pair body FFMA2 + 2 FMNMX + FADD2", 4, out, cyc);
  run<16>("scalar pair body 2 FFMA + 2 FMNMX + 2 FADD", 6, out, cyc);
  run<13>("FFMA.SAT scalar", 2, out, cyc);
  run<6>("2 FFMA.SAT + FADD2", 3, out, cyc);
  run<12>("2 FFMA.SAT + 2 FADD", 4, out, cyc);
  run<14>("bwd pair body FFMA2 + 2 FSET + 2 FFMA2", 5, out, cyc);
  run<15>("bwd 2 FFMA.SAT + 4 FFMA", 6, out, cyc);
  run<8>("FFMA2 + 2 FADD|.| + FADD2", 4, out, cyc);
  run<11>("FFMA2 + 2x(SHF,LOP3) + FADD2", 6, out, cyc);
  return 0;
}
