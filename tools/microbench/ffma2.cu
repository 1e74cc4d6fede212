// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a,
// alone and mixed with integer ALU work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  unsigned long long p0 = threadIdx.x, p1 = p0 + 1, p2 = p0 + 2, p3 = p0 + 3;
  unsigned long long ps = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  int i0 = threadIdx.x, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0 || MODE == 2) {  // 8 scalar FFMA
      a0 = fmaf(a0, s, s); a1 = fmaf(a1, s, s); a2 = fmaf(a2, s, s); a3 = fmaf(a3, s, s);
      a4 = fmaf(a4, s, s); a5 = fmaf(a5, s, s); a6 = fmaf(a6, s, s); a7 = fmaf(a7, s, s);
    }
    if (MODE == 1 || MODE == 3) {  // 4 packed FFMA2 (= 8 fp32 FMAs)
      p0 = fma2(p0, ps, ps); p1 = fma2(p1, ps, ps); p2 = fma2(p2, ps, ps); p3 = fma2(p3, ps, ps);
    }
    if (MODE == 2 || MODE == 3) {  // + 4 integer ALU ops
      i0 = (i0 ^ i) + 1; i1 = (i1 ^ i) + 3; i2 = (i2 ^ i) + 5; i3 = (i3 ^ i) + 7;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(p0 ^ p1 ^ p2 ^ p3) + i0 + i1 + i2 + i3;
}

int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"8xFFMA", "4xFFMA2", "8xFFMA+4xINT(8 ops)", "4xFFMA2+4xINT(8 ops)"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f);
      if (mode == 1) k<1><<<148 * 8, 256>>>(d, iters, 1.0001f);
      if (mode == 2) k<2><<<148 * 8, 256>>>(d, iters, 1.0001f);
      if (mode == 3) k<3><<<148 * 8, 256>>>(d, iters, 1.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1) {
        double fmas = 148.0 * 8 * 256 * (double)iters * 8;
        printf("%-24s %.3f ms  %.2f fp32-FMA/clk/SM (at 1.965 GHz)\n", names[mode], ms,
               fmas / (ms * 1e-3) / 148 / 1.965e9);
      }
    }
  }
  return 0;
}
