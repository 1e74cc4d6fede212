// Microbenchmark: per-SM-sub-partition throughput of the packed fp32x2 instructions (FFMA2 / FMUL2 /
// FADD2) against their scalar forms and in the mixes the backward blend kernel issues, on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmas(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float muls(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float adds(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

// MODE: 0 8xFFMA2  1 8xFMUL2  2 8xFADD2  3 8xFFMA  4 8xFMUL  5 8xFADD
//       6 4xFFMA2+4xFMUL2  7 4xFFMA2+4xFFMA  8 4xFMUL2+4xFADD  9 4xFFMA2 + 4xLOP(int)  10 4xFFMA2+4xIMAD
//       11 4xFMUL2 + 4x MUFU.EX2   12 8xFFMA + 4xLOP  13 8xFSETP-ish (compare+select)
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  f2 p[8]; float a[8]; int n[4];
  const f2 ps = ((f2)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int i = 0; i < 8; ++i) { p[i] = (f2)(threadIdx.x + i) * 0x0000000100000001ull | 0x3f8000003f800000ull; a[i] = threadIdx.x + i; }
  for (int i = 0; i < 4; ++i) n[i] = threadIdx.x * (2 * i + 3);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) p[i] = fma2(p[i], ps, ps);
      if (MODE == 1) p[i] = mul2(p[i], ps);
      if (MODE == 2) p[i] = add2(p[i], ps);
      if (MODE == 3) a[i] = fmas(a[i], s, s);
      if (MODE == 4) a[i] = muls(a[i], s);
      if (MODE == 5) a[i] = adds(a[i], s);
      if (MODE == 6) p[i] = (i & 1) ? fma2(p[i], ps, ps) : mul2(p[i], ps);
      if (MODE == 7) { if (i & 1) p[i] = fma2(p[i], ps, ps); else a[i] = fmas(a[i], s, s); }
      if (MODE == 8) { if (i & 1) p[i] = mul2(p[i], ps); else a[i] = adds(a[i], s); }
      if (MODE == 9) { if (i & 1) p[i] = fma2(p[i], ps, ps); else n[i >> 1] = (n[i >> 1] ^ it) + 1; }
      if (MODE == 10) { if (i & 1) p[i] = fma2(p[i], ps, ps); else n[i >> 1] = n[i >> 1] * 3 + it; }
      if (MODE == 11) { if (i & 1) p[i] = mul2(p[i], ps); else a[i] = exp2f(a[i]); }
      if (MODE == 12) { a[i] = fmas(a[i], s, s); if (i & 1) n[i >> 1] = (n[i >> 1] ^ it) + 1; }
      if (MODE == 13) a[i] = (a[i] > s) ? a[i] - 1.0f : s;
    }
  }
  float r = 0.f;
  for (int i = 0; i < 8; ++i) r += a[i] + (float)(p[i] & 0xffff);
  for (int i = 0; i < 4; ++i) r += n[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(float* d, const char* name, double issue_per_iter) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  // warp instructions issued per scheduler per clock: (148*8 blocks * 8 warps * iters * instr) / (148 SMs * 4 schedulers)
  const double warps = 148.0 * 8 * 8, clk = ms * 1e-3 * 1.965e9;
  printf("%-28s %.3f ms   %.3f warp-instr/clk/scheduler\n", name, ms, warps * iters * issue_per_iter / (148.0 * 4) / clk);
}

int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
  run<0>(d, "8 x FFMA2", 8); run<1>(d, "8 x FMUL2", 8); run<2>(d, "8 x FADD2", 8);
  run<3>(d, "8 x FFMA", 8); run<4>(d, "8 x FMUL", 8); run<5>(d, "8 x FADD", 8);
  run<6>(d, "4 FFMA2 + 4 FMUL2", 8); run<7>(d, "4 FFMA2 + 4 FFMA", 8); run<8>(d, "4 FMUL2 + 4 FADD", 8);
  run<9>(d, "4 FFMA2 + 4x(LOP+IADD)", 12); run<10>(d, "4 FFMA2 + 4 IMAD", 8);
  run<11>(d, "4 FMUL2 + 4 EX2(+fixup)", 8); run<12>(d, "8 FFMA + 4x(LOP+IADD)", 16); run<13>(d, "8 x (FSETP+FADD+SEL)", 24);
  return 0;
}
