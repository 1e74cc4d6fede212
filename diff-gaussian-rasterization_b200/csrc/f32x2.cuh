// f32x2.cuh — Blackwell packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2: one issue slot for two fp32
// lanes; IEEE round-to-nearest per element, i.e. bit-identical to the scalar __f*_rn intrinsics) and the
// small shared-memory / register helpers of the packed blend kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsr {

typedef unsigned long long f2;  // two packed floats (lo, hi)

__device__ __forceinline__ f2 f2_pack(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// acc += a * b, written in place (destination = addend register: no copy on loop-carried accumulators)
__device__ __forceinline__ void f2_fma_acc(f2& acc, f2 a, f2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void f2_mul_acc(f2& acc, f2 a) {
  asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a));
}
__device__ __forceinline__ float f2_lo(f2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float f2_hi(f2 v) { return __uint_as_float((unsigned)(v >> 32)); }

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// Opaque copy: the value stays in a register; without it the compiler rebuilds per-thread shared
// addresses from threadIdx inside the hot loop (5-8 instructions each time it needs one).
__device__ __forceinline__ unsigned pin_reg(unsigned v) {
  unsigned r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

}  // namespace gsr
