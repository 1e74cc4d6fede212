// gsr_common.cuh — shared configuration, buffer layouts and device math for the B200-native
// Gaussian-splat rasterizer core.  No GLM: 3x3 helpers below are column-major like the
// reference's glm::mat3 so that floating-point association matches
// (reference: cuda_rasterizer/forward.cu:74-152, auxiliary.h:41-164).
#pragma once

#include <cuda_runtime.h>
#include <functional>
#include <cstring>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#include "../../include/gsr_b200.h"

namespace gsr {

// ---- compile-time configuration (reference: cuda_rasterizer/config.h:15-17) ------------------
constexpr int kChannels = 3;
constexpr int kTileX = 16;
constexpr int kTileY = 16;
constexpr int kTileThreads = kTileX * kTileY;  // one thread per pixel
constexpr float kAlphaMin = 15.0f / 255.0f;    // reference forward.cu:360 (Inria uses 1/255)
constexpr float kAlphaMax = 0.99f;
constexpr float kTmin = 0.0001f;
constexpr int kAccStride = 16;                 // floats per Gaussian in the backward accumulator
constexpr int kCntStrideMax = 32;              // the per-tile counters / cursors may be spread one per 32-byte
                                               // sector (option "cnt_stride") to relieve same-line atomics

enum Variant { kLight = 0, kFull = 1 };

// accumulator slots (per Gaussian, written by the backward blend kernel with warp-reduced
// atomics, consumed by the per-Gaussian backward kernel)
// With w = G * dL/dalpha per (pixel, Gaussian) pair and (dx, dy) = mean2D - pixel:
enum AccSlot {
  ACC_MX = 0, ACC_MY = 1,         // S1 = sum w dx, S2 = sum w dy          -> dL/dmean2D
  ACC_CA = 2, ACC_CB = 3, ACC_CC = 4,  // S11, S12, S22 = sum w {dx^2, dx dy, dy^2} -> dL/dconic
  ACC_OP = 5,                     // S0 = sum w = dL/dopacity
  ACC_R = 6, ACC_G = 7, ACC_B = 8,  // dL/dcolor (sum alpha*T*dL/dpixel)
  ACC_DEPTH = 9,                  // dL/ddepth (incl. the (depth-gt)^2 term)
  ACC_PGX = 10, ACC_PGY = 11,     // full: Q1, Q2 = sum q {dx, dy}, q = G * (pose weight of the pair)
  ACC_PD = 12,                    // pose: sum alpha*T*dL/dD (light: all pairs; full: front-most)
  ACC_MED = 13                    // light: sum of dL/dmedian over pixels that picked this Gaussian
};

// ---- option flags --------------------------------------------------------------------------
struct Options {
  int exact_ng;
  int tight_tiles;
  int stage_timing;
  int tile_sort;
  int bwd_packed;
  int async_binning;
  int track_headroom_pct;
  int bulk_sh;
  int cnt_stride;
  int bwd_occ;
  int fwd_packed;
  int spec_render;   // 1: forward blend enqueued before the host waits for the duplicate count
  int early_acc_clear;  // 1: the forward clears the backward's accumulator on a side stream (acc_clear_begin)
  int pdl;              // 1: dependent kernels of a frame are launched programmatically (launch_after)
  int exact_median;     // 1: -light backward restores T with the reference's exact arithmetic (render_bwdq_kernel<..., EXACT>)
};
Options& options();  // the calling thread's snapshot (see OptionsCall)
// RAII at the top of every extern "C" entry point: copies the process-wide option defaults into the
// calling thread's snapshot and opens an NVTX range named after the entry point.
struct OptionsCall {
  explicit OptionsCall(const char* entry_point);
  ~OptionsCall();
};
inline int cnt_stride() { const int s = options().cnt_stride; return (s >= 1 && s <= kCntStrideMax) ? s : 1; }

// ---- stage timing (gsr_stage_times in the C ABI) ---------------------------------------------
enum Stage {
  ST_PRE_FWD = 0,   // per-Gaussian forward
  ST_SCAN,          // prefix sum of tiles_touched (+ the count read-back)
  ST_EMIT,          // (tile | depth) key emission
  ST_SORT,          // device radix sort
  ST_RANGES,        // per-tile ranges
  ST_RENDER_FWD,    // forward tile blend
  ST_RENDER_BWD,    // backward tile blend
  ST_PRE_BWD,       // per-Gaussian backward + pose reduction
  ST_MEMSET,        // zero-fills issued by the library
  ST_OTHER,
  ST_COUNT
};
static_assert(ST_COUNT == GSR_STAGE_COUNT, "stage table out of sync with gsr_b200.h");

// RAII scope: records a start/stop event pair on `stream` when options().stage_timing is set;
// `launches` = number of kernels launched inside the scope.
struct StageScope {
  StageScope(int stage, cudaStream_t stream, int launches = 1);
  ~StageScope();
  int slot;
  cudaStream_t stream;
};

// One-time (per kernel and device) request for the maximum shared-memory carve-out: the attribute call
// takes a driver lock, so it is kept off the per-frame path.
void prefer_max_shared_once(const void* kernel);

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define GSR_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      gsr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                  \
                     cudaGetErrorString(_e));                                              \
      return GSR_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)
#define GSR_LAUNCH_OK(debug, stream)                                                       \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(stream);                  \
    if (_e != cudaSuccess) {                                                               \
      gsr::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,              \
                     cudaGetErrorString(_e));                                              \
      return GSR_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

// ---- private buffer layouts ----------------------------------------------------------------
// Every array starts on a 256-byte boundary inside the caller-provided chunk.
struct Carver {
  char* p;
  size_t used;
  explicit Carver(char* base) : p(base), used(0) {}
  template <typename T>
  T* take(size_t count) {
    used = (used + 255) & ~size_t(255);
    T* r = p ? reinterpret_cast<T*>(p + used) : nullptr;
    used += count * sizeof(T);
    return r;
  }
};

// Per-Gaussian state produced by preprocess_fwd and read by binning / blend / backward.
struct GeomState {
  float4* rec;             // [3P] packed blend record:
                           //   rec[3i+0] = (x_pix, y_pix, conic.a, conic.b)
                           //   rec[3i+1] = (conic.c, opacity, power_cut, depth)
                           //   rec[3i+2] = (r, g, b, power_sure)
  float* cov3D;            // [6P] (only when computed from scale/rotation)
  unsigned char* clamped;  // [P] bit k set <=> colour channel k clamped at 0
  uint32_t* tiles_touched; // [P]
  uint2* rect;             // [P] tile rectangle: x = xmin | xmax << 16, y = ymin | ymax << 16
  uint32_t* offsets;       // [P] inclusive scan of tiles_touched
  uint32_t* counters;      // [8]  0: num_rendered, 1: num_related
  float* acc;              // [16P + 16] the backward's accumulator lines + its last-block counter line, cleared by
                           //            the FORWARD on a side stream (acc_clear_begin) so that the backward finds
                           //            them zeroed instead of issuing a memset in front of its blend kernel
  char* scan_temp;
  size_t scan_bytes;
  static size_t carve(GeomState& s, char* base, int P, size_t scan_bytes);
};

struct BinState {
  uint32_t* vals;           // [N] sorted Gaussian index per (tile, depth) entry (always first)
  uint64_t* keys_unsorted;  // [N] tile-local path: (depth bits << 32 | index) per entry, grouped by tile
                            //     radix path: (tile << 32 | depth bits)
  uint64_t* keys;           // [N] radix path only
  uint32_t* vals_unsorted;  // [N] radix path only
  char* sort_temp;          //     radix path only
  size_t sort_bytes;
  static size_t carve(BinState& s, char* base, size_t N, size_t sort_bytes, bool radix);
};

struct ImgState {
  uint2* ranges;            // [tiles] [start, end) into BinState::vals
  uint32_t* n_contrib;      // [HW] 1-based list position of the last blended entry
  float* final_T;           // [HW] (full only)
  uint32_t* first_contrib;  // [HW] 1-based list position of the first blended entry (full only)
  uint32_t* tile_last;      // [tiles] max n_contrib over the tile's pixels
  uint32_t* tile_count;     // [tiles] entries per tile (tile-local binning)
  uint32_t* tile_fill;      // [tiles] scatter cursor per tile
  static size_t carve(ImgState& s, char* base, int HW, int tiles, int variant);
};
// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// A kernel launched with launch_after(true, ...) right behind another kernel on the same stream may be set up —
// and its CTAs scheduled onto free SM slots — while the tail of that kernel is still running; its first statement
// is pdl_wait(), which returns once the preceding kernel has completed and its writes are visible.  The preceding
// kernel calls pdl_trigger() at its start (it fires when every one of its CTAs has been scheduled).  Hides the
// launch latency between the frame's dependent kernels (nine launches of 4-5 us around kernels of 10-100 us on
// the small frames).  Both instructions are no-ops for kernels that were launched the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_after(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                         Args... args) {
  if (!pdl) {
    kernel<<<grid, block, smem, s>>>(args...);
    return;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, args...);   // (errors are picked up by GSR_LAUNCH_OK's cudaGetLastError)
}
#endif

// Accumulator clear moved out of the backward's critical path: the forward enqueues the memset of GeomState::acc on
// a library-owned side stream (ordered after everything already on `stream`, i.e. after the previous frame), where
// it runs underneath the forward's kernels; the backward of the same geometry buffer makes its stream wait for it
// (acc_clear_join returns true) or, if nothing is pending for that buffer (second backward of one forward, ring
// overrun, option off), clears its scratch itself as before.
void acc_clear_begin(const void* geom_key, void* acc, size_t bytes, cudaStream_t stream);
bool acc_clear_join(const void* geom_key, cudaStream_t stream);
// End of the forward: everything enqueued on `stream` from here on is ordered after the pending clear of this
// buffer (it has long finished by then — the point is that the caller's allocator, which only knows `stream`,
// may hand the geometry buffer's memory to somebody else if the caller drops it without running a backward).
void acc_clear_rejoin(const void* geom_key, cudaStream_t stream);

size_t scan_temp_bytes(int P);
size_t sort_temp_bytes(size_t N, int end_bit);

// ---- device math ---------------------------------------------------------------------------
#ifdef __CUDACC__

// ---- pinned arithmetic -----------------------------------------------------------------------
// The blend has hard thresholds (alpha < 15/255, T < 1e-4, T crossing 0.5, radius ceil, tile rect
// casts) and -light's backward restores T from 1 - sum(alpha*T), so a 1-ulp difference in conic /
// alpha against the reference is amplified to ~1e-3 in gradients.  To be bit-identical to the
// reference's sm_100 build the per-Gaussian state and the per-pair power / alpha are written with
// explicit round-to-nearest intrinsics (no compiler-chosen FMA contraction) in exactly the
// association nvcc 12.9 gives the reference sources (read off its SASS; DESIGN.md "bit parity").
#define GSR_MUL(a, b) __fmul_rn((a), (b))
#define GSR_ADD(a, b) __fadd_rn((a), (b))
#define GSR_SUB(a, b) __fsub_rn((a), (b))
#define GSR_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define GSR_DIV(a, b) __fdiv_rn((a), (b))
#define GSR_RCP(a) __frcp_rn((a))
#define GSR_SQRT(a) __fsqrt_rn((a))

// a0*b0 + a1*b1 + a2*b2 the way the reference build evaluates every 3-term product sum:
// the MIDDLE product is rounded, the first and the last are fused.
__device__ __forceinline__ float dot3_mid(float a0, float b0, float a1, float b1, float a2, float b2) {
  return GSR_FMA(a2, b2, GSR_FMA(a0, b0, GSR_MUL(a1, b1)));
}

struct M3 {  // column-major: c[col][row], same convention as glm::mat3
  float c[3][3];
};

// generic column-major product, compiler-contracted (backward only: tolerance 1e-3)
__device__ __forceinline__ M3 m3_mul(const M3& A, const M3& B) {
  M3 R;
#pragma unroll
  for (int col = 0; col < 3; ++col)
#pragma unroll
    for (int row = 0; row < 3; ++row)
      R.c[col][row] = A.c[0][row] * B.c[col][0] + A.c[1][row] * B.c[col][1] + A.c[2][row] * B.c[col][2];
  return R;
}

__device__ __forceinline__ M3 m3_transpose(const M3& A) {
  M3 R;
#pragma unroll
  for (int col = 0; col < 3; ++col)
#pragma unroll
    for (int row = 0; row < 3; ++row) R.c[col][row] = A.c[row][col];
  return R;
}

// row k of  M * (p,1)  for a column-major 4x4:  m[k]*x + m[4+k]*y + m[8+k]*z + m[12+k]
__device__ __forceinline__ float xform_row(const float* m, int k, const float3& p) {
  return GSR_ADD(dot3_mid(m[k], p.x, m[4 + k], p.y, m[8 + k], p.z), m[12 + k]);
}

__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* m) {
  return make_float3(xform_row(m, 0, p), xform_row(m, 1, p), xform_row(m, 2, p));
}

__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* m) {
  return make_float4(xform_row(m, 0, p), xform_row(m, 1, p), xform_row(m, 2, p), xform_row(m, 3, p));
}

// Per-pair Gaussian exponent  -0.5*(A dx^2 + C dy^2) - B dx dy  (forward.cu:351), pinned:
// fma(fma(dx, dx*A, dy*(dy*C)), -0.5, -(dy*(dx*B))).  Shared by forward and backward so both make
// the same skip decisions.
__device__ __forceinline__ float pair_power(float A, float B, float C, float dx, float dy) {
  const float q = GSR_FMA(dx, GSR_MUL(dx, A), GSR_MUL(dy, GSR_MUL(dy, C)));
  return GSR_FMA(q, -0.5f, -GSR_MUL(dy, GSR_MUL(dx, B)));
}

// alpha = min(0.99, opacity * exp(power))
__device__ __forceinline__ float pair_alpha(float opacity, float G) {
  return fminf(kAlphaMax, GSR_MUL(opacity, G));
}

// ndc -> pixel; the reference evaluates this in double (auxiliary.h:41-44) and rounds once.
__device__ __forceinline__ float ndc_to_pix(float v, int S) {
  return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}

// Tile rectangle of a splat (auxiliary.h:46-56).
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy,
                                          uint2& rmin, uint2& rmax) {
  const float r = (float)radius;
  rmin.x = (unsigned)min(gx, max(0, (int)((px - r) / (float)kTileX)));
  rmin.y = (unsigned)min(gy, max(0, (int)((py - r) / (float)kTileY)));
  rmax.x = (unsigned)min(gx, max(0, (int)((px + r + (float)(kTileX - 1)) / (float)kTileX)));
  rmax.y = (unsigned)min(gy, max(0, (int)((py + r + (float)(kTileY - 1)) / (float)kTileY)));
}

// ---- conservative culling of pairs that cannot reach alpha >= 15/255 ---------------------------
// A pair is blended only if power >= power_cut (= log((15/255)/opacity) - 1e-3, see preprocess_fwd),
// i.e. if the pixel lies in the ellipse  A dx^2 + 2 B dx dy + C dy^2 <= -2 power_cut  of the stored
// (rounded) conic.  Its axis-aligned half extents are sqrt(-2 pc C / det), sqrt(-2 pc A / det) with
// det = A C - B^2, evaluated with Kahan's compensated 2x2 determinant so that cancellation in
// needle-shaped splats cannot shrink the box; a relative 1e-4 and absolute 0.02 px pad cover the
// rounding of the per-pair evaluation.  Returns 0: no pixel can pass, 1: box valid, 2: unbounded.
__device__ __forceinline__ int cut_extent(float A, float B, float C, float pc, float& hx, float& hy) {
  if (pc > 0.0f) return 0;
  if (!(pc <= 0.0f)) return 2;  // NaN opacity: keep the reference's behaviour, cull nothing
  const float w = __fmul_rn(B, B);
  const float e = __fmaf_rn(-B, B, w);
  const float f = __fmaf_rn(A, C, -w);
  const float det = __fadd_rn(f, e);
  if (!(det > 0.0f)) return 2;
  const float k = __fdiv_rn(-2.0f * pc, det);
  hx = __fsqrt_ru(__fmul_ru(k, C));
  hy = __fsqrt_ru(__fmul_ru(k, A));
  if (!(hx < 1e30f) || !(hy < 1e30f)) return 2;
  hx = __fmaf_ru(hx, 1.0001f, 0.02f);
  hy = __fmaf_ru(hy, 1.0001f, 0.02f);
  return 1;
}

// 8-bit mask of the 8x4-pixel warp blocks of a 16x16 tile (bit w = block column w & 1, block row
// w >> 1, the pixel-to-warp mapping of the blend kernels) that a splat's cut ellipse may touch.
__device__ __forceinline__ unsigned block_mask8(const float4& r0, const float4& r1, float tx0,
                                                float ty0) {
  float hx, hy;
  const int kind = cut_extent(r0.z, r0.w, r1.x, r1.z, hx, hy);
  if (kind == 0) return 0u;
  if (kind == 2) return 0xFFu;
  const float x0 = r0.x - hx, x1 = r0.x + hx, y0 = r0.y - hy, y1 = r0.y + hy;
  const unsigned col = ((x1 >= tx0 && x0 <= tx0 + 7.0f) ? 1u : 0u) |
                       ((x1 >= tx0 + 8.0f && x0 <= tx0 + 15.0f) ? 2u : 0u);
  unsigned mask = 0u;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float by = ty0 + 4.0f * (float)r;
    if (y1 >= by && y0 <= by + 3.0f) mask |= col << (2 * r);
  }
  return mask;
}

// 16-bit mask of the sixteen 4x4-pixel sub-blocks of a 16x16 tile (bit = 4 * block row + block
// column) that a splat's cut ellipse may touch.  Each half warp of the blend kernels owns one
// sub-block (see sub_block_of) and walks only the entries whose bit is set.
__device__ __forceinline__ unsigned block_mask16(const float4& r0, const float4& r1, float tx0,
                                                 float ty0) {
  float hx, hy;
  const int kind = cut_extent(r0.z, r0.w, r1.x, r1.z, hx, hy);
  if (kind == 0) return 0u;
  if (kind == 2) return 0xFFFFu;
  const float x0 = r0.x - hx, x1 = r0.x + hx, y0 = r0.y - hy, y1 = r0.y + hy;
  unsigned col = 0u, mask = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float bx = tx0 + 4.0f * (float)c;
    if (x1 >= bx && x0 <= bx + 3.0f) col |= 1u << c;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float by = ty0 + 4.0f * (float)r;
    if (y1 >= by && y0 <= by + 3.0f) mask |= col << (4 * r);
  }
  return mask;
}

// Pixel <-> thread mapping of the blend kernels: warp w covers the 8x4 block at column (w & 1),
// row (w >> 1) of the tile; its lower / upper half warp covers the left / right 4x4 sub-block.
__device__ __forceinline__ void pixel_of_thread(int warp, int lane, int& lx, int& ly, int& sub) {
  const int half = lane >> 4;
  lx = (warp & 1) * 8 + half * 4 + (lane & 3);
  ly = (warp >> 1) * 4 + ((lane >> 2) & 3);
  sub = (warp >> 1) * 4 + (warp & 1) * 2 + half;  // bit index in block_mask16
}

__device__ __forceinline__ uint2 pack_rect(uint2 rmin, uint2 rmax) {
  return make_uint2(rmin.x | (rmax.x << 16), rmin.y | (rmax.y << 16));
}

// Coalesced copy of `nrows` rows of m3 floats ([rows, m3] contiguous in global memory) into / out of
// shared memory with row stride m3 + 1 (odd, so that the later one-row-per-thread accesses are bank
// conflict free).  M3 > 0: compile-time row length (divisions become multiply-shifts, and for
// M3 % 4 == 0 a 128-bit vector never straddles rows); M3 == 0: runtime row length.
template <int M3>
__device__ __forceinline__ void rows_to_smem(const float* __restrict__ src, float* smem, int nrows,
                                             int m3_rt, int tid, int nthreads) {
  const int m3 = M3 > 0 ? M3 : m3_rt;
  const int row = m3 + 1;
  const int nfloats = nrows * m3;
  const int nvec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (nfloats >> 2) : 0;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (int v = tid; v < nvec; v += nthreads) {
    const float4 q = __ldg(src4 + v);
    const int f = v << 2;
    if (M3 > 0 && M3 % 4 == 0) {
      const int g = f / m3;
      float* d = smem + g * row + (f - g * m3);
      d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    } else {
      const float vals[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ff = f + k;
        const int g = ff / m3;
        smem[g * row + (ff - g * m3)] = vals[k];
      }
    }
  }
  for (int ff = (nvec << 2) + tid; ff < nfloats; ff += nthreads) {
    const int g = ff / m3;
    smem[g * row + (ff - g * m3)] = __ldg(src + ff);
  }
}

template <int M3>
__device__ __forceinline__ void smem_to_rows(float* __restrict__ dst, const float* smem, int nrows,
                                             int m3_rt, int tid, int nthreads) {
  const int m3 = M3 > 0 ? M3 : m3_rt;
  const int row = m3 + 1;
  const int nfloats = nrows * m3;
  const int nvec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? (nfloats >> 2) : 0;
  float4* dst4 = reinterpret_cast<float4*>(dst);
  for (int v = tid; v < nvec; v += nthreads) {
    const int f = v << 2;
    float4 q;
    if (M3 > 0 && M3 % 4 == 0) {
      const int g = f / m3;
      const float* s = smem + g * row + (f - g * m3);
      q = make_float4(s[0], s[1], s[2], s[3]);
    } else {
      float vals[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ff = f + k;
        const int g = ff / m3;
        vals[k] = smem[g * row + (ff - g * m3)];
      }
      q = make_float4(vals[0], vals[1], vals[2], vals[3]);
    }
    dst4[v] = q;
  }
  for (int ff = (nvec << 2) + tid; ff < nfloats; ff += nthreads) {
    const int g = ff / m3;
    dst[ff] = smem[g * row + (ff - g * m3)];
  }
}

// ---- bulk asynchronous copies (TMA, 1-D) + mbarrier ------------------------------------------------
// The per-Gaussian kernels move each Gaussian's SH row (12*M bytes, 16-byte aligned for M = 16 / 4)
// between global and shared memory with cp.async.bulk: one instruction per row, no register staging,
// completion counted in bytes on an mbarrier (loads) or by bulk groups (stores).  Only rows that are
// needed are moved (culled Gaussians cost no SH traffic).
__device__ __forceinline__ unsigned smem_addr_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_addr_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_addr_u32(bar))
               : "memory");
}
// shared -> global; the issuing thread must have written the row itself (or synchronised) and
// calls bulk_s2g_fence() between its writes and the copy
__device__ __forceinline__ void bulk_s2g_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_addr_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Row stride (floats) of a bulk-copied [rows, M3] slab in shared memory: a multiple of 16 bytes
// (cp.async.bulk) and an ODD multiple, so that one LDS.128 / STS.128 per thread at the same column
// is bank-conflict free across a quarter warp.
__host__ __device__ constexpr int bulk_row_floats(int m3) { return ((m3 / 4) % 2 == 0) ? m3 + 4 : m3; }

// Spherical-harmonics constants (real SH basis up to degree 3, standard values).
__device__ constexpr float kSH0 = 0.28209479177387814f;
__device__ constexpr float kSH1 = 0.4886025119029199f;
__device__ constexpr float kSH2[5] = {1.0925484305920792f, -1.0925484305920792f,
                                      0.31539156525252005f, -1.0925484305920792f,
                                      0.5462742152960396f};
__device__ constexpr float kSH3[7] = {-0.5900435899266435f, 2.890611442640554f,
                                      -0.4570457994644658f, 0.3731763325901154f,
                                      -0.4570457994644658f, 1.445305721320277f,
                                      -0.5900435899266435f};

#endif  // __CUDACC__

// ---- stage launchers (each file implements its kernels + launcher) ---------------------------
struct Camera {
  const float* view;   // [16] device
  const float* proj;   // [16] device
  const float* campos; // [3]  device
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int W, H, grid_x, grid_y;
};

int launch_preprocess_fwd(int P, int D, int M, const float* means3D, const float* scales,
                          float scale_modifier, const float* rotations, const float* opacities,
                          const float* shs, const float* cov3D_precomp,
                          const float* colors_precomp, const Camera& cam, int* radii,
                          GeomState& g, uint32_t* tile_count /* zeroed, or NULL */, bool prefiltered,
                          bool debug, cudaStream_t stream, float* zero_f32 = nullptr /* [P] or NULL */,
                          int* zero_i32 = nullptr /* [P] or NULL */);

// Forward blend launched UNDERNEATH the count read-back (run_binning): when the binning is speculative
// (buffer sized from the previous frame of this context), `launch` is called before the host waits for the
// duplicate count, so that the device still has the scatter, the sort and the blend queued when the host
// wakes up; if the estimate was too small the binning is redone, `reset` clears what the blend accumulates
// and done is false again (the caller then launches the blend as usual).
struct SpecRender {
  std::function<int(const BinState&)> launch;
  std::function<int()> reset;
  bool done = false;
};
int run_binning(int P, const Camera& cam, const int* radii, GeomState& g, gsr_alloc_fn alloc,
                void* alloc_ctx, BinState& b, ImgState& img, int* num_rendered, bool debug,
                cudaStream_t stream,
                struct SpecRender* spec = nullptr);

int launch_render_fwd_light(const Camera& cam, const GeomState& g, const BinState& b,
                            ImgState& img, const float* bg, const float* gt_depth,
                            float* out_color, float* out_depth, float* out_median, float* out_alpha,
                            float* out_var, float* gau_unc, int* gau_px, bool debug,
                            cudaStream_t stream);

// Masked L1 loss fused into the -light forward blend (tracker):
//   L = sum_pix m * (w_color * |C - C_gt|_1 + w_depth * |D - D_gt|),
//   m = (depth_mask == 0 || D_gt > 0) && (alpha > alpha_thresh)        (mask treated as constant)
struct FusedLoss {
  const float* gt_color = nullptr;  // [3,H,W]
  const float* gt_depth = nullptr;  // [H,W]
  float w_color = 0.f, w_depth = 0.f, alpha_thresh = -1.f;
  int depth_mask = 0;
  float* dL_dpix = nullptr;         // out [3,H,W]
  float* dL_ddepth = nullptr;       // out [H,W]
  float* loss_partials = nullptr;   // out [tiles]
};

int launch_render_fwd_light_loss(const Camera& cam, const GeomState& g, const BinState& b,
                                 ImgState& img, const float* bg, float* out_color, float* out_depth,
                                 float* out_median, float* out_alpha, float* out_var,
                                 const FusedLoss& fl, cudaStream_t stream);

int run_binning_static(int P, const Camera& cam, GeomState& g, BinState& b, ImgState& img,
                       uint32_t capacity, uint32_t longest_cap, cudaStream_t stream);

int launch_render_fwd_full(const Camera& cam, const GeomState& g, const BinState& b,
                           ImgState& img, const float* bg, float* out_color, float* out_depth,
                           float* out_unc, bool count_related, bool debug, cudaStream_t stream);

struct BlendGrads {       // per-pixel cotangents
  const float* dL_dpix;   // [3,H,W]
  const float* dL_ddepth; // [H,W]
  const float* dL_dmedian;// [H,W] light
  const float* dL_dvar;   // [H,W] light: depth_var, full: uncertainty
};

int launch_render_bwd(int variant, const Camera& cam, const GeomState& g, const BinState& b,
                      const ImgState& img, const float* bg, const float* gt_depth,
                      const float* alphas /*light*/, const BlendGrads& cot, float* acc,
                      int num_gaussians, int num_entries, bool pose_only, bool debug,
                      cudaStream_t stream);

struct GaussGradOut {
  float* dL_dmean2D; float* dL_dconic; float* dL_dopacity; float* dL_dcolor; float* dL_ddepth;
  float* dL_dmean3D; float* dL_dcov3D; float* dL_dsh; float* dL_dscale; float* dL_drot;
  float* dL_dview;
  float* dL_dcolor_masked;  // optional [P,3]: dL/dcolor with clamped channels zeroed (see gsr_backward_extras)
  float* densify_grad_accum;  // optional in/out [P]: += |dL/dmean2D.xy| for visible Gaussians
  float* densify_denom;       // optional in/out [P]: += 1 for visible Gaussians
  float* max_radii2D;         // optional in/out [P]: max(., radius) for visible Gaussians
};

int launch_preprocess_bwd(int variant, int P, int D, int M, const float* means3D,
                          const int* radii, const float* shs, const float* scales,
                          const float* rotations, float scale_modifier,
                          const float* cov3D_precomp, const Camera& cam, const float* perspec,
                          const GeomState& g, const float* acc, float* pose_partials,
                          const GaussGradOut& out, bool want_gauss, bool want_pose, bool debug,
                          cudaStream_t stream, unsigned int* done_counter /* zeroed, or NULL */);

int preprocess_bwd_blocks(int P);
int launch_preprocess_bwd_partials(int variant, int P, int D, int M, const float* means3D,
                                   const int* radii, const Camera& cam, const float* perspec,
                                   const GeomState& g, float* acc, float* pose_partials,
                                   bool clear_acc, cudaStream_t stream);
int probe_tile_counts(const Camera& cam, GeomState& g, ImgState& img, cudaStream_t stream);

}  // namespace gsr
