// sh_grad_views.cu — summed dL/dSH of several camera views, rebuilt from each view's masked colour
// gradient (3 floats per Gaussian and view).
//
// The SH backward of the reference (cuda_rasterizer/backward.cu:20-139 light, :20-169 full) writes
//   dL_dsh[g][k][c] = basis_k(dir) * dL_dRGB[c] * (clamped[c] ? 0 : 1),   dir = normalize(mean - campos)
// i.e. a rank-1 product per Gaussian and view.  View-level data parallelism therefore does not have
// to all-reduce 3*M floats per Gaussian: ranks all-gather the 3 masked colour gradients (plus the
// camera position) and every rank evaluates the sum over views here.  One thread per Gaussian,
// 3*M accumulators in registers, coalesced 128-bit slab store through shared memory.
// Roofline: HBM, (12 * nviews + 12) bytes in + 12*M bytes out per Gaussian.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kShThreads = 128;

template <int MT>
__global__ void __launch_bounds__(kShThreads)
sh_grad_from_views_kernel(int P, int D, int M, const float* __restrict__ means3D, int nviews,
                          const float* __restrict__ dR_all, size_t view_stride,
                          const float* __restrict__ campos_all, size_t campos_stride,
                          float* __restrict__ dL_dsh) {
  extern __shared__ float sh_smem[];  // [kShThreads][M*3+1]
  constexpr int MAXC = MT > 0 ? MT : 16;
  const int base = blockIdx.x * kShThreads;
  const int idx = base + threadIdx.x;
  const int m3 = M * 3;
  const int row = m3 + 1;
  const int ncoef = min((D + 1) * (D + 1), min(M, MAXC));
  float acc[MAXC * 3];
#pragma unroll
  for (int k = 0; k < MAXC * 3; ++k) acc[k] = 0.f;

  if (idx < P) {
    const float mx = means3D[3 * (size_t)idx], my = means3D[3 * (size_t)idx + 1], mz = means3D[3 * (size_t)idx + 2];
    for (int v = 0; v < nviews; ++v) {
      const float* dR = dR_all + (size_t)v * view_stride + 3 * (size_t)idx;
      const float r0 = dR[0], r1 = dR[1], r2 = dR[2];
      if (r0 == 0.f && r1 == 0.f && r2 == 0.f) continue;  // culled / fully clamped in this view
      const float* cp = campos_all + (size_t)v * campos_stride;
      const float dx = mx - cp[0], dy = my - cp[1], dz = mz - cp[2];
      const float len = sqrtf(dx * dx + dy * dy + dz * dz);
      const float x = dx / len, y = dy / len, z = dz / len;
      float coef[16];
      coef[0] = kSH0;
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      coef[1] = -kSH1 * y; coef[2] = kSH1 * z; coef[3] = -kSH1 * x;
      coef[4] = kSH2[0] * xy; coef[5] = kSH2[1] * yz; coef[6] = kSH2[2] * (2.f * zz - xx - yy);
      coef[7] = kSH2[3] * xz; coef[8] = kSH2[4] * (xx - yy);
      coef[9] = kSH3[0] * y * (3.f * xx - yy);
      coef[10] = kSH3[1] * xy * z;
      coef[11] = kSH3[2] * y * (4.f * zz - xx - yy);
      coef[12] = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
      coef[13] = kSH3[4] * x * (4.f * zz - xx - yy);
      coef[14] = kSH3[5] * z * (xx - yy);
      coef[15] = kSH3[6] * x * (xx - 3.f * yy);
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        if (k < ncoef) {
          acc[3 * k + 0] += coef[k] * r0;
          acc[3 * k + 1] += coef[k] * r1;
          acc[3 * k + 2] += coef[k] * r2;
        }
      }
    }
    float* mine = sh_smem + threadIdx.x * row;
#pragma unroll
    for (int k = 0; k < MAXC * 3; ++k)
      if (k < m3) mine[k] = acc[k];
    for (int k = MAXC * 3; k < m3; ++k) mine[k] = 0.f;  // M > 16: coefficients beyond degree 3
  }
  __syncthreads();
  smem_to_rows<MT * 3>(dL_dsh + (size_t)base * m3, sh_smem, min(kShThreads, P - base), m3, threadIdx.x,
                       kShThreads);
}

// ---- the same sum with the views' gradients read IN PLACE from peer GPUs --------------------------
// View-level data parallelism over NVLink: instead of all-gathering the masked colour gradients
// (12 bytes per Gaussian and rank) into a local buffer and reading that buffer again, every rank reads
// its peers' gradient arenas directly (symmetric memory, P2P loads over NVSwitch) inside the kernel
// that consumes them — the gather IS the kernel's input stream.  A P2P load takes 2-3 us, so all the
// views' triples of a Gaussian are requested before the first one is used.
constexpr int kMaxPtrViews = 16;
constexpr int kChunkViews = 8;   // views fetched per round (8 x 96 float4 = 6 loads in flight per thread)
struct ViewPtrs {
  const float* dR[kMaxPtrViews];      // [P,3] of view v (possibly in a peer GPU's memory)
  const float* campos[kMaxPtrViews];  // [3]
};

template <int MT>
__global__ void __launch_bounds__(kShThreads)
sh_grad_from_view_ptrs_kernel(int P, int D, int M, const float* __restrict__ means3D, int nviews,
                              ViewPtrs vp, float* __restrict__ dL_dsh) {
  extern __shared__ float sh_smem[];  // [kShThreads][M*3+1]
  __shared__ float s_cam[kMaxPtrViews][3];
  __shared__ __align__(16) float s_dr[kChunkViews][kShThreads * 3];   // one chunk of view slabs (12 KB)
  constexpr int MAXC = MT > 0 ? MT : 16;
  const int base = blockIdx.x * kShThreads;
  const int idx = base + threadIdx.x;
  const int m3 = M * 3;
  const int row = m3 + 1;
  const int ncoef = min((D + 1) * (D + 1), min(M, MAXC));
  if (threadIdx.x < nviews * 3) s_cam[threadIdx.x / 3][threadIdx.x % 3] = vp.campos[threadIdx.x / 3][threadIdx.x % 3];
  float acc[MAXC * 3];
#pragma unroll
  for (int k = 0; k < MAXC * 3; ++k) acc[k] = 0.f;
  __syncthreads();

  // Peer memory is not cached locally: the block fetches each view's [128, 3] slab ONCE with coalesced
  // 128-bit loads (all of a chunk's loads in flight together) and the threads then read shared memory.
  const float mx = idx < P ? means3D[3 * (size_t)idx] : 0.f, my = idx < P ? means3D[3 * (size_t)idx + 1] : 0.f,
              mz = idx < P ? means3D[3 * (size_t)idx + 2] : 0.f;
  const int nvalid = min(kShThreads, P - base);
  const int slab4 = (nvalid * 3 + 3) / 4;           // float4s per view slab (P % 4 == 0 is not required:
  const size_t slab_off = (size_t)base * 3;         //  the tail is fetched with scalar loads below)
  const bool vec_ok = (nvalid * 3) % 4 == 0;
  for (int v0 = 0; v0 < nviews; v0 += kChunkViews) {
    const int nv = min(kChunkViews, nviews - v0);
    __syncthreads();
    if (vec_ok) {
      float4 tmp[kChunkViews * 96 / kShThreads];
#pragma unroll
      for (int q = 0; q < kChunkViews * 96 / kShThreads; ++q) {
        const int e = q * kShThreads + threadIdx.x, v = e / 96, i = e % 96;
        tmp[q] = (v < nv && i < slab4)
                     ? *reinterpret_cast<const float4*>(vp.dR[v0 + (v < nv ? v : 0)] + slab_off + 4 * (size_t)i)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < kChunkViews * 96 / kShThreads; ++q) {
        const int e = q * kShThreads + threadIdx.x, v = e / 96, i = e % 96;
        reinterpret_cast<float4*>(s_dr[v])[i] = tmp[q];
      }
    } else {
      for (int e = threadIdx.x; e < nv * nvalid * 3; e += kShThreads)
        s_dr[e / (nvalid * 3)][e % (nvalid * 3)] = vp.dR[v0 + e / (nvalid * 3)][slab_off + e % (nvalid * 3)];
    }
    __syncthreads();
    if (idx < P) {
#pragma unroll
      for (int u = 0; u < kChunkViews; ++u) {
        if (u >= nv) break;
        const float r0 = s_dr[u][3 * threadIdx.x], r1 = s_dr[u][3 * threadIdx.x + 1], r2 = s_dr[u][3 * threadIdx.x + 2];
        if (r0 == 0.f && r1 == 0.f && r2 == 0.f) continue;  // culled / fully clamped in this view
        const int v = v0 + u;
        const float dx = mx - s_cam[v][0], dy = my - s_cam[v][1], dz = mz - s_cam[v][2];
        const float len = sqrtf(dx * dx + dy * dy + dz * dz);
        const float x = dx / len, y = dy / len, z = dz / len;
        float coef[16];
        coef[0] = kSH0;
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        coef[1] = -kSH1 * y; coef[2] = kSH1 * z; coef[3] = -kSH1 * x;
        coef[4] = kSH2[0] * xy; coef[5] = kSH2[1] * yz; coef[6] = kSH2[2] * (2.f * zz - xx - yy);
        coef[7] = kSH2[3] * xz; coef[8] = kSH2[4] * (xx - yy);
        coef[9] = kSH3[0] * y * (3.f * xx - yy);
        coef[10] = kSH3[1] * xy * z;
        coef[11] = kSH3[2] * y * (4.f * zz - xx - yy);
        coef[12] = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
        coef[13] = kSH3[4] * x * (4.f * zz - xx - yy);
        coef[14] = kSH3[5] * z * (xx - yy);
        coef[15] = kSH3[6] * x * (xx - 3.f * yy);
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
          if (k < ncoef) {
            acc[3 * k + 0] += coef[k] * r0;
            acc[3 * k + 1] += coef[k] * r1;
            acc[3 * k + 2] += coef[k] * r2;
          }
        }
      }
    }
  }
  if (idx < P) {
    float* mine = sh_smem + threadIdx.x * row;
#pragma unroll
    for (int k = 0; k < MAXC * 3; ++k)
      if (k < m3) mine[k] = acc[k];
    for (int k = MAXC * 3; k < m3; ++k) mine[k] = 0.f;
  }
  __syncthreads();
  smem_to_rows<MT * 3>(dL_dsh + (size_t)base * m3, sh_smem, min(kShThreads, P - base), m3, threadIdx.x,
                       kShThreads);
}

// ---- in-switch all-reduce of one slice (NVLS) ------------------------------------------------------
// `mc` is the multicast alias of a symmetric buffer.  multimem.ld_reduce returns the SUM of all
// replicas (the NVSwitch pulls the operands and adds them), multimem.st writes the result back into
// every replica: a rank that runs this over its 1/world slice has all-reduced that slice with 1x the
// data volume on its links.  Operates on float4 elements [begin4, end4).
__device__ __forceinline__ float4 mm_ld_reduce4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

constexpr int kNvlsUnroll = 8;   // switch round trips in flight per thread (one takes several microseconds)
__global__ void __launch_bounds__(256)
nvls_reduce_slice_kernel(float* __restrict__ mc, size_t begin4, size_t end4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end4; i0 += stride * kNvlsUnroll) {
    float4 v[kNvlsUnroll];
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u)
      if (i0 + u * stride < end4) v[u] = mm_ld_reduce4(mc + 4 * (i0 + u * stride));
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u)
      if (i0 + u * stride < end4) mm_st4(mc + 4 * (i0 + u * stride), v[u]);
  }
}


// ---- two-shot all-reduce of one slice over plain P2P loads / stores ---------------------------------
// The same job as nvls_reduce_slice_kernel without the switch's reduction engine: a rank reads its 1/world
// slice from every replica (peer pointers of the symmetric buffer), adds the replicas in rank order —
// every element is summed by exactly one rank, so all ranks end up with bit-identical sums — and stores
// the result into every replica.  world x 16 bytes are in flight per thread and unrolled step, so the
// NVLink round trip (2-3 us) is covered by a few hundred KB per SM.
struct PeerPtrs { float* p[kMaxPtrViews]; };

__device__ __forceinline__ float4 ld_sys4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

template <int WORLD, int UNROLL>
__global__ void __launch_bounds__(256)
p2p_reduce_slice_kernel(PeerPtrs pp, int world_rt, size_t begin4, size_t end4) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  constexpr int WMAX = WORLD > 0 ? WORLD : kMaxPtrViews;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end4; i0 += stride * UNROLL) {
    float4 v[UNROLL][WMAX];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + u * stride;
#pragma unroll
      for (int r = 0; r < WMAX; ++r)
        if (r < world && i < end4) v[u][r] = ld_sys4(pp.p[r] + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + u * stride;
      if (i < end4) {
        float4 s = v[u][0];
#pragma unroll
        for (int r = 1; r < WMAX; ++r)
          if (r < world) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
#pragma unroll
        for (int r = 0; r < WMAX; ++r)
          if (r < world) st_sys4(pp.p[r] + 4 * i, s);
      }
    }
  }
}


// ---- all-gather by P2P loads: every view's block is copied from its (peer) replica into a local buffer --------
// A pure copy bound by the NVLink ports: a few CTAs with a deep queue of 128-bit loads are enough, so the kernel
// can run UNDERNEATH the per-Gaussian backward kernel (dp.py launches it on a high-priority stream as soon as the
// masked colour gradients exist) without taking the SMs from it.
constexpr int kGatherUnroll = 8;
__global__ void __launch_bounds__(256)
p2p_gather_kernel(PeerPtrs src, int nviews, size_t count4, float* __restrict__ dst, size_t dst_stride4) {
  const size_t total = (size_t)nviews * count4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * kGatherUnroll) {
    float4 v[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const size_t i = i0 + u * stride;
      if (i < total) {
        const size_t view = i / count4, off = i - view * count4;
        v[u] = ld_sys4(src.p[view] + 4 * off);
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const size_t i = i0 + u * stride;
      if (i < total) {
        const size_t view = i / count4, off = i - view * count4;
        reinterpret_cast<float4*>(dst)[view * dst_stride4 + off] = v[u];
      }
    }
  }
}

}  // namespace
}  // namespace gsr

extern "C" int gsr_sh_grad_from_view_ptrs(int P, int D, int M, const float* means3D, int nviews,
                                          const float* const* dR_ptrs, const float* const* campos_ptrs,
                                          float* dL_dsh, void* stream) {
  using namespace gsr;
  if (P < 0 || M <= 0 || D < 0 || D > 3 || nviews < 0 || nviews > kMaxPtrViews ||
      (P > 0 && (!means3D || !dL_dsh || (nviews > 0 && (!dR_ptrs || !campos_ptrs))))) {
    set_error("gsr_sh_grad_from_view_ptrs: bad arguments (P=%d M=%d D=%d nviews=%d, at most %d views)", P, M, D,
              nviews, kMaxPtrViews);
    return GSR_E_INVALID;
  }
  if (P == 0) return GSR_OK;
  ViewPtrs vp;
  for (int v = 0; v < kMaxPtrViews; ++v) {
    vp.dR[v] = v < nviews ? dR_ptrs[v] : nullptr;
    vp.campos[v] = v < nviews ? campos_ptrs[v] : nullptr;
    if (v < nviews && (!vp.dR[v] || !vp.campos[v])) {
      set_error("gsr_sh_grad_from_view_ptrs: NULL view pointer");
      return GSR_E_INVALID;
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = (P + kShThreads - 1) / kShThreads;
  const size_t smem = sizeof(float) * kShThreads * (size_t)(M * 3 + 1);
  StageScope st(ST_OTHER, s);
#define GSR_SH_PTRS(MT) \
  sh_grad_from_view_ptrs_kernel<MT><<<blocks, kShThreads, smem, s>>>(P, D, M, means3D, nviews, vp, dL_dsh)
  switch (M) {
    case 16: GSR_SH_PTRS(16); break;
    case 9: GSR_SH_PTRS(9); break;
    case 4: GSR_SH_PTRS(4); break;
    case 1: GSR_SH_PTRS(1); break;
    default: GSR_SH_PTRS(0); break;
  }
#undef GSR_SH_PTRS
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}

extern "C" int gsr_nvls_allreduce_slice(float* multicast_ptr, size_t offset_floats, size_t count_floats,
                                        int rank, int world, int max_blocks, void* stream) {
  using namespace gsr;
  if (!multicast_ptr || world <= 0 || rank < 0 || rank >= world || (offset_floats & 3) || (count_floats & 3) ||
      (reinterpret_cast<uintptr_t>(multicast_ptr) & 15)) {
    set_error("gsr_nvls_allreduce_slice: bad arguments (offset / count must be multiples of 4 floats)");
    return GSR_E_INVALID;
  }
  const size_t total4 = count_floats / 4;
  const size_t per = (total4 + (size_t)world - 1) / (size_t)world;
  const size_t b4 = offset_floats / 4 + per * (size_t)rank;
  const size_t e4 = offset_floats / 4 + (per * (size_t)(rank + 1) < total4 ? per * (size_t)(rank + 1) : total4);
  if (b4 >= e4) return GSR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = e4 - b4;
  const size_t want = (n + 256 * kNvlsUnroll - 1) / (256 * kNvlsUnroll);
  int blocks = (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
  if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
  StageScope st(ST_OTHER, s);
  nvls_reduce_slice_kernel<<<blocks, 256, 0, s>>>(multicast_ptr, b4, e4);
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}

extern "C" int gsr_sh_grad_from_views(int P, int D, int M, const float* means3D, int nviews,
                                      const float* dR_all, size_t view_stride,
                                      const float* campos_all, size_t campos_stride, float* dL_dsh,
                                      void* stream) {
  using namespace gsr;
  if (P < 0 || M <= 0 || D < 0 || D > 3 || nviews < 0 ||
      (P > 0 && (!means3D || !dL_dsh || (nviews > 0 && (!dR_all || !campos_all))))) {
    set_error("gsr_sh_grad_from_views: bad arguments (P=%d M=%d D=%d nviews=%d)", P, M, D, nviews);
    return GSR_E_INVALID;
  }
  if (P == 0) return GSR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int blocks = (P + kShThreads - 1) / kShThreads;
  const size_t smem = sizeof(float) * kShThreads * (size_t)(M * 3 + 1);
  StageScope st(ST_OTHER, s);
#define GSR_SH_VIEWS(MT)                                                                    \
  sh_grad_from_views_kernel<MT><<<blocks, kShThreads, smem, s>>>(                           \
      P, D, M, means3D, nviews, dR_all, view_stride, campos_all, campos_stride, dL_dsh)
  switch (M) {
    case 16: GSR_SH_VIEWS(16); break;
    case 9: GSR_SH_VIEWS(9); break;
    case 4: GSR_SH_VIEWS(4); break;
    case 1: GSR_SH_VIEWS(1); break;
    default: GSR_SH_VIEWS(0); break;
  }
#undef GSR_SH_VIEWS
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}

extern "C" int gsr_p2p_allreduce_slice(float* const* replica_ptrs, size_t offset_floats, size_t count_floats,
                                       int rank, int world, int max_blocks, void* stream) {
  using namespace gsr;
  if (!replica_ptrs || world <= 0 || world > kMaxPtrViews || rank < 0 || rank >= world || (offset_floats & 3) ||
      (count_floats & 3)) {
    set_error("gsr_p2p_allreduce_slice: bad arguments (offset / count must be multiples of 4 floats, at most %d ranks)",
              kMaxPtrViews);
    return GSR_E_INVALID;
  }
  PeerPtrs pp;
  for (int r = 0; r < kMaxPtrViews; ++r) {
    pp.p[r] = r < world ? replica_ptrs[r] : nullptr;
    if (r < world && (!pp.p[r] || (reinterpret_cast<uintptr_t>(pp.p[r]) & 15))) {
      set_error("gsr_p2p_allreduce_slice: replica pointers must be non-NULL and 16-byte aligned");
      return GSR_E_INVALID;
    }
  }
  const size_t total4 = count_floats / 4;
  const size_t per = (total4 + (size_t)world - 1) / (size_t)world;
  const size_t b4 = offset_floats / 4 + per * (size_t)rank;
  const size_t e4 = offset_floats / 4 + (per * (size_t)(rank + 1) < total4 ? per * (size_t)(rank + 1) : total4);
  if (b4 >= e4) return GSR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = e4 - b4;
  StageScope st(ST_OTHER, s);
#define GSR_P2P(WT, UT)                                                                   \
  do {                                                                                    \
    const size_t want = (n + 256 * UT - 1) / (256 * UT);                                  \
    int blocks = (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);                     \
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;                       \
    p2p_reduce_slice_kernel<WT, UT><<<blocks, 256, 0, s>>>(pp, world, b4, e4);            \
  } while (0)
  switch (world) {
    case 2: GSR_P2P(2, 8); break;
    case 4: GSR_P2P(4, 4); break;
    case 8: GSR_P2P(8, 2); break;
    default: GSR_P2P(0, 1); break;
  }
#undef GSR_P2P
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}

extern "C" int gsr_p2p_gather(const float* const* src_ptrs, int nviews, size_t count_floats, float* dst,
                              size_t dst_stride_floats, int max_blocks, void* stream) {
  using namespace gsr;
  if (!src_ptrs || !dst || nviews <= 0 || nviews > kMaxPtrViews || (count_floats & 3) || (dst_stride_floats & 3) ||
      dst_stride_floats < count_floats || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    set_error("gsr_p2p_gather: bad arguments (count / stride must be multiples of 4 floats, at most %d views)", kMaxPtrViews);
    return GSR_E_INVALID;
  }
  PeerPtrs pp;
  for (int v = 0; v < kMaxPtrViews; ++v) {
    pp.p[v] = v < nviews ? const_cast<float*>(src_ptrs[v]) : nullptr;
    if (v < nviews && (!pp.p[v] || (reinterpret_cast<uintptr_t>(pp.p[v]) & 15))) {
      set_error("gsr_p2p_gather: source pointers must be non-NULL and 16-byte aligned");
      return GSR_E_INVALID;
    }
  }
  if (count_floats == 0) return GSR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t total = (size_t)nviews * (count_floats / 4);
  const size_t want = (total + 256 * kGatherUnroll - 1) / (256 * kGatherUnroll);
  int blocks = (int)(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
  if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
  StageScope st(ST_OTHER, s);
  p2p_gather_kernel<<<blocks, 256, 0, s>>>(pp, nviews, count_floats / 4, dst, dst_stride_floats / 4);
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}
