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

}  // namespace
}  // namespace gsr

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
