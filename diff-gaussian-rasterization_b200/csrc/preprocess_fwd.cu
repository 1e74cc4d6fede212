// preprocess_fwd.cu — per-Gaussian forward stage (one thread per Gaussian):
// near cull, projection, 3D covariance from scale/rotation, EWA 2D covariance (+0.3 dilation),
// conic, radius, tile rectangle, SH -> RGB, and the packed 48-byte blend record.
//
// Behaviour follows the reference's FORWARD::preprocessCUDA
// (cuda_rasterizer/forward.cu:156-256, computeCov3D :118-152, computeCov2D :74-113,
//  computeColorFromSH :20-71, in_frustum auxiliary.h:139-164); identical in -light and -full.
// The arithmetic keeps the reference's association (column-major 3x3 products evaluated
// A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]) so per-Gaussian state matches bit for bit
// and hard thresholds (radius ceil, tile rect casts) do not flip.
//
// Roofline: HBM. Algorithmic bytes per Gaussian = 44 + 12*M in (236 B at M=16) + 48 (record)
// + 24 (cov3D) + 4 (radii) + 4 (tiles) + 1 (clamped) out.
#include "gsr_common.cuh"

namespace gsr {

namespace {

__device__ __forceinline__ float3 ld3(const float* p, int idx) {
  return make_float3(p[3 * idx], p[3 * idx + 1], p[3 * idx + 2]);
}

// 3D covariance from scale + (un-normalised, as in the reference forward.cu:127) quaternion.
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 s, float mod, const float4 q,
                                                     float* cov3D) {
  M3 S;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) S.c[i][j] = (i == j) ? 1.0f : 0.0f;
  S.c[0][0] = mod * s.x;
  S.c[1][1] = mod * s.y;
  S.c[2][2] = mod * s.z;
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  M3 R;
  R.c[0][0] = 1.f - 2.f * (y * y + z * z); R.c[0][1] = 2.f * (x * y - r * z); R.c[0][2] = 2.f * (x * z + r * y);
  R.c[1][0] = 2.f * (x * y + r * z); R.c[1][1] = 1.f - 2.f * (x * x + z * z); R.c[1][2] = 2.f * (y * z - r * x);
  R.c[2][0] = 2.f * (x * z - r * y); R.c[2][1] = 2.f * (y * z + r * x); R.c[2][2] = 1.f - 2.f * (x * x + y * y);
  const M3 Mm = m3_mul(S, R);
  const M3 Sigma = m3_mul(m3_transpose(Mm), Mm);
  cov3D[0] = Sigma.c[0][0];
  cov3D[1] = Sigma.c[0][1];
  cov3D[2] = Sigma.c[0][2];
  cov3D[3] = Sigma.c[1][1];
  cov3D[4] = Sigma.c[1][2];
  cov3D[5] = Sigma.c[2][2];
}

// EWA projection of the 3D covariance (Zwicker et al. 2002, eqs. 29/31) with the reference's
// 1.3*tanfov clamp and 0.3 px^2 low-pass (forward.cu:74-113).
__device__ __forceinline__ float3 cov2d_ewa(const float3& mean, float fx, float fy, float tanx,
                                            float tany, const float* cov3D, const float* view) {
  float3 t = xform_point_4x3(mean, view);
  const float limx = 1.3f * tanx;
  const float limy = 1.3f * tany;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
  t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
  M3 J;
  J.c[0][0] = fx / t.z; J.c[0][1] = 0.0f;     J.c[0][2] = -(fx * t.x) / (t.z * t.z);
  J.c[1][0] = 0.0f;     J.c[1][1] = fy / t.z; J.c[1][2] = -(fy * t.y) / (t.z * t.z);
  J.c[2][0] = 0.0f;     J.c[2][1] = 0.0f;     J.c[2][2] = 0.0f;
  M3 Wm;
  Wm.c[0][0] = view[0]; Wm.c[0][1] = view[4]; Wm.c[0][2] = view[8];
  Wm.c[1][0] = view[1]; Wm.c[1][1] = view[5]; Wm.c[1][2] = view[9];
  Wm.c[2][0] = view[2]; Wm.c[2][1] = view[6]; Wm.c[2][2] = view[10];
  const M3 T = m3_mul(Wm, J);
  M3 V;
  V.c[0][0] = cov3D[0]; V.c[0][1] = cov3D[1]; V.c[0][2] = cov3D[2];
  V.c[1][0] = cov3D[1]; V.c[1][1] = cov3D[3]; V.c[1][2] = cov3D[4];
  V.c[2][0] = cov3D[2]; V.c[2][1] = cov3D[4]; V.c[2][2] = cov3D[5];
  M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(V)), T);
  cov.c[0][0] += 0.3f;
  cov.c[1][1] += 0.3f;
  return make_float3(cov.c[0][0], cov.c[0][1], cov.c[1][1]);
}

// SH (degree <= 3) -> RGB about the view direction; +0.5, clamp at 0 and remember which
// channels were clamped (forward.cu:20-71).
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float3 pos, const float3 campos,
                                            const float* __restrict__ sh, unsigned char& clamp_bits) {
  float3 dir = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
  const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
  dir.x = dir.x / len;
  dir.y = dir.y / len;
  dir.z = dir.z / len;
#define SHC(k, ch) sh[3 * (k) + (ch)]
  float res[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = kSH0 * SHC(0, ch);
    if (deg > 0) {
      const float x = dir.x, y = dir.y, z = dir.z;
      v = v - kSH1 * y * SHC(1, ch) + kSH1 * z * SHC(2, ch) - kSH1 * x * SHC(3, ch);
      if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, yz = y * z, xz = x * z;
        v = v + kSH2[0] * xy * SHC(4, ch) + kSH2[1] * yz * SHC(5, ch) +
            kSH2[2] * (2.0f * zz - xx - yy) * SHC(6, ch) + kSH2[3] * xz * SHC(7, ch) +
            kSH2[4] * (xx - yy) * SHC(8, ch);
        if (deg > 2) {
          v = v + kSH3[0] * y * (3.0f * xx - yy) * SHC(9, ch) + kSH3[1] * xy * z * SHC(10, ch) +
              kSH3[2] * y * (4.0f * zz - xx - yy) * SHC(11, ch) +
              kSH3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12, ch) +
              kSH3[4] * x * (4.0f * zz - xx - yy) * SHC(13, ch) +
              kSH3[5] * z * (xx - yy) * SHC(14, ch) + kSH3[6] * x * (xx - 3.0f * yy) * SHC(15, ch);
        }
      }
    }
    res[ch] = v + 0.5f;
  }
#undef SHC
  clamp_bits = (unsigned char)((res[0] < 0.f ? 1 : 0) | (res[1] < 0.f ? 2 : 0) | (res[2] < 0.f ? 4 : 0));
  return make_float3(fmaxf(res[0], 0.0f), fmaxf(res[1], 0.0f), fmaxf(res[2], 0.0f));
}

constexpr int kPreThreads = 128;
constexpr int kMaxCoeffs = 16;

// SH coefficients of a block of Gaussians are contiguous in memory ([P, M, 3] fp32): the block
// streams its slab with coalesced 128-bit loads into shared memory (row stride M*3+1 floats to
// spread banks) and each thread then reads its own row.
__global__ void __launch_bounds__(kPreThreads)
preprocess_fwd_kernel(int P, int D, int M, const float* __restrict__ means3D,
                      const float* __restrict__ scales, float scale_modifier,
                      const float* __restrict__ rotations, const float* __restrict__ opacities,
                      const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                      const float* __restrict__ colors_precomp, const float* __restrict__ view,
                      const float* __restrict__ proj, const float* __restrict__ campos_p, int W,
                      int H, float tanx, float tany, float fx, float fy, int grid_x, int grid_y,
                      int* __restrict__ radii, float4* __restrict__ rec, float* __restrict__ cov3Ds,
                      unsigned char* __restrict__ clamped, uint32_t* __restrict__ tiles_touched,
                      bool prefiltered, bool tight_tiles) {
  extern __shared__ float sh_smem[];  // [kPreThreads][M*3+1] when shs != nullptr
  const int base = blockIdx.x * kPreThreads;
  const int idx = base + threadIdx.x;
  const int row = M * 3 + 1;

  if (shs != nullptr && colors_precomp == nullptr) {
    const int nvalid = min(kPreThreads, P - base);
    const int nfloats = nvalid * M * 3;
    const float* src = shs + (size_t)base * M * 3;
    // the slab start is 16-byte aligned whenever base*M*3*4 is (kPreThreads = 128 makes it so)
    const int nvec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (nfloats >> 2) : 0;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int v = threadIdx.x; v < nvec; v += kPreThreads) {
      const float4 q = __ldg(src4 + v);
      const int f = v << 2;
      const float vals[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int ff = f + k;
        const int g = ff / (M * 3);
        sh_smem[g * row + (ff - g * M * 3)] = vals[k];
      }
    }
    for (int ff = (nvec << 2) + threadIdx.x; ff < nfloats; ff += kPreThreads) {
      const int g = ff / (M * 3);
      sh_smem[g * row + (ff - g * M * 3)] = __ldg(src + ff);
    }
    __syncthreads();
  }
  if (idx >= P) return;

  int my_radius_i = 0;
  uint32_t my_tiles = 0;
  unsigned char cbits = 0;

  const float3 p_orig = ld3(means3D, idx);
  const float3 p_view = xform_point_4x3(p_orig, view);
  do {
    if (p_view.z <= 0.2f) {  // near cull (auxiliary.h:154); lateral cull is disabled upstream
      if (prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      break;
    }
    const float4 p_hom = xform_point_4x4(p_orig, proj);
    const float p_w = 1.0f / (p_hom.w + 0.0000001f);
    const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);

    float cov_local[6];
    const float* cov3D;
    if (cov3D_precomp != nullptr) {
      cov3D = cov3D_precomp + (size_t)idx * 6;
    } else {
      const float3 s = ld3(scales, idx);
      const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
      cov3d_from_scale_rot(s, scale_modifier, q, cov_local);
#pragma unroll
      for (int k = 0; k < 6; ++k) cov3Ds[(size_t)idx * 6 + k] = cov_local[k];
      cov3D = cov_local;
    }

    const float3 cov = cov2d_ewa(p_orig, fx, fy, tanx, tany, cov3D, view);
    const float det = (cov.x * cov.z - cov.y * cov.y);
    if (det == 0.0f) break;
    const float det_inv = 1.f / det;
    const float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);

    const float mid = 0.5f * (cov.x + cov.z);
    const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    const float2 pix = make_float2(ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H));
    uint2 rmin, rmax;
    tile_rect(pix.x, pix.y, (int)my_radius, grid_x, grid_y, rmin, rmax);
    if ((rmax.x - rmin.x) * (rmax.y - rmin.y) == 0) break;

    float3 rgb;
    if (colors_precomp == nullptr) {
      const float3 campos = make_float3(campos_p[0], campos_p[1], campos_p[2]);
      rgb = sh_to_rgb(D, p_orig, campos, sh_smem + threadIdx.x * row, cbits);
    } else {
      rgb = ld3(colors_precomp, idx);
    }
    const float opacity = opacities[idx];
    // power_cut: pairs with power < power_cut cannot reach alpha >= 15/255 (opacity*exp(power)
    // is monotone in power); the 1e-3 margin (0.1 % in alpha) is four orders of magnitude above
    // the rounding error of the exact test that still runs for everything above the cut.
    const float power_cut = (opacity > 0.0f) ? (logf(kAlphaMin / opacity) - 1e-3f) : 1.0f;

    my_radius_i = (int)my_radius;
    my_tiles = (rmax.y - rmin.y) * (rmax.x - rmin.x);
    rec[3 * (size_t)idx + 0] = make_float4(pix.x, pix.y, conic.x, conic.y);
    rec[3 * (size_t)idx + 1] = make_float4(conic.z, opacity, power_cut, p_view.z);
    rec[3 * (size_t)idx + 2] = make_float4(rgb.x, rgb.y, rgb.z, 0.0f);
    (void)tight_tiles;
  } while (false);

  radii[idx] = my_radius_i;
  tiles_touched[idx] = my_tiles;
  clamped[idx] = cbits;
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D,
                                    const float* __restrict__ view, unsigned char* present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float3 p = ld3(means3D, idx);
  const float3 pv = xform_point_4x3(p, view);
  present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

}  // namespace

int launch_preprocess_fwd(int P, int D, int M, const float* means3D, const float* scales,
                          float scale_modifier, const float* rotations, const float* opacities,
                          const float* shs, const float* cov3D_precomp,
                          const float* colors_precomp, const Camera& cam, int* radii,
                          GeomState& g, bool prefiltered, bool debug, cudaStream_t stream) {
  if (colors_precomp == nullptr && (shs == nullptr || M <= 0 || M > kMaxCoeffs)) {
    set_error("SH colours need 1 <= M <= %d coefficients (got %d)", kMaxCoeffs, M);
    return GSR_E_INVALID;
  }
  if (colors_precomp == nullptr && (D + 1) * (D + 1) > M) {
    set_error("SH degree %d needs %d coefficients but M = %d", D, (D + 1) * (D + 1), M);
    return GSR_E_INVALID;
  }
  const size_t smem = (shs != nullptr && colors_precomp == nullptr)
                          ? sizeof(float) * kPreThreads * (size_t)(M * 3 + 1) : 0;
  const int blocks = (P + kPreThreads - 1) / kPreThreads;
  preprocess_fwd_kernel<<<blocks, kPreThreads, smem, stream>>>(
      P, D, M, means3D, scales, scale_modifier, rotations, opacities, shs, cov3D_precomp,
      colors_precomp, cam.view, cam.proj, cam.campos, cam.W, cam.H, cam.tan_fovx, cam.tan_fovy,
      cam.focal_x, cam.focal_y, cam.grid_x, cam.grid_y, radii, g.rec, g.cov3D, g.clamped,
      g.tiles_touched, prefiltered, options().tight_tiles != 0);
  GSR_LAUNCH_OK(debug, stream);
  return GSR_OK;
}

}  // namespace gsr

extern "C" int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix,
                                const float* projmatrix, unsigned char* present, void* stream) {
  (void)projmatrix;  // the reference projects but only tests view-space z (auxiliary.h:154)
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
    gsr::set_error("gsr_mark_visible: null pointer or negative P");
    return GSR_E_INVALID;
  }
  if (P == 0) return GSR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  gsr::mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}
