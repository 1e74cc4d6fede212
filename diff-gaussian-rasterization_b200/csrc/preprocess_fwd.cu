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
// Sigma = M^T M with M = S R; every M entry is the single product s_row * R[col][row]; the
// rotation entries and the 3-term sums follow the reference build's association (SASS-derived).
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 s, float mod, const float4 q,
                                                     float* cov3D) {
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  const float xz = GSR_MUL(x, z), rx = GSR_MUL(r, x), rz = GSR_MUL(r, z);
  const float yy = GSR_MUL(y, y), zz = GSR_MUL(z, z);
  const float xz_p_ry = GSR_FMA(r, y, xz);    // x*z + r*y
  const float xz_m_ry = GSR_FMA(-r, y, xz);   // x*z - r*y
  const float yz_m_rx = GSR_FMA(y, z, -rx);   // y*z - r*x
  const float yz_p_rx = GSR_FMA(y, z, rx);    // y*z + r*x
  const float xy_m_rz = GSR_FMA(x, y, -rz);   // x*y - r*z
  const float xy_p_rz = GSR_FMA(x, y, rz);    // x*y + r*z
  const float xx_p_yy = GSR_FMA(x, x, yy);
  const float yy_p_zz = GSR_ADD(yy, zz);
  const float xx_p_zz = GSR_FMA(x, x, zz);
  // R as glm stores it: Rg[col][row]
  float Rg[3][3];
  Rg[0][0] = GSR_SUB(1.f, GSR_ADD(yy_p_zz, yy_p_zz)); Rg[0][1] = GSR_ADD(xy_m_rz, xy_m_rz); Rg[0][2] = GSR_ADD(xz_p_ry, xz_p_ry);
  Rg[1][0] = GSR_ADD(xy_p_rz, xy_p_rz); Rg[1][1] = GSR_SUB(1.f, GSR_ADD(xx_p_zz, xx_p_zz)); Rg[1][2] = GSR_ADD(yz_m_rx, yz_m_rx);
  Rg[2][0] = GSR_ADD(xz_m_ry, xz_m_ry); Rg[2][1] = GSR_ADD(yz_p_rx, yz_p_rx); Rg[2][2] = GSR_SUB(1.f, GSR_ADD(xx_p_yy, xx_p_yy));
  const float sc[3] = {GSR_MUL(mod, s.x), GSR_MUL(mod, s.y), GSR_MUL(mod, s.z)};
  float Mm[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) Mm[c][rr] = GSR_MUL(sc[rr], Rg[c][rr]);
  // Sigma[c][r] = sum_k M[r][k] * M[c][k]
#define SIG(c, rr) dot3_mid(Mm[rr][0], Mm[c][0], Mm[rr][1], Mm[c][1], Mm[rr][2], Mm[c][2])
  cov3D[0] = SIG(0, 0);
  cov3D[1] = SIG(0, 1);
  cov3D[2] = SIG(0, 2);
  cov3D[3] = SIG(1, 1);
  cov3D[4] = SIG(1, 2);
  cov3D[5] = SIG(2, 2);
#undef SIG
}

// EWA projection of the 3D covariance (Zwicker et al. 2002, eqs. 29/31) with the reference's
// 1.3*tanfov clamp and 0.3 px^2 low-pass (forward.cu:74-113).  cov = T^T V^T T with T = W J;
// J's third column and second/first off-diagonal entries are zero, so T[0][r] = W0r*J00 + W2r*J02
// and T[1][r] = W1r*J11 + W2r*J12 (first product rounded, second fused).
__device__ __forceinline__ float3 cov2d_ewa(const float3& mean, float fx, float fy, float tanx,
                                            float tany, const float* cov3D, const float* view) {
  float3 t = xform_point_4x3(mean, view);
  const float limx = GSR_MUL(1.3f, tanx);
  const float limy = GSR_MUL(1.3f, tany);
  const float txtz = GSR_DIV(t.x, t.z);
  const float tytz = GSR_DIV(t.y, t.z);
  t.x = GSR_MUL(fminf(limx, fmaxf(-limx, txtz)), t.z);
  t.y = GSR_MUL(fminf(limy, fmaxf(-limy, tytz)), t.z);
  const float tz2 = GSR_MUL(t.z, t.z);
  const float J00 = GSR_DIV(fx, t.z);
  const float J02 = GSR_DIV(-GSR_MUL(fx, t.x), tz2);
  const float J11 = GSR_DIV(fy, t.z);
  const float J12 = GSR_DIV(-GSR_MUL(fy, t.y), tz2);
  // W[col][row] (glm) = view rotation block: W[0]=(v0,v4,v8) W[1]=(v1,v5,v9) W[2]=(v2,v6,v10)
  const float W0[3] = {view[0], view[4], view[8]};
  const float W1[3] = {view[1], view[5], view[9]};
  const float W2[3] = {view[2], view[6], view[10]};
  float T0[3], T1[3];
#pragma unroll
  for (int rr = 0; rr < 3; ++rr) {
    T0[rr] = GSR_FMA(W2[rr], J02, GSR_MUL(W0[rr], J00));
    T1[rr] = GSR_FMA(W2[rr], J12, GSR_MUL(W1[rr], J11));
  }
  // symmetric V[k][c]
  const float V[3][3] = {{cov3D[0], cov3D[1], cov3D[2]}, {cov3D[1], cov3D[3], cov3D[4]}, {cov3D[2], cov3D[4], cov3D[5]}};
  // A = T^T V^T : A[c][r] = T[r][0]*V[0][c] + T[r][1]*V[1][c] + T[r][2]*V[2][c],  r in {0,1}
  float A0[3], A1[3];  // A?[c] = A[c][r=?]
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    A0[c] = dot3_mid(T0[0], V[0][c], T0[1], V[1][c], T0[2], V[2][c]);
    A1[c] = dot3_mid(T1[0], V[0][c], T1[1], V[1][c], T1[2], V[2][c]);
  }
  // cov[c][r] = A[0][r]*T[c][0] + A[1][r]*T[c][1] + A[2][r]*T[c][2]
  const float c00 = dot3_mid(A0[0], T0[0], A0[1], T0[1], A0[2], T0[2]);
  const float c01 = dot3_mid(A1[0], T0[0], A1[1], T0[1], A1[2], T0[2]);  // col 0, row 1
  const float c11 = dot3_mid(A1[0], T1[0], A1[1], T1[1], A1[2], T1[2]);
  return make_float3(GSR_ADD(c00, 0.3f), c01, GSR_ADD(c11, 0.3f));
}

// SH (degree <= 3) -> RGB about the view direction; +0.5, clamp at 0 and remember which
// channels were clamped (forward.cu:20-71).  Each basis coefficient is built left to right
// ((C*y)*(poly)) and every term is fused into the running sum, as in the reference build.
__device__ __forceinline__ float3 sh_to_rgb(int deg, const float3 pos, const float3 campos,
                                            const float* __restrict__ sh, unsigned char& clamp_bits) {
  const float dx = GSR_SUB(pos.x, campos.x), dy = GSR_SUB(pos.y, campos.y), dz = GSR_SUB(pos.z, campos.z);
  const float len = GSR_SQRT(dot3_mid(dx, dx, dy, dy, dz, dz));
  const float x = GSR_DIV(dx, len), y = GSR_DIV(dy, len), z = GSR_DIV(dz, len);
  float coef[16];
  int ncoef = 1;
  coef[0] = kSH0;
  if (deg > 0) {
    ncoef = 4;
    coef[1] = -GSR_MUL(kSH1, y);
    coef[2] = GSR_MUL(kSH1, z);
    coef[3] = -GSR_MUL(kSH1, x);
    if (deg > 1) {
      ncoef = 9;
      const float xx = GSR_MUL(x, x), yy = GSR_MUL(y, y), zz = GSR_MUL(z, z);
      const float xy = GSR_MUL(x, y), yz = GSR_MUL(y, z), xz = GSR_MUL(x, z);
      const float zz2 = GSR_ADD(zz, zz);
      coef[4] = GSR_MUL(kSH2[0], xy);
      coef[5] = GSR_MUL(kSH2[1], yz);
      coef[6] = GSR_MUL(kSH2[2], GSR_SUB(GSR_SUB(zz2, xx), yy));
      coef[7] = GSR_MUL(kSH2[3], xz);
      const float xx_m_yy = GSR_SUB(xx, yy);
      coef[8] = GSR_MUL(kSH2[4], xx_m_yy);
      if (deg > 2) {
        ncoef = 16;
        const float p4 = GSR_SUB(GSR_FMA(zz, 4.0f, -xx), yy);  // 4zz - xx - yy
        coef[9] = GSR_MUL(GSR_MUL(kSH3[0], y), GSR_FMA(xx, 3.0f, -yy));
        coef[10] = GSR_MUL(GSR_MUL(kSH3[1], xy), z);
        coef[11] = GSR_MUL(GSR_MUL(kSH3[2], y), p4);
        coef[12] = GSR_MUL(GSR_MUL(kSH3[3], z), GSR_FMA(yy, -3.0f, GSR_FMA(xx, -3.0f, zz2)));
        coef[13] = GSR_MUL(GSR_MUL(kSH3[4], x), p4);
        coef[14] = GSR_MUL(GSR_MUL(kSH3[5], z), xx_m_yy);
        coef[15] = GSR_MUL(GSR_MUL(kSH3[6], x), GSR_FMA(yy, -3.0f, xx));
      }
    }
  }
  float res[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = GSR_MUL(kSH0, sh[ch]);
#pragma unroll
    for (int k = 1; k < 16; ++k)
      if (k < ncoef) v = GSR_FMA(coef[k], sh[3 * k + ch], v);
    res[ch] = GSR_ADD(v, 0.5f);
  }
  clamp_bits = (unsigned char)((res[0] < 0.f ? 1 : 0) | (res[1] < 0.f ? 2 : 0) | (res[2] < 0.f ? 4 : 0));
  return make_float3(fmaxf(res[0], 0.0f), fmaxf(res[1], 0.0f), fmaxf(res[2], 0.0f));
}

// same evaluation from a bulk-copied row: the row is pulled into registers with LDS.128 first
template <int MT>
__device__ __forceinline__ float3 sh_to_rgb_row(int deg, const float3 pos, const float3 campos,
                                                const float* __restrict__ row, unsigned char& clamp_bits) {
  float sh[MT * 3];
  const float4* r4 = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int v = 0; v < MT * 3 / 4; ++v) {
    const float4 q = r4[v];
    sh[4 * v + 0] = q.x; sh[4 * v + 1] = q.y; sh[4 * v + 2] = q.z; sh[4 * v + 3] = q.w;
  }
  return sh_to_rgb(deg, pos, campos, sh, clamp_bits);
}

constexpr int kPreThreads = 128;
constexpr int kMaxCoeffs = 16;

// SH coefficients of a block of Gaussians are contiguous in memory ([P, M, 3] fp32): the block
// streams its slab with coalesced 128-bit loads into shared memory (row stride M*3+1 floats to
// spread banks) and each thread then reads its own row.
// 7 CTAs per SM (72 registers); the 8-CTA build (64 registers, 24 bytes spilled) measured slower: 0.067 vs 0.064 ms at C3
template <int MT, bool TMA>  // MT: compile-time number of SH coefficients (16 / 9 / 4 / 1) or 0 = runtime M
__global__ void __launch_bounds__(kPreThreads, 7)
preprocess_fwd_kernel(int P, int D, int M, const float* __restrict__ means3D,
                      const float* __restrict__ scales, float scale_modifier,
                      const float* __restrict__ rotations, const float* __restrict__ opacities,
                      const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                      const float* __restrict__ colors_precomp, const float* __restrict__ view,
                      const float* __restrict__ proj, const float* __restrict__ campos_p, int W,
                      int H, float tanx, float tany, float fx, float fy, int grid_x, int grid_y,
                      int* __restrict__ radii, float4* __restrict__ rec, float* __restrict__ cov3Ds,
                      unsigned char* __restrict__ clamped, uint32_t* __restrict__ tiles_touched,
                      uint2* __restrict__ rects, uint32_t* __restrict__ tile_count,
                      bool prefiltered, bool tight_tiles, uint32_t cs,
                      float* __restrict__ zero_f32, int* __restrict__ zero_i32) {
  // TMA == false: [kPreThreads][M*3+1] slab staged with coalesced loads by the whole block.
  // TMA == true : [kPreThreads][bulk_row_floats(M*3)] rows, each fetched by its own thread with one
  //               cp.async.bulk issued before any arithmetic and awaited only where the colour is
  //               evaluated.
  extern __shared__ __align__(16) float sh_smem[];
  __shared__ uint64_t s_bar;
  pdl_trigger();   // the tile scan behind this kernel may be set up while it runs
  const int base = blockIdx.x * kPreThreads;
  const int idx = base + threadIdx.x;
  constexpr int kBulkRow = bulk_row_floats(MT * 3);
  const int row = TMA ? kBulkRow : M * 3 + 1;
  const bool sh_path = shs != nullptr && colors_precomp == nullptr;

  if (TMA) {
    if (threadIdx.x == 0) mbar_init(&s_bar, kPreThreads);
    __syncthreads();
  } else {
    if (sh_path) {
      rows_to_smem<MT * 3>(shs + (size_t)base * M * 3, sh_smem, min(kPreThreads, P - base), M * 3,
                           threadIdx.x, kPreThreads);
      __syncthreads();
    }
    if (idx >= P) return;
  }

  int my_radius_i = 0;
  uint32_t my_tiles = 0;
  uint2 my_rect = make_uint2(0u, 0u);
  unsigned char cbits = 0;
  bool live = false;
  float3 p_orig = make_float3(0.f, 0.f, 0.f);
  float2 pix = make_float2(0.f, 0.f);
  float3 conic = make_float3(0.f, 0.f, 0.f);
  float opacity = 0.f, power_cut = 0.f, power_sure = 0.f, depth = 0.f;
  uint2 rmin = make_uint2(0u, 0u), rmax = make_uint2(0u, 0u);

  // Every global input of this Gaussian is requested up front — position, scale, rotation, opacity
  // and (TMA) the SH row — so that the loads are in flight together instead of one dependent round
  // trip after the other (position -> cull -> scale / rotation -> opacity -> SH row): the kernel is
  // latency bound (ncu: long-scoreboard stalls 4.8 warps per issue at 40 % occupancy).  Culled
  // Gaussians (3 % at C3) now cost their 28 + 12 M bytes, which is cheaper than the serialisation.
  float3 in_scale = make_float3(0.f, 0.f, 0.f);
  float4 in_rot = make_float4(0.f, 0.f, 0.f, 0.f);
  float in_opacity = 0.f;
  if (idx < P) {
    p_orig = ld3(means3D, idx);
    if (cov3D_precomp == nullptr) {
      in_scale = ld3(scales, idx);
      in_rot = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
    }
    in_opacity = opacities[idx];
  }
  if (TMA) {
    if (idx < P && sh_path) {
      mbar_arrive_expect_tx(&s_bar, (unsigned)(MT * 3 * sizeof(float)));
      bulk_g2s(sh_smem + threadIdx.x * kBulkRow, shs + (size_t)idx * (MT * 3),
               (unsigned)(MT * 3 * sizeof(float)), &s_bar);
    } else {
      mbar_arrive(&s_bar);
    }
  }

  if (idx < P) do {
    const float3 p_view = xform_point_4x3(p_orig, view);
    if (p_view.z <= 0.2f) {  // near cull (auxiliary.h:154); lateral cull is disabled upstream
      if (prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
      break;
    }
    depth = p_view.z;
    const float4 p_hom = xform_point_4x4(p_orig, proj);
    const float p_w = GSR_RCP(GSR_ADD(p_hom.w, 0.0000001f));
    const float3 p_proj = make_float3(GSR_MUL(p_hom.x, p_w), GSR_MUL(p_hom.y, p_w), GSR_MUL(p_hom.z, p_w));

    float cov_local[6];
    const float* cov3D;
    if (cov3D_precomp != nullptr) {
      cov3D = cov3D_precomp + (size_t)idx * 6;
    } else {
      cov3d_from_scale_rot(in_scale, scale_modifier, in_rot, cov_local);
#pragma unroll
      for (int k = 0; k < 6; ++k) cov3Ds[(size_t)idx * 6 + k] = cov_local[k];
      cov3D = cov_local;
    }

    const float3 cov = cov2d_ewa(p_orig, fx, fy, tanx, tany, cov3D, view);
    const float det = GSR_FMA(cov.x, cov.z, -GSR_MUL(cov.y, cov.y));
    if (det == 0.0f) break;
    const float det_inv = GSR_RCP(det);
    conic = make_float3(GSR_MUL(cov.z, det_inv), GSR_MUL(cov.y, -det_inv), GSR_MUL(cov.x, det_inv));

    const float mid = GSR_MUL(0.5f, GSR_ADD(cov.x, cov.z));
    const float disc = GSR_SQRT(fmaxf(0.1f, GSR_FMA(mid, mid, -det)));
    const float lambda1 = GSR_ADD(mid, disc);
    const float lambda2 = GSR_SUB(mid, disc);
    const float my_radius = ceilf(GSR_MUL(3.f, GSR_SQRT(fmaxf(lambda1, lambda2))));
    pix = make_float2(ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H));
    tile_rect(pix.x, pix.y, (int)my_radius, grid_x, grid_y, rmin, rmax);
    if ((rmax.x - rmin.x) * (rmax.y - rmin.y) == 0) break;
    my_radius_i = (int)my_radius;
    live = true;
  } while (false);

  if (live) {
    opacity = in_opacity;
    // power_cut: pairs with power < power_cut cannot reach alpha >= 15/255 (opacity*exp(power)
    // is monotone in power); the 1e-3 margin (0.1 % in alpha) is four orders of magnitude above
    // the rounding error of the exact test that still runs for everything above the cut.
    power_cut = (opacity > 0.0f) ? (logf(kAlphaMin / opacity) - 1e-3f) : 1.0f;
    // power_sure: pairs with power >= power_sure reach alpha >= 15/255 whatever the last bits of exp()
    // are (margin 1e-5 >> the rounding of logf / expf / the product, < 1e-6).  The backward, which
    // evaluates exp with ex2.approx, takes the forward's decision from this bound and re-evaluates the
    // exact expf only for the rare pairs inside [power_cut, power_sure).
    power_sure = (opacity > 0.0f) ? (logf(kAlphaMin / opacity) + 1e-5f) : 1.0f;
    if (tight_tiles) {
      // shrink the reference's 3-sigma rectangle to the tiles the alpha >= 15/255 ellipse can reach:
      // tile t holds pixel centres [16t, 16t+15]
      float hx, hy;
      const int kind = cut_extent(conic.x, conic.y, conic.z, power_cut, hx, hy);
      if (kind == 0) {
        rmax = rmin;
      } else if (kind == 1) {
        const int tx0 = (int)ceilf((pix.x - hx - (float)(kTileX - 1)) / (float)kTileX);
        const int tx1 = (int)floorf((pix.x + hx) / (float)kTileX) + 1;
        const int ty0 = (int)ceilf((pix.y - hy - (float)(kTileY - 1)) / (float)kTileY);
        const int ty1 = (int)floorf((pix.y + hy) / (float)kTileY) + 1;
        rmin.x = (unsigned)max((int)rmin.x, min(tx0, (int)rmax.x));
        rmin.y = (unsigned)max((int)rmin.y, min(ty0, (int)rmax.y));
        rmax.x = (unsigned)min((int)rmax.x, max(tx1, (int)rmin.x));
        rmax.y = (unsigned)min((int)rmax.y, max(ty1, (int)rmin.y));
      }
    }
    my_tiles = (rmax.y - rmin.y) * (rmax.x - rmin.x);
    my_rect = pack_rect(rmin, rmax);
    if (tile_count != nullptr) {  // tile-local binning: per-tile entry counters (fire-and-forget reds)
      for (uint32_t y = rmin.y; y < rmax.y; ++y)
        for (uint32_t x = rmin.x; x < rmax.x; ++x) atomicAdd(tile_count + (y * (uint32_t)grid_x + x) * cs, 1u);
    }
    rec[3 * (size_t)idx + 0] = make_float4(pix.x, pix.y, conic.x, conic.y);
    rec[3 * (size_t)idx + 1] = make_float4(conic.z, opacity, power_cut, depth);
  }
  if (idx < P) {
    radii[idx] = my_radius_i;
    tiles_touched[idx] = my_tiles;
    rects[idx] = my_rect;
    // -light: the forward blend accumulates per-Gaussian statistics with atomics; clearing them here
    // replaces two memset passes
    if (zero_f32 != nullptr) zero_f32[idx] = 0.0f;
    if (zero_i32 != nullptr) zero_i32[idx] = 0;
  }

  // every thread waits: the block's shared memory must outlive the copies in flight
  if (TMA) mbar_wait(&s_bar, 0u);

  if (live) {
    float3 rgb;
    if (colors_precomp == nullptr) {
      const float3 campos = make_float3(campos_p[0], campos_p[1], campos_p[2]);
      if (TMA)
        rgb = sh_to_rgb_row<(MT > 0 ? MT : 1)>(D, p_orig, campos, sh_smem + threadIdx.x * kBulkRow, cbits);
      else
        rgb = sh_to_rgb(D, p_orig, campos, sh_smem + threadIdx.x * row, cbits);
    } else {
      rgb = ld3(colors_precomp, idx);
    }
    rec[3 * (size_t)idx + 2] = make_float4(rgb.x, rgb.y, rgb.z, power_sure);
  }
  if (idx < P) clamped[idx] = cbits;
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
                          GeomState& g, uint32_t* tile_count, bool prefiltered, bool debug,
                          cudaStream_t stream, float* zero_f32, int* zero_i32) {
  if (colors_precomp == nullptr && (shs == nullptr || M <= 0 || M > kMaxCoeffs)) {
    set_error("SH colours need 1 <= M <= %d coefficients (got %d)", kMaxCoeffs, M);
    return GSR_E_INVALID;
  }
  if (colors_precomp == nullptr && (D + 1) * (D + 1) > M) {
    set_error("SH degree %d needs %d coefficients but M = %d", D, (D + 1) * (D + 1), M);
    return GSR_E_INVALID;
  }
  const bool sh_path = shs != nullptr && colors_precomp == nullptr;
  // bulk-copied SH rows need 16-byte aligned rows: M = 16 or 4 on a 16-byte aligned tensor
  const bool tma = sh_path && (M == 16 || M == 4) && (reinterpret_cast<uintptr_t>(shs) & 15) == 0 &&
                   options().bulk_sh != 0;
  const size_t smem = !sh_path ? 0
                      : tma    ? sizeof(float) * kPreThreads * (size_t)bulk_row_floats(M * 3)
                               : sizeof(float) * kPreThreads * (size_t)(M * 3 + 1);
  const int blocks = (P + kPreThreads - 1) / kPreThreads;
  StageScope st(ST_PRE_FWD, stream);
#define GSR_PRE_FWD(MT, TMA)                                                                     \
  prefer_max_shared_once(reinterpret_cast<const void*>(&preprocess_fwd_kernel<MT, TMA>));       \
  preprocess_fwd_kernel<MT, TMA><<<blocks, kPreThreads, smem, stream>>>(                         \
      P, D, M, means3D, scales, scale_modifier, rotations, opacities, shs, cov3D_precomp,        \
      colors_precomp, cam.view, cam.proj, cam.campos, cam.W, cam.H, cam.tan_fovx, cam.tan_fovy,  \
      cam.focal_x, cam.focal_y, cam.grid_x, cam.grid_y, radii, g.rec, g.cov3D, g.clamped,        \
      g.tiles_touched, g.rect, tile_count, prefiltered, options().tight_tiles != 0,            \
      (uint32_t)cnt_stride(), zero_f32, zero_i32)
  if (tma) {
    if (M == 16) { GSR_PRE_FWD(16, true); } else { GSR_PRE_FWD(4, true); }
  } else {
    switch (M) {
      case 16: GSR_PRE_FWD(16, false); break;
      case 9: GSR_PRE_FWD(9, false); break;
      case 4: GSR_PRE_FWD(4, false); break;
      case 1: GSR_PRE_FWD(1, false); break;
      default: GSR_PRE_FWD(0, false); break;
    }
  }
#undef GSR_PRE_FWD
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
