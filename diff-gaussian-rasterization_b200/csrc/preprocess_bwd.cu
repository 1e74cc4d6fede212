// preprocess_bwd.cu — per-Gaussian backward stage (one thread per Gaussian), fusing what the
// reference runs as separate kernels:
//   computeCov2DCUDA        (light backward.cu:144-276, full backward.cu:174-387)
//   BACKWARD::preprocessCUDA (light :348-416, full :459-537) incl. SH backward (:20-139 / :20-169)
//                            and scale/rotation backward (:280-343 / :391-454)
//   pose_gradient_preCUDA   (light :701-751) and the per-Gaussian half of ComputePG (full :990-1072)
// Inputs are the per-Gaussian accumulator records written by render_bwd (16 floats each).
// Every output element is written (zeros for culled Gaussians), so no memset pass is needed.
// The 12 live entries of dL/dviewmatrix are reduced per block and written to `pose_partials`;
// a second tiny kernel adds the partials in a fixed order (deterministic, no atomics).
//
// Roofline: HBM.  Algorithmic bytes per Gaussian: 64 (acc) + 44 + 12*M (inputs) + 24 (cov3D)
// in, 12*M + 4*(3+3+4+1+3+1+3+6+3+4) out  ->  ~ 640 B at M = 16.
#include "gsr_common.cuh"

namespace gsr {

namespace {

constexpr int kBwdThreads = 128;
// The fused last-block pose reduction reads blocks x 48 bytes with ONE block: measured 34 us at 7813 blocks
// (C3) against ~5 us for the 12-CTA finalize kernel, so it is used for small scenes only (C2: 782 blocks),
// where the saved launch matters and the tail is ~1 us.
constexpr int kFuseFinalizeMaxBlocks = 1024;

__device__ __forceinline__ float3 ld3(const float* p, int idx) {
  return make_float3(p[3 * idx], p[3 * idx + 1], p[3 * idx + 2]);
}
__device__ __forceinline__ void st3(float* p, int idx, float a, float b, float c) {
  if (p) { p[3 * idx] = a; p[3 * idx + 1] = b; p[3 * idx + 2] = c; }
}

// d(dir/|dir|)/d(dir) applied to dv (auxiliary.h:100-111)
__device__ __forceinline__ float3 dnormvdv3(float3 v, float3 dv) {
  const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
  float3 r;
  r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
  r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
  r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
  return r;
}

// TMA == true (MT = 16 / 4 on 16-byte aligned tensors): every thread moves its own Gaussian's SH row
// with cp.async.bulk — in only if the Gaussian is visible, awaited right before the SH backward —
// and its dL/dSH row back out with a bulk store; no block-wide staging loops, no barriers.
// MINB (min CTAs per SM) is 5 for the M = 16 bulk-copy variant: 96 registers, measured best (4 CTAs at
// 113-127 registers: +10 % time, 6 CTAs with spills: +3 %); a dense-row variant with ONE bulk store per
// warp instead of one per thread was 8 % slower (4-way bank conflicts on the row accesses).
template <int VARIANT, int MT, bool TMA, int MINB = 1>  // MT: compile-time SH coefficient count (16/9/4/1) or 0 = runtime M
__global__ void __launch_bounds__(kBwdThreads, MINB)
preprocess_bwd_kernel(int P, int D, int M, const float* __restrict__ means3D,
                      const int* __restrict__ radii, const float* __restrict__ shs,
                      const unsigned char* __restrict__ clamped, const float* __restrict__ scales,
                      const float* __restrict__ rotations, float scale_modifier,
                      const float* __restrict__ cov3Ds, const float* __restrict__ view,
                      const float* __restrict__ proj, const float* __restrict__ campos_p,
                      const float* __restrict__ perspec, float fx, float fy, float tanx, float tany,
                      const float* __restrict__ acc, const float4* __restrict__ g_rec, int img_w,
                      int img_h, float* __restrict__ pose_partials, GaussGradOut out,
                      bool want_gauss, bool want_pose, float* __restrict__ acc_clear,
                      unsigned int* __restrict__ done_counter) {
  extern __shared__ __align__(16) float sh_smem[];  // [kBwdThreads][row] SH in, dL/dSH out
  __shared__ float s_pose[kBwdThreads / 32][12];
  __shared__ uint64_t s_bar;
  pdl_wait();   // (may have been launched programmatically behind the blend backward)
  const int base = blockIdx.x * kBwdThreads;
  const int idx = base + threadIdx.x;
  constexpr int kBulkRow = bulk_row_floats(MT * 3);
  constexpr unsigned kRowBytes = (unsigned)(MT * 3 * sizeof(float));
  const int row = TMA ? kBulkRow : M * 3 + 1;
  const bool use_sh = (shs != nullptr) && want_gauss;           // SH coefficients staged in
  const bool sh_out = (shs != nullptr) && out.dL_dsh != nullptr;  // dL/dSH slab staged out
  const int nvalid = min(kBwdThreads, P - base);
  const bool live = idx < P && radii[idx] > 0;
  bool bar_pending = false;

  if (TMA) {
    if (threadIdx.x == 0) mbar_init(&s_bar, kBwdThreads);
    __syncthreads();
    if (live && use_sh) {
      mbar_arrive_expect_tx(&s_bar, kRowBytes);
      bulk_g2s(sh_smem + threadIdx.x * kBulkRow, shs + (size_t)idx * (MT * 3), kRowBytes, &s_bar);
    } else {
      mbar_arrive(&s_bar);
    }
    bar_pending = true;
  } else if (use_sh) {
    rows_to_smem<MT * 3>(shs + (size_t)base * M * 3, sh_smem, nvalid, M * 3, threadIdx.x, kBwdThreads);
    __syncthreads();
  }

  float pose[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) pose[k] = 0.f;

  float* my_sh = sh_smem + threadIdx.x * row;

  // culled Gaussians — and, in -light's tracking mode (map_off), all of them — get exact zeros
  if (idx < P && (!live || !want_gauss)) {
    st3(out.dL_dmean2D, idx, 0.f, 0.f, 0.f);
    if (out.dL_dconic) {
      reinterpret_cast<float4*>(out.dL_dconic)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (out.dL_dopacity) out.dL_dopacity[idx] = 0.f;
    st3(out.dL_dcolor, idx, 0.f, 0.f, 0.f);
    st3(out.dL_dcolor_masked, idx, 0.f, 0.f, 0.f);
    if (out.dL_ddepth) out.dL_ddepth[idx] = 0.f;
    st3(out.dL_dmean3D, idx, 0.f, 0.f, 0.f);
    if (out.dL_dcov3D) {
#pragma unroll
      for (int k = 0; k < 6; ++k) out.dL_dcov3D[(size_t)idx * 6 + k] = 0.f;
    }
    st3(out.dL_dscale, idx, 0.f, 0.f, 0.f);
    if (out.dL_drot) reinterpret_cast<float4*>(out.dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sh_out) {
      if (TMA) {
        // (no wait needed: nothing is in flight towards this thread's own row)
#pragma unroll
        for (int v = 0; v < MT * 3 / 4; ++v) reinterpret_cast<float4*>(my_sh)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int k = 0; k < M * 3; ++k) my_sh[k] = 0.f;
      }
    }
  }

  if (live) {
    const float4* a4 = reinterpret_cast<const float4*>(acc + (size_t)idx * kAccStride);
    const float4 a0 = a4[0], a1 = a4[1], a2 = a4[2], a3 = a4[3];
    if (acc_clear != nullptr) {  // tracker: leave the line zeroed for the next iteration (no memset pass)
      float4* z4 = reinterpret_cast<float4*>(acc_clear + (size_t)idx * kAccStride);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      z4[0] = z; z4[1] = z; z4[2] = z; z4[3] = z;
    }
    // moments of w = G dL/dalpha (render_bwd) -> screen-space gradients with this Gaussian's conic:
    //   dL/dmean2D.x = 0.5 W o (-A S1 - B S2),  dL/dconic = -0.5 o (S11, S12, S22),  dL/dopacity = S0
    const float4 rec0 = g_rec[3 * (size_t)idx + 0], rec1 = g_rec[3 * (size_t)idx + 1];
    const float cA = rec0.z, cB = rec0.w, cC = rec1.x, opac = rec1.y;
    const float half_w_o = 0.5f * (float)img_w * opac, half_h_o = 0.5f * (float)img_h * opac;
    const float g_mx = half_w_o * (-cA * a0.x - cB * a0.y);
    const float g_my = half_h_o * (-cC * a0.y - cB * a0.x);
    const float g_ca = -0.5f * opac * a0.z, g_cb = -0.5f * opac * a0.w, g_cc = -0.5f * opac * a1.x;
    const float g_op = a1.y, g_r = a1.z, g_g = a1.w;
    const float g_b = a2.x, g_depth = a2.y;
    const float g_pgx = half_w_o * (-cA * a2.z - cB * a2.w);
    const float g_pgy = half_h_o * (-cC * a2.w - cB * a2.z);
    const float g_pd = a3.x, g_med = a3.y;

    const float3 m = ld3(means3D, idx);
    // (requesting scales / rotations / clamped here as well, in front of the arithmetic, was measured
    //  slower: 0.148 -> 0.158 ms at C3 — the eight extra live registers spill at the 96-register cap)
    const float4 m_hom = xform_point_4x4(m, proj);
    const float m_w = 1.0f / (m_hom.w + 0.0000001f);
    float3 dq = make_float3(0.f, 0.f, 0.f);  // dRGB.dL/dcolor contracted with d(rgb)/d(campos)

    if (want_gauss) {
      // ---------------- 2D covariance / conic backward ----------------
      const float* cov3D = cov3Ds + (size_t)idx * 6;
      float3 t = xform_point_4x3(m, view);
      const float limx = 1.3f * tanx, limy = 1.3f * tany;
      const float txtz = t.x / t.z, tytz = t.y / t.z;
      t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
      t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
      const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
      const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;

      M3 J;
      J.c[0][0] = fx / t.z; J.c[0][1] = 0.0f;     J.c[0][2] = -(fx * t.x) / (t.z * t.z);
      J.c[1][0] = 0.0f;     J.c[1][1] = fy / t.z; J.c[1][2] = -(fy * t.y) / (t.z * t.z);
      J.c[2][0] = 0.0f;     J.c[2][1] = 0.0f;     J.c[2][2] = 0.0f;
      M3 Wm;
      Wm.c[0][0] = view[0]; Wm.c[0][1] = view[4]; Wm.c[0][2] = view[8];
      Wm.c[1][0] = view[1]; Wm.c[1][1] = view[5]; Wm.c[1][2] = view[9];
      Wm.c[2][0] = view[2]; Wm.c[2][1] = view[6]; Wm.c[2][2] = view[10];
      M3 V;
      V.c[0][0] = cov3D[0]; V.c[0][1] = cov3D[1]; V.c[0][2] = cov3D[2];
      V.c[1][0] = cov3D[1]; V.c[1][1] = cov3D[3]; V.c[1][2] = cov3D[4];
      V.c[2][0] = cov3D[2]; V.c[2][1] = cov3D[4]; V.c[2][2] = cov3D[5];
      const M3 T = m3_mul(Wm, J);
      M3 cov2D = m3_mul(m3_mul(m3_transpose(T), m3_transpose(V)), T);
      const float a = (cov2D.c[0][0] += 0.3f);
      const float b = cov2D.c[0][1];
      const float c = (cov2D.c[1][1] += 0.3f);
      const float denom = a * c - b * b;
      float dL_da = 0, dL_db = 0, dL_dc = 0;
      const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
      float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* u = T.c[0];  // first column of T
      const float* w = T.c[1];  // second column of T
      if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * g_ca + 2 * b * c * g_cb + (denom - a * c) * g_cc);
        dL_dc = denom2inv * (-a * a * g_cc + 2 * a * b * g_cb + (denom - a * c) * g_ca);
        dL_db = denom2inv * 2 * (b * c * g_ca - (denom + 2 * b * b) * g_cb + a * b * g_cc);
        dcov[0] = (u[0] * u[0] * dL_da + u[0] * w[0] * dL_db + w[0] * w[0] * dL_dc);
        dcov[3] = (u[1] * u[1] * dL_da + u[1] * w[1] * dL_db + w[1] * w[1] * dL_dc);
        dcov[5] = (u[2] * u[2] * dL_da + u[2] * w[2] * dL_db + w[2] * w[2] * dL_dc);
        dcov[1] = 2 * u[0] * u[1] * dL_da + (u[0] * w[1] + u[1] * w[0]) * dL_db + 2 * w[0] * w[1] * dL_dc;
        dcov[2] = 2 * u[0] * u[2] * dL_da + (u[0] * w[2] + u[2] * w[0]) * dL_db + 2 * w[0] * w[2] * dL_dc;
        dcov[4] = 2 * u[2] * u[1] * dL_da + (u[1] * w[2] + u[2] * w[1]) * dL_db + 2 * w[1] * w[2] * dL_dc;
      }
      if (out.dL_dcov3D) {
#pragma unroll
        for (int k = 0; k < 6; ++k) out.dL_dcov3D[(size_t)idx * 6 + k] = dcov[k];
      }
      // gradients w.r.t. the upper 2x3 block of T, then J, then the view-space mean t
      float uV[3], wV[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uV[k] = u[0] * V.c[k][0] + u[1] * V.c[k][1] + u[2] * V.c[k][2];
        wV[k] = w[0] * V.c[k][0] + w[1] * V.c[k][1] + w[2] * V.c[k][2];
      }
      float dT0[3], dT1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dT0[k] = 2 * uV[k] * dL_da + wV[k] * dL_db;
        dT1[k] = 2 * wV[k] * dL_dc + uV[k] * dL_db;
      }
      const float dL_dJ00 = Wm.c[0][0] * dT0[0] + Wm.c[0][1] * dT0[1] + Wm.c[0][2] * dT0[2];
      const float dL_dJ02 = Wm.c[2][0] * dT0[0] + Wm.c[2][1] * dT0[1] + Wm.c[2][2] * dT0[2];
      const float dL_dJ11 = Wm.c[1][0] * dT1[0] + Wm.c[1][1] * dT1[1] + Wm.c[1][2] * dT1[2];
      const float dL_dJ12 = Wm.c[2][0] * dT1[0] + Wm.c[2][1] * dT1[1] + Wm.c[2][2] * dT1[2];
      const float tz = 1.f / t.z;
      const float tz2 = tz * tz;
      const float tz3 = tz2 * tz;
      const float dL_dtx = x_grad_mul * -fx * tz2 * dL_dJ02;
      const float dL_dty = y_grad_mul * -fy * tz2 * dL_dJ12;
      const float dL_dtz = -fx * tz2 * dL_dJ00 - fy * tz2 * dL_dJ11 + (2 * fx * t.x) * tz3 * dL_dJ02 +
                           (2 * fy * t.y) * tz3 * dL_dJ12;
      const float3 dmean_cov =
          make_float3(view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz,
                      view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz,
                      view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz);

      // ---------------- depth, median and screen-space mean terms ----------------
      const float mul3 = view[2] * m.x + view[6] * m.y + view[10] * m.z + view[14];
      const float3 dz = make_float3(view[2] - view[3] * mul3, view[6] - view[7] * mul3,
                                    view[10] - view[11] * mul3);
      const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
      const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
      float3 dmean_2d;
      dmean_2d.x = (proj[0] * m_w - proj[3] * mul1) * g_mx + (proj[1] * m_w - proj[3] * mul2) * g_my;
      dmean_2d.y = (proj[4] * m_w - proj[7] * mul1) * g_mx + (proj[5] * m_w - proj[7] * mul2) * g_my;
      dmean_2d.z = (proj[8] * m_w - proj[11] * mul1) * g_mx + (proj[9] * m_w - proj[11] * mul2) * g_my;

      float3 dmean;
      if (VARIANT == kLight) {
        // light: median atomics first, then += cov part, += 2D part, += depth part
        dmean = make_float3(dz.x * g_med * 1.0f, dz.y * g_med * 1.0f, dz.z * g_med * 1.0f);
        dmean.x += dmean_cov.x; dmean.y += dmean_cov.y; dmean.z += dmean_cov.z;
        dmean.x += dmean_2d.x;  dmean.y += dmean_2d.y;  dmean.z += dmean_2d.z;
        dmean.x += dz.x * g_depth; dmean.y += dz.y * g_depth; dmean.z += dz.z * g_depth;
      } else {
        // full: cov part (assigned) + depth part inside the cov2D kernel, then += 2D part
        dmean = make_float3(dmean_cov.x + g_depth * dz.x, dmean_cov.y + g_depth * dz.y,
                            dmean_cov.z + g_depth * dz.z);
        dmean.x += dmean_2d.x;  dmean.y += dmean_2d.y;  dmean.z += dmean_2d.z;
      }

      // ---------------- SH backward ----------------
      if (shs != nullptr) {
        const float3 campos = make_float3(campos_p[0], campos_p[1], campos_p[2]);
        const float3 dir_orig = make_float3(m.x - campos.x, m.y - campos.y, m.z - campos.z);
        const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
        const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
        const unsigned char cb = clamped[idx];
        const float dR[3] = {g_r * ((cb & 1) ? 0.f : 1.f), g_g * ((cb & 2) ? 0.f : 1.f),
                             g_b * ((cb & 4) ? 0.f : 1.f)};
        st3(out.dL_dcolor_masked, idx, dR[0], dR[1], dR[2]);
        float dx_[3] = {0.f, 0.f, 0.f}, dy_[3] = {0.f, 0.f, 0.f}, dz_[3] = {0.f, 0.f, 0.f};
        float coef[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) coef[k] = 0.f;
        coef[0] = kSH0;
        // TMA: wait for this block's rows, then pull the own row into registers with LDS.128
        constexpr int kShRegs = TMA ? (MT > 0 ? MT * 3 : 4) : 4;
        float shr[kShRegs];
        if (TMA) {
          if (bar_pending) { mbar_wait(&s_bar, 0u); bar_pending = false; }
#pragma unroll
          for (int v = 0; v < kShRegs / 4; ++v) {
            const float4 q4 = reinterpret_cast<const float4*>(my_sh)[v];
            shr[4 * v + 0] = q4.x; shr[4 * v + 1] = q4.y; shr[4 * v + 2] = q4.z; shr[4 * v + 3] = q4.w;
          }
        }
#define SH(k, ch) (TMA ? shr[(3 * (k) + (ch)) < kShRegs ? (3 * (k) + (ch)) : 0] : my_sh[3 * (k) + (ch)])
        if (D > 0) {
          coef[1] = -kSH1 * y; coef[2] = kSH1 * z; coef[3] = -kSH1 * x;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            dx_[ch] = -kSH1 * SH(3, ch);
            dy_[ch] = -kSH1 * SH(1, ch);
            dz_[ch] = kSH1 * SH(2, ch);
          }
          if (D > 1) {
            const float xx = x * x, yy = y * y, zz = z * z;
            const float xy = x * y, yz = y * z, xz = x * z;
            coef[4] = kSH2[0] * xy; coef[5] = kSH2[1] * yz; coef[6] = kSH2[2] * (2.f * zz - xx - yy);
            coef[7] = kSH2[3] * xz; coef[8] = kSH2[4] * (xx - yy);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
              dx_[ch] += kSH2[0] * y * SH(4, ch) + kSH2[2] * 2.f * -x * SH(6, ch) + kSH2[3] * z * SH(7, ch) + kSH2[4] * 2.f * x * SH(8, ch);
              dy_[ch] += kSH2[0] * x * SH(4, ch) + kSH2[1] * z * SH(5, ch) + kSH2[2] * 2.f * -y * SH(6, ch) + kSH2[4] * 2.f * -y * SH(8, ch);
              dz_[ch] += kSH2[1] * y * SH(5, ch) + kSH2[2] * 2.f * 2.f * z * SH(6, ch) + kSH2[3] * x * SH(7, ch);
            }
            if (D > 2) {
              coef[9] = kSH3[0] * y * (3.f * xx - yy);
              coef[10] = kSH3[1] * xy * z;
              coef[11] = kSH3[2] * y * (4.f * zz - xx - yy);
              coef[12] = kSH3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
              coef[13] = kSH3[4] * x * (4.f * zz - xx - yy);
              coef[14] = kSH3[5] * z * (xx - yy);
              coef[15] = kSH3[6] * x * (xx - 3.f * yy);
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) {
                dx_[ch] += (kSH3[0] * SH(9, ch) * 3.f * 2.f * xy + kSH3[1] * SH(10, ch) * yz +
                            kSH3[2] * SH(11, ch) * -2.f * xy + kSH3[3] * SH(12, ch) * -3.f * 2.f * xz +
                            kSH3[4] * SH(13, ch) * (-3.f * xx + 4.f * zz - yy) +
                            kSH3[5] * SH(14, ch) * 2.f * xz + kSH3[6] * SH(15, ch) * 3.f * (xx - yy));
                dy_[ch] += (kSH3[0] * SH(9, ch) * 3.f * (xx - yy) + kSH3[1] * SH(10, ch) * xz +
                            kSH3[2] * SH(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                            kSH3[3] * SH(12, ch) * -3.f * 2.f * yz + kSH3[4] * SH(13, ch) * -2.f * xy +
                            kSH3[5] * SH(14, ch) * -2.f * yz + kSH3[6] * SH(15, ch) * -3.f * 2.f * xy);
                dz_[ch] += (kSH3[1] * SH(10, ch) * xy + kSH3[2] * SH(11, ch) * 4.f * 2.f * yz +
                            kSH3[3] * SH(12, ch) * 3.f * (2.f * zz - xx - yy) +
                            kSH3[4] * SH(13, ch) * 4.f * 2.f * xz + kSH3[5] * SH(14, ch) * (xx - yy));
              }
            }
          }
        }
#undef SH
        const int ncoef = (D + 1) * (D + 1);
        if (TMA) {
          if (sh_out) {
            float o[kShRegs];
#pragma unroll
            for (int k = 0; k < kShRegs / 3; ++k) {
              const float ck = (k < ncoef && k < 16) ? coef[k < 16 ? k : 0] : 0.f;
              o[3 * k + 0] = ck * dR[0]; o[3 * k + 1] = ck * dR[1]; o[3 * k + 2] = ck * dR[2];
            }
#pragma unroll
            for (int v = 0; v < kShRegs / 4; ++v)
              reinterpret_cast<float4*>(my_sh)[v] = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
          }
        } else {
          for (int k = 0; k < M; ++k) {
            const float ck = (k < ncoef && k < 16) ? coef[k] : 0.f;
            my_sh[3 * k + 0] = ck * dR[0];
            my_sh[3 * k + 1] = ck * dR[1];
            my_sh[3 * k + 2] = ck * dR[2];
          }
        }
        const float3 dL_ddir = make_float3(dx_[0] * dR[0] + dx_[1] * dR[1] + dx_[2] * dR[2],
                                           dy_[0] * dR[0] + dy_[1] * dR[1] + dy_[2] * dR[2],
                                           dz_[0] * dR[0] + dz_[1] * dR[1] + dz_[2] * dR[2]);
        const float3 dmean_sh = dnormvdv3(dir_orig, dL_ddir);
        dmean.x += dmean_sh.x; dmean.y += dmean_sh.y; dmean.z += dmean_sh.z;

        if (VARIANT == kFull) {
          // d(rgb)/d(campos) = dRGB/d(dir) . d(dir)/d(campos) (full backward.cu:14-30,146-154),
          // contracted with the UNclamped colour gradient sum alpha*T*dL/dpixel (ComputePG part 1)
          const float len3 = len * len * len;
          const float il3 = 1.0f / len3, il = 1.0f / len;
          const float dxdCx = dir_orig.x * dir_orig.x * il3 - il;
          const float dydCx = dir_orig.x * dir_orig.y * il3;
          const float dzdCx = dir_orig.x * dir_orig.z * il3;
          const float dxdCy = dir_orig.x * dir_orig.y * il3;
          const float dydCy = dir_orig.y * dir_orig.y * il3 - il;
          const float dzdCy = dir_orig.y * dir_orig.z * il3;
          const float dxdCz = dir_orig.x * dir_orig.z * il3;
          const float dydCz = dir_orig.y * dir_orig.z * il3;
          const float dzdCz = dir_orig.z * dir_orig.z * il3 - il;
          const float gc[3] = {g_r, g_g, g_b};
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            dq.x += gc[ch] * (dx_[ch] * dxdCx + dy_[ch] * dydCx + dz_[ch] * dzdCx);
            dq.y += gc[ch] * (dx_[ch] * dxdCy + dy_[ch] * dydCy + dz_[ch] * dzdCy);
            dq.z += gc[ch] * (dx_[ch] * dxdCz + dy_[ch] * dydCz + dz_[ch] * dzdCz);
          }
        }
      }

      // ---------------- scale / rotation backward ----------------
      if (scales != nullptr) {
        const float3 sc = ld3(scales, idx);
        const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        M3 R;
        R.c[0][0] = 1.f - 2.f * (y * y + z * z); R.c[0][1] = 2.f * (x * y - r * z); R.c[0][2] = 2.f * (x * z + r * y);
        R.c[1][0] = 2.f * (x * y + r * z); R.c[1][1] = 1.f - 2.f * (x * x + z * z); R.c[1][2] = 2.f * (y * z - r * x);
        R.c[2][0] = 2.f * (x * z - r * y); R.c[2][1] = 2.f * (y * z + r * x); R.c[2][2] = 1.f - 2.f * (x * x + y * y);
        const float3 s = make_float3(scale_modifier * sc.x, scale_modifier * sc.y, scale_modifier * sc.z);
        M3 S;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) S.c[i][j] = (i == j) ? 1.0f : 0.0f;
        S.c[0][0] = s.x; S.c[1][1] = s.y; S.c[2][2] = s.z;
        const M3 Mm = m3_mul(S, R);
        M3 dSig;
        dSig.c[0][0] = dcov[0];        dSig.c[0][1] = 0.5f * dcov[1]; dSig.c[0][2] = 0.5f * dcov[2];
        dSig.c[1][0] = 0.5f * dcov[1]; dSig.c[1][1] = dcov[3];        dSig.c[1][2] = 0.5f * dcov[4];
        dSig.c[2][0] = 0.5f * dcov[2]; dSig.c[2][1] = 0.5f * dcov[4]; dSig.c[2][2] = dcov[5];
        M3 M2;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) M2.c[i][j] = 2.0f * Mm.c[i][j];
        const M3 dL_dM = m3_mul(M2, dSig);
        const M3 Rt = m3_transpose(R);
        M3 dMt = m3_transpose(dL_dM);
        const float dsx = Rt.c[0][0] * dMt.c[0][0] + Rt.c[0][1] * dMt.c[0][1] + Rt.c[0][2] * dMt.c[0][2];
        const float dsy = Rt.c[1][0] * dMt.c[1][0] + Rt.c[1][1] * dMt.c[1][1] + Rt.c[1][2] * dMt.c[1][2];
        const float dsz = Rt.c[2][0] * dMt.c[2][0] + Rt.c[2][1] * dMt.c[2][1] + Rt.c[2][2] * dMt.c[2][2];
        st3(out.dL_dscale, idx, dsx, dsy, dsz);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          dMt.c[0][j] *= s.x;
          dMt.c[1][j] *= s.y;
          dMt.c[2][j] *= s.z;
        }
        float4 dq4;
        dq4.x = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) + 2 * x * (dMt.c[1][2] - dMt.c[2][1]);
        dq4.y = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) + 2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
        dq4.z = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) + 2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
        dq4.w = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) + 2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
        if (out.dL_drot) reinterpret_cast<float4*>(out.dL_drot)[idx] = dq4;  // w.r.t. the raw quaternion
      } else {
        st3(out.dL_dscale, idx, 0.f, 0.f, 0.f);
        if (out.dL_drot) reinterpret_cast<float4*>(out.dL_drot)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
      }

      st3(out.dL_dmean3D, idx, dmean.x, dmean.y, dmean.z);
      st3(out.dL_dmean2D, idx, g_mx, g_my, 0.f);
      // densification statistics of Inria 3DGS / CG-SLAM mapping (add_densification_stats:
      // xyz_gradient_accum += |viewspace grad.xy|, denom += 1, max_radii2D = max(., radii) over the
      // visible Gaussians), accumulated here instead of by five torch kernels over [P]
      if (out.densify_grad_accum != nullptr) {
        out.densify_grad_accum[idx] += sqrtf(g_mx * g_mx + g_my * g_my);
        out.densify_denom[idx] += 1.0f;
      }
      if (out.max_radii2D != nullptr) out.max_radii2D[idx] = fmaxf(out.max_radii2D[idx], (float)radii[idx]);
      if (out.dL_dconic) reinterpret_cast<float4*>(out.dL_dconic)[idx] = make_float4(g_ca, g_cb, 0.f, g_cc);
      if (out.dL_dopacity) out.dL_dopacity[idx] = g_op;
      st3(out.dL_dcolor, idx, g_r, g_g, g_b);
      if (out.dL_ddepth) out.dL_ddepth[idx] = g_depth;
    }

    // ---------------- pose gradient, per-Gaussian contraction ----------------
    if (want_pose) {
      const float p0 = perspec[0], p5 = perspec[5];
      const float gx = (VARIANT == kLight) ? g_mx : g_pgx;
      const float gy = (VARIANT == kLight) ? g_my : g_pgy;
      const float ax = m_w * p0;              // d x_ndc / d v{0,4,8,12}   = ax * (m,1)
      const float by = m_w * p5;              // d y_ndc / d v{1,5,9,13}   = by * (m,1)
      const float cx = m_hom.x * (-m_w * m_w);  // d x_ndc / d v{2,6,10,14} = cx * (m,1)
      const float cy = m_hom.y * (-m_w * m_w);  // d y_ndc / d v{2,6,10,14} = cy * (m,1)
      const float mm[4] = {m.x, m.y, m.z, 1.0f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pose[3 * k + 0] = (ax * mm[k]) * gx;
        pose[3 * k + 1] = (by * mm[k]) * gy;
        pose[3 * k + 2] = (cx * mm[k]) * gx + (cy * mm[k]) * gy + mm[k] * g_pd;
      }
      if (VARIANT == kFull) {
        // colour through the camera position: campos = -R^T t, so
        // d campos_x / d v{0,1,2} = -v{12,13,14}, ... , d campos / d v12 = -(v0,v4,v8) etc.
        pose[0] += dq.x * (-view[12]); pose[1] += dq.x * (-view[13]); pose[2] += dq.x * (-view[14]);
        pose[3] += dq.y * (-view[12]); pose[4] += dq.y * (-view[13]); pose[5] += dq.y * (-view[14]);
        pose[6] += dq.z * (-view[12]); pose[7] += dq.z * (-view[13]); pose[8] += dq.z * (-view[14]);
        pose[9] += dq.x * (-view[0]) + dq.y * (-view[4]) + dq.z * (-view[8]);
        pose[10] += dq.x * (-view[1]) + dq.y * (-view[5]) + dq.z * (-view[9]);
        pose[11] += dq.x * (-view[2]) + dq.y * (-view[6]) + dq.z * (-view[10]);
      }
    }
  }

  // dL/dSH out: one bulk store per thread of its own row (TMA) / coalesced store of the block's slab
  if (TMA) {
    if (bar_pending) mbar_wait(&s_bar, 0u);  // the block's shared memory must outlive the loads in flight
    if (sh_out && idx < P) {
      bulk_s2g_fence();
      bulk_s2g(out.dL_dsh + (size_t)idx * (MT * 3), my_sh, kRowBytes);
    }
  } else if (sh_out) {
    __syncthreads();
    smem_to_rows<MT * 3>(out.dL_dsh + (size_t)base * M * 3, sh_smem, nvalid, M * 3, threadIdx.x,
                         kBwdThreads);
  }

  if (want_pose) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      float s = pose[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_pose[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
      float s = 0.f;
#pragma unroll
      for (int wq = 0; wq < kBwdThreads / 32; ++wq) s += s_pose[wq][threadIdx.x];
      pose_partials[(size_t)blockIdx.x * 12 + threadIdx.x] = s;
    }
  }
  // Fused final reduction (done_counter != NULL): the LAST block to finish adds all blocks' partials
  // in a fixed order (deterministic, no float atomics) and writes the 16-float dL/dviewmatrix — the
  // separate pose_finalize launch disappears.  Entries 3, 7, 11, 15 are 0, as in the reference.
  if (done_counter != nullptr) {
    __shared__ bool s_last;
    __shared__ float s_fin[kBwdThreads / 32][12];
    if (!want_pose) {
      if (blockIdx.x == 0 && threadIdx.x < 16) out.dL_dview[threadIdx.x] = 0.f;
    } else {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
      __syncthreads();
      if (s_last) {
        __threadfence();
        float a[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = 0.f;
        const float4* p4 = reinterpret_cast<const float4*>(pose_partials);
        for (int b = threadIdx.x; b < (int)gridDim.x; b += kBwdThreads) {
          const float4 q0 = __ldcg(p4 + 3 * (size_t)b), q1 = __ldcg(p4 + 3 * (size_t)b + 1), q2 = __ldcg(p4 + 3 * (size_t)b + 2);
          a[0] += q0.x; a[1] += q0.y; a[2] += q0.z; a[3] += q0.w;
          a[4] += q1.x; a[5] += q1.y; a[6] += q1.z; a[7] += q1.w;
          a[8] += q2.x; a[9] += q2.y; a[10] += q2.z; a[11] += q2.w;
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          float s = a[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0) s_fin[warp][k] = s;
        }
        __syncthreads();
        if (threadIdx.x < 12) {
          float t = 0.f;
#pragma unroll
          for (int wq = 0; wq < kBwdThreads / 32; ++wq) t += s_fin[wq][threadIdx.x];
          const int col = threadIdx.x / 3, rowi = threadIdx.x % 3;  // pose[] order: (v0,v1,v2),(v4,v5,v6),...
          out.dL_dview[4 * col + rowi] = t;
          if (rowi == 0) out.dL_dview[4 * col + 3] = 0.f;
        }
      }
    }
  }
  if (TMA && sh_out && idx < P) bulk_s2g_wait_read();
}

// Adds the per-block pose partials in a fixed order (deterministic, no atomics) and scatters the
// 12 live entries into the 16-float column-major dL/dviewmatrix (entries 3,7,11,15 are 0, as in
// the reference).  One CTA per live entry: 1024 threads stride over the blocks, then a tree.
__global__ void __launch_bounds__(1024)
pose_finalize_kernel(int nblocks, const float* __restrict__ partials, float* __restrict__ dL_dview,
                     bool want_pose) {
  __shared__ float s[32];
  const int k = blockIdx.x;  // pose[] order: (v0,v1,v2),(v4,v5,v6),(v8,v9,v10),(v12,v13,v14)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float sum = 0.f;
  if (want_pose)
    for (int b = tid; b < nblocks; b += 1024) sum += partials[(size_t)b * 12 + k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float t = s[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      const int col = k / 3, rowi = k % 3;
      dL_dview[4 * col + rowi] = t;
      if (rowi == 0) dL_dview[4 * col + 3] = 0.f;
    }
  }
}

}  // namespace

int launch_preprocess_bwd(int variant, int P, int D, int M, const float* means3D,
                          const int* radii, const float* shs, const float* scales,
                          const float* rotations, float scale_modifier,
                          const float* cov3D_precomp, const Camera& cam, const float* perspec,
                          const GeomState& g, const float* acc, float* pose_partials,
                          const GaussGradOut& out, bool want_gauss, bool want_pose, bool debug,
                          cudaStream_t stream, unsigned int* done_counter) {
  const int blocks = (P + kBwdThreads - 1) / kBwdThreads;
  // one block adds `blocks` x 48 bytes of partials: worth a launch up to ~1 M Gaussians
  if (blocks > kFuseFinalizeMaxBlocks) done_counter = nullptr;
  const bool tma = shs != nullptr && (M == 16 || M == 4) && options().bulk_sh != 0 &&
                   (reinterpret_cast<uintptr_t>(shs) & 15) == 0 &&
                   (out.dL_dsh == nullptr || (reinterpret_cast<uintptr_t>(out.dL_dsh) & 15) == 0);
  const size_t smem = shs == nullptr ? 0
                      : tma          ? sizeof(float) * kBwdThreads * (size_t)bulk_row_floats(M * 3)
                                     : sizeof(float) * kBwdThreads * (size_t)(M * 3 + 1);
  const float* cov3D = cov3D_precomp != nullptr ? cov3D_precomp : g.cov3D;
  StageScope st(ST_PRE_BWD, stream, done_counter != nullptr ? 1 : 2);
  // With want_gauss == false (-light, map_off: light backward.cu:593,609,654,666;
  // rasterizer_impl.cu:467) the kernel itself writes the zero gradients: no memset passes.
  if (shs == nullptr && out.dL_dsh != nullptr && M > 0) {
    GSR_CUDA_OK(cudaMemsetAsync(out.dL_dsh, 0, sizeof(float) * 3 * (size_t)P * (size_t)M, stream));
  }
  {
#define GSR_PRE_BWD(V, MT, TMA)                                                                  \
  prefer_max_shared_once(reinterpret_cast<const void*>(&preprocess_bwd_kernel<V, MT, TMA, ((MT) == 16 && (TMA)) ? 5 : 1>)); \
  launch_after(options().pdl != 0, preprocess_bwd_kernel<V, MT, TMA, ((MT) == 16 && (TMA)) ? 5 : 1>,   \
      dim3(blocks), dim3(kBwdThreads), smem, stream,                                             \
      P, D, M, means3D, (const int*)radii, shs, (const unsigned char*)g.clamped, scales, rotations, scale_modifier, cov3D, cam.view, \
      cam.proj, cam.campos, perspec, cam.focal_x, cam.focal_y, cam.tan_fovx, cam.tan_fovy, (const float*)acc,  \
      (const float4*)g.rec, cam.W, cam.H, pose_partials, out, want_gauss, want_pose, nullptr, done_counter)
#define GSR_PRE_BWD_M(V)                                                                         \
  if (tma) {   /* 5 CTAs per SM (96 registers); the 6-CTA build (80 registers, 72 bytes spilled) measured equal */ \
    if (M == 16) { GSR_PRE_BWD(V, 16, true); } else { GSR_PRE_BWD(V, 4, true); }                 \
  } else {                                                                                       \
    switch (M) {                                                                                 \
      case 16: GSR_PRE_BWD(V, 16, false); break;                                                 \
      case 9: GSR_PRE_BWD(V, 9, false); break;                                                   \
      case 4: GSR_PRE_BWD(V, 4, false); break;                                                   \
      case 1: GSR_PRE_BWD(V, 1, false); break;                                                   \
      default: GSR_PRE_BWD(V, 0, false); break;                                                  \
    }                                                                                            \
  }
    if (variant == kLight) {
      GSR_PRE_BWD_M(kLight)
    } else {
      GSR_PRE_BWD_M(kFull)
    }
#undef GSR_PRE_BWD_M
#undef GSR_PRE_BWD
    GSR_LAUNCH_OK(debug, stream);
  }
  if (done_counter == nullptr) {
    pose_finalize_kernel<<<12, 1024, 0, stream>>>(blocks, pose_partials, out.dL_dview, want_pose);
    GSR_LAUNCH_OK(debug, stream);
  }
  return GSR_OK;
}

int preprocess_bwd_blocks(int P) { return (P + kBwdThreads - 1) / kBwdThreads; }

// Pose contraction only (tracker): no per-Gaussian gradient is written and the per-block partial
// sums of the 12 live dL/dviewmatrix entries are left in `pose_partials` for the caller's own
// reduction (track_update_kernel).
int launch_preprocess_bwd_partials(int variant, int P, int D, int M, const float* means3D,
                                   const int* radii, const Camera& cam, const float* perspec,
                                   const GeomState& g, float* acc, float* pose_partials,
                                   bool clear_acc, cudaStream_t stream) {
  if (variant != kLight) { set_error("pose-only backward exists for -light only"); return GSR_E_INVALID; }
  const int blocks = preprocess_bwd_blocks(P);
  StageScope st(ST_PRE_BWD, stream, 1);
  preprocess_bwd_kernel<kLight, 0, false><<<blocks, kBwdThreads, 0, stream>>>(
      P, D, M, means3D, radii, nullptr, g.clamped, nullptr, nullptr, 1.0f, g.cov3D, cam.view, cam.proj,
      cam.campos, perspec, cam.focal_x, cam.focal_y, cam.tan_fovx, cam.tan_fovy, acc, g.rec, cam.W,
      cam.H, pose_partials, GaussGradOut{}, false, true, clear_acc ? acc : nullptr, nullptr);
  GSR_LAUNCH_OK(false, stream);
  return GSR_OK;
}

}  // namespace gsr
