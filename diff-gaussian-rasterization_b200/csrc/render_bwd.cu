// render_bwd.cu — backward tile blend: re-walks each tile's entry list back to front, restores
// T by division, and accumulates per-Gaussian gradient sums.
//
// Reference semantics: BACKWARD::renderCUDA of -light (light backward.cu:419-699) and of -full
// (full backward.cu:540-836) plus the per-pair part of -full's ComputePG (:838-1338).
//
// What is different by design (B200-first, same results):
//  * the reference issues ~10 global atomicAdd per (pixel, Gaussian) pair; here the 32 pixels of
//    a warp (a compact 8x4 block) first reduce their 16 partial sums with a transposing
//    butterfly (16 shuffles instead of 80) and 16 lanes then issue one coalesced 64-byte red.add
//    into the Gaussian's accumulator record — whole warps that do not touch a Gaussian skip it;
//  * the pose gradient needs no per-pixel [H*W,16] tensor (light) and no 92-byte-per-pair scratch
//    + second tile walk (full ComputePG): every pose term is (per-pair scalar) x (per-Gaussian
//    vector), so the per-pair scalars are summed per Gaussian here (slots ACC_PGX/PGY/PD) and
//    contracted with the per-Gaussian Jacobians in preprocess_bwd;
//  * the walk starts at the last entry any pixel of the tile actually blended (tile_last), not at
//    the end of the list.
// Bound: FP32 issue + shuffle + L2 atomics, not HBM.
#include "gsr_common.cuh"

namespace gsr {

namespace {

// Sum v[0..15] over the 32 lanes of a warp.  On return lane L holds the total of value (L >> 1)
// in v[0] (both lanes of a pair hold the same number).
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  const bool up16 = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float send = up16 ? v[k] : v[k + 8];
    const float keep = up16 ? v[k + 8] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  const bool up8 = (lane & 8) != 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = up8 ? v[k] : v[k + 4];
    const float keep = up8 ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const bool up4 = (lane & 4) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = up4 ? v[k] : v[k + 2];
    const float keep = up4 ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const bool up2 = (lane & 2) != 0;
  {
    const float send = up2 ? v[0] : v[1];
    const float keep = up2 ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

template <int VARIANT>
__global__ void __launch_bounds__(kTileThreads)
render_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                  const uint32_t* __restrict__ tile_last, int W, int H, int grid_x,
                  const float4* __restrict__ rec, const float* __restrict__ bg,
                  const float* __restrict__ gt_depth,
                  const float* __restrict__ alphas,      // light: T_final = 1 - alphas[pix]
                  const float* __restrict__ final_Ts,    // full
                  const uint32_t* __restrict__ n_contrib,
                  const uint32_t* __restrict__ first_contrib,  // full
                  const float* __restrict__ dL_dpix, const float* __restrict__ dL_ddepths,
                  const float* __restrict__ dL_dmedians,  // light
                  const float* __restrict__ dL_dvars,     // light: depth_var, full: uncertainty
                  float* __restrict__ acc) {
  __shared__ float4 s_r0[kTileThreads];
  __shared__ float4 s_r1[kTileThreads];
  __shared__ float4 s_r2[kTileThreads];
  __shared__ int s_id[kTileThreads];
  __shared__ unsigned char s_mask[kTileThreads];
  __shared__ unsigned char s_list[kTileThreads / 32][kTileThreads];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.y * grid_x + blockIdx.x;
  const int px = blockIdx.x * kTileX + (warp & 1) * 8 + (lane & 7);
  const int py = blockIdx.y * kTileY + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < W && py < H;
  const uint32_t pix_id = (uint32_t)W * (uint32_t)py + (uint32_t)px;
  const float pixfx = (float)px, pixfy = (float)py;

  const uint2 range = ranges[tile];
  const int walk = (int)tile_last[tile];  // entries [0, walk) were used by some pixel
  const int rounds = (walk + kTileThreads - 1) / kTileThreads;

  const size_t HW = (size_t)H * (size_t)W;
  float T_final = 0.f;
  int last_contributor = 0;
  int first = 0;
  float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f, dLd = 0.f, dLv = 0.f, dLm = 0.f, gt = 0.f;
  if (inside) {
    T_final = (VARIANT == kLight) ? (1 - alphas[pix_id]) : final_Ts[pix_id];
    last_contributor = (int)n_contrib[pix_id];
    dLp0 = dL_dpix[0 * HW + pix_id];
    dLp1 = dL_dpix[1 * HW + pix_id];
    dLp2 = dL_dpix[2 * HW + pix_id];
    dLd = dL_ddepths[pix_id];
    dLv = dL_dvars[pix_id];
    gt = gt_depth[pix_id];
    if (VARIANT == kLight) dLm = dL_dmedians[pix_id];
    if (VARIANT == kFull) first = (int)first_contrib[pix_id];
  }
  float T = T_final;
  float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, adr = 0.f, avr = 0.f;
  float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_depth = 0.f, last_var = 0.f;
  const float bg_dot_dpixel = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
  const float ddelx_dx = 0.5f * W;
  const float ddely_dy = 0.5f * H;
  bool mid_once = true;
  const float tile_x0 = (float)(blockIdx.x * kTileX), tile_y0 = (float)(blockIdx.y * kTileY);

  for (int i = 0; i < rounds; ++i) {
    __syncthreads();
    const int progress = i * kTileThreads + tid;
    unsigned my_mask = 0u;
    if (progress < walk) {
      const int id = (int)point_list[range.x + (walk - progress - 1)];
      s_id[tid] = id;
      const float4* r = rec + 3 * (size_t)id;
      const float4 q0 = __ldg(r + 0), q1 = __ldg(r + 1);
      s_r0[tid] = q0;
      s_r1[tid] = q1;
      s_r2[tid] = __ldg(r + 2);
      my_mask = block_mask8(q0, q1, tile_x0, tile_y0);
    }
    s_mask[tid] = (unsigned char)my_mask;
    __syncthreads();

    // per-warp compaction: only entries whose cut ellipse can touch this warp's 8x4 block
    const int nb = min(kTileThreads, walk - i * kTileThreads);
    int cnt = 0;
    for (int c = 0; c * 32 < nb; ++c) {
      const int jj = c * 32 + lane;
      const bool hit = (jj < nb) && ((s_mask[jj] >> warp) & 1u);
      const unsigned ball = __ballot_sync(0xffffffffu, hit);
      if (hit) s_list[warp][cnt + __popc(ball & ((1u << lane) - 1u))] = (unsigned char)jj;
      cnt += __popc(ball);
    }
    __syncwarp();
    for (int k = 0; k < cnt; ++k) {
      const int j = s_list[warp][k];
      const int pos = walk - (i * kTileThreads + j) - 1;  // 0-based list position
      const float4 r0 = s_r0[j];
      const float4 r1 = s_r1[j];
      const float dx = GSR_SUB(r0.x, pixfx), dy = GSR_SUB(r0.y, pixfy);
      const float power = pair_power(r0.z, r0.w, r1.x, dx, dy);
      bool valid = (pos < last_contributor) && !(power > 0.0f) && !(power < r1.z);
      float G = 0.f, alpha = 0.f;
      if (valid) {
        G = expf(power);
        alpha = pair_alpha(r1.y, G);
        valid = !(alpha < kAlphaMin);
      }
      if (!__any_sync(0xffffffffu, valid)) continue;

      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = 0.f;
      if (valid) {
        const float4 r2 = s_r2[j];
        const float o = r1.y;
        const float c_d = r1.w;
        T = T / (1.f - alpha);
        const float aT = alpha * T;

        float dL_dalpha = 0.0f;
        ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0;
        lc0 = r2.x;
        dL_dalpha += (r2.x - ar0) * dLp0;
        ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1;
        lc1 = r2.y;
        dL_dalpha += (r2.y - ar1) * dLp1;
        ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2;
        lc2 = r2.z;
        dL_dalpha += (r2.z - ar2) * dLp2;
        v[ACC_R] = aT * dLp0;
        v[ACC_G] = aT * dLp1;
        v[ACC_B] = aT * dLp2;
        const float colour_part = dL_dalpha;  // sum_ch (c - accum_rec) * dL/dpixel

        const float c_var = (c_d - gt) * (c_d - gt);
        adr = last_alpha * last_depth + (1.f - last_alpha) * adr;
        last_depth = c_d;
        avr = last_alpha * last_var + (1.f - last_alpha) * avr;
        last_var = c_var;
        dL_dalpha += (c_d - adr) * dLd;
        dL_dalpha += (c_var - avr) * dLv;
        v[ACC_DEPTH] = aT * dLd + dLv * aT * 2.f * (c_d - gt);

        dL_dalpha *= T;
        last_alpha = alpha;
        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

        const float dL_dG = o * dL_dalpha;
        const float gdx = G * dx;
        const float gdy = G * dy;
        const float dG_ddelx = -gdx * r0.z - gdy * r0.w;
        const float dG_ddely = -gdy * r1.x - gdx * r0.w;
        v[ACC_MX] = dL_dG * dG_ddelx * ddelx_dx;
        v[ACC_MY] = dL_dG * dG_ddely * ddely_dy;
        v[ACC_CA] = -0.5f * gdx * dx * dL_dG;
        v[ACC_CB] = -0.5f * gdx * dy * dL_dG;
        v[ACC_CC] = -0.5f * gdy * dy * dL_dG;
        v[ACC_OP] = G * dL_dalpha;

        if (VARIANT == kLight) {
          v[ACC_PD] = aT * dLd;
          if (T > 0.5f && mid_once) {
            v[ACC_MED] = dLm;
            mid_once = false;
          }
        } else {
          // pose terms of the reference's ComputePG: colour through ndc without the background
          // term (full backward.cu:746-777, :1028-1072); depth only from the front-most valid
          // contributor of the pixel, because dd_dv* is assigned, not accumulated (:1278-1289).
          float pa = T * colour_part;
          if (pos + 1 == first) {
            pa += dLd * (T * (c_d - adr));
            v[ACC_PD] = aT * dLd;
          }
          v[ACC_PGX] = pa * o * dG_ddelx * ddelx_dx;
          v[ACC_PGY] = pa * o * dG_ddely * ddely_dy;
        }
      }
      const float total = warp_reduce16(v, lane);
      if ((lane & 1) == 0) atomicAdd(acc + (size_t)s_id[j] * kAccStride + (lane >> 1), total);
    }
  }
}

}  // namespace

int launch_render_bwd(int variant, const Camera& cam, const GeomState& g, const BinState& b,
                      const ImgState& img, const float* bg, const float* gt_depth,
                      const float* alphas, const BlendGrads& cot, float* acc, bool debug,
                      cudaStream_t stream) {
  dim3 grid(cam.grid_x, cam.grid_y, 1);
  StageScope st(ST_RENDER_BWD, stream);
  if (variant == kLight) {
    render_bwd_kernel<kLight><<<grid, kTileThreads, 0, stream>>>(
        img.ranges, b.vals, img.tile_last, cam.W, cam.H, cam.grid_x, g.rec, bg, gt_depth, alphas,
        nullptr, img.n_contrib, nullptr, cot.dL_dpix, cot.dL_ddepth, cot.dL_dmedian, cot.dL_dvar,
        acc);
  } else {
    render_bwd_kernel<kFull><<<grid, kTileThreads, 0, stream>>>(
        img.ranges, b.vals, img.tile_last, cam.W, cam.H, cam.grid_x, g.rec, bg, gt_depth, nullptr,
        img.final_T, img.n_contrib, img.first_contrib, cot.dL_dpix, cot.dL_ddepth, nullptr,
        cot.dL_dvar, acc);
  }
  GSR_LAUNCH_OK(debug, stream);
  return GSR_OK;
}

}  // namespace gsr
