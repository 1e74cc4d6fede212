// render_bwd.cu — backward tile blend: re-walks each tile's entry list back to front, restores
// T by division, and accumulates per-Gaussian gradient sums.
//
// Reference semantics: BACKWARD::renderCUDA of -light (light backward.cu:419-699) and of -full
// (full backward.cu:540-836) plus the per-pair part of -full's ComputePG (:838-1338).
//
// What is different by design (B200-first, same results):
//  * the reference issues ~10 global atomicAdd per (pixel, Gaussian) pair; here the 32 pixels of
//    a warp (a compact 8x4 block) first reduce their 14 partial sums through a shared-memory
//    transpose (14 stores, 4 x LDS.128 + one shuffle per reducing lane) and 14 lanes then issue
//    one coalesced red.add into the Gaussian's 64-byte accumulator line — whole warps that do not
//    touch a Gaussian skip it;
//  * mean2D / conic / opacity gradients are accumulated as moments of w = G dL/dalpha over (dx, dy)
//    and converted per Gaussian in preprocess_bwd;
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

constexpr int kRedVals = 14;    // accumulator slots reduced per (warp, Gaussian) hit
constexpr int kRedStride = 36;  // floats per slot row in shared memory (144 B: 16-byte aligned rows
                                // whose bank offsets rotate by 4, conflict-free per quarter warp)

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
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
  __shared__ unsigned short s_mask[kTileThreads];
  __shared__ unsigned char s_list[kTileThreads / 16][kTileThreads];
  __shared__ __align__(16) float s_red[kTileThreads / 32][kRedVals * kRedStride];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.y * grid_x + blockIdx.x;
  int lx, ly, sub;
  pixel_of_thread(warp, lane, lx, ly, sub);
  const int half = lane >> 4;
  const int px = blockIdx.x * kTileX + lx;
  const int py = blockIdx.y * kTileY + ly;
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
      my_mask = block_mask16(q0, q1, tile_x0, tile_y0);
    }
    s_mask[tid] = (unsigned short)my_mask;
    __syncthreads();

    // per-half-warp compaction: only entries whose cut ellipse can touch the half warp's 4x4
    // pixel block; the two halves walk their own lists side by side
    const int nb = min(kTileThreads, walk - i * kTileThreads);
    const int sub_lo = sub - half, sub_hi = sub_lo + 1;
    int cnt_lo = 0, cnt_hi = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int c = 0; c * 32 < nb; ++c) {
      const int jj = c * 32 + lane;
      const unsigned m = (jj < nb) ? (unsigned)s_mask[jj] : 0u;
      const bool hit_lo = (m >> sub_lo) & 1u, hit_hi = (m >> sub_hi) & 1u;
      const unsigned ball_lo = __ballot_sync(0xffffffffu, hit_lo);
      const unsigned ball_hi = __ballot_sync(0xffffffffu, hit_hi);
      if (hit_lo) s_list[2 * warp][cnt_lo + __popc(ball_lo & lt)] = (unsigned char)jj;
      if (hit_hi) s_list[2 * warp + 1][cnt_hi + __popc(ball_hi & lt)] = (unsigned char)jj;
      cnt_lo += __popc(ball_lo);
      cnt_hi += __popc(ball_hi);
    }
    __syncwarp();
    const int cnt = half ? cnt_hi : cnt_lo;
    const int cnt_max = max(cnt_lo, cnt_hi);
    const unsigned char* my_list = s_list[2 * warp + half];
    for (int k = 0; k < cnt_max; ++k) {
      const bool active = k < cnt;
      const int j = active ? (int)my_list[k] : 0;
      const int pos = walk - (i * kTileThreads + j) - 1;  // 0-based list position
      const float4 r0 = s_r0[j];
      const float4 r1 = s_r1[j];
      const float dx = GSR_SUB(r0.x, pixfx), dy = GSR_SUB(r0.y, pixfy);
      const float power = pair_power(r0.z, r0.w, r1.x, dx, dy);
      bool valid = active && (pos < last_contributor) && !(power > 0.0f) && !(power < r1.z);
      float G = 0.f, alpha = 0.f;
      if (valid) {
        G = expf(power);
        alpha = pair_alpha(r1.y, G);
        valid = !(alpha < kAlphaMin);
      }
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      if (vmask == 0u) continue;

      // Per-pair partial sums.  The screen-space mean, conic and opacity gradients are all moments
      // of w = G * dL/dalpha over (dx, dy); they are summed as moments here and turned into
      // dL/dmean2D, dL/dconic by preprocess_bwd with the per-Gaussian conic (fewer per-pair FLOPs).
      float v[kRedVals];
#pragma unroll
      for (int q = 0; q < kRedVals; ++q) v[q] = 0.f;
      if (valid) {
        const float4 r2 = s_r2[j];
        const float c_d = r1.w;
        const float inv = fast_rcp(1.f - alpha);   // 1 - alpha >= 0.01
        T = T * inv;
        const float aT = alpha * T;
        const float one_m_la = 1.f - last_alpha;

        ar0 = last_alpha * lc0 + one_m_la * ar0;
        ar1 = last_alpha * lc1 + one_m_la * ar1;
        ar2 = last_alpha * lc2 + one_m_la * ar2;
        lc0 = r2.x; lc1 = r2.y; lc2 = r2.z;
        const float colour_part = (r2.x - ar0) * dLp0 + (r2.y - ar1) * dLp1 + (r2.z - ar2) * dLp2;
        v[ACC_R] = aT * dLp0;
        v[ACC_G] = aT * dLp1;
        v[ACC_B] = aT * dLp2;

        const float dgt = c_d - gt;
        const float c_var = dgt * dgt;
        adr = last_alpha * last_depth + one_m_la * adr;
        avr = last_alpha * last_var + one_m_la * avr;
        last_depth = c_d;
        last_var = c_var;
        const float depth_part = (c_d - adr) * dLd;
        float dL_dalpha = colour_part + depth_part + (c_var - avr) * dLv;
        const float aTd = aT * dLd;
        v[ACC_DEPTH] = aTd + dLv * aT * 2.f * dgt;

        last_alpha = alpha;
        dL_dalpha = dL_dalpha * T - (T_final * inv) * bg_dot_dpixel;

        const float w = G * dL_dalpha;
        const float wx = w * dx, wy = w * dy;
        v[ACC_OP] = w;
        v[ACC_MX] = wx;        // S1  = sum w dx
        v[ACC_MY] = wy;        // S2  = sum w dy
        v[ACC_CA] = wx * dx;   // S11 = sum w dx^2
        v[ACC_CB] = wx * dy;   // S12 = sum w dx dy
        v[ACC_CC] = wy * dy;   // S22 = sum w dy^2

        if (VARIANT == kLight) {
          v[ACC_PD] = aTd;
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
            pa += T * depth_part;
            v[ACC_PD] = aTd;
          }
          const float q = pa * G;
          v[ACC_PGX] = q * dx;  // Q1
          v[ACC_PGY] = q * dy;  // Q2
        }
      }
      // warp reduction through shared memory: every lane stores its column, 28 lanes each add
      // half a row (4 x LDS.128), pairs combine with one shuffle, 14 lanes issue one coalesced red
      float* red = s_red[warp];
#pragma unroll
      for (int q = 0; q < kRedVals; ++q) red[q * kRedStride + lane] = v[q];
      __syncwarp();
      float sum = 0.f;
      if (lane < 2 * kRedVals) {
        const float4* row = reinterpret_cast<const float4*>(red + (lane >> 1) * kRedStride + (lane & 1) * 16);
        const float4 a = row[0], b = row[1], c = row[2], d = row[3];
        sum = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w)) +
              (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
      }
      // lane (2q + h) now holds slot q summed over half warp h, i.e. over that half's own entry
      const int j_lo = __shfl_sync(0xffffffffu, j, 0), j_hi = __shfl_sync(0xffffffffu, j, 16);
      if (lane < 2 * kRedVals) {
        const int h = lane & 1;
        if ((vmask >> (16 * h)) & 0xFFFFu)
          atomicAdd(acc + (size_t)s_id[h ? j_hi : j_lo] * kAccStride + (lane >> 1), sum);
      }
      __syncwarp();
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
