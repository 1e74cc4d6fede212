// render_fwd.cu — forward tile blend: one 256-thread CTA per 16x16 tile, one pixel per thread,
// front-to-back alpha compositing of the tile's depth-sorted entry list.
//
// Reference semantics: FORWARD::renderCUDA of -full (cuda_rasterizer/forward.cu:261-396) and of
// -light (light forward.cu:261-412).  Differences between the two that are reproduced here:
//   full : the Gaussian that drives T below 1e-4 IS blended, then the pixel stops;
//          outputs colour (+T*bg), depth, "uncertainty" = sum(alpha*T), final_T, n_contrib.
//   light: the Gaussian that would drive T below 1e-4 is NOT blended (Inria behaviour);
//          outputs colour, depth, alpha = sum(alpha*T), median depth (entry at which T crosses
//          0.5), depth_var == 0, n_contrib; per-Gaussian atomics gau_uncertainty /
//          gau_related_pixels at the median crossing.
// The per-pair arithmetic (power, expf, min(0.99, .), the 15/255 cut) is written with the
// reference's association so that hard decisions do not flip.
//
// B200 design: each warp owns a compact 8x4 pixel block (better whole-warp rejection than the
// reference's 16x2 strips); the per-Gaussian state is one packed 48-byte record gathered once per
// entry into shared memory (3 x 128-bit loads; the 48 MB record table of a 1 M scene is
// L2-resident); a conservative per-Gaussian power cut skips expf for pairs that cannot reach
// alpha >= 15/255.  Bound: FP32 issue / MUFU, not HBM.
#include "gsr_common.cuh"

namespace gsr {

namespace {

template <int VARIANT, bool LOSS>
__global__ void __launch_bounds__(kTileThreads)
render_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                  int W, int H, int grid_x, const float4* __restrict__ rec,
                  const float* __restrict__ bg, const float* __restrict__ gt_depth,
                  float* __restrict__ out_color, float* __restrict__ out_depth,
                  float* __restrict__ out_aux0,   // light: alpha      full: uncertainty
                  float* __restrict__ out_median, // light only
                  float* __restrict__ out_var,    // light only (always 0)
                  float* __restrict__ gau_unc, int* __restrict__ gau_px,  // light only
                  uint32_t* __restrict__ n_contrib, float* __restrict__ final_T,
                  uint32_t* __restrict__ first_contrib, uint32_t* __restrict__ tile_last,
                  uint32_t* __restrict__ related_counter, FusedLoss fl) {
  __shared__ float4 s_rec[3][kTileThreads];  // one array: the three rows of an entry are a constant offset apart
  __shared__ int s_id[kTileThreads];
  __shared__ unsigned short s_mask[kTileThreads];                      // sub-block mask per entry
  __shared__ unsigned char s_list[kTileThreads / 16][kTileThreads];    // per-half-warp compacted entries
  __shared__ uint32_t s_red[kTileThreads / 32];
  __shared__ float s_loss[kTileThreads / 32];

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
  const int total = (int)(range.y - range.x);
  const int rounds = (total + kTileThreads - 1) / kTileThreads;
  int todo = total;

  bool done = !inside;
  float T = 1.0f;
  uint32_t last_contributor = 0, first = 0xFFFFFFFFu, valid = 0;
  const float tile_x0 = (float)(blockIdx.x * kTileX), tile_y0 = (float)(blockIdx.y * kTileY);
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f, Wsum = 0.f, Dmed = 0.f;
  float gt = 0.f;
  if (VARIANT == kLight && inside) gt = gt_depth[pix_id];

  for (int i = 0; i < rounds; ++i, todo -= kTileThreads) {
    if (__syncthreads_count(done) == kTileThreads) break;
    const int progress = i * kTileThreads + tid;
    unsigned my_mask = 0u;
    if (progress < total) {
      const int id = (int)point_list[range.x + progress];
      s_id[tid] = id;
      const float4* r = rec + 3 * (size_t)id;
      const float4 q0 = __ldg(r + 0), q1 = __ldg(r + 1);
      s_rec[0][tid] = q0;
      s_rec[1][tid] = q1;
      s_rec[2][tid] = __ldg(r + 2);
      my_mask = block_mask16(q0, q1, tile_x0, tile_y0);
    }
    s_mask[tid] = (unsigned short)my_mask;
    __syncthreads();

    // each half warp keeps only the entries whose cut ellipse can touch its 4x4 pixel block;
    // the two halves then walk their own lists side by side
    const int nb = min(kTileThreads, todo);
    int cnt = 0;  // length of this half warp's list
    if (__any_sync(0xffffffffu, !done)) {
      const int sub_lo = sub - half, sub_hi = sub_lo + 1;
      int cnt_lo = 0, cnt_hi = 0;
      const unsigned lt = (1u << lane) - 1u;
      for (int c = 0; c * 32 < nb; ++c) {
        const int j = c * 32 + lane;
        const unsigned m = (j < nb) ? (unsigned)s_mask[j] : 0u;
        const bool hit_lo = (m >> sub_lo) & 1u, hit_hi = (m >> sub_hi) & 1u;
        const unsigned ball_lo = __ballot_sync(0xffffffffu, hit_lo);
        const unsigned ball_hi = __ballot_sync(0xffffffffu, hit_hi);
        if (hit_lo) s_list[2 * warp][cnt_lo + __popc(ball_lo & lt)] = (unsigned char)j;
        if (hit_hi) s_list[2 * warp + 1][cnt_hi + __popc(ball_hi & lt)] = (unsigned char)j;
        cnt_lo += __popc(ball_lo);
        cnt_hi += __popc(ball_hi);
      }
      __syncwarp();
      cnt = half ? cnt_hi : cnt_lo;
    }
    const unsigned char* my_list = s_list[2 * warp + half];
    if (done) cnt = 0;
    const uint32_t contrib_base = (uint32_t)(i * kTileThreads + 1);
    for (int k = 0; k < cnt; ++k) {
      const int j = my_list[k];
      const uint32_t contributor = contrib_base + (uint32_t)j;  // 1-based list position
      const float4 r0 = s_rec[0][j];
      const float4 r1 = s_rec[1][j];
      const float dx = GSR_SUB(r0.x, pixfx), dy = GSR_SUB(r0.y, pixfy);
      const float power = pair_power(r0.z, r0.w, r1.x, dx, dy);
      if (power > 0.0f) continue;
      if (power < r1.z) continue;  // cannot reach 15/255 (see preprocess_fwd: power_cut)
      const float alpha = pair_alpha(r1.y, expf(power));
      if (alpha < kAlphaMin) continue;

      if (VARIANT == kLight) {
        const float test_T = GSR_MUL(T, GSR_SUB(1.f, alpha));
        if (test_T < kTmin) {
          done = true;
          break;
        }
        const float4 r2 = s_rec[2][j];
        const float depth = r1.w;
        C0 = GSR_FMA(T, GSR_MUL(alpha, r2.x), C0);
        C1 = GSR_FMA(T, GSR_MUL(alpha, r2.y), C1);
        C2 = GSR_FMA(T, GSR_MUL(alpha, r2.z), C2);
        Wsum = GSR_FMA(T, alpha, Wsum);
        D = GSR_FMA(T, GSR_MUL(alpha, depth), D);
        if (T > 0.5f && test_T < 0.5f) {
          Dmed = depth;
          if (!LOSS) {  // per-Gaussian statistics are not produced by the fused-loss (tracking) pass
            const int id = s_id[j];
            const float dg = GSR_SUB(depth, gt);
            atomicAdd(gau_unc + id, GSR_MUL(T, GSR_MUL(alpha, GSR_MUL(dg, dg))));
            atomicAdd(gau_px + id, 1);
          }
        }
        T = test_T;
        last_contributor = contributor;
      } else {
        const float4 r2 = s_rec[2][j];
        const float depth = r1.w;
        C0 = GSR_FMA(T, GSR_MUL(alpha, r2.x), C0);
        C1 = GSR_FMA(T, GSR_MUL(alpha, r2.y), C1);
        C2 = GSR_FMA(T, GSR_MUL(alpha, r2.z), C2);
        D = GSR_FMA(T, GSR_MUL(alpha, depth), D);
        Wsum = GSR_FMA(T, alpha, Wsum);
        first = min(first, contributor);  // entries come in increasing position order
        ++valid;
        T = GSR_MUL(T, GSR_SUB(1.f, alpha));
        last_contributor = contributor;
        if (T < kTmin) {
          done = true;
          break;
        }
      }
    }
  }

  float my_loss = 0.f;
  if (inside) {
    const size_t HW = (size_t)H * (size_t)W;
    n_contrib[pix_id] = last_contributor;
    const float c0 = GSR_FMA(bg[0], T, C0), c1 = GSR_FMA(bg[1], T, C1), c2 = GSR_FMA(bg[2], T, C2);
    if (!LOSS || out_color != nullptr) {
      out_color[0 * HW + pix_id] = c0;
      out_color[1 * HW + pix_id] = c1;
      out_color[2 * HW + pix_id] = c2;
      out_depth[pix_id] = D;
    }
    out_aux0[pix_id] = Wsum;
    if (VARIANT == kLight) {
      if (!LOSS || out_median != nullptr) {
        out_median[pix_id] = Dmed;
        out_var[pix_id] = 0.0f;  // the reference never updates D_var (light forward.cu:317,410)
      }
    } else {
      final_T[pix_id] = T;
      first_contrib[pix_id] = (first == 0xFFFFFFFFu) ? 0u : first;
    }
    if (LOSS) {
      // fused masked-L1 loss and its cotangents (tracker.cu): the images need not leave the chip
      const float gtd = fl.gt_depth[pix_id];
      const bool m = (fl.depth_mask == 0 || gtd > 0.0f) && (Wsum > fl.alpha_thresh);
      const float wc = m ? fl.w_color : 0.f, wd = m ? fl.w_depth : 0.f;
      const float e0 = c0 - fl.gt_color[0 * HW + pix_id], e1 = c1 - fl.gt_color[1 * HW + pix_id],
                  e2 = c2 - fl.gt_color[2 * HW + pix_id], ed = D - gtd;
      fl.dL_dpix[0 * HW + pix_id] = e0 > 0.f ? wc : (e0 < 0.f ? -wc : 0.f);
      fl.dL_dpix[1 * HW + pix_id] = e1 > 0.f ? wc : (e1 < 0.f ? -wc : 0.f);
      fl.dL_dpix[2 * HW + pix_id] = e2 > 0.f ? wc : (e2 < 0.f ? -wc : 0.f);
      fl.dL_ddepth[pix_id] = ed > 0.f ? wd : (ed < 0.f ? -wd : 0.f);
      my_loss = wc * ((fabsf(e0) + fabsf(e1)) + fabsf(e2)) + wd * fabsf(ed);
    }
  }

  // tile-wide max of last_contributor (lets the backward skip the unused list tail) and,
  // for -full, the number of valid pairs (the reference's num_related_gaussians)
  uint32_t m = inside ? last_contributor : 0u;
  uint32_t v = (VARIANT == kFull && inside) ? valid : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    v += __shfl_xor_sync(0xffffffffu, v, o);
    if (LOSS) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
  }
  __syncthreads();
  if (lane == 0) s_red[warp] = m;
  if (LOSS && lane == 0) s_loss[warp] = my_loss;
  if (VARIANT == kFull && related_counter != nullptr && lane == 0 && v != 0)
    atomicAdd(related_counter, v);
  __syncthreads();
  if (tid == 0) {
    uint32_t mm = 0;
#pragma unroll
    for (int k = 0; k < kTileThreads / 32; ++k) mm = max(mm, s_red[k]);
    tile_last[tile] = mm;
    if (LOSS) {  // fixed summation order: the loss value is deterministic
      float l = 0.f;
#pragma unroll
      for (int k = 0; k < kTileThreads / 32; ++k) l += s_loss[k];
      fl.loss_partials[tile] = l;
    }
  }
}

}  // namespace

int launch_render_fwd_light(const Camera& cam, const GeomState& g, const BinState& b,
                            ImgState& img, const float* bg, const float* gt_depth,
                            float* out_color, float* out_depth, float* out_median, float* out_alpha,
                            float* out_var, float* gau_unc, int* gau_px, bool debug,
                            cudaStream_t stream) {
  dim3 grid(cam.grid_x, cam.grid_y, 1);
  StageScope st(ST_RENDER_FWD, stream);
  render_fwd_kernel<kLight, false><<<grid, kTileThreads, 0, stream>>>(
      img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, gt_depth, out_color, out_depth,
      out_alpha, out_median, out_var, gau_unc, gau_px, img.n_contrib, nullptr, nullptr,
      img.tile_last, nullptr, FusedLoss{});
  GSR_LAUNCH_OK(debug, stream);
  return GSR_OK;
}

// -light forward blend with the masked-L1 loss and its cotangents fused into the epilogue
// (tracker.cu).  Only alpha, n_contrib, tile_last, the cotangent images and the per-tile loss
// partials are written; out_color / out_depth / out_median / out_var may be NULL.
int launch_render_fwd_light_loss(const Camera& cam, const GeomState& g, const BinState& b,
                                 ImgState& img, const float* bg, float* out_color, float* out_depth,
                                 float* out_median, float* out_alpha, float* out_var,
                                 const FusedLoss& fl, cudaStream_t stream) {
  dim3 grid(cam.grid_x, cam.grid_y, 1);
  StageScope st(ST_RENDER_FWD, stream);
  render_fwd_kernel<kLight, true><<<grid, kTileThreads, 0, stream>>>(
      img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, fl.gt_depth, out_color, out_depth,
      out_alpha, out_median, out_var, nullptr, nullptr, img.n_contrib, nullptr, nullptr,
      img.tile_last, nullptr, fl);
  GSR_LAUNCH_OK(false, stream);
  return GSR_OK;
}

int launch_render_fwd_full(const Camera& cam, const GeomState& g, const BinState& b,
                           ImgState& img, const float* bg, float* out_color, float* out_depth,
                           float* out_unc, bool count_related, bool debug, cudaStream_t stream) {
  dim3 grid(cam.grid_x, cam.grid_y, 1);
  StageScope st(ST_RENDER_FWD, stream);
  render_fwd_kernel<kFull, false><<<grid, kTileThreads, 0, stream>>>(
      img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, nullptr, out_color, out_depth,
      out_unc, nullptr, nullptr, nullptr, nullptr, img.n_contrib, img.final_T, img.first_contrib,
      img.tile_last, count_related ? (g.counters + 1) : nullptr, FusedLoss{});
  GSR_LAUNCH_OK(debug, stream);
  return GSR_OK;
}

}  // namespace gsr
