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
#include "f32x2.cuh"

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

  pdl_wait();   // (may have been launched programmatically behind the per-tile sort)
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


// =================================================================================================
// Packed quarter-list forward (default): 128 threads per tile, every lane owns TWO vertically adjacent
// pixels and blends them with Blackwell's packed fp32x2 instructions; every quarter warp (8 lanes = a
// 4x4 pixel block) walks its own compacted entry list, the four quarters side by side (see
// render_bwdq_kernel for the statistics: 2.5 M list iterations per C3 frame instead of 4.8 M with
// half-warp lists of one pixel per lane).
// Bit parity with the scalar kernel above and with the reference build is kept by construction:
//  * the packed instructions round to nearest per element, so mul / fma sequences are the ones the
//    scalar kernel writes with __f*_rn intrinsics; exp is the same expf call;
//  * a pixel that an entry does not touch blends alpha = 0, which leaves T and every accumulator
//    exactly unchanged (x + (+-0) = x), so the per-pixel skip decisions need no branches;
//  * the rare per-pixel events — -light's "would fall below 1e-4: stop WITHOUT blending", the median
//    crossing with its two per-Gaussian atomics — are handled in divergent branches.
// =================================================================================================
constexpr int kFwdQThreads = 128;
constexpr int kFwdQWarps = kFwdQThreads / 32;
constexpr int kFwdQBatch = 128;

template <int VARIANT, bool LOSS, bool COUNT>
__global__ void __launch_bounds__(kFwdQThreads, 8)   // (7 CTAs per SM / 71 registers measured equal: 0.263 vs 0.265 ms)
render_fwdq_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
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
  __shared__ float4 s_rec[3][kFwdQBatch];  // one array: the three rows of an entry are a constant offset apart
  __shared__ int s_id[kFwdQBatch];
  __shared__ unsigned short s_mask[kFwdQBatch];
  __shared__ __align__(16) unsigned char s_list[kFwdQWarps][kFwdQBatch][4];  // [k][quarter], see render_bwdq_kernel
  __shared__ uint32_t s_red[kFwdQWarps];
  __shared__ float s_loss[kFwdQWarps];

  pdl_wait();   // (may have been launched programmatically behind the per-tile sort)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int quarter = lane >> 3, ql = lane & 7;
  const int tile = blockIdx.y * grid_x + blockIdx.x;
  reinterpret_cast<uint4*>(&s_list[0][0][0])[tid] = make_uint4(0u, 0u, 0u, 0u);
  const int px = blockIdx.x * kTileX + (warp & 1) * 8 + (quarter & 1) * 4 + (ql & 3);
  const int py0 = blockIdx.y * kTileY + (warp >> 1) * 8 + (quarter >> 1) * 4 + 2 * (ql >> 2);
  const int py1 = py0 + 1;
  const bool in_a = px < W && py0 < H, in_b = px < W && py1 < H;
  const uint32_t pix_a = (uint32_t)W * (uint32_t)py0 + (uint32_t)px;
  const uint32_t pix_b = pix_a + (uint32_t)W;
  const float pxf = (float)px;
  const f2 npy2 = f2_pack(-(float)py0, -(float)py1);
  const float tile_x0 = (float)(blockIdx.x * kTileX), tile_y0 = (float)(blockIdx.y * kTileY);
  const int sub0 = 4 * (2 * (warp >> 1)) + 2 * (warp & 1);  // bit of quarter 0 in block_mask16; +1, +4, +5

  const uint2 range = ranges[tile];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + kFwdQBatch - 1) / kFwdQBatch;

  // per-pixel "still blending" flags as integers (0 / 1): predicates that live across the loop end up
  // packed into byte lanes of a register otherwise
  int live_a = in_a ? 1 : 0, live_b = in_b ? 1 : 0;
  const f2 one2 = f2_pack(1.f, 1.f), mone2 = f2_pack(-1.f, -1.f), mhalf2 = f2_pack(-0.5f, -0.5f);
  f2 T2 = one2;
  f2 C0 = 0ull, C1 = 0ull, C2 = 0ull, D2 = 0ull, W2 = 0ull;
  float dmed_a = 0.f, dmed_b = 0.f;
  uint32_t last_a = 0, last_b = 0, first_a = 0xFFFFFFFFu, first_b = 0xFFFFFFFFu, valid = 0;
  float gt_a = 0.f, gt_b = 0.f;
  if (VARIANT == kLight && !LOSS) {
    if (in_a) gt_a = gt_depth[pix_a];
    if (in_b) gt_b = gt_depth[pix_b];
  }
  const unsigned list_w = pin_reg(smem_u32(&s_list[warp][0][0]));
  const unsigned qshift = pin_reg((unsigned)quarter * 8u);

  for (int i = 0; i < rounds; ++i) {
    if (__syncthreads_count((live_a | live_b) == 0) == kFwdQThreads) break;
    const int progress = i * kFwdQBatch + tid;
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

    const int nb = min(kFwdQBatch, total - i * kFwdQBatch);
    int cnt = 0, cnt_max = 0;
    if (__any_sync(0xffffffffu, (live_a | live_b) != 0)) {
      int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      const unsigned lt = (1u << lane) - 1u;
      for (int c = 0; c * 32 < nb; ++c) {
        const int jj = c * 32 + lane;
        const unsigned m = (jj < nb) ? ((unsigned)s_mask[jj] >> sub0) : 0u;
        const unsigned b0 = __ballot_sync(0xffffffffu, m & 1u), b1 = __ballot_sync(0xffffffffu, m & 2u);
        const unsigned b2 = __ballot_sync(0xffffffffu, m & 16u), b3 = __ballot_sync(0xffffffffu, m & 32u);
        if (m & 1u) s_list[warp][c0 + __popc(b0 & lt)][0] = (unsigned char)jj;
        if (m & 2u) s_list[warp][c1 + __popc(b1 & lt)][1] = (unsigned char)jj;
        if (m & 16u) s_list[warp][c2 + __popc(b2 & lt)][2] = (unsigned char)jj;
        if (m & 32u) s_list[warp][c3 + __popc(b3 & lt)][3] = (unsigned char)jj;
        c0 += __popc(b0); c1 += __popc(b1); c2 += __popc(b2); c3 += __popc(b3);
      }
      __syncwarp();
      cnt = quarter == 0 ? c0 : (quarter == 1 ? c1 : (quarter == 2 ? c2 : c3));
      cnt_max = max(max(c0, c1), max(c2, c3));
    }
    const uint32_t contrib_base = (uint32_t)(i * kFwdQBatch + 1);

    // unrolled by 4: the loop-carried accumulators are then renamed across the copies instead of being
    // copied back into their home registers at every back edge (12 moves per iteration otherwise)
#pragma unroll 4
    for (int k = 0; k < cnt_max; ++k) {
      const int j = (int)((lds_u32(list_w + 4u * (unsigned)k) >> qshift) & 0xFFu);
      const bool active = k < cnt;
      const float4 r0 = s_rec[0][j];
      const float4 r1 = s_rec[1][j];
      const float dx = GSR_SUB(r0.x, pxf);
      const f2 dx2 = f2_pack(dx, dx);
      const f2 dy2 = f2_add(f2_pack(r0.y, r0.y), npy2);
      // pair_power for both pixels: fma(fma(dx, dx*A, dy*(dy*C)), -0.5, -(dy*(dx*B)))
      const f2 qf = f2_fma(dx2, f2_mul(dx2, f2_pack(r0.z, r0.z)), f2_mul(dy2, f2_mul(dy2, f2_pack(r1.x, r1.x))));
      const f2 pw2 = f2_fma(qf, mhalf2, f2_mul(dy2, f2_mul(dx2, f2_pack(-r0.w, -r0.w))));
      const float pw_a = f2_lo(pw2), pw_b = f2_hi(pw2);
      bool va = active && live_a != 0 && !(pw_a > 0.0f) && !(pw_a < r1.z);
      bool vb = active && live_b != 0 && !(pw_b > 0.0f) && !(pw_b < r1.z);
      // No warp-level early-outs: with quarter lists only 7 % of the iterations have no blended pixel, and a
      // straight-line body lets every accumulator be updated in place (the skip edges cost a dozen register
      // copies per iteration).  A pixel the entry does not touch blends alpha = 0: exactly a no-op.
      float al_a = pair_alpha(r1.y, expf(pw_a)), al_b = pair_alpha(r1.y, expf(pw_b));
      va = va && !(al_a < kAlphaMin);
      vb = vb && !(al_b < kAlphaMin);
      if (!va) al_a = 0.f;
      if (!vb) al_b = 0.f;
      const uint32_t contributor = contrib_base + (uint32_t)j;  // 1-based list position
      f2 alpha2 = f2_pack(al_a, al_b);
      f2 om2 = f2_fma(alpha2, mone2, one2);   // 1 - alpha
      f2 tT2 = f2_mul(T2, om2);               // T * (1 - alpha)
      if (VARIANT == kLight) {
        // the Gaussian that would drive T below 1e-4 is NOT blended and the pixel stops: its alpha becomes
        // 0 and 1 - alpha, T (1 - alpha) are formed again (branch-free: the body stays one basic block)
        const bool ta = va && f2_lo(tT2) < kTmin, tb = vb && f2_hi(tT2) < kTmin;
        if (ta) { live_a = 0; al_a = 0.f; }
        if (tb) { live_b = 0; al_b = 0.f; }
        va = va && !ta;
        vb = vb && !tb;
        alpha2 = f2_pack(al_a, al_b);
        om2 = f2_fma(alpha2, mone2, one2);
        tT2 = f2_mul(T2, om2);
      }
      const float4 r2 = s_rec[2][j];
      const float depth = r1.w;
      f2_fma_acc(C0, T2, f2_mul(alpha2, f2_pack(r2.x, r2.x)));
      f2_fma_acc(C1, T2, f2_mul(alpha2, f2_pack(r2.y, r2.y)));
      f2_fma_acc(C2, T2, f2_mul(alpha2, f2_pack(r2.z, r2.z)));
      f2_fma_acc(W2, T2, alpha2);
      f2_fma_acc(D2, T2, f2_mul(alpha2, f2_pack(depth, depth)));
      if (VARIANT == kLight) {
        // median depth: the entry at which T crosses 0.5 (once per pixel), with -light's per-Gaussian statistics
        const bool ma = va && f2_lo(T2) > 0.5f && f2_lo(tT2) < 0.5f, mb = vb && f2_hi(T2) > 0.5f && f2_hi(tT2) < 0.5f;
        dmed_a = ma ? depth : dmed_a;
        dmed_b = mb ? depth : dmed_b;
        if (ma || mb) {
          if (!LOSS) {  // per-Gaussian statistics are not produced by the fused-loss (tracking) pass
            const int id = s_id[j];
            if (ma) {
              const float dg = GSR_SUB(depth, gt_a);
              atomicAdd(gau_unc + id, GSR_MUL(f2_lo(T2), GSR_MUL(al_a, GSR_MUL(dg, dg))));
              atomicAdd(gau_px + id, 1);
            }
            if (mb) {
              const float dg = GSR_SUB(depth, gt_b);
              atomicAdd(gau_unc + id, GSR_MUL(f2_hi(T2), GSR_MUL(al_b, GSR_MUL(dg, dg))));
              atomicAdd(gau_px + id, 1);
            }
          }
        }
        f2_mul_acc(T2, om2);   // = tT2, in place
        if (va) last_a = contributor;
        if (vb) last_b = contributor;
      } else {
        f2_mul_acc(T2, om2);
        if (va) { last_a = contributor; first_a = min(first_a, contributor); }
        if (vb) { last_b = contributor; first_b = min(first_b, contributor); }
        if (COUNT) valid += (va ? 1u : 0u) + (vb ? 1u : 0u);
        // the Gaussian that drives T below 1e-4 IS blended, then the pixel stops
        if (va && f2_lo(T2) < kTmin) live_a = 0;
        if (vb && f2_hi(T2) < kTmin) live_b = 0;
      }
    }
  }

  float my_loss = 0.f;
  const size_t HW = (size_t)H * (size_t)W;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const bool inside = h ? in_b : in_a;
    if (!inside) continue;
    const uint32_t pix_id = h ? pix_b : pix_a;
    const float T = h ? f2_hi(T2) : f2_lo(T2);
    const float c0a = h ? f2_hi(C0) : f2_lo(C0), c1a = h ? f2_hi(C1) : f2_lo(C1), c2a = h ? f2_hi(C2) : f2_lo(C2);
    const float D = h ? f2_hi(D2) : f2_lo(D2), Wsum = h ? f2_hi(W2) : f2_lo(W2);
    n_contrib[pix_id] = h ? last_b : last_a;
    const float c0 = GSR_FMA(bg[0], T, c0a), c1 = GSR_FMA(bg[1], T, c1a), c2 = GSR_FMA(bg[2], T, c2a);
    if (!LOSS || out_color != nullptr) {
      out_color[0 * HW + pix_id] = c0;
      out_color[1 * HW + pix_id] = c1;
      out_color[2 * HW + pix_id] = c2;
      out_depth[pix_id] = D;
    }
    out_aux0[pix_id] = Wsum;
    if (VARIANT == kLight) {
      if (!LOSS || out_median != nullptr) {
        out_median[pix_id] = h ? dmed_b : dmed_a;
        out_var[pix_id] = 0.0f;  // the reference never updates D_var (light forward.cu:317,410)
      }
    } else {
      final_T[pix_id] = T;
      const uint32_t first = h ? first_b : first_a;
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
      my_loss += wc * ((fabsf(e0) + fabsf(e1)) + fabsf(e2)) + wd * fabsf(ed);
    }
  }

  // tile-wide max of last_contributor (lets the backward skip the unused list tail) and, for -full,
  // the number of valid pairs (the reference's num_related_gaussians)
  uint32_t m = max(in_a ? last_a : 0u, in_b ? last_b : 0u);
  uint32_t v = (VARIANT == kFull) ? valid : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    v += __shfl_xor_sync(0xffffffffu, v, o);
    if (LOSS) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
  }
  __syncthreads();
  if (lane == 0) s_red[warp] = m;
  if (LOSS && lane == 0) s_loss[warp] = my_loss;
  if (VARIANT == kFull && COUNT && lane == 0 && v != 0)
    atomicAdd(related_counter, v);
  __syncthreads();
  if (tid == 0) {
    uint32_t mm = 0;
#pragma unroll
    for (int k = 0; k < kFwdQWarps; ++k) mm = max(mm, s_red[k]);
    tile_last[tile] = mm;
    if (LOSS) {  // fixed summation order: the loss value is deterministic
      float l = 0.f;
#pragma unroll
      for (int k = 0; k < kFwdQWarps; ++k) l += s_loss[k];
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
  // "fwd_packed": 1 = two pixels per lane + quarter-warp lists, 0 = one pixel per lane + half-warp lists,
  // 2 (default) = per variant: packed for -full (C3 0.263 vs 0.295 ms, C4 0.559 vs 0.581), scalar for -light
  // (C3 0.324 vs 0.323, C4 0.697 vs 0.629: the branch-free packed body pays for -light's two extra channels
  // and its stop-before-blending rule on every entry)
  if (options().fwd_packed == 1)
    render_fwdq_kernel<kLight, false, false><<<grid, kFwdQThreads, 0, stream>>>(
        img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, gt_depth, out_color, out_depth,
        out_alpha, out_median, out_var, gau_unc, gau_px, img.n_contrib, nullptr, nullptr,
        img.tile_last, nullptr, FusedLoss{});
  else
  launch_after(options().pdl != 0, render_fwd_kernel<kLight, false>, grid, dim3(kTileThreads), 0, stream,
      (const uint2*)img.ranges, (const uint32_t*)b.vals, cam.W, cam.H, cam.grid_x, (const float4*)g.rec, bg, gt_depth, out_color, out_depth,
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
  if (options().fwd_packed == 1)
    render_fwdq_kernel<kLight, true, false><<<grid, kFwdQThreads, 0, stream>>>(
        img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, fl.gt_depth, out_color, out_depth,
        out_alpha, out_median, out_var, nullptr, nullptr, img.n_contrib, nullptr, nullptr,
        img.tile_last, nullptr, fl);
  else
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
  if (options().fwd_packed != 0 && count_related)
    render_fwdq_kernel<kFull, false, true><<<grid, kFwdQThreads, 0, stream>>>(
        img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, nullptr, out_color, out_depth,
        out_unc, nullptr, nullptr, nullptr, nullptr, img.n_contrib, img.final_T, img.first_contrib,
        img.tile_last, g.counters + 1, FusedLoss{});
  else if (options().fwd_packed != 0)
    launch_after(options().pdl != 0, render_fwdq_kernel<kFull, false, false>, grid, dim3(kFwdQThreads), 0, stream,
        (const uint2*)img.ranges, (const uint32_t*)b.vals, cam.W, cam.H, cam.grid_x, (const float4*)g.rec, bg, nullptr, out_color, out_depth,
        out_unc, nullptr, nullptr, nullptr, nullptr, img.n_contrib, img.final_T, img.first_contrib,
        img.tile_last, nullptr, FusedLoss{});
  else
  render_fwd_kernel<kFull, false><<<grid, kTileThreads, 0, stream>>>(
      img.ranges, b.vals, cam.W, cam.H, cam.grid_x, g.rec, bg, nullptr, out_color, out_depth,
      out_unc, nullptr, nullptr, nullptr, nullptr, img.n_contrib, img.final_T, img.first_contrib,
      img.tile_last, count_related ? (g.counters + 1) : nullptr, FusedLoss{});
  GSR_LAUNCH_OK(debug, stream);
  return GSR_OK;
}

}  // namespace gsr
