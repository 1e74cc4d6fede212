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
#include "f32x2.cuh"

namespace gsr {

namespace {

constexpr int kRedVals = 14;    // accumulator slots reduced per (warp, Gaussian) hit
// Reduction scratch layout: s_red[slot][warp * 32 + lane] with rows of (32 * warps + 4) floats — the
// store address is a constant plus 4 * threadIdx.x (cheap to rematerialise, conflict-free), and the
// row stride is 16 bytes modulo 128, so the 4 x LDS.128 of the reducing lanes (lane -> slot lane >> 1,
// half lane & 1) are conflict-free per quarter warp as well.
__host__ __device__ constexpr int red_row(int warps) { return 32 * warps + 4; }

// Reduced slot sets: all 14 accumulator slots, or — for -light's tracking mode (map_off: only the
// pose gradient is wanted) — just the three the pose contraction reads.
template <int VARIANT, bool POSE_ONLY>
struct RedSet {
  // -full never writes ACC_MED (the last slot), so it reduces one slot less
  static constexpr int N = POSE_ONLY ? 3 : (VARIANT == kFull ? kRedVals - 1 : kRedVals);
  __device__ static constexpr int slot(int q) {
    return POSE_ONLY ? (q == 0 ? (int)ACC_MX : (q == 1 ? (int)ACC_MY : (int)ACC_PD)) : q;
  }
};

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// sum of 16 consecutive floats in shared memory (4 x LDS.128, a packed add tree, one scalar add)
__device__ __forceinline__ float row16_sum(unsigned addr) {
  f2 a0, a1, b0, b1, c0, c1, d0, d1;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "r"(addr) : "memory");
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(addr + 16) : "memory");
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(c0), "=l"(c1) : "r"(addr + 32) : "memory");
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d0), "=l"(d1) : "r"(addr + 48) : "memory");
  const f2 t = f2_add(f2_add(f2_add(a0, a1), f2_add(b0, b1)), f2_add(f2_add(c0, c1), f2_add(d0, d1)));
  return f2_lo(t) + f2_hi(t);
}

template <int VARIANT, bool POSE_ONLY>
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
  __shared__ float4 s_rec[3][kTileThreads];
  __shared__ int s_id[kTileThreads];
  __shared__ unsigned short s_mask[kTileThreads];
  __shared__ unsigned char s_list[kTileThreads / 16][kTileThreads];
  __shared__ __align__(16) float s_red[kRedVals][red_row(kTileThreads / 32)];

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
    dLv = dL_dvars != nullptr ? dL_dvars[pix_id] : 0.f;  // NULL = zero cotangent (tracker)
    gt = gt_depth[pix_id];
    if (VARIANT == kLight && dL_dmedians != nullptr) dLm = dL_dmedians[pix_id];
    if (VARIANT == kFull) first = (int)first_contrib[pix_id];
  }
  float T = T_final;
  float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, adr = 0.f, avr = 0.f;
  float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_depth = 0.f, last_var = 0.f;
  const float bg_dot_dpixel = bg[0] * dLp0 + bg[1] * dLp1 + bg[2] * dLp2;
  const float ddelx_dx = 0.5f * W;
  const float ddely_dy = 0.5f * H;
  bool mid_once = true;
  const unsigned red_st = smem_u32(&s_red[0][tid]);
  const unsigned red_ld = smem_u32(&s_red[(lane >> 1) % kRedVals][(tid & ~31) + (lane & 1) * 16]);
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
      s_rec[0][tid] = q0;
      s_rec[1][tid] = q1;
      s_rec[2][tid] = __ldg(r + 2);
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
      const float4 r0 = s_rec[0][j];
      const float4 r1 = s_rec[1][j];
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
        const float4 r2 = s_rec[2][j];
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
      using RS = RedSet<VARIANT, POSE_ONLY>;
#pragma unroll
      for (int q = 0; q < RS::N; ++q) sts_f32(red_st + q * (red_row(kTileThreads / 32) * 4), v[RS::slot(q)]);
      __syncwarp();
      float sum = 0.f;
      if (lane < 2 * RS::N) {
        sum = row16_sum(red_ld);
      }
      // lane (2q + h) now holds slot q summed over half warp h, i.e. over that half's own entry
      const int j_lo = __shfl_sync(0xffffffffu, j, 0), j_hi = __shfl_sync(0xffffffffu, j, 16);
      if (lane < 2 * RS::N) {
        const int h = lane & 1;
        if ((vmask >> (16 * h)) & 0xFFFFu)
          atomicAdd(acc + (size_t)s_id[h ? j_hi : j_lo] * kAccStride + RS::slot(lane >> 1), sum);
      }
      __syncwarp();
    }
  }
}


// =================================================================================================
// Packed variant (default): 128 threads per tile, every lane owns TWO vertically adjacent pixels and
// does their arithmetic with Blackwell's packed fp32x2 instructions (FADD2 / FMUL2 / FFMA2: one
// issue slot for two lanes of work — the blend kernels are issue-bound, not FMA-pipe bound).
// A warp covers an 8x8 pixel block, so an entry is visited by fewer warps than with 8x4 blocks,
// and the per-entry fixed costs (staging reads, cull vote, reduction, atomics) are paid per 64
// pixels instead of per 32.
//  * per-entry data is staged in shared memory already duplicated into (v, v) pairs, so that the
//    packed operands come straight out of LDS.128;
//  * the per-pixel state uses the compositing recurrence  B <- B + alpha (c - B)  ("colour behind
//    the current entry"), which is branch-free for pixels the entry does not touch (alpha = G = 0
//    leaves T and B unchanged and makes every partial sum exactly zero);
//  * power / alpha are evaluated with the same rounding sequence as the forward (packed ops are
//    IEEE round-to-nearest per element), so both passes take identical skip decisions.
// =================================================================================================
constexpr int kBwd2Threads = 128;
constexpr int kBwd2Warps = kBwd2Threads / 32;
constexpr int kBwd2Batch = 128;  // entries staged per round (one per thread)

// 4-bit mask of the four 8x8 pixel blocks of a tile (bit = 2 * block row + block column)
__device__ __forceinline__ unsigned block_mask4(const float4& r0, const float4& r1, float tx0, float ty0) {
  float hx, hy;
  const int kind = cut_extent(r0.z, r0.w, r1.x, r1.z, hx, hy);
  if (kind == 0) return 0u;
  if (kind == 2) return 0xFu;
  const float x0 = r0.x - hx, x1 = r0.x + hx, y0 = r0.y - hy, y1 = r0.y + hy;
  const unsigned col = ((x1 >= tx0 && x0 <= tx0 + 7.0f) ? 1u : 0u) |
                       ((x1 >= tx0 + 8.0f && x0 <= tx0 + 15.0f) ? 2u : 0u);
  unsigned mask = 0u;
  if (y1 >= ty0 && y0 <= ty0 + 7.0f) mask |= col;
  if (y1 >= ty0 + 8.0f && y0 <= ty0 + 15.0f) mask |= col << 2;
  return mask;
}

template <int VARIANT, bool POSE_ONLY>
__global__ void __launch_bounds__(kBwd2Threads, 8)
render_bwd2_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
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
  // staged entry, duplicated into pairs: q0 = (xg, pc | yg, yg)  q1 = (A, A | -B, -B)
  //   q2 = (C, C | o, o)  q3 = (depth, depth | r, r)  q4 = (g, g | b, b)
  __shared__ ulonglong2 s_q[5][kBwd2Batch];  // one array: an entry's five vectors are constant offsets apart
  __shared__ int s_id[kBwd2Batch];
  __shared__ unsigned char s_mask[kBwd2Batch];
  __shared__ unsigned char s_list[kBwd2Warps][kBwd2Batch];
  __shared__ __align__(16) float s_red[kRedVals][red_row(kBwd2Warps)];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.y * grid_x + blockIdx.x;
  const int px = blockIdx.x * kTileX + (warp & 1) * 8 + (lane & 7);
  const int py0 = blockIdx.y * kTileY + (warp >> 1) * 8 + 2 * (lane >> 3);
  const int py1 = py0 + 1;
  const bool in_a = px < W && py0 < H, in_b = px < W && py1 < H;
  const uint32_t pix_a = (uint32_t)W * (uint32_t)py0 + (uint32_t)px;
  const uint32_t pix_b = pix_a + (uint32_t)W;
  const float pxf = (float)px;
  const f2 npy2 = f2_pack(-(float)py0, -(float)py1);
  const float tile_x0 = (float)(blockIdx.x * kTileX), tile_y0 = (float)(blockIdx.y * kTileY);

  const uint2 range = ranges[tile];
  const int walk = (int)tile_last[tile];  // entries [0, walk) were used by some pixel
  const int rounds = (walk + kBwd2Batch - 1) / kBwd2Batch;
  const size_t HW = (size_t)H * (size_t)W;

  // per-pixel constants and state, packed (a, b)
  float Tf_a = 0.f, Tf_b = 0.f, g0a = 0.f, g0b = 0.f, g1a = 0.f, g1b = 0.f, g2a = 0.f, g2b = 0.f;
  float gda = 0.f, gdb = 0.f, gva = 0.f, gvb = 0.f, gta = 0.f, gtb = 0.f, gma = 0.f, gmb = 0.f;
  int lc_a = 0, lc_b = 0, first_a = 0, first_b = 0;
  if (in_a) {
    Tf_a = (VARIANT == kLight) ? (1 - alphas[pix_a]) : final_Ts[pix_a];
    lc_a = (int)n_contrib[pix_a];
    g0a = dL_dpix[pix_a]; g1a = dL_dpix[HW + pix_a]; g2a = dL_dpix[2 * HW + pix_a];
    gda = dL_ddepths[pix_a]; gva = dL_dvars != nullptr ? dL_dvars[pix_a] : 0.f; gta = gt_depth[pix_a];
    if (VARIANT == kLight && dL_dmedians != nullptr) gma = dL_dmedians[pix_a];
    if (VARIANT == kFull) first_a = (int)first_contrib[pix_a];
  }
  if (in_b) {
    Tf_b = (VARIANT == kLight) ? (1 - alphas[pix_b]) : final_Ts[pix_b];
    lc_b = (int)n_contrib[pix_b];
    g0b = dL_dpix[pix_b]; g1b = dL_dpix[HW + pix_b]; g2b = dL_dpix[2 * HW + pix_b];
    gdb = dL_ddepths[pix_b]; gvb = dL_dvars != nullptr ? dL_dvars[pix_b] : 0.f; gtb = gt_depth[pix_b];
    if (VARIANT == kLight && dL_dmedians != nullptr) gmb = dL_dmedians[pix_b];
    if (VARIANT == kFull) first_b = (int)first_contrib[pix_b];
  }
  const f2 Tf2 = f2_pack(Tf_a, Tf_b);
  const f2 dLp0 = f2_pack(g0a, g0b), dLp1 = f2_pack(g1a, g1b), dLp2 = f2_pack(g2a, g2b);
  const f2 dLd = f2_pack(gda, gdb), dLv = f2_pack(gva, gvb), dLv_x2 = f2_pack(2.f * gva, 2.f * gvb);
  const f2 ngt2 = f2_pack(-gta, -gtb);
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  const f2 nbgdot = f2_pack(-(bg0 * g0a + bg1 * g1a + bg2 * g2a), -(bg0 * g0b + bg1 * g1b + bg2 * g2b));
  const f2 one2 = f2_pack(1.f, 1.f), mone2 = f2_pack(-1.f, -1.f), mhalf2 = f2_pack(-0.5f, -0.5f);
  f2 T2 = Tf2;
  f2 Bc0 = 0ull, Bc1 = 0ull, Bc2 = 0ull, Bd = 0ull, Bv = 0ull;  // colour / depth / var "behind" the entry
  bool mid_a = true, mid_b = true;
  const unsigned red_st = smem_u32(&s_red[0][tid]);
  const unsigned red_ld = smem_u32(&s_red[(lane >> 1) % kRedVals][(tid & ~31) + (lane & 1) * 16]);

  for (int i = 0; i < rounds; ++i) {
    __syncthreads();
    const int progress = i * kBwd2Batch + tid;
    unsigned my_mask = 0u;
    float4 q0, q1, q2;
    int id = -1;
    if (progress < walk) {
      id = (int)point_list[range.x + (walk - progress - 1)];
      const float4* r = rec + 3 * (size_t)id;
      q0 = __ldg(r + 0); q1 = __ldg(r + 1); q2 = __ldg(r + 2);
    }
    if (id >= 0) {
      s_id[tid] = id;
      my_mask = block_mask4(q0, q1, tile_x0, tile_y0);
      s_q[0][tid] = make_ulonglong2(f2_pack(q0.x, q1.z), f2_pack(q0.y, q0.y));
      s_q[1][tid] = make_ulonglong2(f2_pack(q0.z, q0.z), f2_pack(-q0.w, -q0.w));
      s_q[2][tid] = make_ulonglong2(f2_pack(q1.x, q1.x), f2_pack(q1.y, q1.y));
      s_q[3][tid] = make_ulonglong2(f2_pack(q1.w, q1.w), f2_pack(q2.x, q2.x));
      s_q[4][tid] = make_ulonglong2(f2_pack(q2.y, q2.y), f2_pack(q2.z, q2.z));
    }
    s_mask[tid] = (unsigned char)my_mask;
    __syncthreads();

    // per-warp compaction: only entries whose cut ellipse can touch this warp's 8x8 block
    const int nb = min(kBwd2Batch, walk - i * kBwd2Batch);
    int cnt = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int c = 0; c * 32 < nb; ++c) {
      const int jj = c * 32 + lane;
      const bool hit = (jj < nb) && ((s_mask[jj] >> warp) & 1u);
      const unsigned ball = __ballot_sync(0xffffffffu, hit);
      if (hit) s_list[warp][cnt + __popc(ball & lt)] = (unsigned char)jj;
      cnt += __popc(ball);
    }
    __syncwarp();

    for (int k = 0; k < cnt; ++k) {
      const int j = s_list[warp][k];
      const int pos = walk - (i * kBwd2Batch + j) - 1;  // 0-based list position
      const ulonglong2* eq = &s_q[0][j];
      const ulonglong2 e0 = eq[0], e1 = eq[kBwd2Batch], e2 = eq[2 * kBwd2Batch];
      const float dx = GSR_SUB(f2_lo(e0.x), pxf);
      const float pc = f2_hi(e0.x);
      const f2 dx2 = f2_pack(dx, dx);
      const f2 dy2 = f2_add(e0.y, npy2);
      // pair_power for both pixels, same rounding sequence as the scalar version:
      //   fma(fma(dx, dx*A, dy*(dy*C)), -0.5, -(dy*(dx*B)))
      const f2 qf = f2_fma(dx2, f2_mul(dx2, e1.x), f2_mul(dy2, f2_mul(dy2, e2.x)));
      const f2 pw2 = f2_fma(qf, mhalf2, f2_mul(dy2, f2_mul(dx2, e1.y)));
      const float pw_a = f2_lo(pw2), pw_b = f2_hi(pw2);
      bool va = (pos < lc_a) && !(pw_a > 0.0f) && !(pw_a < pc);
      bool vb = (pos < lc_b) && !(pw_b > 0.0f) && !(pw_b < pc);
      if (!__any_sync(0xffffffffu, va || vb)) continue;
      const float o = f2_lo(e2.y);
      // both exponentials unconditionally (no divergence; a warp that gets here almost always needs
      // them), results masked afterwards
      float Ga = expf(pw_a), Gb = expf(pw_b);
      float al_a = pair_alpha(o, Ga), al_b = pair_alpha(o, Gb);
      va = va && !(al_a < kAlphaMin);
      vb = vb && !(al_b < kAlphaMin);
      if (!__any_sync(0xffffffffu, va || vb)) continue;
      if (!va) { Ga = 0.f; al_a = 0.f; }
      if (!vb) { Gb = 0.f; al_b = 0.f; }
      const f2 G2 = f2_pack(Ga, Gb), alpha2 = f2_pack(al_a, al_b);

      const ulonglong2 e3 = eq[3 * kBwd2Batch], e4 = eq[4 * kBwd2Batch];
      const f2 om2 = f2_fma(alpha2, mone2, one2);  // 1 - alpha (>= 0.01)
      const f2 inv2 = f2_pack(fast_rcp(f2_lo(om2)), fast_rcp(f2_hi(om2)));
      T2 = f2_mul(T2, inv2);
      const f2 aT2 = f2_mul(alpha2, T2);
      // c - B for colour, depth and var
      const f2 d0 = f2_fma(Bc0, mone2, e3.y), d1 = f2_fma(Bc1, mone2, e4.x), d2 = f2_fma(Bc2, mone2, e4.y);
      const f2 dgt2 = f2_add(e3.x, ngt2);
      const f2 cvar2 = f2_mul(dgt2, dgt2);
      const f2 dd = f2_fma(Bd, mone2, e3.x), dv = f2_fma(Bv, mone2, cvar2);
      const f2 colour_part = f2_fma(d2, dLp2, f2_fma(d1, dLp1, f2_mul(d0, dLp0)));
      const f2 depth_part = f2_mul(dd, dLd);
      f2 dLa = f2_fma(dv, dLv, f2_add(colour_part, depth_part));
      // B <- B + alpha (c - B)
      Bc0 = f2_fma(alpha2, d0, Bc0);
      Bc1 = f2_fma(alpha2, d1, Bc1);
      Bc2 = f2_fma(alpha2, d2, Bc2);
      Bd = f2_fma(alpha2, dd, Bd);
      Bv = f2_fma(alpha2, dv, Bv);
      const f2 aTd = f2_mul(aT2, dLd);
      // dL/dalpha = T * (...) - T_final / (1 - alpha) * (bg . dL/dpixel)
      dLa = f2_fma(f2_mul(Tf2, inv2), nbgdot, f2_mul(dLa, T2));
      const f2 w2 = f2_mul(G2, dLa);
      const f2 wx = f2_mul(w2, dx2), wy = f2_mul(w2, dy2);

      f2 v[kRedVals];
      v[ACC_MX] = wx;
      v[ACC_MY] = wy;
      v[ACC_CA] = f2_mul(wx, dx2);
      v[ACC_CB] = f2_mul(wx, dy2);
      v[ACC_CC] = f2_mul(wy, dy2);
      v[ACC_OP] = w2;
      v[ACC_R] = f2_mul(aT2, dLp0);
      v[ACC_G] = f2_mul(aT2, dLp1);
      v[ACC_B] = f2_mul(aT2, dLp2);
      v[ACC_DEPTH] = f2_fma(f2_mul(aT2, dgt2), dLv_x2, aTd);
      v[ACC_PGX] = 0ull;
      v[ACC_PGY] = 0ull;
      v[ACC_PD] = 0ull;
      v[ACC_MED] = 0ull;
      if (VARIANT == kLight) {
        v[ACC_PD] = aTd;
        // median: the first valid entry met from the back whose restored T exceeds 0.5
        float med_a = 0.f, med_b = 0.f;
        if (va && mid_a && f2_lo(T2) > 0.5f) { med_a = gma; mid_a = false; }
        if (vb && mid_b && f2_hi(T2) > 0.5f) { med_b = gmb; mid_b = false; }
        v[ACC_MED] = f2_pack(med_a, med_b);
      } else {
        // pose terms of the reference's ComputePG: colour through ndc without the background
        // term (full backward.cu:746-777, :1028-1072); depth only from the front-most valid
        // contributor of the pixel, because dd_dv* is assigned, not accumulated (:1278-1289).
        const bool fa = va && (pos + 1 == first_a), fb = vb && (pos + 1 == first_b);
        const f2 fsel = f2_pack(fa ? 1.f : 0.f, fb ? 1.f : 0.f);
        const f2 pa = f2_mul(T2, f2_fma(fsel, depth_part, colour_part));
        v[ACC_PD] = f2_mul(fsel, aTd);
        const f2 q = f2_mul(pa, G2);
        v[ACC_PGX] = f2_mul(q, dx2);
        v[ACC_PGY] = f2_mul(q, dy2);
      }
      // warp reduction through shared memory (both pixels of a lane are added first)
      using RS = RedSet<VARIANT, POSE_ONLY>;
#pragma unroll
      for (int qn = 0; qn < RS::N; ++qn)
        sts_f32(red_st + qn * (red_row(kBwd2Warps) * 4), f2_lo(v[RS::slot(qn)]) + f2_hi(v[RS::slot(qn)]));
      __syncwarp();
      float sum = 0.f;
      if (lane < 2 * RS::N) {
        sum = row16_sum(red_ld);
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (lane < 2 * RS::N && (lane & 1) == 0)
        atomicAdd(acc + (size_t)s_id[j] * kAccStride + RS::slot(lane >> 1), sum);
      __syncwarp();
    }
  }
}


// =================================================================================================
// Quarter-list variant of the packed kernel (default): same pixel arithmetic (two vertically adjacent
// pixels per lane, fp32x2), but every QUARTER warp — 8 lanes = a 4x4 pixel block — walks its OWN
// compacted entry list, the four quarters side by side.  An 8x8 block is hit by 3.2 M (block, entry)
// pairs per C3 frame with only 31 % of its 64 pixel slots blended; its four 4x4 blocks have 55 % of
// their slots blended, and walking their lists in lock step needs 2.3 M full-cost iterations instead
// of 3.2 M (C4: 4.0 M instead of 7.0 M; tools/blend_stats.py).  The shared-memory reduction then
// produces four 8-lane sums per slot (one per quarter / Gaussian) instead of one 32-lane sum:
// 13 x 4 (slot, quarter) sums are spread over the 32 lanes in two passes of 2 x LDS.128 each, and
// each pass ends in one red.add instruction whose lanes address up to four accumulator lines.
// Further per-iteration savings against render_bwd2_kernel: exp(power) is ex2.approx(power * log2 e)
// (the backward's tolerance is 1e-3; the forward keeps expf and decides what is blended through
// n_contrib), list positions come from one subtraction, the first-contributor select is one
// compare-and-select per pixel.
// =================================================================================================
constexpr int kBwdQThreads = 128;
constexpr int kBwdQWarps = kBwdQThreads / 32;
constexpr int kBwdQBatch = 128;  // entries staged per round (one per thread)

__device__ __forceinline__ float fast_exp(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}
// sum of 8 consecutive floats in shared memory (2 x LDS.128)
__device__ __forceinline__ float row8_sum(unsigned addr) {
  f2 a0, a1, b0, b1;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "r"(addr) : "memory");
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(b0), "=l"(b1) : "r"(addr + 16) : "memory");
  const f2 t = f2_add(f2_add(a0, a1), f2_add(b0, b1));
  return f2_lo(t) + f2_hi(t);
}

// EXACT (option "exact_median", -light only): exp(power) with expf and T restored with an IEEE division, i.e. the
// reference's own arithmetic for the chain that decides which entry receives the median-depth gradient
// (DESIGN.md section 2): the 2-in-3000 random cases where the approximate chain sends that gradient to the
// neighbouring entry disappear, at ~13 % more instructions per iteration.
template <int VARIANT, bool POSE_ONLY, int MINB, bool EXACT = false>
__global__ void __launch_bounds__(kBwdQThreads, MINB)
render_bwdq_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
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
  // staged entry, duplicated into pairs: q0 = (xg, pc | yg, yg)  q1 = (A, A | -B, -B)
  //   q2 = (C, C | o, ps)  q3 = (depth, depth | r, r)  q4 = (g, g | b, b)     (pc / ps = power_cut / power_sure)
  __shared__ ulonglong2 s_q[5][kBwdQBatch];
  __shared__ int s_id[kBwdQBatch];
  __shared__ unsigned short s_mask[kBwdQBatch];
  // per-warp entry lists of the four quarters, interleaved: byte [k][q] = k-th entry of quarter q, so
  // that ONE broadcast LDS.32 per iteration fetches the four quarters' entries
  __shared__ __align__(16) unsigned char s_list[kBwdQWarps][kBwdQBatch][4];
  __shared__ __align__(16) float s_red[kRedVals][red_row(kBwdQWarps)];

  pdl_trigger();   // the per-Gaussian backward kernel behind this one may be set up while it runs
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int quarter = lane >> 3, ql = lane & 7;
  // lists start zeroed: slots beyond a quarter's length are read (and ignored) by its lanes
  reinterpret_cast<uint4*>(&s_list[0][0][0])[tid] = make_uint4(0u, 0u, 0u, 0u);
  const int tile = blockIdx.y * grid_x + blockIdx.x;
  // warp -> 8x8 block (column warp & 1, row warp >> 1); quarter -> 4x4 block inside it; lane -> a 1x2 column
  const int px = blockIdx.x * kTileX + (warp & 1) * 8 + (quarter & 1) * 4 + (ql & 3);
  const int py0 = blockIdx.y * kTileY + (warp >> 1) * 8 + (quarter >> 1) * 4 + 2 * (ql >> 2);
  const int py1 = py0 + 1;
  const bool in_a = px < W && py0 < H, in_b = px < W && py1 < H;
  const uint32_t pix_a = (uint32_t)W * (uint32_t)py0 + (uint32_t)px;
  const uint32_t pix_b = pix_a + (uint32_t)W;
  const float pxf = (float)px;
  const f2 npy2 = f2_pack(-(float)py0, -(float)py1);
  const float tile_x0 = (float)(blockIdx.x * kTileX), tile_y0 = (float)(blockIdx.y * kTileY);
  // bit of this warp's quarter 0 in block_mask16 (bit = 4 * block row + block column of the 4x4 blocks);
  // quarters 1, 2, 3 are bits +1, +4, +5
  const int sub0 = 4 * (2 * (warp >> 1)) + 2 * (warp & 1);

  const uint2 range = ranges[tile];
  const int walk = (int)tile_last[tile];  // entries [0, walk) were used by some pixel
  const int rounds = (walk + kBwdQBatch - 1) / kBwdQBatch;
  const size_t HW = (size_t)H * (size_t)W;

  float Tf_a = 0.f, Tf_b = 0.f, g0a = 0.f, g0b = 0.f, g1a = 0.f, g1b = 0.f, g2a = 0.f, g2b = 0.f;
  float gda = 0.f, gdb = 0.f, gva = 0.f, gvb = 0.f, gta = 0.f, gtb = 0.f, gma = 0.f, gmb = 0.f;
  int lc_a = 0, lc_b = 0, fm1_a = -1, fm1_b = -1;   // fm1: 0-based position of the front-most contributor
  if (in_a) {
    Tf_a = (VARIANT == kLight) ? (1 - alphas[pix_a]) : final_Ts[pix_a];
    lc_a = (int)n_contrib[pix_a];
    g0a = dL_dpix[pix_a]; g1a = dL_dpix[HW + pix_a]; g2a = dL_dpix[2 * HW + pix_a];
    gda = dL_ddepths[pix_a]; gva = dL_dvars != nullptr ? dL_dvars[pix_a] : 0.f; gta = gt_depth[pix_a];
    if (VARIANT == kLight && dL_dmedians != nullptr) gma = dL_dmedians[pix_a];
    if (VARIANT == kFull) fm1_a = (int)first_contrib[pix_a] - 1;
  }
  if (in_b) {
    Tf_b = (VARIANT == kLight) ? (1 - alphas[pix_b]) : final_Ts[pix_b];
    lc_b = (int)n_contrib[pix_b];
    g0b = dL_dpix[pix_b]; g1b = dL_dpix[HW + pix_b]; g2b = dL_dpix[2 * HW + pix_b];
    gdb = dL_ddepths[pix_b]; gvb = dL_dvars != nullptr ? dL_dvars[pix_b] : 0.f; gtb = gt_depth[pix_b];
    if (VARIANT == kLight && dL_dmedians != nullptr) gmb = dL_dmedians[pix_b];
    if (VARIANT == kFull) fm1_b = (int)first_contrib[pix_b] - 1;
  }
  const f2 dLp0 = f2_pack(g0a, g0b), dLp1 = f2_pack(g1a, g1b), dLp2 = f2_pack(g2a, g2b);
  const f2 dLd = f2_pack(gda, gdb), dLv = f2_pack(gva, gvb), dLv_x2 = f2_pack(2.f * gva, 2.f * gvb);
  const f2 ngt2 = f2_pack(-gta, -gtb);
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  // -T_final * (bg . dL/dpixel): the background term of dL/dalpha is this times 1 / (1 - alpha)
  const f2 tfbg2 = f2_pack(-Tf_a * (bg0 * g0a + bg1 * g1a + bg2 * g2a), -Tf_b * (bg0 * g0b + bg1 * g1b + bg2 * g2b));
  const f2 one2 = f2_pack(1.f, 1.f), mone2 = f2_pack(-1.f, -1.f), mhalf2 = f2_pack(-0.5f, -0.5f);
  f2 T2 = f2_pack(Tf_a, Tf_b);
  f2 Bc0 = 0ull, Bc1 = 0ull, Bc2 = 0ull, Bd = 0ull, Bv = 0ull;  // colour / depth / var "behind" the entry
  bool mid_a = true, mid_b = true;
  using RS = RedSet<VARIANT, POSE_ONLY>;
  constexpr unsigned kRowBytes = red_row(kBwdQWarps) * 4;
  const unsigned red_st = pin_reg(smem_u32(&s_red[0][tid]));
  // reducing lane: slot (lane & 7) [+ 8 in the second pass] of its OWN quarter: conflict-free (row stride
  // 16 bytes mod 128), and the lane already holds that quarter's entry index
  const unsigned red_ld = pin_reg(smem_u32(&s_red[ql][(tid & ~31) + quarter * 8]));
  const unsigned list_w = pin_reg(smem_u32(&s_list[warp][0][0]));
  const unsigned qshift = pin_reg((unsigned)quarter * 8u);
  const unsigned ql4 = (unsigned)ql * 4u;            // byte offset of this lane's slot in an accumulator line

  for (int i = 0; i < rounds; ++i) {
    __syncthreads();
    const int progress = i * kBwdQBatch + tid;
    unsigned my_mask = 0u;
    if (progress < walk) {
      const int id = (int)point_list[range.x + (walk - progress - 1)];
      const float4* r = rec + 3 * (size_t)id;
      const float4 q0 = __ldg(r + 0), q1 = __ldg(r + 1), q2 = __ldg(r + 2);
      s_id[tid] = id;
      my_mask = block_mask16(q0, q1, tile_x0, tile_y0);
      s_q[0][tid] = make_ulonglong2(f2_pack(q0.x, q1.z), f2_pack(q0.y, q0.y));
      s_q[1][tid] = make_ulonglong2(f2_pack(q0.z, q0.z), f2_pack(-q0.w, -q0.w));
      s_q[2][tid] = make_ulonglong2(f2_pack(q1.x, q1.x), f2_pack(q1.y, q2.w));
      s_q[3][tid] = make_ulonglong2(f2_pack(q1.w, q1.w), f2_pack(q2.x, q2.x));
      s_q[4][tid] = make_ulonglong2(f2_pack(q2.y, q2.y), f2_pack(q2.z, q2.z));
    }
    s_mask[tid] = (unsigned short)my_mask;
    __syncthreads();

    // per-quarter compaction: entries whose cut ellipse can touch the quarter's 4x4 block
    const int nb = min(kBwdQBatch, walk - i * kBwdQBatch);
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
    const int cnt = quarter == 0 ? c0 : (quarter == 1 ? c1 : (quarter == 2 ? c2 : c3));
    const int cnt_max = max(max(c0, c1), max(c2, c3));
    const int posbase = walk - i * kBwdQBatch - 1;   // 0-based list position of staged entry j = posbase - j

    for (int k = 0; k < cnt_max; ++k) {
      const bool active = k < cnt;
      // one broadcast load: the four quarters' k-th entries; slots past a list's end hold an older
      // (in-range) index and are masked by `active`
      const int j = (int)((lds_u32(list_w + 4u * (unsigned)k) >> qshift) & 0xFFu);
      const int pos = posbase - j;
      const ulonglong2* eq = &s_q[0][j];
      const ulonglong2 e0 = eq[0], e1 = eq[kBwdQBatch], e2 = eq[2 * kBwdQBatch];
      const float dx = GSR_SUB(f2_lo(e0.x), pxf);
      const float pc = f2_hi(e0.x);
      const f2 dx2 = f2_pack(dx, dx);
      const f2 dy2 = f2_add(e0.y, npy2);
      // pair_power for both pixels, same rounding sequence as the forward
      const f2 qf = f2_fma(dx2, f2_mul(dx2, e1.x), f2_mul(dy2, f2_mul(dy2, e2.x)));
      const f2 pw2 = f2_fma(qf, mhalf2, f2_mul(dy2, f2_mul(dx2, e1.y)));
      const float pw_a = f2_lo(pw2), pw_b = f2_hi(pw2);
      bool va = active && (pos < lc_a) && !(pw_a > 0.0f) && !(pw_a < pc);
      bool vb = active && (pos < lc_b) && !(pw_b > 0.0f) && !(pw_b < pc);
      if (!__any_sync(0xffffffffu, va || vb)) continue;
      const float o = f2_lo(e2.y), ps = f2_hi(e2.y);
      float Ga = EXACT ? expf(pw_a) : fast_exp(pw_a), Gb = EXACT ? expf(pw_b) : fast_exp(pw_b);
      float al_a = pair_alpha(o, Ga), al_b = pair_alpha(o, Gb);
      // the forward blended this pair iff min(0.99, o * expf(power)) >= 15/255: certain for power >=
      // power_sure; the few pairs below it repeat the forward's exact evaluation (divergent, rare)
      if (EXACT) {
        va = va && !(al_a < kAlphaMin);
        vb = vb && !(al_b < kAlphaMin);
      } else if ((va && pw_a < ps) || (vb && pw_b < ps)) {
        if (va && pw_a < ps) { Ga = expf(pw_a); al_a = pair_alpha(o, Ga); va = !(al_a < kAlphaMin); }
        if (vb && pw_b < ps) { Gb = expf(pw_b); al_b = pair_alpha(o, Gb); vb = !(al_b < kAlphaMin); }
      }
      const unsigned vmask = __ballot_sync(0xffffffffu, va || vb);
      if (vmask == 0u) continue;
      if (!va) { Ga = 0.f; al_a = 0.f; }
      if (!vb) { Gb = 0.f; al_b = 0.f; }
      const f2 G2 = f2_pack(Ga, Gb), alpha2 = f2_pack(al_a, al_b);

      const ulonglong2 e3 = eq[3 * kBwdQBatch], e4 = eq[4 * kBwdQBatch];
      const f2 om2 = f2_fma(alpha2, mone2, one2);  // 1 - alpha (>= 0.01)
      const f2 inv2 = f2_pack(fast_rcp(f2_lo(om2)), fast_rcp(f2_hi(om2)));
      if (EXACT) T2 = f2_pack(__fdiv_rn(f2_lo(T2), f2_lo(om2)), __fdiv_rn(f2_hi(T2), f2_hi(om2)));   // as the reference: T / (1 - alpha)
      else T2 = f2_mul(T2, inv2);
      const f2 aT2 = f2_mul(alpha2, T2);
      // c - B for colour, depth and var
      const f2 d0 = f2_fma(Bc0, mone2, e3.y), d1 = f2_fma(Bc1, mone2, e4.x), d2 = f2_fma(Bc2, mone2, e4.y);
      const f2 dgt2 = f2_add(e3.x, ngt2);
      const f2 cvar2 = f2_mul(dgt2, dgt2);
      const f2 dd = f2_fma(Bd, mone2, e3.x), dv = f2_fma(Bv, mone2, cvar2);
      const f2 colour_part = f2_fma(d2, dLp2, f2_fma(d1, dLp1, f2_mul(d0, dLp0)));
      const f2 depth_part = f2_mul(dd, dLd);
      f2 dLa = f2_fma(dv, dLv, f2_add(colour_part, depth_part));
      // B <- B + alpha (c - B)
      Bc0 = f2_fma(alpha2, d0, Bc0);
      Bc1 = f2_fma(alpha2, d1, Bc1);
      Bc2 = f2_fma(alpha2, d2, Bc2);
      Bd = f2_fma(alpha2, dd, Bd);
      Bv = f2_fma(alpha2, dv, Bv);
      const f2 aTd = f2_mul(aT2, dLd);
      // dL/dalpha = T * (...) - T_final / (1 - alpha) * (bg . dL/dpixel)
      dLa = f2_fma(inv2, tfbg2, f2_mul(dLa, T2));
      const f2 w2 = f2_mul(G2, dLa);
      const f2 wx = f2_mul(w2, dx2), wy = f2_mul(w2, dy2);

      f2 v[kRedVals];
      v[ACC_MX] = wx;
      v[ACC_MY] = wy;
      v[ACC_CA] = f2_mul(wx, dx2);
      v[ACC_CB] = f2_mul(wx, dy2);
      v[ACC_CC] = f2_mul(wy, dy2);
      v[ACC_OP] = w2;
      v[ACC_R] = f2_mul(aT2, dLp0);
      v[ACC_G] = f2_mul(aT2, dLp1);
      v[ACC_B] = f2_mul(aT2, dLp2);
      v[ACC_DEPTH] = f2_fma(f2_mul(aT2, dgt2), dLv_x2, aTd);
      v[ACC_PGX] = 0ull;
      v[ACC_PGY] = 0ull;
      v[ACC_PD] = 0ull;
      v[ACC_MED] = 0ull;
      if (VARIANT == kLight) {
        v[ACC_PD] = aTd;
        // median: the first valid entry met from the back whose restored T exceeds 0.5
        float med_a = 0.f, med_b = 0.f;
        if (va && mid_a && f2_lo(T2) > 0.5f) { med_a = gma; mid_a = false; }
        if (vb && mid_b && f2_hi(T2) > 0.5f) { med_b = gmb; mid_b = false; }
        v[ACC_MED] = f2_pack(med_a, med_b);
      } else {
        // pose terms of the reference's ComputePG: colour through ndc without the background
        // term (full backward.cu:746-777, :1028-1072); depth only from the front-most valid
        // contributor of the pixel, because dd_dv* is assigned, not accumulated (:1278-1289).
        const bool fa = va & (pos == fm1_a), fb = vb & (pos == fm1_b);
        const f2 fsel = f2_pack(fa ? 1.f : 0.f, fb ? 1.f : 0.f);
        const f2 pa = f2_mul(T2, f2_fma(fsel, depth_part, colour_part));
        v[ACC_PD] = f2_mul(fsel, aTd);
        const f2 q = f2_mul(pa, G2);
        v[ACC_PGX] = f2_mul(q, dx2);
        v[ACC_PGY] = f2_mul(q, dy2);
      }
      // reduction through shared memory: every lane stores its column (both pixels added first), then
      // lane (quarter, l) adds the 8 values of its own quarter for slot l (second pass: slot 8 + l) and
      // issues the red into the quarter's Gaussian
#pragma unroll
      for (int qn = 0; qn < RS::N; ++qn)
        sts_f32(red_st + qn * kRowBytes, f2_lo(v[RS::slot(qn)]) + f2_hi(v[RS::slot(qn)]));
      __syncwarp();
      if ((vmask >> qshift) & 0xFFu) {
        char* line = reinterpret_cast<char*>(acc + (size_t)s_id[j] * kAccStride);
        if (POSE_ONLY) {
          if (ql4 < 4u * RS::N) atomicAdd(reinterpret_cast<float*>(line) + RS::slot((int)(ql4 >> 2)), row8_sum(red_ld));
        } else {
          float* dst = reinterpret_cast<float*>(line + ql4);   // slot ql, second pass slot ql + 8
          if (RS::N >= 8 || ql4 < 4u * RS::N) atomicAdd(dst, row8_sum(red_ld));
          if (RS::N > 8 && ql4 < 4u * (RS::N - 8)) atomicAdd(dst + 8, row8_sum(red_ld + 8 * kRowBytes));
        }
      }
      __syncwarp();
    }
  }
}


// =================================================================================================
// Octet variant: FOUR vertically adjacent pixels per lane (two fp32x2 pairs), a 4-lane group = a 4x4
// pixel block, a warp = a 16x8 half tile whose EIGHT groups walk their own compacted entry lists side
// by side; one warp per CTA (no block-level barrier at all, every warp stages its tile's entries itself).
// Why: the quarter-list kernel spends ~171 instructions per list iteration for 64 pixel slots; most of
// that is per-lane fixed cost (list fetch, entry loads, predicates, the 13-slot store / reduce / red).
// With four pixels per lane the same fixed cost covers 128 pixel slots: the packed arithmetic doubles,
// the reduction and the scalar bookkeeping do not; eight 4x4 lists in lock step need 1.22 M iterations
// per C3 frame against 2.33 M (tools/blend_stats.py).  The per-pixel arithmetic (power, alpha, skip
// decisions) is the quarter-list kernel's, instruction for instruction.
// =================================================================================================
constexpr int kBwdOBatch = 64;  // entries staged per round: two per lane
constexpr int kBwdORow = 40;    // floats per reduction row: 160 bytes = 32 mod 128 -> conflict-free LDS.128

struct OctConst { f2 dLp0, dLp1, dLp2, dLd, dLv, ngt2, tfbg2; };
struct OctState { f2 T2, Bc0, Bc1, Bc2, Bd, Bv; };

// one pixel pair's share of an entry: updates the pair's state, adds its partial sums to v (ACCUM) or
// initialises v with them
template <int VARIANT, bool ACCUM>
__device__ __forceinline__ void oct_pair(OctState& S, const OctConst& K, const f2 G2, const f2 alpha2, const f2 dx2,
                                         const f2 dy2, const ulonglong2& e3, const ulonglong2& e4, const f2 fsel,
                                         f2 (&v)[kRedVals]) {
  const f2 one2 = f2_pack(1.f, 1.f), mone2 = f2_pack(-1.f, -1.f), two2 = f2_pack(2.f, 2.f);
  const f2 om2 = f2_fma(alpha2, mone2, one2);  // 1 - alpha (>= 0.01)
  const f2 inv2 = f2_pack(fast_rcp(f2_lo(om2)), fast_rcp(f2_hi(om2)));
  S.T2 = f2_mul(S.T2, inv2);
  const f2 aT2 = f2_mul(alpha2, S.T2);
  const f2 d0 = f2_fma(S.Bc0, mone2, e3.y), d1 = f2_fma(S.Bc1, mone2, e4.x), d2 = f2_fma(S.Bc2, mone2, e4.y);
  const f2 dgt2 = f2_add(e3.x, K.ngt2);
  const f2 cvar2 = f2_mul(dgt2, dgt2);
  const f2 dd = f2_fma(S.Bd, mone2, e3.x), dv = f2_fma(S.Bv, mone2, cvar2);
  const f2 colour_part = f2_fma(d2, K.dLp2, f2_fma(d1, K.dLp1, f2_mul(d0, K.dLp0)));
  const f2 depth_part = f2_mul(dd, K.dLd);
  f2 dLa = f2_fma(dv, K.dLv, f2_add(colour_part, depth_part));
  S.Bc0 = f2_fma(alpha2, d0, S.Bc0);
  S.Bc1 = f2_fma(alpha2, d1, S.Bc1);
  S.Bc2 = f2_fma(alpha2, d2, S.Bc2);
  S.Bd = f2_fma(alpha2, dd, S.Bd);
  S.Bv = f2_fma(alpha2, dv, S.Bv);
  const f2 aTd = f2_mul(aT2, K.dLd);
  dLa = f2_fma(inv2, K.tfbg2, f2_mul(dLa, S.T2));
  const f2 w2 = f2_mul(G2, dLa);
  const f2 wx = f2_mul(w2, dx2), wy = f2_mul(w2, dy2);
  const f2 dvar = f2_mul(f2_mul(aT2, dgt2), K.dLv);   // alpha T (depth - gt) dL/dvar
  if (ACCUM) {
    v[ACC_MX] = f2_add(v[ACC_MX], wx);
    v[ACC_MY] = f2_add(v[ACC_MY], wy);
    f2_fma_acc(v[ACC_CA], wx, dx2);
    f2_fma_acc(v[ACC_CB], wx, dy2);
    f2_fma_acc(v[ACC_CC], wy, dy2);
    v[ACC_OP] = f2_add(v[ACC_OP], w2);
    f2_fma_acc(v[ACC_R], aT2, K.dLp0);
    f2_fma_acc(v[ACC_G], aT2, K.dLp1);
    f2_fma_acc(v[ACC_B], aT2, K.dLp2);
    v[ACC_DEPTH] = f2_add(v[ACC_DEPTH], f2_fma(dvar, two2, aTd));
  } else {
    v[ACC_MX] = wx;
    v[ACC_MY] = wy;
    v[ACC_CA] = f2_mul(wx, dx2);
    v[ACC_CB] = f2_mul(wx, dy2);
    v[ACC_CC] = f2_mul(wy, dy2);
    v[ACC_OP] = w2;
    v[ACC_R] = f2_mul(aT2, K.dLp0);
    v[ACC_G] = f2_mul(aT2, K.dLp1);
    v[ACC_B] = f2_mul(aT2, K.dLp2);
    v[ACC_DEPTH] = f2_fma(dvar, two2, aTd);
  }
  if (VARIANT == kLight) {
    if (ACCUM) {
      v[ACC_PD] = f2_add(v[ACC_PD], aTd);
    } else {
      v[ACC_PD] = aTd;
      v[ACC_MED] = 0ull;   // filled in by the caller (needs both pairs' restored T)
      v[ACC_PGX] = 0ull;
      v[ACC_PGY] = 0ull;
    }
  } else {
    // pose terms of the reference's ComputePG (see render_bwdq_kernel)
    const f2 pa = f2_mul(S.T2, f2_fma(fsel, depth_part, colour_part));
    const f2 q = f2_mul(pa, G2);
    if (ACCUM) {
      f2_fma_acc(v[ACC_PD], fsel, aTd);
      f2_fma_acc(v[ACC_PGX], q, dx2);
      f2_fma_acc(v[ACC_PGY], q, dy2);
    } else {
      v[ACC_PD] = f2_mul(fsel, aTd);
      v[ACC_PGX] = f2_mul(q, dx2);
      v[ACC_PGY] = f2_mul(q, dy2);
      v[ACC_MED] = 0ull;
    }
  }
}

// sum of 4 consecutive floats in shared memory (one LDS.128)
__device__ __forceinline__ float row4_sum(unsigned addr) {
  f2 a0, a1;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "r"(addr) : "memory");
  const f2 t = f2_add(a0, a1);
  return f2_lo(t) + f2_hi(t);
}

template <int VARIANT, bool POSE_ONLY>
__global__ void __launch_bounds__(32, 16)
render_bwdo_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
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
  // staged entry, duplicated into pairs (layout of render_bwdq_kernel)
  __shared__ ulonglong2 s_q[5][kBwdOBatch];
  __shared__ int s_id[kBwdOBatch];
  // entry lists of the eight groups, interleaved: byte [k][e] = k-th entry of group e
  __shared__ __align__(16) unsigned char s_list[kBwdOBatch][8];
  __shared__ __align__(16) float s_red[kRedVals][kBwdORow];

  const int lane = threadIdx.x;
  const int e = lane >> 2, l = lane & 3;     // group (4x4 block) and column inside it
  reinterpret_cast<uint4*>(&s_list[0][0])[lane] = make_uint4(0u, 0u, 0u, 0u);   // 64 x 8 bytes = 32 x 16
  const int tile_x = (int)blockIdx.x >> 1, half = (int)blockIdx.x & 1;          // half: rows 8 half .. 8 half + 7
  const int tile = blockIdx.y * grid_x + tile_x;
  const int px = tile_x * kTileX + (e & 3) * 4 + l;
  const int py0 = blockIdx.y * kTileY + half * 8 + (e >> 2) * 4;
  const bool in_x = px < W;
  const uint32_t pix0 = (uint32_t)W * (uint32_t)py0 + (uint32_t)px;
  const float pxf = (float)px;
  const f2 npyA = f2_pack(-(float)py0, -(float)(py0 + 1)), npyB = f2_pack(-(float)(py0 + 2), -(float)(py0 + 3));
  const float tile_x0 = (float)(tile_x * kTileX), tile_y0 = (float)(blockIdx.y * kTileY);

  const uint2 range = ranges[tile];
  const int walk = (int)tile_last[tile];  // entries [0, walk) were used by some pixel
  const int rounds = (walk + kBwdOBatch - 1) / kBwdOBatch;
  const size_t HW = (size_t)H * (size_t)W;

  float Tf[4], g0[4], g1[4], g2[4], gd[4], gv[4], gt[4], gm[4];
  int lc[4], fm1[4];   // fm1: 0-based position of the front-most contributor
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    Tf[r] = g0[r] = g1[r] = g2[r] = gd[r] = gv[r] = gt[r] = gm[r] = 0.f;
    lc[r] = 0; fm1[r] = -1;
    if (in_x && py0 + r < H) {
      const uint32_t pix = pix0 + (uint32_t)r * (uint32_t)W;
      Tf[r] = (VARIANT == kLight) ? (1 - alphas[pix]) : final_Ts[pix];
      lc[r] = (int)n_contrib[pix];
      g0[r] = dL_dpix[pix]; g1[r] = dL_dpix[HW + pix]; g2[r] = dL_dpix[2 * HW + pix];
      gd[r] = dL_ddepths[pix]; gv[r] = dL_dvars != nullptr ? dL_dvars[pix] : 0.f; gt[r] = gt_depth[pix];
      if (VARIANT == kLight && dL_dmedians != nullptr) gm[r] = dL_dmedians[pix];
      if (VARIANT == kFull) fm1[r] = (int)first_contrib[pix] - 1;
    }
  }
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  OctConst KA, KB;
  KA.dLp0 = f2_pack(g0[0], g0[1]); KA.dLp1 = f2_pack(g1[0], g1[1]); KA.dLp2 = f2_pack(g2[0], g2[1]);
  KA.dLd = f2_pack(gd[0], gd[1]); KA.dLv = f2_pack(gv[0], gv[1]); KA.ngt2 = f2_pack(-gt[0], -gt[1]);
  KA.tfbg2 = f2_pack(-Tf[0] * (bg0 * g0[0] + bg1 * g1[0] + bg2 * g2[0]), -Tf[1] * (bg0 * g0[1] + bg1 * g1[1] + bg2 * g2[1]));
  KB.dLp0 = f2_pack(g0[2], g0[3]); KB.dLp1 = f2_pack(g1[2], g1[3]); KB.dLp2 = f2_pack(g2[2], g2[3]);
  KB.dLd = f2_pack(gd[2], gd[3]); KB.dLv = f2_pack(gv[2], gv[3]); KB.ngt2 = f2_pack(-gt[2], -gt[3]);
  KB.tfbg2 = f2_pack(-Tf[2] * (bg0 * g0[2] + bg1 * g1[2] + bg2 * g2[2]), -Tf[3] * (bg0 * g0[3] + bg1 * g1[3] + bg2 * g2[3]));
  OctState SA, SB;
  SA.T2 = f2_pack(Tf[0], Tf[1]); SB.T2 = f2_pack(Tf[2], Tf[3]);
  SA.Bc0 = SA.Bc1 = SA.Bc2 = SA.Bd = SA.Bv = 0ull;
  SB.Bc0 = SB.Bc1 = SB.Bc2 = SB.Bd = SB.Bv = 0ull;
  bool mid0 = true, mid1 = true, mid2 = true, mid3 = true;
  const f2 mhalf2 = f2_pack(-0.5f, -0.5f);
  using RS = RedSet<VARIANT, POSE_ONLY>;
  constexpr unsigned kRowBytes = kBwdORow * 4;
  const unsigned red_st = pin_reg(smem_u32(&s_red[0][lane]));
  // reducing lane (e, l): slots l, l + 4, l + 8, l + 12 of its OWN group's four values
  const unsigned red_ld = pin_reg(smem_u32(&s_red[l][4 * e]));
  const unsigned list_w = pin_reg(smem_u32(&s_list[0][4 * (e >> 2)]));   // the word holding this group's byte
  const unsigned qshift = pin_reg((unsigned)(e & 3) * 8u);
  const unsigned gshift = pin_reg((unsigned)e * 4u);                     // this group's lanes in a ballot
  const unsigned lt = (1u << lane) - 1u;
  const int mshift = 8 * half;                                           // this warp's byte of block_mask16

  for (int i = 0; i < rounds; ++i) {
    __syncwarp();
    // ---- stage 64 entries (two per lane), keep their block masks in registers -----------------------
    unsigned m_s[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const int t = lane + 32 * s2;
      const int progress = i * kBwdOBatch + t;
      m_s[s2] = 0u;
      if (progress < walk) {
        const int id = (int)point_list[range.x + (walk - progress - 1)];
        const float4* r = rec + 3 * (size_t)id;
        const float4 q0 = __ldg(r + 0), q1 = __ldg(r + 1), q2 = __ldg(r + 2);
        s_id[t] = id;
        m_s[s2] = (block_mask16(q0, q1, tile_x0, tile_y0) >> mshift) & 0xFFu;
        s_q[0][t] = make_ulonglong2(f2_pack(q0.x, q1.z), f2_pack(q0.y, q0.y));
        s_q[1][t] = make_ulonglong2(f2_pack(q0.z, q0.z), f2_pack(-q0.w, -q0.w));
        s_q[2][t] = make_ulonglong2(f2_pack(q1.x, q1.x), f2_pack(q1.y, q2.w));
        s_q[3][t] = make_ulonglong2(f2_pack(q1.w, q1.w), f2_pack(q2.x, q2.x));
        s_q[4][t] = make_ulonglong2(f2_pack(q2.y, q2.y), f2_pack(q2.z, q2.z));
      }
    }
    // ---- per-group compaction: entries whose cut ellipse can touch the group's 4x4 block ------------
    int c[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) c[b] = 0;
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool hit = (m_s[s2] >> b) & 1u;
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        if (hit) s_list[c[b] + __popc(ball & lt)][b] = (unsigned char)(lane + 32 * s2);
        c[b] += __popc(ball);
      }
    }
    __syncwarp();
    int cnt = c[0], cnt_max = c[0];
#pragma unroll
    for (int b = 1; b < 8; ++b) {
      cnt = (e == b) ? c[b] : cnt;
      cnt_max = max(cnt_max, c[b]);
    }
    const int posbase = walk - i * kBwdOBatch - 1;   // 0-based list position of staged entry j = posbase - j

    for (int k = 0; k < cnt_max; ++k) {
      const bool active = k < cnt;
      const int j = (int)((lds_u32(list_w + 8u * (unsigned)k) >> qshift) & 0xFFu);
      const int pos = posbase - j;
      const ulonglong2* eq = &s_q[0][j];
      const ulonglong2 e0 = eq[0], e1 = eq[kBwdOBatch], e2 = eq[2 * kBwdOBatch];
      const float dx = GSR_SUB(f2_lo(e0.x), pxf);
      const float pc = f2_hi(e0.x);
      const f2 dx2 = f2_pack(dx, dx);
      const f2 dyA = f2_add(e0.y, npyA), dyB = f2_add(e0.y, npyB);
      // pair_power for the four pixels, same rounding sequence as the forward
      const f2 dxA = f2_mul(dx2, e1.x), dxB = f2_mul(dx2, e1.y);
      const f2 qfA = f2_fma(dx2, dxA, f2_mul(dyA, f2_mul(dyA, e2.x)));
      const f2 qfB = f2_fma(dx2, dxA, f2_mul(dyB, f2_mul(dyB, e2.x)));
      const f2 pwA = f2_fma(qfA, mhalf2, f2_mul(dyA, dxB));
      const f2 pwB = f2_fma(qfB, mhalf2, f2_mul(dyB, dxB));
      const float p0 = f2_lo(pwA), p1 = f2_hi(pwA), p2 = f2_lo(pwB), p3 = f2_hi(pwB);
      bool v0 = active && (pos < lc[0]) && !(p0 > 0.0f) && !(p0 < pc);
      bool v1 = active && (pos < lc[1]) && !(p1 > 0.0f) && !(p1 < pc);
      bool v2 = active && (pos < lc[2]) && !(p2 > 0.0f) && !(p2 < pc);
      bool v3 = active && (pos < lc[3]) && !(p3 > 0.0f) && !(p3 < pc);
      if (!__any_sync(0xffffffffu, v0 || v1 || v2 || v3)) continue;
      const float o = f2_lo(e2.y), ps = f2_hi(e2.y);
      float G0 = fast_exp(p0), G1 = fast_exp(p1), G2 = fast_exp(p2), G3 = fast_exp(p3);
      float a0 = pair_alpha(o, G0), a1 = pair_alpha(o, G1), a2 = pair_alpha(o, G2), a3 = pair_alpha(o, G3);
      // the forward blended this pair iff min(0.99, o * expf(power)) >= 15/255: certain for power >=
      // power_sure; the few pairs below it repeat the forward's exact evaluation (divergent, rare)
      if ((v0 && p0 < ps) || (v1 && p1 < ps) || (v2 && p2 < ps) || (v3 && p3 < ps)) {
        if (v0 && p0 < ps) { G0 = expf(p0); a0 = pair_alpha(o, G0); v0 = !(a0 < kAlphaMin); }
        if (v1 && p1 < ps) { G1 = expf(p1); a1 = pair_alpha(o, G1); v1 = !(a1 < kAlphaMin); }
        if (v2 && p2 < ps) { G2 = expf(p2); a2 = pair_alpha(o, G2); v2 = !(a2 < kAlphaMin); }
        if (v3 && p3 < ps) { G3 = expf(p3); a3 = pair_alpha(o, G3); v3 = !(a3 < kAlphaMin); }
      }
      const unsigned vmask = __ballot_sync(0xffffffffu, v0 || v1 || v2 || v3);
      if (vmask == 0u) continue;
      if (!v0) { G0 = 0.f; a0 = 0.f; }
      if (!v1) { G1 = 0.f; a1 = 0.f; }
      if (!v2) { G2 = 0.f; a2 = 0.f; }
      if (!v3) { G3 = 0.f; a3 = 0.f; }

      const ulonglong2 e3 = eq[3 * kBwdOBatch], e4 = eq[4 * kBwdOBatch];
      f2 fselA = 0ull, fselB = 0ull;
      if (VARIANT == kFull) {
        fselA = f2_pack((v0 & (pos == fm1[0])) ? 1.f : 0.f, (v1 & (pos == fm1[1])) ? 1.f : 0.f);
        fselB = f2_pack((v2 & (pos == fm1[2])) ? 1.f : 0.f, (v3 & (pos == fm1[3])) ? 1.f : 0.f);
      }
      f2 v[kRedVals];
      oct_pair<VARIANT, false>(SA, KA, f2_pack(G0, G1), f2_pack(a0, a1), dx2, dyA, e3, e4, fselA, v);
      oct_pair<VARIANT, true>(SB, KB, f2_pack(G2, G3), f2_pack(a2, a3), dx2, dyB, e3, e4, fselB, v);
      if (VARIANT == kLight) {
        // median: the first valid entry met from the back whose restored T exceeds 0.5
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
        if (v0 && mid0 && f2_lo(SA.T2) > 0.5f) { m0 = gm[0]; mid0 = false; }
        if (v1 && mid1 && f2_hi(SA.T2) > 0.5f) { m1 = gm[1]; mid1 = false; }
        if (v2 && mid2 && f2_lo(SB.T2) > 0.5f) { m2 = gm[2]; mid2 = false; }
        if (v3 && mid3 && f2_hi(SB.T2) > 0.5f) { m3 = gm[3]; mid3 = false; }
        v[ACC_MED] = f2_pack(m0 + m2, m1 + m3);
      }

      // reduction through shared memory: every lane stores its four pixels' sum per slot, then lane (e, l)
      // adds its own group's four values for slots l, l + 4, l + 8, l + 12 and issues the red into the
      // group's Gaussian
#pragma unroll
      for (int qn = 0; qn < RS::N; ++qn)
        sts_f32(red_st + qn * kRowBytes, f2_lo(v[RS::slot(qn)]) + f2_hi(v[RS::slot(qn)]));
      __syncwarp();
      if ((vmask >> gshift) & 0xFu) {
        float* line = acc + (size_t)s_id[j] * kAccStride;
        if (POSE_ONLY) {
          if (l < RS::N) atomicAdd(line + RS::slot(l), row4_sum(red_ld));
        } else {
          float* dst = line + l;
#pragma unroll
          for (int p4 = 0; p4 < (RS::N + 3) / 4; ++p4)
            if (4 * p4 + 4 <= RS::N || l + 4 * p4 < RS::N) atomicAdd(dst + 4 * p4, row4_sum(red_ld + 4 * p4 * kRowBytes));
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace

int launch_render_bwd(int variant, const Camera& cam, const GeomState& g, const BinState& b,
                      const ImgState& img, const float* bg, const float* gt_depth,
                      const float* alphas, const BlendGrads& cot, float* acc, int num_gaussians,
                      int num_entries, bool pose_only, bool debug, cudaStream_t stream) {
  dim3 grid(cam.grid_x, cam.grid_y, 1);
  StageScope st(ST_RENDER_BWD, stream);
  // "bwd_packed": 0 scalar 8x4 kernel, 1 packed 8x8 kernel (one list per warp), 3 packed kernel with
  // quarter-warp lists, 2 (default) = 3.  "bwd_occ": CTAs per SM the quarter kernel is compiled for
  // (8 = 64 registers, 7 = 72 registers).
  const int mode = options().bwd_packed;
  const bool quarter = mode == 3 || mode == 2;
  const bool packed = mode == 1;
  const bool octet = mode == 4;
  // measured (C3 / C4, B200): -full 0.609 / 1.23 ms at 7 CTAs per SM (72 registers) vs 0.667 / 1.33 at 8 (the
  // 64-register build rematerialises shared addresses in the loop); -light 0.607 vs 0.620 at C3.
  // "bwd_occ" = 8 forces the 64-register build, anything else the 72-register one.
  const bool occ7 = options().bwd_occ != 8;
  (void)num_gaussians; (void)num_entries;
#define GSR_BWD_ARGS(FT, FC)                                                                       \
  img.ranges, b.vals, img.tile_last, cam.W, cam.H, cam.grid_x, g.rec, bg, gt_depth, alphas, FT,    \
      img.n_contrib, FC, cot.dL_dpix, cot.dL_ddepth, cot.dL_dmedian, cot.dL_dvar, acc
#define GSR_BWDQ(V, PO, FT, FC)                                                                    \
  do {                                                                                             \
    if (V == kLight && !PO && options().exact_median != 0)                                         \
      render_bwdq_kernel<kLight, false, 7, true><<<grid, kBwdQThreads, 0, stream>>>(GSR_BWD_ARGS(FT, FC)); \
    else if (occ7) render_bwdq_kernel<V, PO, 7><<<grid, kBwdQThreads, 0, stream>>>(GSR_BWD_ARGS(FT, FC)); \
    else render_bwdq_kernel<V, PO, 8><<<grid, kBwdQThreads, 0, stream>>>(GSR_BWD_ARGS(FT, FC));    \
  } while (0)
  const dim3 grid_o(2 * cam.grid_x, cam.grid_y, 1);   // octet kernel: one warp per half tile
  if (octet) {
    if (variant == kLight && pose_only)
      render_bwdo_kernel<kLight, true><<<grid_o, 32, 0, stream>>>(GSR_BWD_ARGS(nullptr, nullptr));
    else if (variant == kLight)
      render_bwdo_kernel<kLight, false><<<grid_o, 32, 0, stream>>>(GSR_BWD_ARGS(nullptr, nullptr));
    else
      render_bwdo_kernel<kFull, false><<<grid_o, 32, 0, stream>>>(GSR_BWD_ARGS(img.final_T, img.first_contrib));
  } else if (variant == kLight) {
    if (quarter && pose_only)
      GSR_BWDQ(kLight, true, nullptr, nullptr);
    else if (quarter)
      GSR_BWDQ(kLight, false, nullptr, nullptr);
    else if (packed && pose_only)
      render_bwd2_kernel<kLight, true><<<grid, kBwd2Threads, 0, stream>>>(GSR_BWD_ARGS(nullptr, nullptr));
    else if (packed)
      render_bwd2_kernel<kLight, false><<<grid, kBwd2Threads, 0, stream>>>(GSR_BWD_ARGS(nullptr, nullptr));
    else if (pose_only)
      render_bwd_kernel<kLight, true><<<grid, kTileThreads, 0, stream>>>(GSR_BWD_ARGS(nullptr, nullptr));
    else
      render_bwd_kernel<kLight, false><<<grid, kTileThreads, 0, stream>>>(GSR_BWD_ARGS(nullptr, nullptr));
  } else {
    if (quarter)
      GSR_BWDQ(kFull, false, img.final_T, img.first_contrib);
    else if (packed)
      render_bwd2_kernel<kFull, false><<<grid, kBwd2Threads, 0, stream>>>(GSR_BWD_ARGS(img.final_T, img.first_contrib));
    else
      render_bwd_kernel<kFull, false><<<grid, kTileThreads, 0, stream>>>(GSR_BWD_ARGS(img.final_T, img.first_contrib));
  }
#undef GSR_BWDQ
#undef GSR_BWD_ARGS
  GSR_LAUNCH_OK(debug, stream);
  return GSR_OK;
}

}  // namespace gsr
