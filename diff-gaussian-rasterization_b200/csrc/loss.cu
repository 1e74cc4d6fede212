// loss.cu — RGB-D L1 loss and its per-pixel cotangents in ONE pass over the rendered images
// (helper next to the hot path, not part of the reference surface).
//
// A CG-SLAM / SplaTAM style mapping step evaluates, after rasterize_gaussians,
//   L = w_c sum |C - C_gt| + w_d sum_m |D - D_gt| + (light) w_0 sum_m |D_med - D_gt| + w_1 sum depth_var
//                                                  (full)  w_0 sum (1 - O)
// with torch: ~15 element-wise / reduction kernels over 2 M pixels forward and the same again in
// autograd, i.e. every image is read and written several times (measured 0.17 ms of device time per
// 1080p frame next to a 1.26 ms rasterizer frame).  This kernel reads the rendered images and the
// ground-truth frame once — in the formats RGB-D datasets ship, 8-bit colour and 16-bit depth, or
// fp32 — and writes the loss and the cotangent images the rasterizer's backward consumes.
// The loss is summed in a fixed order (per-block partials, last block adds them): deterministic.
// Roofline: HBM, 4 B x (channels in + channels out) + ground truth bytes per pixel.
#include "gsr_common.cuh"

namespace gsr {
namespace {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float sgn(float e) { return e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f); }

template <int VARIANT, bool U8, bool I16>
__global__ void __launch_bounds__(kLossThreads)
rgbd_l1_kernel(int HW, const float* __restrict__ color, const float* __restrict__ depth,
               const float* __restrict__ aux0, const float* __restrict__ aux1,
               const void* __restrict__ gt_color, const void* __restrict__ gt_depth, gsr_rgbd_l1 prm,
               float* __restrict__ dL_dcolor, float* __restrict__ dL_ddepth, float* __restrict__ dL_daux0,
               float* __restrict__ dL_daux1, float* __restrict__ loss, float* __restrict__ partials,
               unsigned int* __restrict__ counter) {
  __shared__ float s_part[kLossThreads / 32];
  __shared__ bool s_last;
  const int i = blockIdx.x * kLossThreads + threadIdx.x;
  float l = 0.f;
  if (i < HW) {
    float gc[3], gd;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      gc[c] = U8 ? prm.color_scale * (float)static_cast<const unsigned char*>(gt_color)[(size_t)c * HW + i]
                 : static_cast<const float*>(gt_color)[(size_t)c * HW + i];
    gd = I16 ? prm.depth_scale * (float)static_cast<const short*>(gt_depth)[i]
             : static_cast<const float*>(gt_depth)[i];
    const float md = (prm.depth_mask != 0 && !(gd > 0.f)) ? 0.f : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e = color[(size_t)c * HW + i] - gc[c];
      l += prm.w_color * fabsf(e);
      dL_dcolor[(size_t)c * HW + i] = prm.w_color * sgn(e);
    }
    const float ed = depth[i] - gd;
    l += md * prm.w_depth * fabsf(ed);
    dL_ddepth[i] = md * prm.w_depth * sgn(ed);
    if (VARIANT == kLight) {
      const float em = aux0[i] - gd;              // median depth
      l += md * prm.w_aux0 * fabsf(em);
      dL_daux0[i] = md * prm.w_aux0 * sgn(em);
      l += prm.w_aux1 * aux1[i];                  // depth_var enters linearly
      dL_daux1[i] = prm.w_aux1;
    } else {
      l += prm.w_aux0 * (1.0f - aux0[i]);         // silhouette: accumulated opacity should reach 1
      dL_daux0[i] = -prm.w_aux0;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) s_part[warp] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) t += s_part[w];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kLossThreads) t += __ldcg(partials + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) s_part[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kLossThreads / 32; ++w) s += s_part[w];
      loss[0] = s;
      *counter = 0u;  // ready for the next call on the same scratch
    }
  }
}

}  // namespace
}  // namespace gsr

extern "C" size_t gsr_rgbd_l1_scratch_floats(int width, int height) {
  const size_t hw = (size_t)(width > 0 ? width : 0) * (size_t)(height > 0 ? height : 0);
  return (hw + gsr::kLossThreads - 1) / gsr::kLossThreads + 16;
}

extern "C" int gsr_rgbd_l1_loss(int variant, int width, int height, const float* color, const float* depth,
                                const float* aux0, const float* aux1, const void* gt_color,
                                int gt_color_is_u8, const void* gt_depth, int gt_depth_is_i16,
                                const gsr_rgbd_l1* prm, float* dL_dcolor, float* dL_ddepth, float* dL_daux0,
                                float* dL_daux1, float* loss, float* scratch, void* stream) {
  using namespace gsr;
  OptionsCall oc("gsr_rgbd_l1_loss");
  const bool light = variant == kLight;
  if ((variant != kLight && variant != kFull) || width <= 0 || height <= 0 || !color || !depth || !aux0 ||
      (light && !aux1) || !gt_color || !gt_depth || !prm || !dL_dcolor || !dL_ddepth || !dL_daux0 ||
      (light && !dL_daux1) || !loss || !scratch) {
    set_error("gsr_rgbd_l1_loss: bad arguments");
    return GSR_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int HW = width * height;
  const int blocks = (HW + kLossThreads - 1) / kLossThreads;
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch);
  float* partials = scratch + 16;
  StageScope st(ST_OTHER, s, 2);
  GSR_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
#define GSR_L1(V, U8, I16)                                                                           \
  rgbd_l1_kernel<V, U8, I16><<<blocks, kLossThreads, 0, s>>>(HW, color, depth, aux0, aux1, gt_color, \
                                                             gt_depth, *prm, dL_dcolor, dL_ddepth,  \
                                                             dL_daux0, dL_daux1, loss, partials, counter)
  const bool u8 = gt_color_is_u8 != 0, i16 = gt_depth_is_i16 != 0;
  if (light) {
    if (u8 && i16) GSR_L1(kLight, true, true); else if (u8) GSR_L1(kLight, true, false);
    else if (i16) GSR_L1(kLight, false, true); else GSR_L1(kLight, false, false);
  } else {
    if (u8 && i16) GSR_L1(kFull, true, true); else if (u8) GSR_L1(kFull, true, false);
    else if (i16) GSR_L1(kFull, false, true); else GSR_L1(kFull, false, false);
  }
#undef GSR_L1
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}
