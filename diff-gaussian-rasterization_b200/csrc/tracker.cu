// tracker.cu — camera-pose tracking iterations on the device.
//
// CG-SLAM's tracking loop keeps the map fixed and runs, per frame, K iterations of
//     render(-light, map_off=True) -> L1 colour + depth loss -> backward -> dL/dviewmatrix
//     -> (autograd through the pose parametrisation) -> Adam step on the camera pose
// (reference README.md:60-95: the -light package exists for exactly this loop; its backward's
// `map_off` switch, L/cuda_rasterizer/backward.cu:593-666, drops every map gradient).  Through the
// reference surface each iteration costs one blocking read-back in the forward, ~10 small torch
// kernels for the loss, the autograd graph and the optimiser — at 640x480 / 100 k Gaussians that
// host work is longer than the rasterizer itself.
//
// Here one iteration is eight kernels and one memset with no host interaction, captured once in a CUDA graph and
// replayed K times:
//     preprocess_fwd -> scan_tiles -> scatter_entries -> sort_tiles        (static-capacity binning)
//     -> render_fwd<light, fused loss>   (writes cotangents + alpha, per-tile loss partials)
//     -> render_bwd2<light, pose only>   (3 accumulator slots per Gaussian)
//     -> preprocess_bwd<light>(pose contraction only)
//     -> track_update   (pose-gradient reduction, chain rule to quaternion + translation, Adam step,
//                        next iteration's view / projection / camera position, loss reduction)
// The binning buffer is sized from a probing forward (+50 %); ranges are clamped on the device, so
// an overflow can never leave the buffers — it raises a flag, and the run is repeated with a larger
// buffer from the saved start pose.
//
// Pose parametrisation: world-to-camera  W2C = [R(q/|q|) t; 0 1],  q = (w, x, y, z)  (the
// parametrisation of SplaTAM / CG-SLAM style trackers); viewmatrix = W2C^T in the reference's
// layout, so  dL/dR[r][c] = dL_dview[4c + r],  dL/dt[r] = dL_dview[12 + r].
#include <cstring>
#include <new>
#include <vector>

#include "gsr_common.cuh"

namespace gsr {

namespace {

struct PoseState {      // device resident
  float q[4];
  float t[3];
  float pad0;
  float m[8];           // Adam first moment  (q0..q3, t0..t2, unused)
  float v[8];           // Adam second moment
  float grad[8];        // dL/dq, dL/dt of the last iteration
  float dview[16];      // dL/dviewmatrix of the last iteration (reference layout)
  float twist[8];       // dL/d(xi) of the last iteration, xi = (omega, v): W2C' = exp(xi^) W2C
  int step;             // Adam step count
  int iter;             // iterations run since the last reset of the loss history
  int pad1[2];
};

struct UpdateParams {
  float lr_rot, lr_trans, beta1, beta2, eps;
  int max_hist;
};

__device__ __forceinline__ void quat_to_R(const float* q, float n_inv, float R[3][3]) {
  const float w = q[0] * n_inv, x = q[1] * n_inv, y = q[2] * n_inv, z = q[3] * n_inv;
  R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - w * z); R[0][2] = 2.f * (x * z + w * y);
  R[1][0] = 2.f * (x * y + w * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - w * x);
  R[2][0] = 2.f * (x * z - w * y); R[2][1] = 2.f * (y * z + w * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
}

// view (W2C^T, i.e. flat[4c + r] = W2C[r][c]), proj = (Persp * W2C) in the same flat layout,
// campos = -R^T t.  `persp` is the flat perspective matrix (flat[4k + r] = Persp[r][k]).
__device__ void write_camera(const float* q, const float* t, const float* __restrict__ persp,
                             float* view, float* proj, float* campos) {
  const float n_inv = rsqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  float R[3][3];
  quat_to_R(q, n_inv, R);
  float Wm[4][4];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Wm[r][c] = R[r][c];
    Wm[r][3] = t[r];
  }
  Wm[3][0] = 0.f; Wm[3][1] = 0.f; Wm[3][2] = 0.f; Wm[3][3] = 1.f;
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      view[4 * c + r] = Wm[r][c];
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += persp[4 * k + r] * Wm[k][c];
      proj[4 * c + r] = s;
    }
  for (int c = 0; c < 3; ++c) campos[c] = -(R[0][c] * t[0] + R[1][c] * t[1] + R[2][c] * t[2]);
}

__global__ void init_camera_kernel(const PoseState* ps, const float* persp, float* view, float* proj,
                                   float* campos) {
  if (threadIdx.x == 0 && blockIdx.x == 0) write_camera(ps->q, ps->t, persp, view, proj, campos);
}

// One CTA.  Warps 0..11 add the per-block pose partials of preprocess_bwd (fixed order), warp 12..15
// add the per-tile loss partials; thread 0 then applies the chain rule, the Adam step and writes the
// next camera; all threads finally clear the per-tile counters for the next iteration's preprocess.
constexpr int kUpdThreads = 512;
__global__ void __launch_bounds__(kUpdThreads)
track_update_kernel(int nblocks, const float* __restrict__ partials, int tiles,
                    const float* __restrict__ loss_partials, PoseState* ps,
                    const float* __restrict__ persp, float* view, float* proj, float* campos,
                    float* loss_hist, uint32_t* tile_count, uint32_t* counters, UpdateParams up, int cs) {
  __shared__ float s_g[12];
  __shared__ float s_l[4];
  __shared__ PoseState s_ps;   // the pose state travels through shared memory: one coalesced load /
  __shared__ float s_persp[16];  // store instead of dozens of dependent global accesses by thread 0
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kPsWords = (int)(sizeof(PoseState) / 4);
  if (tid >= 384 && tid < 384 + kPsWords)
    reinterpret_cast<uint32_t*>(&s_ps)[tid - 384] = reinterpret_cast<const uint32_t*>(ps)[tid - 384];
  if (tid >= 480 && tid < 496) s_persp[tid - 480] = persp[tid - 480];
  if (warp < 12) {
    float s = 0.f;
    for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * 12 + warp];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_g[warp] = s;
  } else {
    float s = 0.f;
    for (int b = (warp - 12) * 32 + lane; b < tiles; b += 128) s += loss_partials[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_l[warp - 12] = s;
  }
  __syncthreads();
  if (tid == 0) {
    const int it = s_ps.iter;
    if (it < up.max_hist) loss_hist[it] = (s_l[0] + s_l[1]) + (s_l[2] + s_l[3]);
    s_ps.iter = it + 1;
    // dL/dR[r][c] = g[3c + r], dL/dt[r] = g[9 + r]
    float dR[3][3], dt[3];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) dR[r][c] = s_g[3 * c + r];
    for (int r = 0; r < 3; ++r) dt[r] = s_g[9 + r];
    for (int c = 0; c < 4; ++c) {
      for (int r = 0; r < 3; ++r) s_ps.dview[4 * c + r] = s_g[3 * c + r];
      s_ps.dview[4 * c + 3] = 0.f;
    }
    // SE(3) tangent-space projection (left perturbation W2C' = exp(xi^) W2C, xi = (omega, v)):
    //   dL/dv = dL/dt,   dL/domega = sum_c R[:,c] x dL/dR[:,c] + t x dL/dt
    {
      float Rm[3][3];
      const float ni = rsqrtf(s_ps.q[0] * s_ps.q[0] + s_ps.q[1] * s_ps.q[1] + s_ps.q[2] * s_ps.q[2] +
                              s_ps.q[3] * s_ps.q[3]);
      quat_to_R(s_ps.q, ni, Rm);
      float om[3] = {0.f, 0.f, 0.f};
      for (int c = 0; c < 4; ++c) {
        const float a0 = c < 3 ? Rm[0][c] : s_ps.t[0], a1 = c < 3 ? Rm[1][c] : s_ps.t[1],
                    a2 = c < 3 ? Rm[2][c] : s_ps.t[2];
        const float b0 = c < 3 ? dR[0][c] : dt[0], b1 = c < 3 ? dR[1][c] : dt[1], b2 = c < 3 ? dR[2][c] : dt[2];
        om[0] += a1 * b2 - a2 * b1;
        om[1] += a2 * b0 - a0 * b2;
        om[2] += a0 * b1 - a1 * b0;
      }
      for (int k = 0; k < 3; ++k) { s_ps.twist[k] = om[k]; s_ps.twist[3 + k] = dt[k]; }
    }
    float q[4] = {s_ps.q[0], s_ps.q[1], s_ps.q[2], s_ps.q[3]};
    const float n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const float n_inv = rsqrtf(n2);
    const float w = q[0] * n_inv, x = q[1] * n_inv, y = q[2] * n_inv, z = q[3] * n_inv;
    // gradient w.r.t. the unit quaternion
    float gh[4];
    gh[0] = 2.f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
    gh[1] = 2.f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.f * x * dR[1][1] - w * dR[1][2] +
                   z * dR[2][0] + w * dR[2][1] - 2.f * x * dR[2][2]);
    gh[2] = 2.f * (-2.f * y * dR[0][0] + x * dR[0][1] + w * dR[0][2] + x * dR[1][0] + z * dR[1][2] -
                   w * dR[2][0] + z * dR[2][1] - 2.f * y * dR[2][2]);
    gh[3] = 2.f * (-2.f * z * dR[0][0] - w * dR[0][1] + x * dR[0][2] + w * dR[1][0] - 2.f * z * dR[1][1] +
                   y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);
    // through q_hat = q / |q|
    const float qh[4] = {w, x, y, z};
    const float dot = qh[0] * gh[0] + qh[1] * gh[1] + qh[2] * gh[2] + qh[3] * gh[3];
    float g[7];
    for (int k = 0; k < 4; ++k) g[k] = (gh[k] - qh[k] * dot) * n_inv;
    for (int k = 0; k < 3; ++k) g[4 + k] = dt[k];
    // Adam (torch.optim.Adam semantics, no weight decay / amsgrad)
    const int step = s_ps.step + 1;
    s_ps.step = step;
    const float bc1 = 1.f - powf(up.beta1, (float)step), bc2 = 1.f - powf(up.beta2, (float)step);
    float p[7] = {q[0], q[1], q[2], q[3], s_ps.t[0], s_ps.t[1], s_ps.t[2]};
    for (int k = 0; k < 7; ++k) {
      s_ps.grad[k] = g[k];
      const float mk = up.beta1 * s_ps.m[k] + (1.f - up.beta1) * g[k];
      const float vk = up.beta2 * s_ps.v[k] + (1.f - up.beta2) * g[k] * g[k];
      s_ps.m[k] = mk;
      s_ps.v[k] = vk;
      const float lr = k < 4 ? up.lr_rot : up.lr_trans;
      const float denom = sqrtf(vk) / sqrtf(bc2) + up.eps;
      p[k] -= (lr / bc1) * (mk / denom);
    }
    for (int k = 0; k < 4; ++k) s_ps.q[k] = p[k];
    for (int k = 0; k < 3; ++k) s_ps.t[k] = p[4 + k];
    write_camera(p, p + 4, s_persp, view, proj, campos);
    counters[0] = 0u; counters[1] = 0u; counters[2] = 0u;  // [3] (overflow flag) is sticky
  }
  for (int i = tid; i < tiles * cs; i += kUpdThreads) tile_count[i] = 0u;
  __syncthreads();
  if (tid < kPsWords) reinterpret_cast<uint32_t*>(ps)[tid] = reinterpret_cast<const uint32_t*>(&s_ps)[tid];
}

}  // namespace

}  // namespace gsr

using namespace gsr;

struct gsr_tracker {
  int P = 0, D = 0, M = 0, W = 0, H = 0;
  float tanx = 0.f, tany = 0.f;
  cudaStream_t stream = nullptr;
  cudaEvent_t caller_ev = nullptr;  // orders `stream` after the caller's stream (gsr_tracker_run)
  // borrowed device pointers
  const float *means3D = nullptr, *shs = nullptr, *colors = nullptr, *opac = nullptr, *scales = nullptr,
              *rots = nullptr, *cov3D = nullptr, *bg = nullptr, *gt_color = nullptr, *gt_depth = nullptr;
  float scale_modifier = 1.f;
  // owned device memory
  char* geom_buf = nullptr; char* img_buf = nullptr; char* bin_buf = nullptr;
  size_t bin_bytes = 0;
  uint32_t capacity = 0, longest_cap = 0;
  int* radii = nullptr;
  float *alpha = nullptr, *dL_dpix = nullptr, *dL_ddepth = nullptr, *loss_partials = nullptr;
  float* scratch = nullptr;     // acc [16P] + pose partials
  float* cam = nullptr;         // view[16] proj[16] campos[4] persp[16]
  PoseState* ps = nullptr;
  float* loss_hist = nullptr;   // [max_hist]
  int max_hist = 0;
  GeomState g; BinState b; ImgState img; Camera camera;
  // graph cache
  cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
  struct Key {
    const void* ptrs[11]; float scale_modifier; gsr_track_params params; uint32_t capacity, longest_cap;
    int opts[2]; int packed;
  } key;
  bool key_valid = false;
  int num_rendered = 0;
};

namespace {

void tracker_free(gsr_tracker* t) {
  if (!t) return;
  if (t->exec) cudaGraphExecDestroy(t->exec);
  if (t->graph) cudaGraphDestroy(t->graph);
  cudaFree(t->geom_buf); cudaFree(t->img_buf); cudaFree(t->bin_buf); cudaFree(t->radii);
  cudaFree(t->alpha); cudaFree(t->dL_dpix); cudaFree(t->dL_ddepth); cudaFree(t->loss_partials);
  cudaFree(t->scratch); cudaFree(t->cam); cudaFree(t->ps); cudaFree(t->loss_hist);
  if (t->caller_ev) cudaEventDestroy(t->caller_ev);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
}

int ensure_binning(gsr_tracker* t, uint32_t capacity, uint32_t longest_cap) {
  if (capacity <= t->capacity && longest_cap <= t->longest_cap) return GSR_OK;
  if (capacity < t->capacity) capacity = t->capacity;
  if (longest_cap < t->longest_cap) longest_cap = t->longest_cap;
  const size_t need = BinState::carve(t->b, nullptr, capacity, 0, false);
  if (need > t->bin_bytes) {
    GSR_CUDA_OK(cudaStreamSynchronize(t->stream));
    cudaFree(t->bin_buf);
    t->bin_buf = nullptr;
    GSR_CUDA_OK(cudaMalloc(&t->bin_buf, need));
    t->bin_bytes = need;
  }
  BinState::carve(t->b, t->bin_buf, capacity, 0, false);
  t->capacity = capacity;
  t->longest_cap = longest_cap;
  return GSR_OK;
}

// the front half of an iteration, shared by the probe and the captured graph
int enqueue_preprocess(gsr_tracker* t) {
  return launch_preprocess_fwd(t->P, t->D, t->M, t->means3D, t->scales, t->scale_modifier, t->rots,
                               t->opac, t->shs, t->cov3D, t->colors, t->camera, t->radii, t->g,
                               t->img.tile_count, false, false, t->stream);
}

int enqueue_iteration(gsr_tracker* t, const gsr_track_params& prm, int packed_entries) {
  cudaStream_t s = t->stream;
  int rc = enqueue_preprocess(t);
  if (rc != GSR_OK) return rc;
  rc = run_binning_static(t->P, t->camera, t->g, t->b, t->img, t->capacity, t->longest_cap, s);
  if (rc != GSR_OK) return rc;
  FusedLoss fl;
  fl.gt_color = t->gt_color; fl.gt_depth = t->gt_depth;
  fl.w_color = prm.w_color; fl.w_depth = prm.w_depth;
  fl.alpha_thresh = prm.alpha_thresh; fl.depth_mask = prm.use_depth_mask;
  fl.dL_dpix = t->dL_dpix; fl.dL_ddepth = t->dL_ddepth; fl.loss_partials = t->loss_partials;
  rc = launch_render_fwd_light_loss(t->camera, t->g, t->b, t->img, t->bg, nullptr, nullptr, nullptr,
                                    t->alpha, nullptr, fl, s);
  if (rc != GSR_OK) return rc;
  float* acc = t->scratch;
  float* partials = t->scratch + (size_t)t->P * kAccStride;
  // (clearing acc inside the pose-contraction kernel instead was measured slower: 8.8 -> 13.5 us
  // for that kernel against a 2.9 us memset node)
  GSR_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)t->P * kAccStride * sizeof(float), s));
  BlendGrads cot{t->dL_dpix, t->dL_ddepth, nullptr, nullptr};
  rc = launch_render_bwd(kLight, t->camera, t->g, t->b, t->img, t->bg, t->gt_depth, t->alpha, cot, acc,
                         t->P, packed_entries, /*pose_only=*/true, false, s);
  if (rc != GSR_OK) return rc;
  const float* persp = t->cam + 36;
  rc = launch_preprocess_bwd_partials(kLight, t->P, t->D, t->M, t->means3D, t->radii, t->camera, persp,
                                      t->g, acc, partials, /*clear_acc=*/false, s);
  if (rc != GSR_OK) return rc;
  const int nblocks = preprocess_bwd_blocks(t->P);
  UpdateParams up{prm.lr_rot, prm.lr_trans, prm.beta1, prm.beta2, prm.eps, t->max_hist};
  track_update_kernel<<<1, kUpdThreads, 0, s>>>(nblocks, partials, t->camera.grid_x * t->camera.grid_y,
                                                t->loss_partials, t->ps, persp, t->cam, t->cam + 16,
                                                t->cam + 32, t->loss_hist, t->img.tile_count,
                                                t->g.counters, up, cnt_stride());
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}

}  // namespace

extern "C" {

gsr_tracker* gsr_tracker_create(int P, int D, int M, int width, int height, float tan_fovx,
                                float tan_fovy, const float* perspec_matrix_host, int max_iterations) {
  set_error("%s", "");
  if (P <= 0 || width <= 0 || height <= 0 || !perspec_matrix_host || max_iterations <= 0) {
    set_error("gsr_tracker_create: bad arguments");
    return nullptr;
  }
  gsr_tracker* t = new (std::nothrow) gsr_tracker();
  if (!t) { set_error("gsr_tracker_create: out of host memory"); return nullptr; }
  t->P = P; t->D = D; t->M = M; t->W = width; t->H = height; t->tanx = tan_fovx; t->tany = tan_fovy;
  t->max_hist = max_iterations;
  bool ok = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&t->caller_ev, cudaEventDisableTiming) == cudaSuccess;
  const int HW = width * height;
  Camera& c = t->camera;
  c.tan_fovx = tan_fovx; c.tan_fovy = tan_fovy;
  c.focal_y = height / (2.0f * tan_fovy); c.focal_x = width / (2.0f * tan_fovx);
  c.W = width; c.H = height;
  c.grid_x = (width + kTileX - 1) / kTileX; c.grid_y = (height + kTileY - 1) / kTileY;
  const int tiles = c.grid_x * c.grid_y;
  const size_t geom_need = GeomState::carve(t->g, nullptr, P, 0);
  const size_t img_need = ImgState::carve(t->img, nullptr, HW, tiles, kLight);
  ok = ok && cudaMalloc(&t->geom_buf, geom_need) == cudaSuccess;
  ok = ok && cudaMalloc(&t->img_buf, img_need) == cudaSuccess;
  ok = ok && cudaMalloc(&t->radii, sizeof(int) * (size_t)P) == cudaSuccess;
  ok = ok && cudaMalloc(&t->alpha, sizeof(float) * (size_t)HW) == cudaSuccess;
  ok = ok && cudaMalloc(&t->dL_dpix, sizeof(float) * 3 * (size_t)HW) == cudaSuccess;
  ok = ok && cudaMalloc(&t->dL_ddepth, sizeof(float) * (size_t)HW) == cudaSuccess;
  ok = ok && cudaMalloc(&t->loss_partials, sizeof(float) * (size_t)tiles) == cudaSuccess;
  ok = ok && cudaMalloc(&t->scratch, sizeof(float) * gsr_backward_scratch_floats(P)) == cudaSuccess;
  ok = ok && cudaMalloc(&t->cam, sizeof(float) * 64) == cudaSuccess;
  ok = ok && cudaMalloc(&t->ps, sizeof(PoseState)) == cudaSuccess;
  ok = ok && cudaMalloc(&t->loss_hist, sizeof(float) * (size_t)max_iterations) == cudaSuccess;
  if (!ok) {
    set_error("gsr_tracker_create: CUDA allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    tracker_free(t);
    return nullptr;
  }
  GeomState::carve(t->g, t->geom_buf, P, 0);
  ImgState::carve(t->img, t->img_buf, HW, tiles, kLight);
  c.view = t->cam; c.proj = t->cam + 16; c.campos = t->cam + 32;
  cudaMemsetAsync(t->cam, 0, sizeof(float) * 64, t->stream);
  cudaMemcpyAsync(t->cam + 36, perspec_matrix_host, sizeof(float) * 16, cudaMemcpyHostToDevice, t->stream);
  cudaMemsetAsync(t->ps, 0, sizeof(PoseState), t->stream);
  cudaMemsetAsync(t->g.counters, 0, 8 * sizeof(uint32_t), t->stream);
  cudaMemsetAsync(t->img.tile_count, 0, sizeof(uint32_t) * (size_t)tiles * kCntStrideMax, t->stream);
  if (cudaStreamSynchronize(t->stream) != cudaSuccess) {
    set_error("gsr_tracker_create: initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    tracker_free(t);
    return nullptr;
  }
  return t;
}

void gsr_tracker_destroy(gsr_tracker* t) {
  if (t && t->stream) cudaStreamSynchronize(t->stream);
  tracker_free(t);
}

int gsr_tracker_set_scene(gsr_tracker* t, const float* means3D, const float* shs,
                          const float* colors_precomp, const float* opacities, const float* scales,
                          float scale_modifier, const float* rotations, const float* cov3D_precomp,
                          const float* background) {
  set_error("%s", "");
  if (!t || !means3D || !opacities || !background || (!shs && !colors_precomp) ||
      (!cov3D_precomp && (!scales || !rotations))) {
    set_error("gsr_tracker_set_scene: NULL required input");
    return GSR_E_INVALID;
  }
  t->means3D = means3D; t->shs = shs; t->colors = colors_precomp; t->opac = opacities;
  t->scales = scales; t->scale_modifier = scale_modifier; t->rots = rotations;
  t->cov3D = cov3D_precomp; t->bg = background;
  return GSR_OK;
}

int gsr_tracker_set_frame(gsr_tracker* t, const float* gt_color, const float* gt_depth) {
  set_error("%s", "");
  if (!t || !gt_color || !gt_depth) { set_error("gsr_tracker_set_frame: NULL input"); return GSR_E_INVALID; }
  t->gt_color = gt_color; t->gt_depth = gt_depth;
  return GSR_OK;
}

int gsr_tracker_set_pose(gsr_tracker* t, const float* quat_wxyz, const float* trans) {
  set_error("%s", "");
  if (!t || !quat_wxyz || !trans) { set_error("gsr_tracker_set_pose: NULL input"); return GSR_E_INVALID; }
  PoseState h;
  memset(&h, 0, sizeof(h));
  for (int k = 0; k < 4; ++k) h.q[k] = quat_wxyz[k];
  for (int k = 0; k < 3; ++k) h.t[k] = trans[k];
  GSR_CUDA_OK(cudaMemcpyAsync(t->ps, &h, sizeof(h), cudaMemcpyHostToDevice, t->stream));
  init_camera_kernel<<<1, 32, 0, t->stream>>>(t->ps, t->cam + 36, t->cam, t->cam + 16, t->cam + 32);
  GSR_LAUNCH_OK(false, t->stream);
  GSR_CUDA_OK(cudaStreamSynchronize(t->stream));  // `h` lives on this stack frame
  return GSR_OK;
}

int gsr_tracker_run(gsr_tracker* t, const gsr_track_params* params, int iterations,
                    float* loss_history, gsr_track_result* result, void* caller_stream) {
  set_error("%s", "");
  OptionsCall oc("gsr_tracker_run");
  if (!t || !params || iterations <= 0 || iterations > t->max_hist) {
    set_error("gsr_tracker_run: bad arguments (iterations must be in [1, max_iterations])");
    return GSR_E_INVALID;
  }
  if (!t->means3D || !t->gt_color) { set_error("gsr_tracker_run: scene / frame not set"); return GSR_E_INVALID; }
  cudaStream_t s = t->stream;
  // order the private stream after the caller's: the borrowed scene / frame tensors are usually the
  // output of asynchronous work on that stream (see the stream contract in gsr_b200.h)
  GSR_CUDA_OK(cudaEventRecord(t->caller_ev, (cudaStream_t)caller_stream));
  GSR_CUDA_OK(cudaStreamWaitEvent(s, t->caller_ev, 0));
  const int tiles = t->camera.grid_x * t->camera.grid_y;
  PoseState start;
  GSR_CUDA_OK(cudaMemcpyAsync(&start, t->ps, sizeof(start), cudaMemcpyDeviceToHost, s));
  // probe: one per-Gaussian forward + tile scan at the start pose sizes the binning buffer
  uint32_t h[4] = {0, 0, 0, 0};
  {
    GSR_CUDA_OK(cudaMemsetAsync(t->g.counters, 0, 8 * sizeof(uint32_t), s));
    GSR_CUDA_OK(cudaMemsetAsync(t->img.tile_count, 0, sizeof(uint32_t) * (size_t)tiles * kCntStrideMax, s));
    int rc = enqueue_preprocess(t);
    if (rc != GSR_OK) return rc;
    rc = probe_tile_counts(t->camera, t->g, t->img, s);
    if (rc != GSR_OK) return rc;
    GSR_CUDA_OK(cudaMemcpyAsync(h, t->g.counters, sizeof(h), cudaMemcpyDeviceToHost, s));
    GSR_CUDA_OK(cudaStreamSynchronize(s));
  }
  t->num_rendered = (int)h[0];
  // head room over the probe's counts ("track_headroom_pct", default +50 %; negative values are a
  // test hook that forces the overflow / retry path)
  const int pct = options().track_headroom_pct;
  uint32_t want_cap = (uint32_t)((double)h[0] * (100 + pct) / 100.0) + (pct >= 0 ? 4096u : 16u);
  uint32_t want_long = pct >= 0 ? (h[2] * 2 < 256u ? 256u : h[2] * 2) : (h[2] / 2 + 1);
  int retries = 0;
  for (;;) {
    int rc = ensure_binning(t, want_cap, want_long);
    if (rc != GSR_OK) return rc;
    const int packed_entries = (int)h[0];
    gsr_tracker::Key key;
    memset(&key, 0, sizeof(key));
    const void* ptrs[11] = {t->means3D, t->shs, t->colors, t->opac, t->scales, t->rots, t->cov3D, t->bg,
                            t->gt_color, t->gt_depth, t->bin_buf};
    memcpy(key.ptrs, ptrs, sizeof(ptrs));
    key.scale_modifier = t->scale_modifier; key.params = *params;
    key.capacity = t->capacity; key.longest_cap = t->longest_cap;
    key.opts[0] = options().tight_tiles; key.opts[1] = options().bwd_packed + 16 * cnt_stride();
    key.packed = (double)packed_entries >= 1.6 * (double)t->P;
    if (!t->key_valid || memcmp(&key, &t->key, sizeof(key)) != 0 || !t->exec) {
      if (t->exec) { cudaGraphExecDestroy(t->exec); t->exec = nullptr; }
      if (t->graph) { cudaGraphDestroy(t->graph); t->graph = nullptr; }
      const int saved_timing = options().stage_timing;
      options().stage_timing = 0;  // no event records inside a capture
      GSR_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
      rc = enqueue_iteration(t, *params, packed_entries);
      cudaGraph_t gr = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(s, &gr);
      options().stage_timing = saved_timing;
      if (rc != GSR_OK) { if (gr) cudaGraphDestroy(gr); return rc; }
      if (ce != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(ce)); return GSR_E_CUDA; }
      t->graph = gr;
      GSR_CUDA_OK(cudaGraphInstantiate(&t->exec, t->graph, 0));
      t->key = key;
      t->key_valid = true;
    }
    // reset the per-run device state: counters, tile counters, loss history cursor, overflow flag
    start.iter = 0;
    GSR_CUDA_OK(cudaMemcpyAsync(t->ps, &start, sizeof(start), cudaMemcpyHostToDevice, s));
    init_camera_kernel<<<1, 32, 0, s>>>(t->ps, t->cam + 36, t->cam, t->cam + 16, t->cam + 32);
    GSR_LAUNCH_OK(false, s);
    GSR_CUDA_OK(cudaMemsetAsync(t->g.counters, 0, 8 * sizeof(uint32_t), s));
    GSR_CUDA_OK(cudaMemsetAsync(t->img.tile_count, 0, sizeof(uint32_t) * (size_t)tiles * kCntStrideMax, s));
    for (int i = 0; i < iterations; ++i) GSR_CUDA_OK(cudaGraphLaunch(t->exec, s));
    uint32_t flag[4] = {0, 0, 0, 0};
    GSR_CUDA_OK(cudaMemcpyAsync(flag, t->g.counters, sizeof(flag), cudaMemcpyDeviceToHost, s));
    GSR_CUDA_OK(cudaStreamSynchronize(s));
    if (flag[3] == 0u) break;
    if (++retries > 4) { set_error("gsr_tracker_run: binning buffer overflow persisted"); return GSR_E_ALLOC; }
    want_cap = t->capacity * 2;
    want_long = t->longest_cap * 2 > 8192u ? 8192u : t->longest_cap * 2;
    if (t->longest_cap >= 8192u && want_cap <= t->capacity) {
      set_error("gsr_tracker_run: a tile list exceeds 8192 entries (not supported by the tracker)");
      return GSR_E_INVALID;
    }
  }
  PoseState fin;
  GSR_CUDA_OK(cudaMemcpyAsync(&fin, t->ps, sizeof(fin), cudaMemcpyDeviceToHost, s));
  if (loss_history)
    GSR_CUDA_OK(cudaMemcpyAsync(loss_history, t->loss_hist, sizeof(float) * (size_t)iterations,
                                cudaMemcpyDeviceToHost, s));
  GSR_CUDA_OK(cudaStreamSynchronize(s));
  if (result) {
    for (int k = 0; k < 4; ++k) result->q[k] = fin.q[k];
    for (int k = 0; k < 3; ++k) result->t[k] = fin.t[k];
    for (int k = 0; k < 16; ++k) result->last_dL_dview[k] = fin.dview[k];
    for (int k = 0; k < 7; ++k) result->last_grad[k] = fin.grad[k];
    for (int k = 0; k < 6; ++k) result->last_twist_grad[k] = fin.twist[k];
    result->iterations = iterations;
    result->num_rendered = t->num_rendered;
    result->retries = retries;
    result->kernels_per_iteration = 9;
  }
  return GSR_OK;
}

}  // extern "C"
