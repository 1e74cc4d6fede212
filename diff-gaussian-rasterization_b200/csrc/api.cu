// api.cu — extern "C" entry points of libgsr_b200.so (declared in include/gsr_b200.h) and the
// host-side orchestration of the stages.  Replaces CudaRasterizer::Rasterizer::{forward,backward,
// markVisible} (reference cuda_rasterizer/rasterizer_impl.cu:197-350,354-495 light;
// :349-500,504-666 full).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "gsr_common.cuh"

namespace gsr {

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
// Process-wide defaults (gsr_set_option).  Every entry point takes a thread-local SNAPSHOT of them
// when it starts (OptionsCall), and everything below reads the snapshot: a concurrent
// gsr_set_option from another thread can never change the switches in the middle of a call, and a
// value one thread's call is using is never written by another thread.
Options g_opts = {/*exact_ng=*/0, /*tight_tiles=*/1, /*stage_timing=*/0, /*tile_sort=*/1, /*bwd_packed=*/2, /*async_binning=*/1, /*track_headroom_pct=*/50, /*bulk_sh=*/1, /*cnt_stride=*/8, /*bwd_occ=*/0, /*fwd_packed=*/2, /*spec_render=*/1, /*early_acc_clear=*/1, /*pdl=*/1, /*exact_median=*/0};

// Stage timer: a pool of event pairs filled by StageScope and drained by gsr_stage_times().
struct StageTimer {
  struct Rec { int stage; int launches; cudaEvent_t a, b; };
  std::mutex mu;
  std::vector<Rec> pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> free_pairs;
  double ms[ST_COUNT] = {0};
  long long scopes[ST_COUNT] = {0};
  long long launches[ST_COUNT] = {0};

  int begin(int stage, int nlaunch, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(mu);
    if (pending.size() >= 16384) drain_locked();
    Rec r{stage, nlaunch, nullptr, nullptr};
    if (!free_pairs.empty()) {
      r.a = free_pairs.back().first; r.b = free_pairs.back().second;
      free_pairs.pop_back();
    } else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    cudaEventRecord(r.a, s);
    pending.push_back(r);
    return (int)pending.size() - 1;
  }
  void end(int slot, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(mu);
    if (slot >= 0 && slot < (int)pending.size()) cudaEventRecord(pending[slot].b, s);
  }
  void drain_locked() {
    for (Rec& r : pending) {
      float t = 0.f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
        ms[r.stage] += t;
        scopes[r.stage] += 1;
        launches[r.stage] += r.launches;
      } else {
        cudaGetLastError();
      }
      free_pairs.emplace_back(r.a, r.b);
    }
    pending.clear();
  }
};
StageTimer& stage_timer() {
  static StageTimer* t = new StageTimer();  // leaked on purpose: events outlive static destruction
  return *t;
}
const char* const kStageNames[ST_COUNT] = {"preprocess_fwd", "scan", "emit_keys", "radix_sort",
                                           "tile_ranges", "render_fwd", "render_bwd",
                                           "preprocess_bwd", "memset", "other"};
thread_local Options t_opts = g_opts;
}  // namespace

// Every stage is an NVTX range (visible in Nsight Systems / ncu --nvtx) and, with the
// "stage_timing" option, a CUDA-event pair on the launching stream.
StageScope::StageScope(int stage, cudaStream_t s, int nlaunch) : slot(-1), stream(s) {
  nvtxRangePushA(kStageNames[stage]);
  g_launches.fetch_add(nlaunch, std::memory_order_relaxed);
  if (t_opts.stage_timing) slot = stage_timer().begin(stage, nlaunch, s);
}
StageScope::~StageScope() {
  if (slot >= 0) stage_timer().end(slot, stream);
  nvtxRangePop();
}

Options& options() { return t_opts; }

void prefer_max_shared_once(const void* kernel) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;   // (kernel, device) pairs already configured
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
  std::lock_guard<std::mutex> lk(mu);
  for (const auto& d : done)
    if (d.first == kernel && d.second == dev) return;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess)
    cudaGetLastError();
  done.emplace_back(kernel, dev);
}

OptionsCall::OptionsCall(const char* entry_point) {
  t_opts = g_opts;
  nvtxRangePushA(entry_point);
}
OptionsCall::~OptionsCall() { nvtxRangePop(); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

struct FwdArgs {
  gsr_alloc_fn geom_alloc; void* geom_ctx;
  gsr_alloc_fn binning_alloc; void* binning_ctx;
  gsr_alloc_fn img_alloc; void* img_ctx;
  int P, D, M;
  const float* background; int width, height;
  const float* means3D; const float* shs; const float* colors_precomp; const float* opacities;
  const float* scales; float scale_modifier; const float* rotations; const float* cov3D_precomp;
  const float* viewmatrix; const float* projmatrix; const float* cam_pos;
  float tan_fovx, tan_fovy; int prefiltered;
  const float* gt_depth; int* radii; int debug; cudaStream_t stream;
  float* gau_unc; int* gau_px;   // -light: per-Gaussian statistics, zeroed by preprocess_fwd (NULL for -full)
};

int check_common(const char* who, int P, int width, int height, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* scales,
                 const float* rotations, const float* cov3D_precomp, const float* view,
                 const float* proj, const float* campos) {
  if (P < 0 || width <= 0 || height <= 0) {
    set_error("%s: bad sizes P=%d W=%d H=%d", who, P, width, height);
    return GSR_E_INVALID;
  }
  if (P == 0) return GSR_OK;
  if (!means3D || !view || !proj || !campos) {
    set_error("%s: means3D / viewmatrix / projmatrix / cam_pos must not be NULL", who);
    return GSR_E_INVALID;
  }
  if (!shs && !colors_precomp) {
    // the reference throws std::runtime_error("For non-RGB, provide precomputed Gaussian
    // colors!") only for NUM_CHANNELS != 3; with neither input it would read a null SH pointer.
    set_error("%s: provide SH coefficients or precomputed colours", who);
    return GSR_E_INVALID;
  }
  if (!cov3D_precomp && (!scales || !rotations)) {
    set_error("%s: provide scales+rotations or a precomputed 3D covariance", who);
    return GSR_E_INVALID;
  }
  if (!cov3D_precomp && (reinterpret_cast<uintptr_t>(rotations) & 15) != 0) {
    set_error("%s: rotations must be 16-byte aligned (read with 128-bit loads)", who);
    return GSR_E_INVALID;
  }
  return GSR_OK;
}

// 128-bit stores of the per-Gaussian backward
int check_grad_alignment(const char* who, const float* rotations, const float* cov3D_precomp,
                         const float* dL_dconic, const float* dL_drot) {
  if ((reinterpret_cast<uintptr_t>(dL_dconic) & 15) != 0 || (reinterpret_cast<uintptr_t>(dL_drot) & 15) != 0 ||
      (!cov3D_precomp && (reinterpret_cast<uintptr_t>(rotations) & 15) != 0)) {
    set_error("%s: rotations, dL_dconic and dL_drot must be 16-byte aligned (128-bit accesses)", who);
    return GSR_E_INVALID;
  }
  return GSR_OK;
}

Camera make_camera(const float* view, const float* proj, const float* campos, float tanx,
                   float tany, int W, int H) {
  Camera c;
  c.view = view; c.proj = proj; c.campos = campos;
  c.tan_fovx = tanx; c.tan_fovy = tany;
  c.focal_y = H / (2.0f * tany);
  c.focal_x = W / (2.0f * tanx);
  c.W = W; c.H = H;
  c.grid_x = (W + kTileX - 1) / kTileX;
  c.grid_y = (H + kTileY - 1) / kTileY;
  return c;
}

// The accumulator clear of the coming backward is forked off right in front of the forward blend kernel: the
// memset then runs underneath that kernel (issue-bound, HBM idle) instead of next to the HBM-bound per-Gaussian
// kernel, where it gained nothing (measured).  The geometry buffer's base address is the key the backward looks up.
void begin_acc_clear_once(bool& begun, const GeomState& g, int P, cudaStream_t s) {
  if (begun || options().early_acc_clear == 0) return;
  begun = true;
  acc_clear_begin(g.rec, g.acc, ((size_t)P * kAccStride + 16) * sizeof(float), s);
}

// shared front half of both forwards: allocate state, preprocess, bin
int forward_front(const FwdArgs& a, int variant, Camera& cam, GeomState& g, BinState& b,
                  ImgState& img, int* num_rendered, SpecRender* spec = nullptr) {
  cam = make_camera(a.viewmatrix, a.projmatrix, a.cam_pos, a.tan_fovx, a.tan_fovy, a.width, a.height);
  const int HW = a.width * a.height;
  const int tiles = cam.grid_x * cam.grid_y;

  const size_t scan_bytes = scan_temp_bytes(a.P);
  const size_t geom_need = GeomState::carve(g, nullptr, a.P, scan_bytes);
  char* geom_chunk = a.geom_alloc(a.geom_ctx, geom_need);
  if (!geom_chunk) { set_error("geometry allocator returned NULL for %zu bytes", geom_need); return GSR_E_ALLOC; }
  GeomState::carve(g, geom_chunk, a.P, scan_bytes);

  const size_t img_need = ImgState::carve(img, nullptr, HW, tiles, variant);
  char* img_chunk = a.img_alloc(a.img_ctx, img_need);
  if (!img_chunk) { set_error("image allocator returned NULL for %zu bytes", img_need); return GSR_E_ALLOC; }
  ImgState::carve(img, img_chunk, HW, tiles, variant);

  const bool tile_local = options().tile_sort != 0;
  {
    // the only zero-fill of the forward: the per-tile entry counters preprocess_fwd adds to.  (The
    // counter block is initialised by scan_tiles_kernel, -light's per-Gaussian statistics by
    // preprocess_fwd itself; the radix path still clears the counter block here.)
    StageScope st(ST_MEMSET, a.stream, 1);
    if (tile_local)
      GSR_CUDA_OK(cudaMemsetAsync(img.tile_count, 0, sizeof(uint32_t) * (size_t)tiles * cnt_stride(), a.stream));
    else
      GSR_CUDA_OK(cudaMemsetAsync(g.counters, 0, 8 * sizeof(uint32_t), a.stream));
  }
  int rc = launch_preprocess_fwd(a.P, a.D, a.M, a.means3D, a.scales, a.scale_modifier, a.rotations,
                                 a.opacities, a.shs, a.cov3D_precomp, a.colors_precomp, cam,
                                 a.radii, g, tile_local ? img.tile_count : nullptr,
                                 a.prefiltered != 0, a.debug != 0, a.stream, a.gau_unc, a.gau_px);
  if (rc != GSR_OK) return rc;
  return run_binning(a.P, cam, a.radii, g, a.binning_alloc, a.binning_ctx, b, img, num_rendered,
                     a.debug != 0, a.stream, options().spec_render != 0 ? spec : nullptr);
}

int rederive(int variant, int P, int R, int width, int height, char* geom_buffer,
             char* binning_buffer, char* img_buffer, GeomState& g, BinState& b, ImgState& img,
             const Camera& cam) {
  if (!geom_buffer || !img_buffer || (R > 0 && !binning_buffer)) {
    set_error("backward: NULL state buffer");
    return GSR_E_INVALID;
  }
  GeomState::carve(g, geom_buffer, P, 0);
  BinState::carve(b, binning_buffer, (size_t)(R > 0 ? R : 0), 0, false);
  ImgState::carve(img, img_buffer, width * height, cam.grid_x * cam.grid_y, variant);
  return GSR_OK;
}

}  // namespace
}  // namespace gsr

using namespace gsr;

namespace gsr {
namespace {
// masked colour gradient of one view, straight from the blend backward's accumulator lines:
// dL/dRGB with the channels the forward clamped at 0 zeroed, zeros for culled Gaussians (what
// preprocess_bwd writes into GaussGradOut::dL_dcolor_masked).  12 bytes out per Gaussian.
__global__ void __launch_bounds__(256)
masked_color_kernel(int P, const int* __restrict__ radii, const unsigned char* __restrict__ clamped,
                    const float* __restrict__ acc, float* __restrict__ out) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= P) return;
  float r = 0.f, g = 0.f, b = 0.f;
  if (radii[idx] > 0) {
    const float* line = acc + (size_t)idx * kAccStride;
    const unsigned char cb = clamped[idx];
    r = (cb & 1) ? 0.f : line[ACC_R];
    g = (cb & 2) ? 0.f : line[ACC_G];
    b = (cb & 4) ? 0.f : line[ACC_B];
  }
  out[3 * (size_t)idx] = r; out[3 * (size_t)idx + 1] = g; out[3 * (size_t)idx + 2] = b;
}

// extras -> GaussGradOut; with masked_color_ready_event: early kernel + event, the per-Gaussian kernel skips it
int apply_extras(const gsr_backward_extras* extras, GaussGradOut& out, int P, const int* radii, const GeomState& g,
                 const float* acc, bool have_sh, cudaStream_t s) {
  if (extras == nullptr) return GSR_OK;
  out.dL_dcolor_masked = extras->dL_dcolor_masked;
  if (extras->skip_sh_grad) out.dL_dsh = nullptr;
  if (extras->densify_grad_accum != nullptr && extras->densify_denom != nullptr) {
    out.densify_grad_accum = extras->densify_grad_accum;
    out.densify_denom = extras->densify_denom;
  }
  out.max_radii2D = extras->max_radii2D;
  if (extras->masked_color_ready_event != nullptr && extras->dL_dcolor_masked != nullptr && have_sh) {
    {
      StageScope st(ST_PRE_BWD, s);
      masked_color_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, radii, g.clamped, acc, extras->dL_dcolor_masked);
      GSR_LAUNCH_OK(false, s);
    }
    GSR_CUDA_OK(cudaEventRecord((cudaEvent_t)extras->masked_color_ready_event, s));
    out.dL_dcolor_masked = nullptr;
  }
  return GSR_OK;
}
}  // namespace
}  // namespace gsr

extern "C" {

int gsr_abi_version(void) { return GSR_B200_ABI_VERSION; }

const char* gsr_last_error(void) { return g_err; }

size_t gsr_backward_scratch_floats(int P) {
  const size_t p = P > 0 ? (size_t)P : 0;
  return p * kAccStride + ((p + 127) / 128) * 12 + 64;
}

static int* option_slot(const char* key) {
  if (!key) return nullptr;
  if (!strcmp(key, "exact_ng")) return &g_opts.exact_ng;
  if (!strcmp(key, "tight_tiles")) return &g_opts.tight_tiles;
  if (!strcmp(key, "stage_timing")) return &g_opts.stage_timing;
  if (!strcmp(key, "tile_sort")) return &g_opts.tile_sort;
  if (!strcmp(key, "bwd_packed")) return &g_opts.bwd_packed;
  if (!strcmp(key, "async_binning")) return &g_opts.async_binning;
  if (!strcmp(key, "track_headroom_pct")) return &g_opts.track_headroom_pct;
  if (!strcmp(key, "bulk_sh")) return &g_opts.bulk_sh;
  if (!strcmp(key, "cnt_stride")) return &g_opts.cnt_stride;
  if (!strcmp(key, "bwd_occ")) return &g_opts.bwd_occ;
  if (!strcmp(key, "fwd_packed")) return &g_opts.fwd_packed;
  if (!strcmp(key, "spec_render")) return &g_opts.spec_render;
  if (!strcmp(key, "early_acc_clear")) return &g_opts.early_acc_clear;
  if (!strcmp(key, "pdl")) return &g_opts.pdl;
  if (!strcmp(key, "exact_median")) return &g_opts.exact_median;
  return nullptr;
}

int gsr_set_option(const char* key, int value) {
  int* s = option_slot(key);
  if (!s) { set_error("unknown option '%s'", key ? key : "(null)"); return GSR_E_INVALID; }
  const int prev = *s;
  *s = value;
  return prev;
}

int gsr_get_option(const char* key) {
  int* s = option_slot(key);
  if (!s) { set_error("unknown option '%s'", key ? key : "(null)"); return GSR_E_INVALID; }
  return *s;
}

long long gsr_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int gsr_stage_times(double* ms, long long* scopes, long long* launches, int reset) {
  StageTimer& t = stage_timer();
  std::lock_guard<std::mutex> lk(t.mu);
  t.drain_locked();
  for (int i = 0; i < ST_COUNT; ++i) {
    if (ms) ms[i] = t.ms[i];
    if (scopes) scopes[i] = t.scopes[i];
    if (launches) launches[i] = t.launches[i];
    if (reset) { t.ms[i] = 0; t.scopes[i] = 0; t.launches[i] = 0; }
  }
  return GSR_OK;
}

const char* gsr_stage_name(int stage) {
  return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : "";
}

int gsr_light_forward(
    gsr_alloc_fn geom_alloc, void* geom_ctx, gsr_alloc_fn binning_alloc, void* binning_ctx,
    gsr_alloc_fn img_alloc, void* img_ctx, int P, int D, int M, const float* background,
    int width, int height, const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
    float* out_depth, float* out_median_depth, float* out_alpha, const float* gt_depth,
    float* out_depth_var, float* gau_uncertainty, int* gau_related_pixels, int* radii, int debug,
    void* stream, int* num_rendered) {
  g_err[0] = 0;
  OptionsCall oc("gsr_light_forward");
  int rc = check_common("gsr_light_forward", P, width, height, means3D, shs, colors_precomp, scales,
                        rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos);
  if (rc != GSR_OK) return rc;
  if (!out_color || !out_depth || !out_median_depth || !out_alpha || !out_depth_var ||
      !num_rendered || !geom_alloc || !binning_alloc || !img_alloc ||
      (P > 0 && (!radii || !gau_uncertainty || !gau_related_pixels || !opacities || !gt_depth || !background))) {
    set_error("gsr_light_forward: NULL output / allocator / required input");
    return GSR_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t HW = (size_t)width * height;
  *num_rendered = 0;
  if (P == 0) {  // the reference short-circuits and returns zero-filled outputs
    GSR_CUDA_OK(cudaMemsetAsync(out_color, 0, 3 * HW * sizeof(float), s));
    GSR_CUDA_OK(cudaMemsetAsync(out_depth, 0, HW * sizeof(float), s));
    GSR_CUDA_OK(cudaMemsetAsync(out_median_depth, 0, HW * sizeof(float), s));
    GSR_CUDA_OK(cudaMemsetAsync(out_alpha, 0, HW * sizeof(float), s));
    GSR_CUDA_OK(cudaMemsetAsync(out_depth_var, 0, HW * sizeof(float), s));
    return GSR_OK;
  }
  FwdArgs a{geom_alloc, geom_ctx, binning_alloc, binning_ctx, img_alloc, img_ctx, P, D, M,
            background, width, height, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
            tan_fovy, prefiltered, gt_depth, radii, debug, s, gau_uncertainty, gau_related_pixels};
  Camera cam; GeomState g; BinState b; ImgState img;
  SpecRender spec;
  bool acc_begun = false;
  spec.launch = [&](const BinState& bs) {
    begin_acc_clear_once(acc_begun, g, P, s);
    return launch_render_fwd_light(cam, g, bs, img, background, gt_depth, out_color, out_depth, out_median_depth,
                                   out_alpha, out_depth_var, gau_uncertainty, gau_related_pixels, debug != 0, s);
  };
  spec.reset = [&]() {   // the per-Gaussian statistics are accumulated with atomics by the blend
    StageScope st(ST_MEMSET, s, 2);
    GSR_CUDA_OK(cudaMemsetAsync(gau_uncertainty, 0, sizeof(float) * (size_t)P, s));
    GSR_CUDA_OK(cudaMemsetAsync(gau_related_pixels, 0, sizeof(int) * (size_t)P, s));
    return (int)GSR_OK;
  };
  rc = forward_front(a, kLight, cam, g, b, img, num_rendered, &spec);
  if (rc != GSR_OK) return rc;
  if (!spec.done) rc = spec.launch(b);
  if (acc_begun) acc_clear_rejoin(g.rec, s);
  return rc;
}

int gsr_full_forward(
    gsr_alloc_fn geom_alloc, void* geom_ctx, gsr_alloc_fn binning_alloc, void* binning_ctx,
    gsr_alloc_fn img_alloc, void* img_ctx, int P, int D, int M, const float* background,
    int width, int height, const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
    float* out_depth, float* out_uncertainty, int* radii, const float* gt_depth, void* stream,
    int* num_rendered, int* num_related) {
  g_err[0] = 0;
  OptionsCall oc("gsr_full_forward");
  (void)gt_depth;  // loaded but unused by the reference's forward (forward.cu:313-317)
  int rc = check_common("gsr_full_forward", P, width, height, means3D, shs, colors_precomp, scales,
                        rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos);
  if (rc != GSR_OK) return rc;
  if (!out_color || !out_depth || !out_uncertainty || !num_rendered || !num_related ||
      !geom_alloc || !binning_alloc || !img_alloc || (P > 0 && (!radii || !opacities || !background))) {
    set_error("gsr_full_forward: NULL output / allocator / required input");
    return GSR_E_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t HW = (size_t)width * height;
  *num_rendered = 0;
  *num_related = 0;
  if (P == 0) {
    GSR_CUDA_OK(cudaMemsetAsync(out_color, 0, 3 * HW * sizeof(float), s));
    GSR_CUDA_OK(cudaMemsetAsync(out_depth, 0, HW * sizeof(float), s));
    GSR_CUDA_OK(cudaMemsetAsync(out_uncertainty, 0, HW * sizeof(float), s));
    return GSR_OK;
  }
  FwdArgs a{geom_alloc, geom_ctx, binning_alloc, binning_ctx, img_alloc, img_ctx, P, D, M,
            background, width, height, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx,
            tan_fovy, prefiltered, gt_depth, radii, 0, s, nullptr, nullptr};
  Camera cam; GeomState g; BinState b; ImgState img;
  const bool count = options().exact_ng != 0;
  SpecRender spec;
  bool acc_begun = false;
  spec.launch = [&](const BinState& bs) {
    begin_acc_clear_once(acc_begun, g, P, s);
    return launch_render_fwd_full(cam, g, bs, img, background, out_color, out_depth, out_uncertainty, count, false, s);
  };
  spec.reset = []() { return (int)GSR_OK; };   // num_related is reset by the rescan of the redone binning
  rc = forward_front(a, kFull, cam, g, b, img, num_rendered, &spec);
  if (rc != GSR_OK) return rc;
  if (!spec.done) {
    rc = spec.launch(b);
    if (rc != GSR_OK) return rc;
  }
  if (acc_begun) acc_clear_rejoin(g.rec, s);
  if (count) {
    uint32_t ng = 0;
    GSR_CUDA_OK(cudaMemcpyAsync(&ng, g.counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    GSR_CUDA_OK(cudaStreamSynchronize(s));
    *num_related = (int)ng;
  }
  return GSR_OK;
}

int gsr_light_backward(
    int P, int D, int M, int R, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* alphas,
    const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* cam_pos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
    char* binning_buffer, char* img_buffer, const float* dL_dpix, const float* dL_dpix_depth,
    const float* dL_dpix_median_depth, const float* dL_dpix_depth_var, float* dL_dmean2D,
    float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth, float* dL_dmean3D,
    float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug,
    const float* perspec_matrix, float* dL_dview, const float* gt_depth, int track_off,
    int map_off, float* scratch, void* stream, const gsr_backward_extras* extras) {
  g_err[0] = 0;
  OptionsCall oc("gsr_light_backward");
  (void)colors_precomp;  // colours come from the packed record written by the forward
  cudaStream_t s = (cudaStream_t)stream;
  if (!dL_dview) { set_error("gsr_light_backward: dL_dview is NULL"); return GSR_E_INVALID; }
  if (P <= 0) {
    GSR_CUDA_OK(cudaMemsetAsync(dL_dview, 0, 16 * sizeof(float), s));
    return P == 0 ? GSR_OK : GSR_E_INVALID;
  }
  if (!means3D || !radii || !alphas || !dL_dpix || !dL_dpix_depth || !dL_dpix_median_depth ||
      !dL_dpix_depth_var || !gt_depth || !perspec_matrix || !scratch || !background ||
      !viewmatrix || !projmatrix || !cam_pos) {
    set_error("gsr_light_backward: NULL required input");
    return GSR_E_INVALID;
  }
  Camera cam = make_camera(viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, width, height);
  GeomState g; BinState b; ImgState img;
  int rc = check_grad_alignment("gsr_light_backward", rotations, cov3D_precomp, dL_dconic, dL_drot);
  if (rc != GSR_OK) return rc;
  rc = rederive(kLight, P, R, width, height, geom_buffer, binning_buffer, img_buffer, g, b, img, cam);
  if (rc != GSR_OK) return rc;
  // scratch: acc [16 P] | last-block counter (one 64-byte line) | pose partials; one memset clears
  // the accumulator lines and the counter
  float* acc = scratch;
  unsigned int* done_counter = reinterpret_cast<unsigned int*>(scratch + (size_t)P * kAccStride);
  float* partials = scratch + (size_t)P * kAccStride + 16;
  if (options().early_acc_clear != 0 && acc_clear_join(g.rec, s)) {
    // cleared by the forward on a side stream (GeomState::acc): no memset in front of the blend kernel
    acc = g.acc;
    done_counter = reinterpret_cast<unsigned int*>(g.acc + (size_t)P * kAccStride);
  } else {
    StageScope st(ST_MEMSET, s);
    GSR_CUDA_OK(cudaMemsetAsync(acc, 0, ((size_t)P * kAccStride + 16) * sizeof(float), s));
  }
  const bool want_gauss = !map_off, want_pose = !track_off;
  if (want_gauss || want_pose) {
    BlendGrads cot{dL_dpix, dL_dpix_depth, dL_dpix_median_depth, dL_dpix_depth_var};
    rc = launch_render_bwd(kLight, cam, g, b, img, background, gt_depth, alphas, cot, acc, P, R,
                           /*pose_only=*/!want_gauss, debug != 0, s);
    if (rc != GSR_OK) return rc;
  }
  GaussGradOut out{dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D,
                   dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dview, nullptr, nullptr, nullptr, nullptr};
  rc = apply_extras(extras, out, P, radii, g, acc, shs != nullptr && want_gauss, s);
  if (rc != GSR_OK) return rc;
  return launch_preprocess_bwd(kLight, P, D, M, means3D, radii, shs, scales, rotations,
                               scale_modifier, cov3D_precomp, cam, perspec_matrix, g, acc, partials,
                               out, want_gauss, want_pose, debug != 0, s, done_counter);
}

int gsr_full_backward(
    int P, int D, int M, int R, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
    float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* img_buffer,
    const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dpix_uncertainty,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_ddepth,
    float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
    const float* perspec_matrix, float* dL_dview, const float* gt_depth, float* scratch,
    void* stream, const gsr_backward_extras* extras) {
  g_err[0] = 0;
  OptionsCall oc("gsr_full_backward");
  (void)colors_precomp;
  cudaStream_t s = (cudaStream_t)stream;
  if (!dL_dview) { set_error("gsr_full_backward: dL_dview is NULL"); return GSR_E_INVALID; }
  if (P <= 0) {
    GSR_CUDA_OK(cudaMemsetAsync(dL_dview, 0, 16 * sizeof(float), s));
    return P == 0 ? GSR_OK : GSR_E_INVALID;
  }
  if (!means3D || !radii || !dL_dpix || !dL_dpix_depth || !dL_dpix_uncertainty || !gt_depth ||
      !perspec_matrix || !scratch || !background || !viewmatrix || !projmatrix || !cam_pos) {
    set_error("gsr_full_backward: NULL required input");
    return GSR_E_INVALID;
  }
  Camera cam = make_camera(viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, width, height);
  GeomState g; BinState b; ImgState img;
  int rc = check_grad_alignment("gsr_full_backward", rotations, cov3D_precomp, dL_dconic, dL_drot);
  if (rc != GSR_OK) return rc;
  rc = rederive(kFull, P, R, width, height, geom_buffer, binning_buffer, img_buffer, g, b, img, cam);
  if (rc != GSR_OK) return rc;
  // scratch: acc [16 P] | last-block counter (one 64-byte line) | pose partials; one memset clears
  // the accumulator lines and the counter
  float* acc = scratch;
  unsigned int* done_counter = reinterpret_cast<unsigned int*>(scratch + (size_t)P * kAccStride);
  float* partials = scratch + (size_t)P * kAccStride + 16;
  if (options().early_acc_clear != 0 && acc_clear_join(g.rec, s)) {
    // cleared by the forward on a side stream (GeomState::acc): no memset in front of the blend kernel
    acc = g.acc;
    done_counter = reinterpret_cast<unsigned int*>(g.acc + (size_t)P * kAccStride);
  } else {
    StageScope st(ST_MEMSET, s);
    GSR_CUDA_OK(cudaMemsetAsync(acc, 0, ((size_t)P * kAccStride + 16) * sizeof(float), s));
  }
  BlendGrads cot{dL_dpix, dL_dpix_depth, nullptr, dL_dpix_uncertainty};
  rc = launch_render_bwd(kFull, cam, g, b, img, background, gt_depth, nullptr, cot, acc, P, R, false, false, s);
  if (rc != GSR_OK) return rc;
  GaussGradOut out{dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_ddepth, dL_dmean3D,
                   dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dview, nullptr, nullptr, nullptr, nullptr};
  rc = apply_extras(extras, out, P, radii, g, acc, shs != nullptr, s);
  if (rc != GSR_OK) return rc;
  return launch_preprocess_bwd(kFull, P, D, M, means3D, radii, shs, scales, rotations,
                               scale_modifier, cov3D_precomp, cam, perspec_matrix, g, acc, partials,
                               out, true, true, false, s, done_counter);
}

}  // extern "C"

namespace gsr {
namespace {
__global__ void decode_geometry_kernel(int P, const float4* __restrict__ rec,
                                       const float* __restrict__ cov3D_in,
                                       const uint32_t* __restrict__ tiles_in,
                                       const unsigned char* __restrict__ clamped_in, float* depths,
                                       float* means2D, float* conic_opacity, float* rgb,
                                       float* cov3D, uint32_t* tiles, unsigned char* clamped) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 r0 = rec[3 * (size_t)i], r1 = rec[3 * (size_t)i + 1], r2 = rec[3 * (size_t)i + 2];
  if (depths) depths[i] = r1.w;
  if (means2D) { means2D[2 * i] = r0.x; means2D[2 * i + 1] = r0.y; }
  if (conic_opacity) {
    conic_opacity[4 * i] = r0.z; conic_opacity[4 * i + 1] = r0.w;
    conic_opacity[4 * i + 2] = r1.x; conic_opacity[4 * i + 3] = r1.y;
  }
  if (rgb) { rgb[3 * i] = r2.x; rgb[3 * i + 1] = r2.y; rgb[3 * i + 2] = r2.z; }
  if (cov3D) for (int k = 0; k < 6; ++k) cov3D[6 * (size_t)i + k] = cov3D_in[6 * (size_t)i + k];
  if (tiles) tiles[i] = tiles_in[i];
  if (clamped) {
    const unsigned char c = clamped_in[i];
    clamped[3 * i] = c & 1; clamped[3 * i + 1] = (c >> 1) & 1; clamped[3 * i + 2] = (c >> 2) & 1;
  }
}
}  // namespace
}  // namespace gsr

extern "C" int gsr_decode_geometry(const char* geom_buffer, int P, float* depths, float* means2D,
                                   float* conic_opacity, float* rgb, float* cov3D,
                                   uint32_t* tiles_touched, unsigned char* clamped, void* stream) {
  if (!geom_buffer || P < 0) { set_error("gsr_decode_geometry: bad arguments"); return GSR_E_INVALID; }
  if (P == 0) return GSR_OK;
  GeomState g;
  GeomState::carve(g, const_cast<char*>(geom_buffer), P, 0);
  cudaStream_t s = (cudaStream_t)stream;
  gsr::decode_geometry_kernel<<<(P + 255) / 256, 256, 0, s>>>(
      P, g.rec, g.cov3D, g.tiles_touched, g.clamped, depths, means2D, conic_opacity, rgb, cov3D,
      tiles_touched, clamped);
  GSR_LAUNCH_OK(false, s);
  return GSR_OK;
}
