// binning.cu — tile binning: prefix sum of tiles_touched, (tile | depth) key emission,
// device radix sort, per-tile [start, end) ranges.
//
// Reference: cuda_rasterizer/rasterizer_impl.cu — InclusiveSum + blocking D2H (:283-287 light,
// :430-435 full), duplicateWithKeys (:71-112), getHigherMsb (:36-51), SortPairs over bits
// [0, 32+bit) (:309-314 / :457-462), identifyTileRanges (:117-139).
// Ordering contract (what the blend kernels observe): entries of a tile are ordered by the
// raw IEEE bits of view-space depth, ties by ascending Gaussian index (the reference emits
// duplicates in index order and CUB's radix sort is stable).
//
// Roofline: HBM. Algorithmic bytes: scan 8*P; emit 20*P + 12*N; sort (24*passes + 8)*N;
// ranges 8*N + 8*tiles.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gsr_common.cuh"

namespace gsr {

size_t scan_temp_bytes(int P) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
  return bytes;
}

size_t sort_temp_bytes(size_t N, int end_bit) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (uint32_t*)nullptr, (uint32_t*)nullptr, (int)N, 0, end_bit);
  return bytes;
}

size_t GeomState::carve(GeomState& s, char* base, int P, size_t scan_bytes) {
  Carver c(base);
  s.rec = c.take<float4>(3 * (size_t)P);
  s.cov3D = c.take<float>(6 * (size_t)P);
  s.clamped = c.take<unsigned char>((size_t)P);
  s.tiles_touched = c.take<uint32_t>((size_t)P);
  s.rect = c.take<uint2>((size_t)P);
  s.offsets = c.take<uint32_t>((size_t)P);
  s.counters = c.take<uint32_t>(8);
  s.scan_temp = c.take<char>(scan_bytes);
  s.scan_bytes = scan_bytes;
  return c.used + 256;
}

size_t BinState::carve(BinState& s, char* base, size_t N, size_t sort_bytes) {
  Carver c(base);
  s.keys_unsorted = c.take<uint64_t>(N);
  s.keys = c.take<uint64_t>(N);
  s.vals_unsorted = c.take<uint32_t>(N);
  s.vals = c.take<uint32_t>(N);
  s.sort_temp = c.take<char>(sort_bytes);
  s.sort_bytes = sort_bytes;
  return c.used + 256;
}

size_t ImgState::carve(ImgState& s, char* base, int HW, int tiles, int variant) {
  Carver c(base);
  s.ranges = c.take<uint2>((size_t)tiles);
  s.tile_last = c.take<uint32_t>((size_t)tiles);
  s.n_contrib = c.take<uint32_t>((size_t)HW);
  if (variant == kFull) {
    s.final_T = c.take<float>((size_t)HW);
    s.first_contrib = c.take<uint32_t>((size_t)HW);
  } else {
    s.final_T = nullptr;
    s.first_contrib = nullptr;
  }
  return c.used + 256;
}

namespace {

// smallest b with (n >> b) == 0, i.e. the number of bits needed for tile ids < n
// (same value the reference's getHigherMsb produces)
int bits_for(uint32_t n) {
  int b = 0;
  while (b < 32 && (n >> b)) ++b;
  return b;
}

__global__ void emit_keys_kernel(int P, const float4* __restrict__ rec,
                                 const uint32_t* __restrict__ offsets,
                                 const uint32_t* __restrict__ tiles_touched,
                                 const uint2* __restrict__ rects, int grid_x,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  if (tiles_touched[idx] == 0) return;
  uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
  const uint2 rc = rects[idx];  // the rectangle preprocess_fwd counted
  const uint2 rmin = make_uint2(rc.x & 0xFFFFu, rc.y & 0xFFFFu);
  const uint2 rmax = make_uint2(rc.x >> 16, rc.y >> 16);
  const uint64_t depth_bits = (uint64_t)__float_as_uint(rec[3 * (size_t)idx + 1].w);
  for (uint32_t y = rmin.y; y < rmax.y; ++y) {
    for (uint32_t x = rmin.x; x < rmax.x; ++x) {
      const uint64_t key = ((uint64_t)(y * (uint32_t)grid_x + x) << 32) | depth_bits;
      keys[off] = key;
      vals[off] = (uint32_t)idx;
      ++off;
    }
  }
}

__global__ void tile_ranges_kernel(int L, const uint64_t* __restrict__ keys,
                                   uint2* __restrict__ ranges) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> 32);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = (uint32_t)idx;
      ranges[cur].x = (uint32_t)idx;
    }
  }
  if (idx == L - 1) ranges[cur].y = (uint32_t)L;
}

}  // namespace

int run_binning(int P, const Camera& cam, const int* /*radii*/, GeomState& g, gsr_alloc_fn alloc,
                void* alloc_ctx, BinState& b, ImgState& img, int* num_rendered, bool debug,
                cudaStream_t stream) {
  const int tiles = cam.grid_x * cam.grid_y;
  {
    StageScope st(ST_SCAN, stream);
    GSR_CUDA_OK(cub::DeviceScan::InclusiveSum(g.scan_temp, g.scan_bytes, g.tiles_touched,
                                              g.offsets, P, stream));
    GSR_LAUNCH_OK(debug, stream);
  }

  // The one host<->device synchronisation of the forward: the duplicate count sizes the
  // binning buffer (the reference blocks in the same place).
  uint32_t N = 0;
  GSR_CUDA_OK(cudaMemcpyAsync(&N, g.offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                              stream));
  GSR_CUDA_OK(cudaStreamSynchronize(stream));
  *num_rendered = (int)N;

  const int end_bit = 32 + bits_for((uint32_t)tiles);
  const size_t sort_bytes = sort_temp_bytes(N, end_bit);
  const size_t need = BinState::carve(b, nullptr, N, sort_bytes);
  char* chunk = alloc(alloc_ctx, need);
  if (chunk == nullptr && need > 0) {
    set_error("binning allocator returned NULL for %zu bytes", need);
    return GSR_E_ALLOC;
  }
  BinState::carve(b, chunk, N, sort_bytes);

  {
    StageScope st(ST_MEMSET, stream);
    GSR_CUDA_OK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, stream));
  }
  if (N == 0) return GSR_OK;

  {
    StageScope st(ST_EMIT, stream);
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.rec, g.offsets, g.tiles_touched,
                                                          g.rect, cam.grid_x, b.keys_unsorted,
                                                          b.vals_unsorted);
    GSR_LAUNCH_OK(debug, stream);
  }
  {
    StageScope st(ST_SORT, stream);
    GSR_CUDA_OK(cub::DeviceRadixSort::SortPairs(b.sort_temp, b.sort_bytes, b.keys_unsorted, b.keys,
                                                b.vals_unsorted, b.vals, (int)N, 0, end_bit,
                                                stream));
    GSR_LAUNCH_OK(debug, stream);
  }
  {
    StageScope st(ST_RANGES, stream);
    tile_ranges_kernel<<<(N + 255) / 256, 256, 0, stream>>>((int)N, b.keys, img.ranges);
    GSR_LAUNCH_OK(debug, stream);
  }
  return GSR_OK;
}

}  // namespace gsr
