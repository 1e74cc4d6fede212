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
#include <atomic>
#include <mutex>

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
  s.acc = c.take<float>((size_t)P * kAccStride + 16);   // before scan_temp: the backward re-derives with scan_bytes = 0
  s.scan_temp = c.take<char>(scan_bytes);
  s.scan_bytes = scan_bytes;
  return c.used + 256;
}

size_t BinState::carve(BinState& s, char* base, size_t N, size_t sort_bytes, bool radix) {
  Carver c(base);
  s.vals = c.take<uint32_t>(N);  // first, so that the backward finds it whatever path the forward took
  s.keys_unsorted = c.take<uint64_t>(N);
  s.keys = radix ? c.take<uint64_t>(N) : nullptr;
  s.vals_unsorted = radix ? c.take<uint32_t>(N) : nullptr;
  s.sort_temp = radix ? c.take<char>(sort_bytes) : nullptr;
  s.sort_bytes = sort_bytes;
  return c.used + 256;
}

size_t ImgState::carve(ImgState& s, char* base, int HW, int tiles, int variant) {
  Carver c(base);
  s.ranges = c.take<uint2>((size_t)tiles);
  s.tile_last = c.take<uint32_t>((size_t)tiles);
  s.tile_count = c.take<uint32_t>((size_t)tiles * kCntStrideMax);
  s.tile_fill = c.take<uint32_t>((size_t)tiles * kCntStrideMax);
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


// ---- tile-local binning (default path) ---------------------------------------------------------
// Instead of six HBM passes of a device-wide 45-bit radix sort, entries are scattered straight into
// their tile's segment and every tile is sorted on chip by one CTA:
//   (preprocess_fwd) one red per (Gaussian, tile) duplicate on a per-tile counter
//   scan_tiles     one CTA: exclusive scan of the tile counters -> ranges, total, longest list
//   scatter        slot = atomicAdd(cursor[tile]) (cursor starts at the range start);
//                  entry = depth bits << 32 | index
//   sort_tiles     one CTA per tile: bitonic sort of the 64-bit entries in shared memory
// The 64-bit entries are unique (index in the low word), so the result is the reference's order
// (depth bits ascending, ties by ascending Gaussian index) regardless of the scatter order.
// HBM traffic: 20 B per duplicate instead of 152 B.  Lists longer than kTileSortCap fall back to
// the radix path.
constexpr int kTileSortCap = 8192;     // entries per tile the shared-memory sort accepts (64 KB)
constexpr int kTileSortThreads = 256;

// one CTA of 1024 threads; counters[0] = total, counters[2] = longest tile list
// `capacity` / `longest_cap` (0xFFFFFFFF = unlimited) serve the sync-free "static" binning of the
// tracker (run_binning_static): ranges are clamped to the buffer and to the longest list the sort
// was launched for, so that every later kernel stays inside the buffers whatever this frame holds,
// and counters[3] is raised when something was cut (the host checks it after the fact).
// The counters of a chunk of 8192 tiles are read coalesced (thread t reads tiles t, t + 1024, ...: all
// eight loads in flight), transposed through shared memory so that every thread scans eight CONSECUTIVE
// tiles serially, one warp-shuffle + one shared-memory step scan the 1024 thread totals, and the range
// starts go back through shared memory to coalesced stores — 4 barriers per 8192 tiles.
// reset_counters: this kernel is the first writer of the frame's counter block and initialises all of
// it (no memset pass); the tracker passes false, its overflow flag is sticky across iterations.
constexpr int kScanPer = 8;
__global__ void __launch_bounds__(1024)
scan_tiles_kernel(int tiles, const uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges,
                  uint32_t* __restrict__ tile_fill, uint32_t* __restrict__ counters,
                  uint32_t capacity, uint32_t longest_cap, int cs, bool reset_counters) {
  __shared__ uint32_t s_cnt[1024 * (kScanPer + 1)];   // one pad word per 8: the stride-8 reads become stride 9
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry, s_max;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto slot = [](int i) { return i + (i >> 3); };
  pdl_wait();   // (launched programmatically behind preprocess_fwd, whose counters it scans)
  if (tid == 0) { s_carry = 0; s_max = 0; }
  uint32_t local_max = 0;
  for (int base = 0; base < tiles; base += 1024 * kScanPer) {
    uint32_t v[kScanPer];
#pragma unroll
    for (int r = 0; r < kScanPer; ++r) {
      const int t = base + r * 1024 + tid;
      v[r] = (t < tiles) ? tile_count[(size_t)t * cs] : 0u;
    }
    __syncthreads();   // previous chunk's readers are done with s_cnt; s_carry is visible
#pragma unroll
    for (int r = 0; r < kScanPer; ++r) s_cnt[slot(r * 1024 + tid)] = v[r];
    __syncthreads();
    uint32_t c[kScanPer];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
      c[k] = s_cnt[slot(tid * kScanPer + k)];
      local_max = max(local_max, c[k]);
      sum += c[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    uint32_t start = carry + (warp ? s_warp[warp - 1] : 0u) + incl - sum;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {   // exclusive starts replace the counts in place
      s_cnt[slot(tid * kScanPer + k)] = start;
      start += c[k];
    }
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_warp[31];
#pragma unroll
    for (int r = 0; r < kScanPer; ++r) {
      const int t = base + r * 1024 + tid;
      if (t < tiles) {
        const uint32_t st = s_cnt[slot(r * 1024 + tid)];
        ranges[t] = make_uint2(min(st, capacity), min(st + min(v[r], longest_cap), capacity));
        tile_fill[(size_t)t * cs] = st;  // scatter cursor
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if (lane == 0) atomicMax(&s_max, local_max);
  __syncthreads();
  if (tid == 0) {
    const bool cut = s_carry > capacity || s_max > longest_cap;
    counters[0] = s_carry;
    counters[2] = s_max;
    if (reset_counters) {
      counters[1] = 0u;                 // num_related, accumulated by the -full forward blend
      counters[3] = cut ? 1u : 0u;
      counters[4] = counters[5] = counters[6] = counters[7] = 0u;
    } else if (cut) {
      counters[3] = 1u;
    }
  }
}

// One thread per Gaussian.  tile_fill[t] starts at the tile's range start (scan_tiles_kernel), so
// one atomic yields the slot.  A splat that covers many tiles is handed to the whole warp (lane i
// takes tiles i, i+32, ...): a serial loop of dependent atomic -> store round trips in one thread
// would otherwise set the kernel's duration.
__global__ void scatter_entries_kernel(int P, const float4* __restrict__ rec,
                                       const uint32_t* __restrict__ tiles_touched,
                                       const uint2* __restrict__ rects, int grid_x,
                                       uint32_t* __restrict__ tile_fill,
                                       uint64_t* __restrict__ entries, uint32_t capacity, uint32_t cs) {
  // `capacity` = entries the buffer can hold.  It only bites when the buffer was sized from the
  // previous frame's count (speculative launch, see run_binning) and this frame has more: the
  // surplus is dropped and the host, which sees the real count, redoes the binning.
  pdl_trigger();   // the per-tile sort behind this kernel may be set up while it runs
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  uint32_t n = (idx < P) ? tiles_touched[idx] : 0u;
  uint2 rc = make_uint2(0u, 0u);
  uint64_t entry = 0ull;
  if (n != 0u) {
    rc = rects[idx];
    entry = ((uint64_t)__float_as_uint(rec[3 * (size_t)idx + 1].w) << 32) | (uint32_t)idx;
  }
  constexpr uint32_t kWide = 8;  // splats with more tiles than this are processed by the whole warp
  unsigned wide = __ballot_sync(0xffffffffu, n > kWide);
  while (wide) {
    const int src = __ffs(wide) - 1;
    wide &= wide - 1;
    const uint32_t rx = __shfl_sync(0xffffffffu, rc.x, src), ry = __shfl_sync(0xffffffffu, rc.y, src);
    const uint32_t e_lo = __shfl_sync(0xffffffffu, (uint32_t)entry, src);
    const uint32_t e_hi = __shfl_sync(0xffffffffu, (uint32_t)(entry >> 32), src);
    const uint32_t x0 = rx & 0xFFFFu, x1 = rx >> 16, y0 = ry & 0xFFFFu, y1 = ry >> 16;
    const uint32_t w = x1 - x0, total = w * (y1 - y0);
    const uint64_t e = ((uint64_t)e_hi << 32) | e_lo;
    // four atomics in flight per lane before the dependent stores (a wide splat is otherwise a
    // chain of atomic round trips)
    for (uint32_t q0 = lane; q0 < total; q0 += 128) {
      uint32_t slot[4];
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u) {
        const uint32_t q = q0 + 32u * u;
        if (q < total) slot[u] = atomicAdd(tile_fill + ((y0 + q / w) * (uint32_t)grid_x + x0 + q % w) * cs, 1u);
      }
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u)
        if (q0 + 32u * u < total && slot[u] < capacity) entries[slot[u]] = e;
    }
  }
  if (n != 0u && n <= kWide) {
    const uint32_t x0 = rc.x & 0xFFFFu, x1 = rc.x >> 16, y0 = rc.y & 0xFFFFu, y1 = rc.y >> 16;
    const uint32_t w = x1 - x0;
    (void)y1;
    uint32_t slots[kWide];
#pragma unroll
    for (uint32_t u = 0; u < kWide; ++u)  // issue all atomics first, then the dependent stores
      if (u < n) slots[u] = atomicAdd(tile_fill + ((y0 + u / w) * (uint32_t)grid_x + x0 + u % w) * cs, 1u);
#pragma unroll
    for (uint32_t u = 0; u < kWide; ++u)
      if (u < n && slots[u] < capacity) entries[slots[u]] = entry;
  }
}

// Bitonic sort of one tile's entries with E elements per thread held in registers (element index
// e = m * 256 + tid) over p = 2^LOGP >= n slots.  The network is fully unrolled (j, k are compile
// time constants); compare-exchange partners at distance j are reached with warp shuffles
// (j < 32), through shared memory (32 <= j < 256) or inside the thread (j >= 256), so only 6 of
// the 36 steps of a 256-entry tile need a barrier.  Threads beyond p leave at once.
__device__ __forceinline__ uint64_t cex_pick(uint64_t a, uint64_t o, bool keep_min) {
  const bool a_lt = a < o;
  return (a_lt == keep_min) ? a : o;
}

template <int E, int LOGP>
__device__ __forceinline__ void bitonic_sort_regs(uint64_t* s, const uint64_t* __restrict__ entries,
                                                  uint32_t* __restrict__ out, int n, int tid) {
  constexpr int T = kTileSortThreads;
  if (E == 1 && tid >= ((1 << LOGP) < 32 ? 32 : (1 << LOGP))) return;  // whole idle warps leave
  uint64_t v[E];
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const int e = m * T + tid;
    v[m] = (e < n) ? entries[e] : ~0ull;
  }
#pragma unroll
  for (int lk = 1; lk <= LOGP; ++lk) {
    const int k = 1 << lk;
#pragma unroll
    for (int lj = lk - 1; lj >= 0; --lj) {
      const int j = 1 << lj;
      if (j >= T) {
        const int JM = j / T;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          if ((m & JM) == 0 && (m | JM) < E) {
            const bool up = ((m * T + tid) & k) == 0;
            const uint64_t a = v[m], b = v[m | JM];
            const bool sw = (a > b) == up;
            v[m] = sw ? b : a;
            v[m | JM] = sw ? a : b;
          }
        }
      } else if (j >= 32) {
        __syncthreads();
#pragma unroll
        for (int m = 0; m < E; ++m) s[m * T + tid] = v[m];
        __syncthreads();
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int e = m * T + tid;
          const bool keep_min = ((e & j) == 0) == ((e & k) == 0);
          v[m] = cex_pick(v[m], s[e ^ j], keep_min);
        }
      } else {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int e = m * T + tid;
          const bool keep_min = ((e & j) == 0) == ((e & k) == 0);
          v[m] = cex_pick(v[m], __shfl_xor_sync(0xffffffffu, v[m], j), keep_min);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < E; ++m) {
    const int e = m * T + tid;
    if (e < n) out[e] = (uint32_t)v[m];
  }
}

// generic shared-memory network for long lists (2048 < n <= kTileSortCap)
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, const uint64_t* __restrict__ entries,
                                                  uint32_t* __restrict__ out, int n, int tid) {
  int p = 2;
  while (p < n) p <<= 1;
  for (int i = tid; i < p; i += kTileSortThreads) s[i] = (i < n) ? entries[i] : ~0ull;
  __syncthreads();
  for (int k = 2; k <= p; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (p >> 1); t += kTileSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const uint64_t a = s[i], b = s[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { s[i] = b; s[l] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += kTileSortThreads) out[i] = (uint32_t)s[i];
}

// built for 6 CTAs per SM (40 registers, no spills): 0.045 ms at C3 against 0.047 at 4 CTAs (64 registers) and 0.053
// with the compiler's own choice of 80 registers / 3 CTAs (profiles/r02_ab_pre_occ.txt)
__global__ void __launch_bounds__(kTileSortThreads, 6)
sort_tiles_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ entries,
                  uint32_t* __restrict__ vals, uint32_t capacity, int smem_entries) {
  extern __shared__ __align__(16) unsigned char sort_smem_raw[];
  uint64_t* s = reinterpret_cast<uint64_t*>(sort_smem_raw);
  pdl_wait();      // launched programmatically behind the scatter
  pdl_trigger();   // ... and the forward blend behind this kernel
  const uint2 range = ranges[blockIdx.x];
  const int n = (int)(range.y - range.x);
  if (n == 0) return;
  // speculative launch with too small a buffer / too little shared memory: the host redoes it
  if (range.y > capacity || n > smem_entries) return;
  const int tid = threadIdx.x;
  const uint64_t* in = entries + range.x;
  uint32_t* out = vals + range.x;
  if (n == 1) {
    if (tid == 0) out[0] = (uint32_t)in[0];
  } else if (n <= 2) { bitonic_sort_regs<1, 1>(s, in, out, n, tid);
  } else if (n <= 4) { bitonic_sort_regs<1, 2>(s, in, out, n, tid);
  } else if (n <= 8) { bitonic_sort_regs<1, 3>(s, in, out, n, tid);
  } else if (n <= 16) { bitonic_sort_regs<1, 4>(s, in, out, n, tid);
  } else if (n <= 32) { bitonic_sort_regs<1, 5>(s, in, out, n, tid);
  } else if (n <= 64) { bitonic_sort_regs<1, 6>(s, in, out, n, tid);
  } else if (n <= 128) { bitonic_sort_regs<1, 7>(s, in, out, n, tid);
  } else if (n <= 256) { bitonic_sort_regs<1, 8>(s, in, out, n, tid);
  } else if (n <= 512) { bitonic_sort_regs<2, 9>(s, in, out, n, tid);
  } else if (n <= 1024) { bitonic_sort_regs<4, 10>(s, in, out, n, tid);
  } else if (n <= 2048) { bitonic_sort_regs<8, 11>(s, in, out, n, tid);
  } else {
    bitonic_sort_smem(s, in, out, n, tid);
  }
}

// Per-DEVICE read-back resources: pinned slots for the duplicate count (so the copy is truly
// asynchronous), their events (events belong to the device they were created on) and the one-time
// kernel attribute of sort_tiles_kernel.  Created lazily for the device that is current at the call.
struct DeviceSlots {
  static constexpr int kSlots = 32;
  uint32_t* host = nullptr;  // [kSlots][4] pinned
  cudaEvent_t ev[kSlots];
  std::atomic<unsigned> next{0};
  bool ok = false;
  bool sort_attr_set = false;
  void init() {
    if (cudaHostAlloc(reinterpret_cast<void**>(&host), sizeof(uint32_t) * 4 * kSlots, cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    for (int i = 0; i < kSlots; ++i)
      if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    ok = true;
  }
};
constexpr int kMaxDevices = 64;
std::mutex g_ctx_mu;
DeviceSlots* g_dev_slots[kMaxDevices] = {nullptr};

DeviceSlots* device_slots(int* device_out = nullptr) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (device_out) *device_out = dev;
  if (dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (g_dev_slots[dev] == nullptr) {
    g_dev_slots[dev] = new DeviceSlots();  // leaked on purpose (outlives static destruction)
    g_dev_slots[dev]->init();
  }
  return g_dev_slots[dev];
}

// ---- accumulator clear on a side stream (see gsr_common.cuh) -------------------------------------------------
struct AccClear {
  static constexpr int kRing = 16;
  cudaStream_t side = nullptr;
  cudaEvent_t fork[kRing], done[kRing];
  const void* key[kRing];
  bool pending[kRing];
  unsigned next = 0;
  bool ok = false;
  void init() {
    if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return; }
    for (int i = 0; i < kRing; ++i) {
      key[i] = nullptr; pending[i] = false;
      if (cudaEventCreateWithFlags(&fork[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    }
    ok = true;
  }
};
AccClear* g_acc_clear[kMaxDevices] = {nullptr};

AccClear* acc_clear_state() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (dev < 0 || dev >= kMaxDevices) return nullptr;
  if (g_acc_clear[dev] == nullptr) {   // (caller holds g_ctx_mu)
    g_acc_clear[dev] = new AccClear();   // leaked on purpose (outlives static destruction)
    g_acc_clear[dev]->init();
  }
  return g_acc_clear[dev]->ok ? g_acc_clear[dev] : nullptr;
}

// Speculative-binning size estimate, one per (device, width, height, P) context: the sizes seen on
// that context's previous frame size this frame's binning buffer before the count is known.  A
// context speculates only while its estimate is `stable` (the last frame's counts would have fitted
// the estimate made from the frame before), so an alternating or drifting workload falls back to the
// synchronous path instead of mis-speculating every frame.
struct BinContext {
  int device, W, H, P;
  uint32_t hint_entries, hint_longest;
  bool stable;
  uint64_t stamp;
};
constexpr int kMaxContexts = 64;
BinContext g_ctx[kMaxContexts];
int g_ctx_n = 0;
uint64_t g_ctx_clock = 0;

uint32_t spec_capacity(uint32_t hint_n) { return hint_n + hint_n / 4 + 4096; }
uint32_t spec_longest(uint32_t hint_l) {
  uint32_t l = hint_l * 2 < (uint32_t)kTileSortThreads ? (uint32_t)kTileSortThreads : hint_l * 2;
  return l > (uint32_t)kTileSortCap ? (uint32_t)kTileSortCap : l;
}

// -> true and the estimate when this context may speculate
bool context_hint(int device, int W, int H, int P, uint32_t* n, uint32_t* l) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  for (int i = 0; i < g_ctx_n; ++i) {
    BinContext& c = g_ctx[i];
    if (c.device == device && c.W == W && c.H == H && c.P == P) {
      c.stamp = ++g_ctx_clock;
      *n = c.hint_entries; *l = c.hint_longest;
      return c.stable && c.hint_entries > 0 && c.hint_longest <= (uint32_t)kTileSortCap;
    }
  }
  return false;
}

void context_update(int device, int W, int H, int P, uint32_t n, uint32_t l) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  BinContext* c = nullptr;
  for (int i = 0; i < g_ctx_n; ++i)
    if (g_ctx[i].device == device && g_ctx[i].W == W && g_ctx[i].H == H && g_ctx[i].P == P) c = &g_ctx[i];
  if (c == nullptr) {
    if (g_ctx_n < kMaxContexts) {
      c = &g_ctx[g_ctx_n++];
    } else {  // evict the least recently used context
      c = &g_ctx[0];
      for (int i = 1; i < g_ctx_n; ++i) if (g_ctx[i].stamp < c->stamp) c = &g_ctx[i];
    }
    *c = BinContext{device, W, H, P, n, l, /*stable=*/true, ++g_ctx_clock};
    return;
  }
  int p = kTileSortThreads;
  while (p < (int)spec_longest(c->hint_longest)) p <<= 1;
  c->stable = c->hint_entries > 0 && n <= spec_capacity(c->hint_entries) && l <= (uint32_t)p;
  c->hint_entries = n;
  c->hint_longest = l;
  c->stamp = ++g_ctx_clock;
}

int launch_tile_sort(const Camera& cam, int P, const GeomState& g, const ImgState& img, BinState& b,
                     uint32_t capacity, uint32_t longest_cap, bool debug, cudaStream_t stream) {
  const int tiles = cam.grid_x * cam.grid_y;
  {
    StageScope st(ST_EMIT, stream);
    scatter_entries_kernel<<<(P + 255) / 256, 256, 0, stream>>>(
        P, g.rec, g.tiles_touched, g.rect, cam.grid_x, img.tile_fill, b.keys_unsorted, capacity,
        (uint32_t)cnt_stride());
    GSR_LAUNCH_OK(debug, stream);
  }
  {
    StageScope st(ST_SORT, stream);
    int p = kTileSortThreads;  // the register network exchanges 256*E entries through smem
    while (p < (int)longest_cap) p <<= 1;
    const size_t smem = (size_t)p * sizeof(uint64_t);
    // the opt-in to > 48 KB of dynamic shared memory is per device (and only needed for long lists)
    DeviceSlots* ds = smem > 48 * 1024 ? device_slots() : nullptr;
    if (smem > 48 * 1024 && (ds == nullptr || !ds->sort_attr_set)) {
      GSR_CUDA_OK(cudaFuncSetAttribute(sort_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kTileSortCap * (int)sizeof(uint64_t)));
      if (ds != nullptr) ds->sort_attr_set = true;
    }
    launch_after(options().pdl != 0, sort_tiles_kernel, dim3(tiles), dim3(kTileSortThreads), smem, stream,
                 (const uint2*)img.ranges, (const uint64_t*)b.keys_unsorted, b.vals, capacity, p);
    GSR_LAUNCH_OK(debug, stream);
  }
  return GSR_OK;
}

}  // namespace

// Sync-free binning into a caller-owned buffer of `capacity` entries whose longest tile list may
// not exceed `longest_cap` (tracker: sizes come from a probing frame plus head room).  Nothing is
// read back; overflow raises g.counters[3] and truncates the frame's lists (memory-safe).
void acc_clear_begin(const void* geom_key, void* acc, size_t bytes, cudaStream_t stream) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  AccClear* a = acc_clear_state();
  if (a == nullptr) return;
  int i = -1;
  for (int k = 0; k < AccClear::kRing; ++k)
    if (a->key[k] == geom_key) i = k;            // the same buffer again: its old entry is stale
  if (i < 0) i = (int)(a->next++ % AccClear::kRing);
  a->key[i] = nullptr; a->pending[i] = false;
  if (cudaEventRecord(a->fork[i], stream) != cudaSuccess || cudaStreamWaitEvent(a->side, a->fork[i], 0) != cudaSuccess ||
      cudaMemsetAsync(acc, 0, bytes, a->side) != cudaSuccess || cudaEventRecord(a->done[i], a->side) != cudaSuccess) {
    cudaGetLastError();   // the backward clears its own scratch
    return;
  }
  a->key[i] = geom_key; a->pending[i] = true;
}

void acc_clear_rejoin(const void* geom_key, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  AccClear* a = acc_clear_state();
  if (a == nullptr) return;
  for (int k = 0; k < AccClear::kRing; ++k)
    if (a->key[k] == geom_key && a->pending[k]) {
      if (cudaStreamWaitEvent(stream, a->done[k], 0) != cudaSuccess) cudaGetLastError();
      return;
    }
}

bool acc_clear_join(const void* geom_key, cudaStream_t stream) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  AccClear* a = acc_clear_state();
  if (a == nullptr) return false;
  for (int k = 0; k < AccClear::kRing; ++k) {
    if (a->key[k] == geom_key && a->pending[k]) {
      a->pending[k] = false;   // one backward per clear
      a->key[k] = nullptr;
      if (cudaStreamWaitEvent(stream, a->done[k], 0) != cudaSuccess) { cudaGetLastError(); return false; }
      return true;
    }
  }
  return false;
}

int run_binning_static(int P, const Camera& cam, GeomState& g, BinState& b, ImgState& img,
                       uint32_t capacity, uint32_t longest_cap, cudaStream_t stream) {
  const int tiles = cam.grid_x * cam.grid_y;
  if (longest_cap > (uint32_t)kTileSortCap) longest_cap = kTileSortCap;
  {
    StageScope st(ST_SCAN, stream, 1);
    scan_tiles_kernel<<<1, 1024, 0, stream>>>(tiles, img.tile_count, img.ranges, img.tile_fill,
                                              g.counters, capacity, longest_cap, cnt_stride(), false);
    GSR_LAUNCH_OK(false, stream);
  }
  return launch_tile_sort(cam, P, g, img, b, capacity, longest_cap, false, stream);
}

// Tile scan only (no clamps): leaves the duplicate count in g.counters[0] and the longest tile list
// in g.counters[2] for the caller to read back.
int probe_tile_counts(const Camera& cam, GeomState& g, ImgState& img, cudaStream_t stream) {
  scan_tiles_kernel<<<1, 1024, 0, stream>>>(cam.grid_x * cam.grid_y, img.tile_count, img.ranges,
                                            img.tile_fill, g.counters, 0xFFFFFFFFu, 0xFFFFFFFFu, cnt_stride(), false);
  GSR_LAUNCH_OK(false, stream);
  return GSR_OK;
}

int run_binning(int P, const Camera& cam, const int* /*radii*/, GeomState& g, gsr_alloc_fn alloc,
                void* alloc_ctx, BinState& b, ImgState& img, int* num_rendered, bool debug,
                cudaStream_t stream, SpecRender* spec) {
  const int tiles = cam.grid_x * cam.grid_y;
  const bool tile_local = options().tile_sort != 0;
  if (spec) spec->done = false;
  uint32_t N = 0, longest = 0;
  bool sorted_speculatively = false;

  if (tile_local) {
    // The one host<->device synchronisation of the forward: the duplicate count sizes the binning
    // buffer and is returned to the caller (the reference blocks in the same place,
    // rasterizer_impl.cu:287).  To keep the GPU busy while the host waits, the count is copied into
    // pinned memory asynchronously and — when a previous frame left a size estimate — the scatter,
    // the per-tile sort and (spec) the forward blend are enqueued first, on a buffer sized from that
    // estimate; the scan clamps the ranges to that buffer, the kernels are guarded, and if the estimate
    // turns out too small the binning (and the blend) is simply redone.
    int device = 0;
    DeviceSlots* dsp = device_slots(&device);
    const bool slots_ok = dsp != nullptr && dsp->ok;
    uint32_t hint_n = 0, hint_l = 0;
    const bool speculate = slots_ok && options().async_binning != 0 &&
                           context_hint(device, cam.W, cam.H, P, &hint_n, &hint_l);
    uint32_t cap = 0, lcap = 0, lpad = 0;
    if (speculate) {
      cap = spec_capacity(hint_n);
      lcap = spec_longest(hint_l);
      int p = kTileSortThreads;
      while (p < (int)lcap) p <<= 1;
      lpad = (uint32_t)p;   // the sort is launched for the next power of two
    }
    {
      // img.tile_count was filled by preprocess_fwd (one red per duplicate)
      StageScope st(ST_SCAN, stream, 1);
      launch_after(options().pdl != 0, scan_tiles_kernel, dim3(1), dim3(1024), 0, stream, tiles,
                   (const uint32_t*)img.tile_count, img.ranges, img.tile_fill, g.counters,
                   speculate ? cap : 0xFFFFFFFFu, speculate ? lpad : 0xFFFFFFFFu, cnt_stride(), true);
      GSR_LAUNCH_OK(debug, stream);
    }
    uint32_t h[4] = {0, 0, 0, 0};
    if (slots_ok) {
      DeviceSlots& cs = *dsp;
      const unsigned slot = cs.next.fetch_add(1) % DeviceSlots::kSlots;
      uint32_t* hp = cs.host + 4 * slot;
      GSR_CUDA_OK(cudaMemcpyAsync(hp, g.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
      GSR_CUDA_OK(cudaEventRecord(cs.ev[slot], stream));
      if (speculate) {
        const size_t need = BinState::carve(b, nullptr, cap, 0, false);
        char* chunk = alloc(alloc_ctx, need);
        if (chunk == nullptr) { set_error("binning allocator returned NULL for %zu bytes", need); return GSR_E_ALLOC; }
        BinState::carve(b, chunk, cap, 0, false);
        int rc = launch_tile_sort(cam, P, g, img, b, cap, lcap, debug, stream);
        if (rc != GSR_OK) return rc;
        if (spec != nullptr) {
          rc = spec->launch(b);
          if (rc != GSR_OK) return rc;
          spec->done = true;
        }
      }
      GSR_CUDA_OK(cudaEventSynchronize(cs.ev[slot]));
      h[0] = hp[0]; h[2] = hp[2];
      if (speculate) {
        sorted_speculatively = h[0] <= cap && h[2] <= lpad;
        if (!sorted_speculatively) {
          if (spec != nullptr && spec->done) {
            const int rc = spec->reset();
            if (rc != GSR_OK) return rc;
            spec->done = false;
          }
          // estimate too small: reset the scatter cursors (the scan rewrites them) and fall through
          StageScope st(ST_SCAN, stream, 1);
          scan_tiles_kernel<<<1, 1024, 0, stream>>>(tiles, img.tile_count, img.ranges, img.tile_fill,
                                                    g.counters, 0xFFFFFFFFu, 0xFFFFFFFFu, cnt_stride(), true);
          GSR_LAUNCH_OK(debug, stream);
        }
      }
    } else {
      GSR_CUDA_OK(cudaMemcpyAsync(h, g.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
      GSR_CUDA_OK(cudaStreamSynchronize(stream));
    }
    N = h[0];
    longest = h[2];
    context_update(device, cam.W, cam.H, P, N, longest);
  } else {
    StageScope st(ST_SCAN, stream);
    GSR_CUDA_OK(cub::DeviceScan::InclusiveSum(g.scan_temp, g.scan_bytes, g.tiles_touched,
                                              g.offsets, P, stream));
    GSR_LAUNCH_OK(debug, stream);
  }
  if (!tile_local) {
    GSR_CUDA_OK(cudaMemcpyAsync(&N, g.offsets + (P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                stream));
    GSR_CUDA_OK(cudaStreamSynchronize(stream));
  }
  *num_rendered = (int)N;
  if (sorted_speculatively) return GSR_OK;

  const bool use_tile_sort = tile_local && longest <= (uint32_t)kTileSortCap;
  const int end_bit = 32 + bits_for((uint32_t)tiles);
  const size_t sort_bytes = use_tile_sort ? 0 : sort_temp_bytes(N, end_bit);
  const size_t need = BinState::carve(b, nullptr, N, sort_bytes, !use_tile_sort);
  char* chunk = alloc(alloc_ctx, need);
  if (chunk == nullptr && need > 0) {
    set_error("binning allocator returned NULL for %zu bytes", need);
    return GSR_E_ALLOC;
  }
  BinState::carve(b, chunk, N, sort_bytes, !use_tile_sort);

  if (use_tile_sort) {
    if (N == 0) return GSR_OK;  // ranges were written (all empty) by scan_tiles_kernel
    return launch_tile_sort(cam, P, g, img, b, N, longest, debug, stream);
  }

  // ---- radix path (reference structure; also the fallback for very long tile lists) ----
  if (tile_local) {  // the offsets were not computed yet
    StageScope st(ST_SCAN, stream);
    GSR_CUDA_OK(cub::DeviceScan::InclusiveSum(g.scan_temp, g.scan_bytes, g.tiles_touched,
                                              g.offsets, P, stream));
    GSR_LAUNCH_OK(debug, stream);
  }
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
