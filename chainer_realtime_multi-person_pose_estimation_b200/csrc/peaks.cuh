// peaks.cuh -- compute_peaks_from_heatmaps (CPU branch), pose_detector.py:75-110.
//
//   smooth_nms_kernel : per (image, joint channel) tile: 21-tap separable Gaussian
//                       (scipy.ndimage.gaussian_filter semantics: 'reflect' = symmetric
//                       extension, axis-0 pass then axis-1 pass, float64 accumulate in scipy's
//                       order  x0*w0 + sum_{j=-r..-1} (x[j]+x[-j])*w[j],  float32 store after
//                       each pass), then strict '>' against the 4 axial neighbours (zero
//                       outside the image), threshold '>' in float32, append (key, score).
//   sort_peaks_kernel : per image bitonic sort by key = (channel, y, x) so that peak ids equal
//                       the reference's np.nonzero order (channel-major, row-major).
//
// All float64 arithmetic uses explicit __dadd_rn/__dmul_rn so nvcc cannot contract it to FMA:
// results are bit-identical to scipy (tests/test_gpu_postprocess.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "upsample.cuh"

namespace opb {

struct PeakD {       // device peak table row
  double x, y;       // integer-valued on the fused path
  float score;       // smoothed heatmap value (float32, as in the reference)
  int type;          // joint channel
};

struct PeakKey {
  uint32_t key;      // (c*H + y)*W + x
  float score;
};

constexpr int PK_TX = 64, PK_TY = 16, PK_R_MAX = 16;   // tile of the peak kernel (32-row tiles measured no faster)
constexpr int PK_THREADS = 256;  // (128-thread blocks were measured slower: 1.03 vs 0.74 ms)
constexpr int PK_R_FAST = 10;   // radius of sigma = 2.5 (entity.py:75): compile-time specialisation

__device__ __forceinline__ int reflect_index(int i, int n) {
  // scipy 'reflect' (d c b a | a b c d | d c b a); loop handles radius > n
  while (i < 0 || i >= n) {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
  }
  return i;
}

struct GaussTaps {
  double w[2 * PK_R_MAX + 1];
  int radius;
};

// Maxima of the raw heat maps over PK_CELL x PK_CELL pixel cells: cell_max[plane][cy][cx].
// One coalesced pass (each pixel read once); smooth_nms_kernel tests the cells overlapping its
// input window BEFORE loading it, so tiles far from every blob cost nothing.
constexpr int PK_CELL = 8;
__global__ void __launch_bounds__(256)
cell_max_kernel(const float* __restrict__ heat, int c_total, int c_use, int H, int W, float* __restrict__ cell_max,
                int cells_y, int cells_x) {
  const int plane = blockIdx.z;
  const int img = plane / c_use, c = plane - img * c_use;
  const int x0 = blockIdx.x * PK_TX, y0 = blockIdx.y * PK_TY;
  const float* src = heat + (static_cast<size_t>(img) * c_total + c) * H * W;
  constexpr int CELLS = (PK_TY / PK_CELL) * (PK_TX / PK_CELL);       // cells per tile
  constexpr int TPC = 256 / CELLS;                                    // threads per cell (power of two <= 32)
  static_assert(CELLS * TPC == 256 && TPC >= 1 && TPC <= 32 && (64 % TPC) == 0, "cell mapping");
  const int cell = threadIdx.x / TPC, sub = threadIdx.x % TPC;
  const int cy = cell / (PK_TX / PK_CELL), cx = cell % (PK_TX / PK_CELL);
  float m = -3.0e38f;
#pragma unroll
  for (int k = 0; k < 64 / TPC; ++k) {
    const int e = sub * (64 / TPC) + k;                                // 64 pixels of the cell
    const int y = y0 + cy * PK_CELL + (e >> 3), x = x0 + cx * PK_CELL + (e & 7);
    if (y < H && x < W) m = fmaxf(m, __ldg(src + static_cast<size_t>(y) * W + x));
  }
#pragma unroll
  for (int o = TPC / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int gy = blockIdx.y * (PK_TY / PK_CELL) + cy, gx = blockIdx.x * (PK_TX / PK_CELL) + cx;
  if (sub == 0 && gy < cells_y && gx < cells_x)
    cell_max[(static_cast<size_t>(plane) * cells_y + gy) * cells_x + gx] = m;
}

// heat: [n_img][c_total][H][W] f32, only the first `c_use` channels of each image are processed.
//
// Exactness with little float64 work:
//   1. tiles whose 3x3 tile neighbourhood stays below the threshold cannot hold a peak
//      (taps >= 0, sum 1) and exit before touching memory (tile_max pre-pass);
//   2. the remaining tiles are smoothed in float32 (FFMA) to find CANDIDATES: pixels that pass
//      the peak test with a slack of delta = 1e-5 * max|input|, > 2x the float32 error bound
//      (2 passes x 21 roundings x 2^-24), so every true peak is a candidate;
//   3. each candidate is re-evaluated EXACTLY: the float64 two-pass sums (scipy's operation
//      order, float32 store between the passes) of the pixel and its four neighbours, and
//      the reference's strict comparisons decide.  Results are bit-identical to the
//      all-float64 kernel; B200's scalar fp64 rate (~1/8 of fp32) is paid only per candidate.
// mode 0 with the default passes: exactly the profiled token stream (its SASS is pinned with tools/sass_diff.py)
#define OPB_PK_NAME(n) n
#define OPB_PK_SPLIT 0
#define OPB_PK_LOWRES 0
#include "peaks_smooth_nms.inc"
#undef OPB_PK_LOWRES
#define OPB_PK_LOWRES 1
#include "peaks_smooth_nms.inc"
#undef OPB_PK_LOWRES
#define OPB_PK_LOWRES 2
#include "peaks_smooth_nms.inc"
#undef OPB_PK_LOWRES
#undef OPB_PK_SPLIT
#undef OPB_PK_NAME
// experimental: both smoothing passes on all threads (OPB_PEAKS_V2=1), for the materialised-map modes
#define OPB_PK_NAME(n) n##_v2
#define OPB_PK_SPLIT 1
#define OPB_PK_LOWRES 2
#include "peaks_smooth_nms.inc"
#undef OPB_PK_LOWRES
#undef OPB_PK_SPLIT
#undef OPB_PK_NAME

inline size_t smooth_nms_smem_bytes(int radius) {
  const int IN_W = PK_TX + 2 + 2 * radius, IN_H = PK_TY + 2 + 2 * radius;
  return sizeof(double) * 8 * 3 * (2 * radius + 3) +
         sizeof(float) * (static_cast<size_t>(IN_H) * IN_W + (PK_TY + 2) * IN_W + (PK_TY + 2) * (PK_TX + 2)) +
         sizeof(int) * PK_TY * PK_TX;
}

// One block per image: bitonic sort of (key, score) in shared memory, then emit the peak table
// (type, x, y, score), the per-type start offsets and the identity index list.
// status[img] |= 1 when the append list overflowed.
__global__ void __launch_bounds__(1024)
sort_peaks_kernel(const PeakKey* __restrict__ keys, int* __restrict__ counts, int cap, int H, int W, int n_types,
                  PeakD* __restrict__ peaks, int* __restrict__ idx_list, int* __restrict__ type_start,
                  int* __restrict__ status) {
  extern __shared__ unsigned long long s_kv[];   // key << 32 | score bits
  const int img = blockIdx.x;
  int n = counts[img];
  if (n > cap) {
    if (threadIdx.x == 0) { atomicOr(&status[img], 1); }
    n = cap;
  }
  int npow = 1;
  while (npow < n) npow <<= 1;
  for (int i = threadIdx.x; i < npow; i += blockDim.x) {
    if (i < n) {
      const PeakKey k = keys[static_cast<size_t>(img) * cap + i];
      s_kv[i] = (static_cast<unsigned long long>(k.key) << 32) | __float_as_uint(k.score);
    } else {
      s_kv[i] = ~0ull;
    }
  }
  __syncthreads();
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s_kv[i], b = s_kv[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { s_kv[i] = b; s_kv[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  const uint32_t plane = static_cast<uint32_t>(H) * W;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long kv = s_kv[i];
    const uint32_t key = static_cast<uint32_t>(kv >> 32);
    const int c = key / plane;
    const uint32_t rem = key - c * plane;
    PeakD p;
    p.type = c;
    p.y = static_cast<double>(rem / W);
    p.x = static_cast<double>(rem % W);
    p.score = __uint_as_float(static_cast<uint32_t>(kv & 0xffffffffu));
    peaks[static_cast<size_t>(img) * cap + i] = p;
    idx_list[static_cast<size_t>(img) * cap + i] = i;
    // first peak of its type?
    const int prev_c = (i == 0) ? -1 : static_cast<int>(static_cast<uint32_t>(s_kv[i - 1] >> 32) / plane);
    for (int t = prev_c + 1; t <= c; ++t) type_start[img * (n_types + 1) + t] = i;
    if (i == n - 1)
      for (int t = c + 1; t <= n_types; ++t) type_start[img * (n_types + 1) + t] = n;
  }
  if (n == 0)
    for (int t = threadIdx.x; t <= n_types; t += blockDim.x) type_start[img * (n_types + 1) + t] = 0;
  __syncthreads();
  if (threadIdx.x == 0) counts[img] = n;
}

}  // namespace opb
