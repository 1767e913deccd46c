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

constexpr int PK_TX = 64, PK_TY = 16, PK_R_MAX = 16;

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

// Per-tile maxima of the raw heat maps: tile_max[(img*c_use + c)][ty][tx] over PK_TY x PK_TX pixels.
// One coalesced pass (each pixel read once); lets smooth_nms_kernel drop inactive tiles before
// it touches memory.
__global__ void __launch_bounds__(256)
tile_max_kernel(const float* __restrict__ heat, int c_total, int c_use, int H, int W, float* __restrict__ tile_max) {
  const int plane = blockIdx.z;
  const int img = plane / c_use, c = plane - img * c_use;
  const int x0 = blockIdx.x * PK_TX, y0 = blockIdx.y * PK_TY;
  const float* src = heat + (static_cast<size_t>(img) * c_total + c) * H * W;
  float m = -3.0e38f;
  for (int i = threadIdx.x; i < PK_TY * PK_TX; i += blockDim.x) {
    const int y = y0 + i / PK_TX, x = x0 + i % PK_TX;
    if (y < H && x < W) m = fmaxf(m, __ldg(src + static_cast<size_t>(y) * W + x));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s_m[8];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_m[i]);
    tile_max[(static_cast<size_t>(plane) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = m;
  }
}

// heat: [n_img][c_total][H][W] f32, only the first `c_use` channels of each image are processed.
// A tile whose whole input window (tile + halo) stays below the threshold cannot contain a
// peak (the taps are non-negative and sum to 1), so it is skipped: on real heat maps (sparse
// blobs) most tiles exit after the load.  Inputs are widened to float64 once, in shared memory.
__global__ void __launch_bounds__(256)
smooth_nms_kernel(const float* __restrict__ heat, int c_total, int c_use, int H, int W, GaussTaps taps,
                  float thresh, PeakKey* __restrict__ out, int* __restrict__ counts, int cap,
                  const float* __restrict__ tile_max) {
  const int R = taps.radius;
  // smoothed <= max(window) * (sum of taps ~ 1): a 0.1 % guard band covers the rounding
  const float skip_below = (thresh > 0.f) ? thresh * 0.999f : thresh * 1.001f - 1e-30f;
  if (tile_max != nullptr && R + 1 <= PK_TY) {
    // the window (tile + R+1 halo, reflected at the borders) lies inside the 3x3 tile neighbourhood
    const float* tm = tile_max + static_cast<size_t>(blockIdx.z) * gridDim.y * gridDim.x;
    float m = -3.0e38f;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int ty = min(max(static_cast<int>(blockIdx.y) + dy, 0), static_cast<int>(gridDim.y) - 1);
        const int tx = min(max(static_cast<int>(blockIdx.x) + dx, 0), static_cast<int>(gridDim.x) - 1);
        m = fmaxf(m, __ldg(tm + ty * gridDim.x + tx));
      }
    if (!(m > skip_below)) return;
  }
  const int IN_W = PK_TX + 2 + 2 * R;   // columns of the input / pass-1 tile
  const int IN_H = PK_TY + 2 + 2 * R;
  const int O_H = PK_TY + 2, O_W = PK_TX + 2;
  extern __shared__ double smd[];
  double* s_in = smd;                      // [IN_H][IN_W]  input widened to float64
  double* s_1 = s_in + IN_H * IN_W;        // [O_H][IN_W]   axis-0 pass, rounded to float32, widened
  float* s_2 = reinterpret_cast<float*>(s_1 + O_H * IN_W);   // [O_H][O_W] after the axis-1 pass

  const int plane = blockIdx.z;
  const int img = plane / c_use, c = plane - img * c_use;
  const int x0 = blockIdx.x * PK_TX, y0 = blockIdx.y * PK_TY;
  const float* src = heat + (static_cast<size_t>(img) * c_total + c) * H * W;

  float vmax = -3.0e38f;
#pragma unroll 4
  for (int i = threadIdx.x; i < IN_H * IN_W; i += blockDim.x) {
    const int r = i / IN_W, q = i - r * IN_W;
    const int gy = reflect_index(y0 - 1 - R + r, H);
    const int gx = reflect_index(x0 - 1 - R + q, W);
    const float v = __ldg(src + static_cast<size_t>(gy) * W + gx);
    vmax = fmaxf(vmax, v);
    s_in[i] = static_cast<double>(v);
  }
  if (!__syncthreads_or(vmax > skip_below)) return;

  // axis-0 pass: output rows y0-1 .. y0+PK_TY, all IN_W columns
  for (int i = threadIdx.x; i < O_H * IN_W; i += blockDim.x) {
    const int r = i / IN_W, q = i - r * IN_W;
    const double* col = s_in + (r + R) * IN_W + q;
    double acc = __dmul_rn(col[0], taps.w[R]);
    for (int j = -R; j < 0; ++j)
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(col[j * IN_W], col[-j * IN_W]), taps.w[R + j]));
    s_1[i] = static_cast<double>(static_cast<float>(acc));   // float32 store between the passes
  }
  __syncthreads();
  // axis-1 pass: output columns x0-1 .. x0+PK_TX
  for (int i = threadIdx.x; i < O_H * O_W; i += blockDim.x) {
    const int r = i / O_W, q = i - r * O_W;
    const double* p = s_1 + r * IN_W + q + R;
    double acc = __dmul_rn(p[0], taps.w[R]);
    for (int j = -R; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(p[j], p[-j]), taps.w[R + j]));
    s_2[i] = static_cast<float>(acc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PK_TY * PK_TX; i += blockDim.x) {
    const int r = i / PK_TX, q = i - r * PK_TX;
    const int y = y0 + r, x = x0 + q;
    if (y >= H || x >= W) continue;
    const float g = s_2[(r + 1) * O_W + q + 1];
    if (!(g > thresh)) continue;
    const float up = (y > 0) ? s_2[r * O_W + q + 1] : 0.f;
    const float dn = (y < H - 1) ? s_2[(r + 2) * O_W + q + 1] : 0.f;
    const float lf = (x > 0) ? s_2[(r + 1) * O_W + q] : 0.f;
    const float rt = (x < W - 1) ? s_2[(r + 1) * O_W + q + 2] : 0.f;
    if (g > up && g > dn && g > lf && g > rt) {
      const int slot = atomicAdd(&counts[img], 1);
      if (slot < cap) {
        PeakKey k;
        k.key = static_cast<uint32_t>((static_cast<size_t>(c) * H + y) * W + x);
        k.score = g;
        out[static_cast<size_t>(img) * cap + slot] = k;
      }
    }
  }
}

inline size_t smooth_nms_smem_bytes(int radius) {
  const int IN_W = PK_TX + 2 + 2 * radius, IN_H = PK_TY + 2 + 2 * radius;
  return sizeof(double) * (static_cast<size_t>(IN_H) * IN_W + (PK_TY + 2) * IN_W) +
         sizeof(float) * (PK_TY + 2) * (PK_TX + 2);
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
