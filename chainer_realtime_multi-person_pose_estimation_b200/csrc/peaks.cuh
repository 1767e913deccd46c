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
template <int RT>
__global__ void __launch_bounds__(256)   // launched with PK_THREADS threads
smooth_nms_kernel(const float* __restrict__ heat, int c_total, int c_use, int H, int W, GaussTaps taps,
                  float thresh, PeakKey* __restrict__ out, int* __restrict__ counts, int cap,
                  const float* __restrict__ tile_max) {
  const int R = (RT > 0) ? RT : taps.radius;
  const float skip_below = (thresh > 0.f) ? thresh * 0.999f : thresh * 1.001f - 1e-30f;
  if (tile_max != nullptr) {
    // cells overlapping the (clamped) input window [y0-1-R, y0+PK_TY+R] x [x0-1-R, x0+PK_TX+R]; reflected
    // border samples are copies of in-image pixels within R of the border, i.e. inside the clamped window
    const int cells_y = (H + PK_CELL - 1) / PK_CELL, cells_x = (W + PK_CELL - 1) / PK_CELL;
    const int wy0 = max(static_cast<int>(blockIdx.y) * PK_TY - 1 - R, 0), wy1 = min(static_cast<int>(blockIdx.y) * PK_TY + PK_TY + R, H - 1);
    const int wx0 = max(static_cast<int>(blockIdx.x) * PK_TX - 1 - R, 0), wx1 = min(static_cast<int>(blockIdx.x) * PK_TX + PK_TX + R, W - 1);
    const int cy0 = wy0 / PK_CELL, cy1 = wy1 / PK_CELL, cx0 = wx0 / PK_CELL, cx1 = wx1 / PK_CELL;
    const int ncx = cx1 - cx0 + 1, ncell = (cy1 - cy0 + 1) * ncx;
    const float* cm = tile_max + static_cast<size_t>(blockIdx.z) * cells_y * cells_x;
    float m = -3.0e38f;
    for (int i = threadIdx.x; i < ncell; i += blockDim.x)
      m = fmaxf(m, __ldg(cm + (cy0 + i / ncx) * cells_x + cx0 + i % ncx));
    if (!__syncthreads_or(m > skip_below)) return;
  }
  const int IN_W = PK_TX + 2 + 2 * R, IN_H = PK_TY + 2 + 2 * R;
  const int O_H = PK_TY + 2, O_W = PK_TX + 2;
  const int P1W = 2 * R + 3;                 // exact pass-1 values needed per candidate row
  extern __shared__ double smd[];
  double* s_scr = smd;                                        // [8 warps][3][P1W] exact pass-1 scratch
  float* s_in = reinterpret_cast<float*>(s_scr + 8 * 3 * P1W);   // [IN_H][IN_W]
  float* s_1 = s_in + IN_H * IN_W;                            // [O_H][IN_W]  float32 approx, axis 0
  float* s_2 = s_1 + O_H * IN_W;                              // [O_H][O_W]   float32 approx, axis 1
  int* s_cand = reinterpret_cast<int*>(s_2 + O_H * O_W);      // [PK_TY*PK_TX]
  __shared__ int s_ncand;
  __shared__ unsigned int s_absmax;
  __shared__ float s_tapf[2 * PK_R_MAX + 1];

  const int plane = blockIdx.z;
  const int img = plane / c_use, c = plane - img * c_use;
  const int x0 = blockIdx.x * PK_TX, y0 = blockIdx.y * PK_TY;
  const float* src = heat + (static_cast<size_t>(img) * c_total + c) * H * W;
  if (threadIdx.x == 0) { s_ncand = 0; s_absmax = 0u; }
  if (threadIdx.x < 2 * R + 1) s_tapf[threadIdx.x] = static_cast<float>(taps.w[threadIdx.x]);
  __syncthreads();

  float vmax = -3.0e38f, amax = 0.f;
  {  // 2 rows x 128 columns per sweep: the column index (and its reflection) is fixed per thread
    const int col = threadIdx.x & 127, rsub = threadIdx.x >> 7;
    if (col < IN_W) {
      const int gx = reflect_index(x0 - 1 - R + col, W);
      for (int r = rsub; r < IN_H; r += (blockDim.x >> 7)) {
        const int gy = reflect_index(y0 - 1 - R + r, H);
        const float v = __ldg(src + static_cast<size_t>(gy) * W + gx);
        vmax = fmaxf(vmax, v);
        amax = fmaxf(amax, fabsf(v));
        s_in[r * IN_W + col] = v;
      }
    }
  }
  atomicMax(&s_absmax, __float_as_uint(amax));
  if (!__syncthreads_or(vmax > skip_below)) return;
  const float delta = 1e-5f * __uint_as_float(s_absmax) + 1e-30f;

  // float32 approximation, axis 0 then axis 1.  Register sliding windows: a thread owns one
  // column (axis 0) / one 8-output row segment (axis 1), so each input is loaded once.
  if (RT == PK_R_FAST) {
    constexpr int R2 = 2 * PK_R_FAST + 1;
    float tp[R2];
#pragma unroll
    for (int k = 0; k < R2; ++k) tp[k] = s_tapf[k];
    if (threadIdx.x < IN_W) {
      const float* col = s_in + threadIdx.x;
      float win[R2];
#pragma unroll
      for (int k = 0; k < R2 - 1; ++k) win[k] = col[k * IN_W];
#pragma unroll
      for (int r = 0; r < PK_TY + 2; ++r) {
        win[(r + R2 - 1) % R2] = col[(r + R2 - 1) * IN_W];
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < R2; ++k) acc = fmaf(win[(r + k) % R2], tp[k], acc);
        s_1[r * IN_W + threadIdx.x] = acc;
      }
    }
    __syncthreads();
    constexpr int SEG = 8, NSEG = (PK_TX + 2 + SEG - 1) / SEG;
    for (int t = threadIdx.x; t < (PK_TY + 2) * NSEG; t += blockDim.x) {
      const int r = t / NSEG, q0 = (t - r * NSEG) * SEG;
      const float* p = s_1 + r * IN_W + q0;
      float win[R2 + SEG - 1];
#pragma unroll
      for (int k = 0; k < R2 + SEG - 1; ++k) win[k] = (q0 + k < IN_W) ? p[k] : 0.f;
#pragma unroll
      for (int o = 0; o < SEG; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < R2; ++k) acc = fmaf(win[o + k], tp[k], acc);
        if (q0 + o < O_W) s_2[r * O_W + q0 + o] = acc;
      }
    }
  } else {
    for (int i = threadIdx.x; i < O_H * IN_W; i += blockDim.x) {
      const int r = i / IN_W, q = i - r * IN_W;
      const float* col = s_in + r * IN_W + q;
      float acc = 0.f;
      for (int k = 0; k <= 2 * R; ++k) acc = fmaf(col[k * IN_W], s_tapf[k], acc);
      s_1[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < O_H * O_W; i += blockDim.x) {
      const int r = i / O_W, q = i - r * O_W;
      const float* p = s_1 + r * IN_W + q;
      float acc = 0.f;
      for (int k = 0; k <= 2 * R; ++k) acc = fmaf(p[k], s_tapf[k], acc);
      s_2[i] = acc;
    }
  }
  __syncthreads();
  // candidates
  for (int i = threadIdx.x; i < PK_TY * PK_TX; i += blockDim.x) {
    const int r = i / PK_TX, q = i - r * PK_TX;
    const int y = y0 + r, x = x0 + q;
    if (y >= H || x >= W) continue;
    const float g = s_2[(r + 1) * O_W + q + 1];
    if (!(g > thresh - delta)) continue;
    const float up = (y > 0) ? s_2[r * O_W + q + 1] : 0.f;
    const float dn = (y < H - 1) ? s_2[(r + 2) * O_W + q + 1] : 0.f;
    const float lf = (x > 0) ? s_2[(r + 1) * O_W + q] : 0.f;
    const float rt = (x < W - 1) ? s_2[(r + 1) * O_W + q + 2] : 0.f;
    const float d2 = 2.f * delta;
    if (g > up - d2 && g > dn - d2 && g > lf - d2 && g > rt - d2) s_cand[atomicAdd(&s_ncand, 1)] = i;
  }
  __syncthreads();
  // exact re-evaluation, one warp per candidate
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* scr = s_scr + warp * 3 * P1W;
  const int ncand = s_ncand;
  for (int ci = warp; ci < ncand; ci += (blockDim.x >> 5)) {
    const int i = s_cand[ci];
    const int r = i / PK_TX, q = i - r * PK_TX;
    // exact pass-1 values P1[rr][cc]: O-rows r..r+2 (s_in rows +R), s_in columns q .. q+2R+2
    for (int e = lane; e < 3 * P1W; e += 32) {
      const int rr = e / P1W, cc = e - rr * P1W;
      const float* col = s_in + (r + rr + R) * IN_W + q + cc;
      double acc = __dmul_rn(static_cast<double>(col[0]), taps.w[R]);
      for (int j = -R; j < 0; ++j) {
        const double pair = __dadd_rn(static_cast<double>(col[j * IN_W]), static_cast<double>(col[-j * IN_W]));
        acc = __dadd_rn(acc, __dmul_rn(pair, taps.w[R + j]));
      }
      scr[e] = static_cast<double>(static_cast<float>(acc));   // float32 store between the passes
    }
    __syncwarp();
    // lanes 0..4: centre, up, down, left, right -> (row in scr, first column in scr)
    float val = 0.f;
    if (lane < 5) {
      const int rr = (lane == 1) ? 0 : (lane == 2) ? 2 : 1;
      const int c0 = (lane == 3) ? 0 : (lane == 4) ? 2 : 1;
      const double* p = scr + rr * P1W + c0 + R;
      double acc = __dmul_rn(p[0], taps.w[R]);
      for (int j = -R; j < 0; ++j) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(p[j], p[-j]), taps.w[R + j]));
      val = static_cast<float>(acc);
    }
    const float g = __shfl_sync(0xffffffffu, val, 0);
    float up = __shfl_sync(0xffffffffu, val, 1), dn = __shfl_sync(0xffffffffu, val, 2);
    float lf = __shfl_sync(0xffffffffu, val, 3), rt = __shfl_sync(0xffffffffu, val, 4);
    if (lane == 0) {
      const int y = y0 + r, x = x0 + q;
      if (y == 0) up = 0.f;
      if (y == H - 1) dn = 0.f;
      if (x == 0) lf = 0.f;
      if (x == W - 1) rt = 0.f;
      if (g > thresh && g > up && g > dn && g > lf && g > rt) {
        const int slot = atomicAdd(&counts[img], 1);
        if (slot < cap) {
          PeakKey k;
          k.key = static_cast<uint32_t>((static_cast<size_t>(c) * H + y) * W + x);
          k.score = g;
          out[static_cast<size_t>(img) * cap + slot] = k;
        }
      }
    }
    __syncwarp();
  }
}

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
