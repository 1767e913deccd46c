// peaks_sep.cuh -- compute_peaks_from_heatmaps (pose_detector.py:75-110) straight from the NETWORK-resolution heat
// maps, without ever forming the full-resolution map or its 21-tap smoothing pass per pixel.
//
// F.resize_images (:502) followed by gaussian_filter (:86) is a linear, separable operator of the low-resolution plane L:
//     S = Cy . L . Cx^T,    Cy[y][r] = sum_j g[j] * (bilinear weight of L-row r in map row reflect(y + j)),
// and because the map is ~7x larger than L while the Gaussian reaches 10 map pixels, every row of Cy (and Cx) has at most
// PKS_TAPS = 6 non-zero entries.  The CANDIDATE search therefore costs 6 + 6/7 fused multiply-adds per map pixel instead
// of 4 (bilinear) + 42 (two 21-tap passes): each thread owns one map column, keeps the six horizontally filtered
// L-values T[r][x] its current rows depend on in registers (a sliding window that advances once per ~7 rows) and walks
// down a 32-row cell; left / right neighbours come from warp shuffles (a warp covers 30 columns + 2 halo lanes).  Cells
// whose supporting L samples are all below the threshold are skipped (real heat maps are almost everywhere below it).
//
// Exactness is unchanged from peaks.cuh: the float32 operator value only SELECTS candidates (pixels that pass the peak
// test with a slack of 1e-5 * max|L|, far above the float32 error of twelve multiply-adds and of the reordering); every
// candidate is then re-evaluated with the reference's own operation sequence -- the bit-exact bilinear samples
// (ac_sample, upsample.cuh) of its (2R+3)^2 neighbourhood, scipy's float64 two-pass sums with the float32 store in
// between, strict '>' against the four neighbours (zero outside the image) and the float32 threshold -- by one warp.
// Results are bit-identical to smooth_nms_kernel on the materialised map (tests/test_gpu_postprocess_batch.py,
// tests/test_emu_postprocess.py).
#pragma once
#include "peaks.cuh"

namespace opb {

constexpr int PKS_TAPS = 6;          // non-zeros per row of the combined (bilinear o Gaussian) operator
constexpr int PKS_SEG = 32;          // rows walked between two candidate-evaluation phases
constexpr int PKS_MAX_WARPS = 10;

struct SepAxes {                     // device arrays, one float[8] record per output row / column:
  const float* wy;                   //   [0..5] weights, [6] = first L-row (int bits), [7] unused
  const float* wx;
};

// shared memory: [H] row records (float4 x 2) | candidate list | per-warp exact scratch
inline size_t smooth_nms_sep_smem_bytes(int H, int n_warps) {
  const int R = PK_R_FAST, P1W = 2 * R + 3;
  return static_cast<size_t>(H) * 32 + sizeof(int) * 30 * n_warps * PKS_SEG +
         static_cast<size_t>(n_warps) * (sizeof(double) * 3 * P1W + sizeof(float) * (P1W * P1W + 1));
}

// Work decomposition: persistent blocks; every warp pulls (plane, CELL) items (cell = 30 columns x PKS_SEG rows) from a
// global counter and processes each item on its own -- skip test, walk, exact re-evaluation of the cell's candidates --
// with no block-wide barrier after the row records (the same for every plane) are loaded.  (The first version walked a whole 300-column strip per block in lockstep:
// with 38 % of the cells of the benchmark's 8-person maps active, almost every block-wide segment had one active warp
// and the others waited at the barrier -- 44 % fewer instructions bought 16 % of the time.)
__global__ void __launch_bounds__(PKS_MAX_WARPS * 32)
smooth_nms_sep_kernel(const float* __restrict__ heat_lo, int c_total, int c_use, int n_planes, int h_lo, int w_lo, int H,
                      int W, GaussTaps taps, float thresh, SepAxes ax, PeakKey* __restrict__ out, int* __restrict__ counts,
                      int cap, int* __restrict__ next_item) {
  constexpr int R = PK_R_FAST, P1W = 2 * R + 3;
  const int n_warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ __align__(16) double smd[];
  float4* s_row = reinterpret_cast<float4*>(smd);                                   // [H][2]
  int* s_cand_all = reinterpret_cast<int*>(s_row + 2 * H);                          // [n_warps][30 * PKS_SEG]
  double* s_scr = reinterpret_cast<double*>(s_cand_all + 30 * n_warps * PKS_SEG + ((30 * n_warps * PKS_SEG) & 1));
  float* s_win = reinterpret_cast<float*>(s_scr + n_warps * 3 * P1W);               // [n_warps][P1W * P1W + 1]
  __shared__ int s_cnt[PKS_MAX_WARPS];

  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) s_row[i] = __ldg(reinterpret_cast<const float4*>(ax.wy) + i);
  if (lane == 0) s_cnt[warp] = 0;
  __syncthreads();

  const float skip_below = (thresh > 0.f) ? thresh * 0.999f : thresh * 1.001f - 1e-30f;
  const int n_cg = (W + 29) / 30, n_seg = (H + PKS_SEG - 1) / PKS_SEG, n_cells = n_cg * n_seg;
  int* s_cand = s_cand_all + warp * 30 * PKS_SEG;
  double* scr = s_scr + warp * 3 * P1W;
  float* win = s_win + warp * (P1W * P1W + 1);
  const double sy = ac_step(h_lo, H), sx = ac_step(w_lo, W);

  const long long n_items = static_cast<long long>(n_planes) * n_cells;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(next_item, 1);            // persistent blocks: (plane, cell) items from a global counter
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int plane = item / n_cells, cell = item - plane * n_cells;
    const int img = plane / c_use, c = plane - img * c_use;
    const float* __restrict__ L = heat_lo + (static_cast<size_t>(img) * c_total + c) * h_lo * w_lo;
    const int seg = cell / n_cg, cg = cell - seg * n_cg;     // consecutive items: neighbouring column groups of one segment
    const int y0 = seg * PKS_SEG, y1 = min(y0 + PKS_SEG, H);
    const int x_cell = cg * 30;
    const int x = x_cell + lane - 1;                          // lanes 0 and 31 are halo columns

    // ---- the L samples this cell (and the one-pixel ring around it whose values enter the peak test) depends on:
    // skip test and the candidate slack.  A pixel can only be a candidate if its own operator value reaches the
    // threshold, and that value is a convex combination (taps and bilinear weights >= 0, sum 1 up to rounding) of the L
    // samples under its 6 x 6 support.
    float delta;
    {
      const int ya = max(y0 - 1, 0), yb = min(y1, H - 1), xa = max(x_cell - 1, 0), xb = min(x_cell + 30, W - 1);
      const int r0 = __float_as_int(s_row[2 * ya + 1].z), r1 = min(__float_as_int(s_row[2 * yb + 1].z) + PKS_TAPS - 1, h_lo - 1);
      const int c0 = __float_as_int(__ldg(ax.wx + 8 * xa + 6));
      const int c1 = min(__float_as_int(__ldg(ax.wx + 8 * xb + 6)) + PKS_TAPS - 1, w_lo - 1);
      const int ncol = c1 - c0 + 1, nl = (r1 - r0 + 1) * ncol;
      float vmax = -3.0e38f, amax = 0.f;
      for (int i = lane; i < nl; i += 32) {
        const float v = __ldg(L + (r0 + i / ncol) * w_lo + c0 + i % ncol);
        vmax = fmaxf(vmax, v);
        amax = fmaxf(amax, fabsf(v));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      }
      if (thresh > 0.f && !(vmax > skip_below)) continue;     // warp-uniform
      // slack: 1e-5 * max|L| of the support, far above the float32 error of twelve multiply-adds and of the reordering
      delta = 1e-5f * amax + 1e-30f;
    }

    // ---- this thread's column operator
    const bool in_img = (x >= 0) && (x < W);
    const int xc = min(max(x, 0), W - 1);
    float wxr[PKS_TAPS];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ax.wx) + 2 * xc), b = __ldg(reinterpret_cast<const float4*>(ax.wx) + 2 * xc + 1);
      wxr[0] = a.x; wxr[1] = a.y; wxr[2] = a.z; wxr[3] = a.w; wxr[4] = b.x; wxr[5] = b.y;
    }
    const float* __restrict__ Lx = L + __float_as_int(__ldg(ax.wx + 8 * xc + 6));
    auto t_of = [&](int r) {                                  // T[r][x] = sum_c Cx[x][c] L[r][c]
      const float* q = Lx + r * w_lo;
      float acc = wxr[0] * __ldg(q);
#pragma unroll
      for (int k = 1; k < PKS_TAPS; ++k) acc = fmaf(wxr[k], __ldg(q + k), acc);
      return acc;
    };
    float t[PKS_TAPS];
    const int ys = max(y0 - 1, 0);
    int cur = __float_as_int(s_row[2 * ys + 1].z);            // first L-row under the sliding window
#pragma unroll
    for (int k = 0; k < PKS_TAPS; ++k) t[k] = t_of(cur + k);
    auto s_of = [&](int y) {                                  // S[y][x] (float32 operator value)
      const float4 a = s_row[2 * y], b = s_row[2 * y + 1];
      const int nb = __float_as_int(b.z);
      while (cur < nb) {                                      // warp-uniform: the window moves down one L-row
#pragma unroll
        for (int k = 0; k < PKS_TAPS - 1; ++k) t[k] = t[k + 1];
        ++cur;
        t[PKS_TAPS - 1] = t_of(cur + PKS_TAPS - 1);
      }
      float acc = a.x * t[0];
      acc = fmaf(a.y, t[1], acc);
      acc = fmaf(a.z, t[2], acc);
      acc = fmaf(a.w, t[3], acc);
      acc = fmaf(b.x, t[4], acc);
      acc = fmaf(b.y, t[5], acc);
      return acc;
    };
    float s_up = 0.f, s_c;
    if (y0 > 0) { s_up = s_of(y0 - 1); s_c = s_of(y0); }
    else s_c = s_of(0);
    const bool owner = in_img && lane >= 1 && lane <= 30;
    const float d2 = 2.f * delta;
    for (int y = y0; y < y1; ++y) {
      const float s_dn = (y + 1 < H) ? s_of(y + 1) : 0.f;
      float lf = __shfl_up_sync(0xffffffffu, s_c, 1), rt = __shfl_down_sync(0xffffffffu, s_c, 1);
      if (x == 0) lf = 0.f;
      if (x == W - 1) rt = 0.f;
      const float up = (y > 0) ? s_up : 0.f;
      if (owner && s_c > thresh - delta && s_c > up - d2 && s_c > s_dn - d2 && s_c > lf - d2 && s_c > rt - d2)
        s_cand[atomicAdd(&s_cnt[warp], 1)] = (y << 16) | (x - x_cell);
      s_up = s_c;
      s_c = s_dn;
    }
    __syncwarp();
    // ---- exact re-evaluation of this cell's candidates, the whole warp per candidate
    const int ncand = s_cnt[warp];
    for (int ci = 0; ci < ncand; ++ci) {
      const int code = s_cand[ci];
      const int py = code >> 16, px = x_cell + (code & 0xffff);
      // the (2R+3)^2 neighbourhood of the upsampled map, bit-exact (rows py-1-R.., columns px-1-R.., reflected)
      for (int e = lane; e < P1W * P1W; e += 32) {
        const int i = e / P1W, j = e - i * P1W;
        const AcAxis ay = ac_axis_frac(reflect_index(py - 1 - R + i, H), h_lo, H, sy);
        const AcAxis axx = ac_axis_frac(reflect_index(px - 1 - R + j, W), w_lo, W, sx);
        win[e] = ac_sample(L, w_lo, ac_tap(axx, ay, w_lo));
      }
      __syncwarp();
      // exact axis-0 pass for map rows py-1, py, py+1 (scipy's order, float32 store between the passes)
      for (int e = lane; e < 3 * P1W; e += 32) {
        const int rr = e / P1W, cc = e - rr * P1W;
        const float* col = win + (rr + R) * P1W + cc;
        double acc = __dmul_rn(static_cast<double>(col[0]), taps.w[R]);
        for (int j = -R; j < 0; ++j) {
          const double pair = __dadd_rn(static_cast<double>(col[j * P1W]), static_cast<double>(col[-j * P1W]));
          acc = __dadd_rn(acc, __dmul_rn(pair, taps.w[R + j]));
        }
        scr[e] = static_cast<double>(static_cast<float>(acc));
      }
      __syncwarp();
      float val = 0.f;
      if (lane < 5) {   // centre, up, down, left, right -> (row in scr, first column in scr)
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
        if (py == 0) up = 0.f;
        if (py == H - 1) dn = 0.f;
        if (px == 0) lf = 0.f;
        if (px == W - 1) rt = 0.f;
        if (g > thresh && g > up && g > dn && g > lf && g > rt) {
          const int slot = atomicAdd(&counts[img], 1);
          if (slot < cap) {
            PeakKey k;
            k.key = static_cast<uint32_t>((static_cast<size_t>(c) * H + py) * W + px);
            k.score = g;
            out[static_cast<size_t>(img) * cap + slot] = k;
          }
        }
      }
      __syncwarp();
    }
    if (lane == 0) s_cnt[warp] = 0;
    __syncwarp();
  }
}

}  // namespace opb
