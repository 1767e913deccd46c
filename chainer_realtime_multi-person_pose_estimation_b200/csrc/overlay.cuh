// overlay.cuh -- draw_person_pose (pose_detector.py:520-553) on the device, bit-exact with OpenCV's rasteriser.
//
// The reference draws, per person, 17 limbs with cv2.line(canvas, p1, p2, color, 2) (ear-shoulder limbs 9 and 13 are
// skipped) and then, per person, up to 18 joints with cv2.circle(canvas, (x, y), 3, color, -1); later primitives
// overwrite earlier ones.  OpenCV (imgproc/drawing.cpp, 4.13) rasterises them as
//   line, thickness 2, LINE_8   ThickLine: half-width vector dp = (cvRound(dy*r), cvRound(dx*r)), r = 2^16/|p1-p0| in
//                               16.16 fixed point; FillConvexPoly of the 4-corner polygon (its edges with the
//                               fixed-point DDA `Line2` after clipLine, then the scan-line fill with 16.16 edge
//                               walkers), then a filled circle of radius 1 at both end points;
//   circle, radius 3, filled    the midpoint circle's horizontal spans, clipped to the image.
// Here every primitive is rasterised by ONE thread with the same integer / double operation sequence (single IEEE
// operations, no contraction) into a per-pixel PRIORITY plane (atomicMax of the primitive's 1-based index in the
// reference's drawing order); a second kernel paints each covered pixel with the colour of its highest-priority
// primitive -- exactly the pixel the sequential overwrite order leaves.  All joints are inside the image by
// construction (they are scaled peak positions), which is the case OpenCV's own pre-clipping leaves untouched; the entry
// point rejects joints outside the image.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

constexpr int OV_SHIFT = 16;
constexpr long long OV_ONE = 1ll << OV_SHIFT;
constexpr int OV_LIMBS = 19, OV_JOINTS = 18;

struct OverlayTables {
  int limb_a[OV_LIMBS], limb_b[OV_LIMBS];      // params['limbs_point'] (entity.py:85-105)
  uint8_t limb_color[OV_LIMBS][3];             // pose_detector.py:524-529 (written to channels 0,1,2 as given)
  uint8_t joint_color[OV_JOINTS][3];           // :531-535
};

struct OvCanvas {
  unsigned int* prio;
  int H, W;
  unsigned int tag;
  __device__ __forceinline__ void put(long long x, long long y) const {
    if (x >= 0 && x < W && y >= 0 && y < H) atomicMax(prio + static_cast<size_t>(y) * W + x, tag);
  }
  __device__ __forceinline__ void hline(int y, int xa, int xb) const {   // caller guarantees 0 <= y < H and clipped xa..xb
    for (int x = xa; x <= xb; ++x) atomicMax(prio + static_cast<size_t>(y) * W + x, tag);
  }
};

// cv::clipLine(Size2l, Point2l&, Point2l&) on the 16.16-scaled image rectangle
__device__ inline bool ov_clip_line(long long Ws, long long Hs, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = Ws - 1, bottom = Hs - 1;
  if (Ws <= 0 || Hs <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - y1), static_cast<double>(x2 - x1)), static_cast<double>(y2 - y1)));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - y2), static_cast<double>(x2 - x1)), static_cast<double>(y2 - y1)));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - x1), static_cast<double>(y2 - y1)), static_cast<double>(x2 - x1)));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - x2), static_cast<double>(y2 - y1)), static_cast<double>(x2 - x1)));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// drawing.cpp Line2: fixed-point DDA between two 16.16 points
__device__ inline void ov_line2(const OvCanvas& cv, long long x1, long long y1, long long x2, long long y2) {
  if (!ov_clip_line(static_cast<long long>(cv.W) << OV_SHIFT, static_cast<long long>(cv.H) << OV_SHIFT, x1, y1, x2, y2)) return;
  long long dx = x2 - x1, dy = y2 - y1;
  const long long ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
  long long x_step, y_step;
  int ecount;
  if (ax > ay) {
    if (dx < 0) {
      dy = -dy;
      long long t = x1; x1 = x2; x2 = t;
      t = y1; y1 = y2; y2 = t;
    }
    x_step = OV_ONE;
    y_step = (dy * OV_ONE) / (ax | 1);       // (dy << XY_SHIFT) / (ax | 1), C++ truncating division
    ecount = static_cast<int>((x2 - x1) >> OV_SHIFT);
  } else {
    if (dy < 0) {
      dx = -dx;
      long long t = x1; x1 = x2; x2 = t;
      t = y1; y1 = y2; y2 = t;
    }
    x_step = (dx * OV_ONE) / (ay | 1);
    y_step = OV_ONE;
    ecount = static_cast<int>((y2 - y1) >> OV_SHIFT);
  }
  x1 += OV_ONE >> 1;
  y1 += OV_ONE >> 1;
  cv.put((x2 + (OV_ONE >> 1)) >> OV_SHIFT, (y2 + (OV_ONE >> 1)) >> OV_SHIFT);
  if (ax > ay) {
    x1 >>= OV_SHIFT;
    while (ecount >= 0) {
      cv.put(x1, y1 >> OV_SHIFT);
      x1++;
      y1 += y_step;
      ecount--;
    }
  } else {
    y1 >>= OV_SHIFT;
    while (ecount >= 0) {
      cv.put(x1 >> OV_SHIFT, y1);
      x1 += x_step;
      y1++;
      ecount--;
    }
  }
  (void)x_step;
}

// drawing.cpp FillConvexPoly(img, v, 4, color, LINE_8, XY_SHIFT)
__device__ inline void ov_fill_convex4(const OvCanvas& cv, const long long (&vx)[4], const long long (&vy)[4]) {
  constexpr int npts = 4;
  const long long delta = OV_ONE >> 1;
  long long xmin = vx[0], xmax = vx[0], ymin = vy[0], ymax = vy[0];
  int imin = 0;
  long long px = vx[npts - 1], py = vy[npts - 1];
  for (int i = 0; i < npts; ++i) {
    if (vy[i] < ymin) { ymin = vy[i]; imin = i; }
    ymax = vy[i] > ymax ? vy[i] : ymax;
    xmax = vx[i] > xmax ? vx[i] : xmax;
    xmin = vx[i] < xmin ? vx[i] : xmin;
    ov_line2(cv, px, py, vx[i], vy[i]);
    px = vx[i];
    py = vy[i];
  }
  xmin = (xmin + delta) >> OV_SHIFT;
  xmax = (xmax + delta) >> OV_SHIFT;
  ymin = (ymin + delta) >> OV_SHIFT;
  ymax = (ymax + delta) >> OV_SHIFT;
  if (xmax < 0 || ymax < 0 || xmin >= cv.W || ymin >= cv.H) return;
  if (ymax > cv.H - 1) ymax = cv.H - 1;
  int e_idx[2] = {imin, imin}, e_di[2] = {1, npts - 1}, e_ye[2] = {static_cast<int>(ymin), static_cast<int>(ymin)};
  long long e_x[2] = {-OV_ONE, -OV_ONE}, e_dx[2] = {0, 0};
  int y = static_cast<int>(ymin), edges = npts;
  do {
    for (int i = 0; i < 2; ++i) {
      if (y >= e_ye[i]) {
        int idx0 = e_idx[i];
        const int di = e_di[i];
        int idx = idx0 + di;
        if (idx >= npts) idx -= npts;
        for (; edges-- > 0;) {
          const int ty = static_cast<int>((vy[idx] + delta) >> OV_SHIFT);
          if (ty > y) {
            const long long xs = vx[idx0], xe = vx[idx];
            e_ye[i] = ty;
            e_dx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
            e_x[i] = xs;
            e_idx[i] = idx;
            break;
          }
          idx0 = idx;
          idx += di;
          if (idx >= npts) idx -= npts;
        }
      }
    }
    if (edges < 0) break;
    if (y >= 0) {
      int left = 0, right = 1;
      if (e_x[0] > e_x[1]) { left = 1; right = 0; }
      int xx1 = static_cast<int>((e_x[left] + delta) >> OV_SHIFT);
      int xx2 = static_cast<int>((e_x[right] + delta) >> OV_SHIFT);
      if (xx2 >= 0 && xx1 < cv.W) {
        if (xx1 < 0) xx1 = 0;
        if (xx2 >= cv.W) xx2 = cv.W - 1;
        cv.hline(y, xx1, xx2);
      }
    }
    e_x[0] += e_dx[0];
    e_x[1] += e_dx[1];
  } while (++y <= static_cast<int>(ymax));
}

// drawing.cpp Circle(img, center, radius, color, fill = 1)
__device__ inline void ov_circle_fill(const OvCanvas& cv, int cx, int cy, int radius) {
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  const int W = cv.W, H = cv.H;
  while (dx >= dy) {
    const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
    int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
    if (x11 < W && x12 >= 0 && y21 < H && y22 >= 0) {
      x11 = max(x11, 0);
      x12 = min(x12, W - 1);
      if (static_cast<unsigned>(y11) < static_cast<unsigned>(H)) cv.hline(y11, x11, x12);
      if (static_cast<unsigned>(y12) < static_cast<unsigned>(H)) cv.hline(y12, x11, x12);
      if (x21 < W && x22 >= 0) {
        x21 = max(x21, 0);
        x22 = min(x22, W - 1);
        if (static_cast<unsigned>(y21) < static_cast<unsigned>(H)) cv.hline(y21, x21, x22);
        if (static_cast<unsigned>(y22) < static_cast<unsigned>(H)) cv.hline(y22, x21, x22);
      }
    }
    dy++;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

// drawing.cpp ThickLine(img, p0, p1, color, thickness = 2, LINE_8, flags = 3, shift = 0)
__device__ inline void ov_thick_line2(const OvCanvas& cv, int x0, int y0, int x1, int y1) {
  const long long p0x = static_cast<long long>(x0) << OV_SHIFT, p0y = static_cast<long long>(y0) << OV_SHIFT;
  const long long p1x = static_cast<long long>(x1) << OV_SHIFT, p1y = static_cast<long long>(y1) << OV_SHIFT;
  const double inv = 1.0 / 65536.0;
  const double dx = __dmul_rn(static_cast<double>(p0x - p1x), inv), dy = __dmul_rn(static_cast<double>(p1y - p0y), inv);
  double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  const long long th = 2ll << (OV_SHIFT - 1);
  if (fabs(r) > 2.220446049250313e-16) {
    r = __ddiv_rn(static_cast<double>(th), __dsqrt_rn(r));
    const long long dpx = __double2ll_rn(__dmul_rn(dy, r)), dpy = __double2ll_rn(__dmul_rn(dx, r));
    const long long vx[4] = {p0x + dpx, p0x - dpx, p1x - dpx, p1x + dpx};
    const long long vy[4] = {p0y + dpy, p0y - dpy, p1y - dpy, p1y + dpy};
    ov_fill_convex4(cv, vx, vy);
  }
  const int rad = static_cast<int>((th + (OV_ONE >> 1)) >> OV_SHIFT);
  ov_circle_fill(cv, x0, y0, rad);
  ov_circle_fill(cv, x1, y1, rad);
}

// records of the pipeline -> the reference's `poses.round().astype('i')` (:513-514 then :539):
// x = rint(double(peak x) * sx), y likewise, v = 2 for a present joint; absent joints are (0, 0, 0)
__global__ void overlay_poses_from_records_kernel(const int* __restrict__ rec_x, const int* __restrict__ rec_y,
                                                  const int* __restrict__ rec_id, int person_stride_ints, int n_persons,
                                                  double sx, double sy, int* __restrict__ poses) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_persons * OV_JOINTS) return;
  const int p = i / OV_JOINTS, j = i - p * OV_JOINTS;
  const bool has = rec_id[static_cast<size_t>(p) * person_stride_ints + j] >= 0;
  poses[3 * i + 0] = has ? __double2int_rn(__dmul_rn(static_cast<double>(rec_x[static_cast<size_t>(p) * person_stride_ints + j]), sx)) : 0;
  poses[3 * i + 1] = has ? __double2int_rn(__dmul_rn(static_cast<double>(rec_y[static_cast<size_t>(p) * person_stride_ints + j]), sy)) : 0;
  poses[3 * i + 2] = has ? 2 : 0;
}

// one thread per potential primitive: [n_poses * 19 limbs | n_poses * 18 joints] in the reference's drawing order.
// `bad` is raised when a drawn joint lies outside the image.
__global__ void overlay_raster_kernel(const int* __restrict__ poses, int n_poses, OverlayTables tb, unsigned int* __restrict__ prio,
                                      int H, int W, int* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_limb = n_poses * OV_LIMBS, n_all = n_limb + n_poses * OV_JOINTS;
  if (i >= n_all) return;
  OvCanvas cv{prio, H, W, static_cast<unsigned int>(i) + 1u};
  if (i < n_limb) {
    const int p = i / OV_LIMBS, l = i - p * OV_LIMBS;
    if (l == 9 || l == 13) return;
    const int* a = poses + (static_cast<size_t>(p) * OV_JOINTS + tb.limb_a[l]) * 3;
    const int* b = poses + (static_cast<size_t>(p) * OV_JOINTS + tb.limb_b[l]) * 3;
    if (a[2] == 0 || b[2] == 0) return;
    if (a[0] < 0 || a[0] >= W || a[1] < 0 || a[1] >= H || b[0] < 0 || b[0] >= W || b[1] < 0 || b[1] >= H) { atomicOr(bad, 1); return; }
    ov_thick_line2(cv, a[0], a[1], b[0], b[1]);
  } else {
    const int k = i - n_limb;
    const int* a = poses + static_cast<size_t>(k) * 3;
    if (a[2] == 0) return;
    if (a[0] < 0 || a[0] >= W || a[1] < 0 || a[1] >= H) { atomicOr(bad, 1); return; }
    ov_circle_fill(cv, a[0], a[1], 3);
  }
}

// canvas = orig_img.copy(), then every covered pixel takes the colour of its last primitive
__global__ void __launch_bounds__(256)
overlay_paint_kernel(const uint8_t* __restrict__ img, const unsigned int* __restrict__ prio, int n_pix, int n_poses,
                     OverlayTables tb, uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const unsigned int t = prio[i];
  uint8_t c0 = img[3 * static_cast<size_t>(i)], c1 = img[3 * static_cast<size_t>(i) + 1], c2 = img[3 * static_cast<size_t>(i) + 2];
  if (t) {
    const int k = static_cast<int>(t) - 1, n_limb = n_poses * OV_LIMBS;
    const uint8_t* col = (k < n_limb) ? tb.limb_color[k % OV_LIMBS] : tb.joint_color[(k - n_limb) % OV_JOINTS];
    c0 = col[0]; c1 = col[1]; c2 = col[2];
  }
  out[3 * static_cast<size_t>(i)] = c0;
  out[3 * static_cast<size_t>(i) + 1] = c1;
  out[3 * static_cast<size_t>(i) + 2] = c2;
}

}  // namespace opb
