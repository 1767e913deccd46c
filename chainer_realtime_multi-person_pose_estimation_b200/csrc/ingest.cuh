// ingest.cuh -- device-side image ingest: cv2.resize(orig_img, (input_w, input_h)) of
// pose_detector.py:493 (default INTER_LINEAR on uint8 BGR), bit-exact.
//
// OpenCV's 8-bit path (imgproc/resize.cpp) is fixed point: per axis
//   f = float((d + 0.5) * scale - 0.5), s = floor(f), f -= s        (scale = 1 / (dst/src) in double)
//   columns clamp (s < 0 -> s = 0, f = 0;  s >= w-1 -> s = w-1, f = 0); rows keep f and clip the two
//   source rows separately;  coefficients = rint((1-f)*2048), rint(f*2048)  (int16)
//   horizontal   D = S[s]*a0 + S[s+1]*a1                                   (int32)
//   vertical     dst = (((b0*(D0>>4))>>16) + ((b1*(D1>>4))>>16) + 2) >> 2
// Every float/double step below is a single IEEE operation, so the device reproduces the host bits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

__device__ __forceinline__ void cv_linear_axis(int d, double scale, int src_n, bool clamp, int& s0, int& s1, int& c0,
                                               int& c1) {
  float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (clamp) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
  }
  c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  s0 = min(max(s, 0), src_n - 1);
  s1 = min(max(s + 1, 0), src_n - 1);
}

// src [n][h0][w0][3] uint8 -> dst [n][h][w][3] uint8; grid (ceil(w/32), ceil(h/8), n), block (32, 8)
__global__ void __launch_bounds__(256)
resize_linear_u8_kernel(const uint8_t* __restrict__ src, int h0, int w0, uint8_t* __restrict__ dst, int h, int w,
                        double scale_x, double scale_y) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, n = blockIdx.z;
  if (x >= w || y >= h) return;
  int sx0, sx1, a0, a1, sy0, sy1, b0, b1;
  cv_linear_axis(x, scale_x, w0, true, sx0, sx1, a0, a1);
  cv_linear_axis(y, scale_y, h0, false, sy0, sy1, b0, b1);
  const uint8_t* r0 = src + (static_cast<size_t>(n) * h0 + sy0) * w0 * 3;
  const uint8_t* r1 = src + (static_cast<size_t>(n) * h0 + sy1) * w0 * 3;
  uint8_t* o = dst + ((static_cast<size_t>(n) * h + y) * w + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int d0 = r0[sx0 * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
    const int d1 = r1[sx0 * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
    o[c] = static_cast<uint8_t>((((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2);
  }
}

// pad_image (pose_detector.py:46-55): copy src [h][w][3] into the top-left corner of dst [ph][pw][3] and fill the
// bottom / right margin with the per-channel pad value (104, 117, 123 in detect_precise, :445).
__global__ void __launch_bounds__(256)
pad_image_u8_kernel(const uint8_t* __restrict__ src, int h, int w, uint8_t* __restrict__ dst, int ph, int pw, int v0,
                    int v1, int v2) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= pw || y >= ph) return;
  uint8_t* o = dst + (static_cast<size_t>(y) * pw + x) * 3;
  if (x < w && y < h) {
    const uint8_t* i = src + (static_cast<size_t>(y) * w + x) * 3;
    o[0] = i[0]; o[1] = i[1]; o[2] = i[2];
  } else {
    o[0] = static_cast<uint8_t>(v0); o[1] = static_cast<uint8_t>(v1); o[2] = static_cast<uint8_t>(v2);
  }
}

}  // namespace opb
