// conv_first.cuh -- conv1_1 (3 -> 64, 3x3, pad 1) + ReLU (models/CocoPoseNet.py:136) fused with
// preprocess (pose_detector.py:426-431: float32, /255, -0.5, BGR kept).  K = 27 is too small for
// a tensor-core tile and the layer is 0.17 % of the FLOPs: direct fp32 CUDA-core kernel writing
// NHWC fp16 (and the lo plane in parity mode).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

// wt: [27][64] fp32 (k = (r*3+s)*3 + c), bias [64].  One thread = one pixel x 16 output channels.
// x_u8: [N][H][W][3] uint8 BGR, or x_f32: [N][3][H][W] float32 (already preprocessed).
__global__ void __launch_bounds__(256)
conv_first_kernel(const uint8_t* __restrict__ x_u8, const float* __restrict__ x_f32, const float* __restrict__ wt,
                  const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W, int cstride,
                  int lo_off) {
  __shared__ float s_w[27 * 64];
  __shared__ float s_b[64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) s_w[i] = wt[i];
  if (threadIdx.x < 64) s_b[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int cg = threadIdx.x >> 6;                 // 0..3 -> channels cg*16 .. +15
  const int lane_pix = threadIdx.x & 63;
  const size_t total = static_cast<size_t>(N) * H * W;
  for (size_t pix = static_cast<size_t>(blockIdx.x) * 64 + lane_pix; pix < total;
       pix += static_cast<size_t>(gridDim.x) * 64) {
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    float in[27];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yy = y + r - 1, xx = x + s - 1;
        const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = 0.f;
          if (ok) {
            if (x_u8) {
              const float t = static_cast<float>(x_u8[((static_cast<size_t>(n) * H + yy) * W + xx) * 3 + c]);
              v = __fsub_rn(__fdiv_rn(t, 255.f), 0.5f);
            } else {
              v = x_f32[((static_cast<size_t>(n) * 3 + c) * H + yy) * W + xx];
            }
          }
          in[(r * 3 + s) * 3 + c] = v;
        }
      }
    }
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = s_b[cg * 16 + j];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4* w4 = reinterpret_cast<const float4*>(s_w + k * 64 + cg * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 wv = w4[j];
        acc[4 * j + 0] = fmaf(in[k], wv.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(in[k], wv.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(in[k], wv.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(in[k], wv.w, acc[4 * j + 3]);
      }
    }
    __align__(16) __half hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = fmaxf(acc[j], 0.f);
      hi[j] = __float2half_rn(v);
      lo[j] = __float2half_rn(v - __half2float(hi[j]));
    }
    __half* o = out + pix * cstride + cg * 16;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(o + 8) = *reinterpret_cast<const uint4*>(hi + 8);
    if (lo_off) {
      *reinterpret_cast<uint4*>(o + lo_off) = *reinterpret_cast<const uint4*>(lo);
      *reinterpret_cast<uint4*>(o + lo_off + 8) = *reinterpret_cast<const uint4*>(lo + 8);
    }
  }
}

}  // namespace opb
