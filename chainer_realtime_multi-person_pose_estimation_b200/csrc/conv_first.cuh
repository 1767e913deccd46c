// conv_first.cuh -- conv1_1 (3 -> 64, 3x3, pad 1) + ReLU (models/CocoPoseNet.py:136) fused with
// preprocess (pose_detector.py:426-431: float32, /255, -0.5, BGR kept).  K = 27 is too small for
// a tensor-core tile and the layer is 0.17 % of the FLOPs: direct fp32 CUDA-core kernel writing
// NHWC fp16 (and the lo plane in parity mode).
//
// Block = 8 x 32 output pixels.  The uint8 halo tile is normalised once through a 256-entry
// look-up table (exactly float32(v)/255 - 0.5, as the reference computes it) into shared
// memory; each thread produces 2 horizontally adjacent pixels x 32 output channels so that
// every broadcast weight load (LDS.128) feeds 8 FMAs.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

constexpr int CF_TH = 8, CF_TW = 32;

// wt: [27][64] fp32 (k = (r*3+s)*3 + c), bias [64].
// x_u8: [N][H][W][3] uint8 BGR, or x_f32: [N][3][H][W] float32 (already preprocessed).
__global__ void __launch_bounds__(256, 2)
conv_first_kernel(const uint8_t* __restrict__ x_u8, const float* __restrict__ x_f32, const float* __restrict__ wt,
                  const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W, int cstride,
                  int lo_off) {
  __shared__ __align__(16) float s_w[27 * 64];
  __shared__ float s_b[64];
  __shared__ float s_lut[256];
  __shared__ float s_in[(CF_TH + 2) * (CF_TW + 2) * 3];
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * 64; i += 256) s_w[i] = wt[i];
  if (tid < 64) s_b[tid] = bias[tid];
  s_lut[tid] = __fsub_rn(__fdiv_rn(static_cast<float>(tid), 255.f), 0.5f);
  const int tiles_x = (W + CF_TW - 1) / CF_TW, tiles_y = (H + CF_TH - 1) / CF_TH;
  const int half = tid >> 7;           // output channels half*32 .. +31; warp-uniform so that every
                                       // weight LDS.128 is a single broadcast wavefront
  const int pair = tid & 127;          // 0..127
  const int py = pair >> 4;            // 0..7
  const int px = (pair & 15) * 2;      // 0,2,..,30
  const int total_tiles = N * tiles_y * tiles_x;
  constexpr int IN_ELEMS = (CF_TH + 2) * (CF_TW + 2) * 3;
  constexpr int PER_THREAD = (IN_ELEMS + 255) / 256;
  float pre[PER_THREAD];   // next tile's (normalised) inputs, in flight while this tile computes
  auto prefetch = [&](int tile) {
    const int n = tile / (tiles_y * tiles_x);
    const int rem = tile - n * (tiles_y * tiles_x);
    const int y0 = (rem / tiles_x) * CF_TH, x0 = (rem % tiles_x) * CF_TW;
#pragma unroll
    for (int e = 0; e < PER_THREAD; ++e) {
      const int i = tid + e * 256;
      float v = 0.f;
      if (i < IN_ELEMS) {
        const int c = i % 3;
        const int q = i / 3;
        const int xx = x0 - 1 + q % (CF_TW + 2), yy = y0 - 1 + q / (CF_TW + 2);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          if (x_u8) v = __int_as_float(static_cast<int>(x_u8[((static_cast<size_t>(n) * H + yy) * W + xx) * 3 + c]));
          else v = x_f32[((static_cast<size_t>(n) * 3 + c) * H + yy) * W + xx];
        } else {
          v = x_u8 ? __int_as_float(-1) : 0.f;   // -1 marks zero padding on the uint8 path
        }
      }
      pre[e] = v;
    }
  };
  if (static_cast<int>(blockIdx.x) < total_tiles) prefetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_y * tiles_x);
    const int rem = tile - n * (tiles_y * tiles_x);
    const int y0 = (rem / tiles_x) * CF_TH, x0 = (rem % tiles_x) * CF_TW;
    __syncthreads();   // LUT / weights ready; previous tile's s_in consumed
#pragma unroll
    for (int e = 0; e < PER_THREAD; ++e) {
      const int i = tid + e * 256;
      if (i < IN_ELEMS) {
        float v = pre[e];
        if (x_u8) { const int b = __float_as_int(v); v = (b < 0) ? 0.f : s_lut[b]; }
        s_in[i] = v;
      }
    }
    __syncthreads();
    if (tile + static_cast<int>(gridDim.x) < total_tiles) prefetch(tile + gridDim.x);
    float in[3][4][3];   // rows py..py+2, cols px..px+3 of the halo tile
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) in[r][q][c] = s_in[((py + r) * (CF_TW + 2) + px + q) * 3 + c];
    float acc0[32], acc1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { acc0[j] = s_b[half * 32 + j]; acc1[j] = acc0[j]; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float a0 = in[r][s][c], a1 = in[r][s + 1][c];
          const float4* w4 = reinterpret_cast<const float4*>(s_w + ((r * 3 + s) * 3 + c) * 64 + half * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 wv = w4[j];
            acc0[4 * j + 0] = fmaf(a0, wv.x, acc0[4 * j + 0]);
            acc0[4 * j + 1] = fmaf(a0, wv.y, acc0[4 * j + 1]);
            acc0[4 * j + 2] = fmaf(a0, wv.z, acc0[4 * j + 2]);
            acc0[4 * j + 3] = fmaf(a0, wv.w, acc0[4 * j + 3]);
            acc1[4 * j + 0] = fmaf(a1, wv.x, acc1[4 * j + 0]);
            acc1[4 * j + 1] = fmaf(a1, wv.y, acc1[4 * j + 1]);
            acc1[4 * j + 2] = fmaf(a1, wv.z, acc1[4 * j + 2]);
            acc1[4 * j + 3] = fmaf(a1, wv.w, acc1[4 * j + 3]);
          }
        }
      }
    }
    const int y = y0 + py;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int x = x0 + px + p;
      if (y >= H || x >= W) continue;
      const float* acc = p ? acc1 : acc0;
      __half* o = out + ((static_cast<size_t>(n) * H + y) * W + x) * cstride + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fmaxf(acc[g * 8 + j], 0.f);
          hi[j] = __float2half_rn(v);
          lo[j] = __float2half_rn(v - __half2float(hi[j]));
        }
        *reinterpret_cast<uint4*>(o + g * 8) = *reinterpret_cast<const uint4*>(hi);
        if (lo_off) *reinterpret_cast<uint4*>(o + lo_off + g * 8) = *reinterpret_cast<const uint4*>(lo);
      }
    }
  }
}

}  // namespace opb
