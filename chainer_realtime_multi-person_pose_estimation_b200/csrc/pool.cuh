// pool.cuh -- F.max_pooling_2d(h, ksize=2, stride=2) (models/CocoPoseNet.py:138,141,146) on
// NHWC fp16 activations; in parity mode the value is hi+lo and both planes are carried.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace opb {

// in [N][H][W][cstride] -> out [N][H/2][W/2][cstride]; C valid channels (multiple of 8);
// lo_off = 0 (fast) or channel offset of the lo plane (parity)
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W, int C, int cstride,
                  int lo_off) {
  const int Ho = H >> 1, Wo = W >> 1;
  const int vec_per_pix = C >> 3;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * vec_per_pix;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_pix);
    size_t pix = i / vec_per_pix;
    const int xo = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int yo = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    const __half* base = in + ((static_cast<size_t>(n) * H + 2 * yo) * W + 2 * xo) * cstride + v * 8;
    const size_t dx = cstride, dy = static_cast<size_t>(W) * cstride;
    __half* o = out + ((static_cast<size_t>(n) * Ho + yo) * Wo + xo) * cstride + v * 8;
    if (lo_off == 0) {
      uint4 a = *reinterpret_cast<const uint4*>(base), b = *reinterpret_cast<const uint4*>(base + dx);
      uint4 c = *reinterpret_cast<const uint4*>(base + dy), d = *reinterpret_cast<const uint4*>(base + dy + dx);
      uint4 r;
      __half2* rh = reinterpret_cast<__half2*>(&r);
      const __half2 *ah = reinterpret_cast<const __half2*>(&a), *bh = reinterpret_cast<const __half2*>(&b),
                    *ch = reinterpret_cast<const __half2*>(&c), *dh = reinterpret_cast<const __half2*>(&d);
#pragma unroll
      for (int k = 0; k < 4; ++k) rh[k] = __hmax2(__hmax2(ah[k], bh[k]), __hmax2(ch[k], dh[k]));
      *reinterpret_cast<uint4*>(o) = r;
    } else {
      __align__(16) __half hi[4][8], lo[4][8];
      const size_t offs[4] = {0, dx, dy, dy + dx};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<uint4*>(hi[q]) = *reinterpret_cast<const uint4*>(base + offs[q]);
        *reinterpret_cast<uint4*>(lo[q]) = *reinterpret_cast<const uint4*>(base + offs[q] + lo_off);
      }
      __align__(16) __half rh[8], rl[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int best = 0;
        float bv = __half2float(hi[0][k]) + __half2float(lo[0][k]);
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float t = __half2float(hi[q][k]) + __half2float(lo[q][k]);
          if (t > bv) { bv = t; best = q; }
        }
        rh[k] = hi[best][k];
        rl[k] = lo[best][k];
      }
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(rh);
      *reinterpret_cast<uint4*>(o + lo_off) = *reinterpret_cast<const uint4*>(rl);
    }
  }
}

}  // namespace opb
