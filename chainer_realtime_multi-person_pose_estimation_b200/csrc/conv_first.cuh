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

#include "conv_tcgen05.cuh"

namespace opb {

constexpr int CF_TH = 8, CF_TW = 32;

// wt: [27][64] fp32 (k = (r*3+s)*3 + c), bias [64].
// x_u8: [N][H][W][3] uint8 BGR, or x_f32: [N][3][H][W] float32 (already preprocessed).
__global__ void __launch_bounds__(256, 2)
conv_first_kernel(const uint8_t* __restrict__ x_u8, const float* __restrict__ x_f32, const float* __restrict__ wt,
                  const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W, int cstride,
                  int lo_off, float u8_denom, int comp) {
  __shared__ __align__(16) float s_w[27 * 64];
  __shared__ float s_b[64];
  __shared__ float s_lut[256];
  __shared__ float s_in[(CF_TH + 2) * (CF_TW + 2) * 3];
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * 64; i += 256) s_w[i] = wt[i];
  if (tid < 64) s_b[tid] = bias[tid];
  s_lut[tid] = __fsub_rn(__fdiv_rn(static_cast<float>(tid), u8_denom), 0.5f);   // 255: pose (:429), 256: face / hand
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
        if (comp) {   // compensated precision: correction plane [fp8(lo * 2^11) 64 B | fp8(v) 64 B] of the single 64-channel chunk
          float v[8], l[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = fmaxf(acc[g * 8 + j], 0.f); l[j] = (v[j] - __half2float(hi[j])) * kCompLoScale; }
          uint8_t* c = reinterpret_cast<uint8_t*>(o - half * 32 + lo_off) + half * 32 + g * 8;
          *reinterpret_cast<uint2*>(c) = make_uint2(f32x4_to_act8x4(l[0], l[1], l[2], l[3]), f32x4_to_act8x4(l[4], l[5], l[6], l[7]));
          *reinterpret_cast<uint2*>(c + 64) = make_uint2(f32x4_to_act8x4(v[0], v[1], v[2], v[3]), f32x4_to_act8x4(v[4], v[5], v[6], v[7]));
        } else if (lo_off) *reinterpret_cast<uint4*>(o + lo_off + g * 8) = *reinterpret_cast<const uint4*>(lo);
      }
    }
  }
}

}  // namespace opb

// ------------------------------------------------------------------------------------------------
// Tensor-core variant (fast precision, uint8 frames): the 27-tap im2col row of every pixel is built
// by its own thread (look-up table -> fp16) directly in the canonical K-major SWIZZLE_128B layout,
// so conv1_1 becomes two tcgen05.mma (M = 128 pixels, N = 64 channels, K = 2 x 16) per 16 x 8
// tile instead of 1728 FMAs per pixel; the kernel is then bound by its 128 B/pixel store.
// One CTA = 128 threads, 64 TMEM columns, ~26 KB shared memory -> up to 8 CTAs per SM hide the
// (synchronous) build -> MMA -> epilogue latency of each tile.
#include "conv_tcgen05.cuh"

namespace opb {

// wth: [64][32] fp16 (row = output channel, k = (r*3+s)*3 + c, k >= 27 zero); bias [64] fp32
// Output: the 128 pixels x 128 B of a tile are staged in shared memory in the SWIZZLE_128B box layout (row = pixel
// hl * 8 + wl) and leave through ONE TMA store per plane (box {64 ch, 8 px, 16 rows}; clipped at the image edge by the
// tensor map) -- full 128-byte lines instead of 32 scattered 16-byte stores per warp instruction.
__global__ void __launch_bounds__(128)
conv_first_tc_kernel(const uint8_t* __restrict__ x_u8, const __half* __restrict__ wth, const float* __restrict__ bias,
                     const __grid_constant__ CUtensorMap tmOut, int N, int H, int W, float u8_denom) {
  __shared__ __align__(1024) uint8_t sA[128 * 128];
  __shared__ __align__(1024) uint8_t sB[64 * 128];
  __shared__ __align__(1024) uint8_t sOut[128 * 128];
  __shared__ __half s_lut[260];
  __shared__ __align__(4) __half s_val[18 * 10 * 3 + 4];   // normalised halo tile, [row][col][c]
  __shared__ float s_bias[64];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  // one-time setup
  for (int i = tid; i < 257; i += 128)
    s_lut[i] = (i < 256) ? __float2half_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i), u8_denom), 0.5f)) : __float2half(0.f);
  if (tid < 64) {
    s_bias[tid] = bias[tid];
    // weight row `tid`: 4 chunks of 8 halfs, physical chunk = logical ^ (row & 7)
    const uint4* src = reinterpret_cast<const uint4*>(wth + tid * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sB + tid * 128 + ((j ^ (tid & 7)) * 16)) = src[j];
#pragma unroll
    for (int j = 4; j < 8; ++j) *reinterpret_cast<uint4*>(sB + tid * 128 + ((j ^ (tid & 7)) * 16)) = make_uint4(0, 0, 0, 0);
  }
  {  // the chunks 4..7 of A are never read (only k-steps 0 and 1 are issued) but keep them finite
#pragma unroll
    for (int j = 4; j < 8; ++j) *reinterpret_cast<uint4*>(sA + tid * 128 + ((j ^ (tid & 7)) * 16)) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) { ptx::mbar_init(&s_bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<64>(&s_tmem);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sA), 1024);
  const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sB), 1024);
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, 64);

  const int tiles_x = (W + 7) >> 3, tiles_y = (H + 15) >> 4;
  const int total = N * tiles_y * tiles_x;
  const int hl = tid >> 3, wl = tid & 7;
  uint32_t parity = 0;
  // The 540 halo bytes of the NEXT tile are fetched into registers while this tile is built, multiplied and stored, so the
  // global-load latency (long-scoreboard 4.0 per issue with 3-8 CTAs of 4 warps per SM) is off the per-tile critical path.
  constexpr int kHaloPerThread = (18 * 10 * 3 + 127) / 128;
  int pre[kHaloPerThread];
  auto fetch_halo = [&](int t) {
    const int n = t / (tiles_y * tiles_x);
    const int rem = t - n * (tiles_y * tiles_x);
    const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
#pragma unroll
    for (int k = 0; k < kHaloPerThread; ++k) {
      const int i = tid + 128 * k;
      int v = 256;                                    // 256 = outside the image (zero padding)
      if (i < 18 * 10 * 3) {
        const int c = i % 3, q = i / 3;
        const int xx = x0 - 1 + q % 10, yy = y0 - 1 + q / 10;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x_u8[((static_cast<size_t>(n) * H + yy) * W + xx) * 3 + c];
      }
      pre[k] = v;
    }
  };
  if (static_cast<int>(blockIdx.x) < total) fetch_halo(blockIdx.x);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / (tiles_y * tiles_x);
    const int rem = tile - n * (tiles_y * tiles_x);
    const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
    // (1) halo tile, already normalised to fp16 through the look-up table (entry 256 = zero padding)
#pragma unroll
    for (int k = 0; k < kHaloPerThread; ++k) {
      const int i = tid + 128 * k;
      if (i < 18 * 10 * 3) s_val[i] = s_lut[pre[k]];
    }
    __syncthreads();
    if (tile + static_cast<int>(gridDim.x) < total) fetch_halo(tile + gridDim.x);
    // (2) this thread's im2col row: k = (r*3+s)*3 + c = 9 consecutive halfs of each of 3 halo rows.  Each row segment
    //     is fetched as six aligned 32-bit words and funnel-shifted by its parity; the 27 halfs are then packed with
    //     compile-time byte permutes (row r starts at the odd position 9r).
    {
      uint32_t seg[3][5];               // seg[r][j] = halfs (2j, 2j+1) of row r's 9-half segment (half 9 is junk)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int o = ((hl + r) * 10 + wl) * 3;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_val) + (o >> 1);
        const uint32_t sh = (o & 1) * 16;
        uint32_t w[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) w[j] = wp[j];
#pragma unroll
        for (int j = 0; j < 5; ++j) seg[r][j] = __funnelshift_r(w[j], w[j + 1], sh);
      }
      const uint32_t LO_LO = 0x5410, HI_LO = 0x5432;   // __byte_perm selectors: (a.lo, b.lo), (a.hi, b.lo)
      uint32_t pk[16];
      pk[0] = seg[0][0]; pk[1] = seg[0][1]; pk[2] = seg[0][2]; pk[3] = seg[0][3];
      pk[4] = __byte_perm(seg[0][4], seg[1][0], LO_LO);            // (a8, b0)
      pk[5] = __byte_perm(seg[1][0], seg[1][1], HI_LO);            // (b1, b2)
      pk[6] = __byte_perm(seg[1][1], seg[1][2], HI_LO);
      pk[7] = __byte_perm(seg[1][2], seg[1][3], HI_LO);
      pk[8] = __byte_perm(seg[1][3], seg[1][4], HI_LO);            // (b7, b8)
      pk[9] = seg[2][0]; pk[10] = seg[2][1]; pk[11] = seg[2][2]; pk[12] = seg[2][3];
      pk[13] = seg[2][4] & 0xffffu;                                 // (c8, 0)
      pk[14] = 0u; pk[15] = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(sA + tid * 128 + ((j ^ (tid & 7)) * 16)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    ptx::fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core (async proxy)
    ptx::tc_fence_before();
    __syncthreads();
    // (3) two MMAs, one elected thread
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      ptx::mma_f16_ss(tmem, a_desc, b_desc, IDESC, 0u);
      ptx::mma_f16_ss_acc(tmem, a_desc + 2, b_desc + 2, IDESC);
      ptx::mma_commit(&s_bar);
    }
    ptx::mbar_wait(&s_bar, parity);
    parity ^= 1;
    ptx::tc_fence_after();
    // (4) epilogue: bias + ReLU + fp16 -> swizzled staging tile -> one TMA store
    if (tid == 0) ptx::tma_store_wait_read();      // the previous tile's store has finished reading sOut
    __syncthreads();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float f[32];
      tmem_load_group<32>(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, f);
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __half2 t = __floats2half2_rn(fmaxf(f[2 * i] + s_bias[c0 + 2 * i], 0.f),
                                            fmaxf(f[2 * i + 1] + s_bias[c0 + 2 * i + 1], 0.f));
        h[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(sOut + tid * 128 + (((c0 >> 3) + g) ^ (tid & 7)) * 16) =
            make_uint4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();                   // s_val / sA / TMEM are free for the next tile; sOut is complete
    if (tid == 0) {
      ptx::tma_store_4d(&tmOut, sOut, 0, x0, y0, n);
      ptx::tma_store_commit();
    }
  }
  if (tid == 0) ptx::tma_store_wait_all();
  if (warp == 0) ptx::tmem_dealloc<64>(tmem);
}

// ------------------------------------------------------------------------------------------------
// EXACT tensor-core conv1_1 for the precisions that must stay inside the 1e-3 map tolerance (compensated, parity).
// preprocess (pose_detector.py:426-431) is affine in the uint8 pixel, x = u/255 - 0.5, and zero padding sets x (not u)
// to zero, so
//     conv(x, W)[o] = sum_{taps in the image} ( sum_c u_c * W[o,c,tap]/255  -  0.5 * sum_c W[o,c,tap] ).
// The im2col row of a pixel therefore holds the 27 raw uint8 values (exact in fp16; 0 where padded) followed by 9
// in-image indicators (1 / 0), K = 36 -> 48, and the weight matrix holds W/255 and -0.5*sum_c W per tap, each as
// hi + lo fp16 (pre-scaled by the exact power of two 2^S; the epilogue multiplies by 2^-S).  Every product is exact
// (8 x 11 significant bits) and the result agrees with the fp32 reference to ~1e-7 relative: 6 MMAs of N = 64 per
// 16 x 8 tile instead of 1728 fp32 FMAs per pixel.
// wth: [2][64][64] fp16 = {hi, lo} x (row = output channel, k as above, k >= 36 zero); out_mode 0: fp16 only,
// 1: parity (hi | lo planes), 2: compensated (hi | correction bytes).
constexpr int kFirstTcxSmem = 1024 + 128 * 128 + 2 * 64 * 128 + 2 * 128 * 128;
__global__ void __launch_bounds__(128)
conv_first_tcx_kernel(const uint8_t* __restrict__ x_u8, const __half* __restrict__ wth, const float* __restrict__ bias,
                      const __grid_constant__ CUtensorMap tmOut, int N, int H, int W, int out_mode, float acc_scale) {
  extern __shared__ uint8_t smem_raw[];   // kFirstTcxSmem bytes: sA 16 KB | sB 16 KB | sOut 32 KB, 1024-byte aligned
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = sA + 128 * 128;
  uint8_t* sOut = sB + 2 * 64 * 128;      // [hi plane | lo / correction plane], SWIZZLE_128B box layout
  __shared__ __align__(4) __half s_val[18 * 10 * 3 + 4];   // raw pixel values of the halo tile as fp16, [row][col][c]
  __shared__ float s_bias[64];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 64) s_bias[tid] = bias[tid];
  {  // weight rows: 2 x 64 rows of 64 halfs (128 B), 16-byte chunk j of row r stored at chunk j ^ (r & 7)
    const int r = tid & 63, part = tid >> 6;
    const uint4* src = reinterpret_cast<const uint4*>(wth + (part * 64 + r) * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(sB + part * 8192 + r * 128 + ((j ^ (r & 7)) * 16)) = src[j];
  }
  {  // chunks 6, 7 of A (k = 48..63) are never read; keep them finite
#pragma unroll
    for (int j = 6; j < 8; ++j) *reinterpret_cast<uint4*>(sA + tid * 128 + ((j ^ (tid & 7)) * 16)) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) { ptx::mbar_init(&s_bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<64>(&s_tmem);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(sA), 1024);
  const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(sB), 1024);
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, 64);

  const int tiles_x = (W + 7) >> 3, tiles_y = (H + 15) >> 4;
  const int total = N * tiles_y * tiles_x;
  const int hl = tid >> 3, wl = tid & 7;
  uint32_t parity = 0;
  // next tile's halo bytes prefetched into registers (see conv_first_tc_kernel)
  constexpr int kHaloPerThread = (18 * 10 * 3 + 127) / 128;
  int pre[kHaloPerThread];
  auto fetch_halo = [&](int t) {
    const int n = t / (tiles_y * tiles_x);
    const int rem = t - n * (tiles_y * tiles_x);
    const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
#pragma unroll
    for (int k = 0; k < kHaloPerThread; ++k) {
      const int i = tid + 128 * k;
      int v = 0;                                      // zero padding
      if (i < 18 * 10 * 3) {
        const int c = i % 3, q = i / 3;
        const int xx = x0 - 1 + q % 10, yy = y0 - 1 + q / 10;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x_u8[((static_cast<size_t>(n) * H + yy) * W + xx) * 3 + c];
      }
      pre[k] = v;
    }
  };
  if (static_cast<int>(blockIdx.x) < total) fetch_halo(blockIdx.x);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / (tiles_y * tiles_x);
    const int rem = tile - n * (tiles_y * tiles_x);
    const int y0 = (rem / tiles_x) * 16, x0 = (rem % tiles_x) * 8;
#pragma unroll
    for (int k = 0; k < kHaloPerThread; ++k) {
      const int i = tid + 128 * k;
      if (i < 18 * 10 * 3) s_val[i] = __ushort2half_rn(static_cast<unsigned short>(pre[k]));
    }
    __syncthreads();
    if (tile + static_cast<int>(gridDim.x) < total) fetch_halo(tile + gridDim.x);
    {
      uint32_t seg[3][5];               // seg[r][j] = halfs (2j, 2j+1) of row r's 9-half segment (half 9 is junk)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int o = ((hl + r) * 10 + wl) * 3;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_val) + (o >> 1);
        const uint32_t sh = (o & 1) * 16;
        uint32_t w[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) w[j] = wp[j];
#pragma unroll
        for (int j = 0; j < 5; ++j) seg[r][j] = __funnelshift_r(w[j], w[j + 1], sh);
      }
      const uint32_t LO_LO = 0x5410, HI_LO = 0x5432;   // __byte_perm selectors: (a.lo, b.lo), (a.hi, b.lo)
      // in-image indicators of the 9 taps (row r: y0+hl+r-1, column s: x0+wl+s-1), fp16 1.0 = 0x3C00
      const int y = y0 + hl, x = x0 + wl;
      uint32_t ind[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int sx = 0; sx < 3; ++sx)
          ind[r * 3 + sx] = (y + r - 1 >= 0 && y + r - 1 < H && x + sx - 1 >= 0 && x + sx - 1 < W) ? 0x3C00u : 0u;
      uint32_t pk[24];
      pk[0] = seg[0][0]; pk[1] = seg[0][1]; pk[2] = seg[0][2]; pk[3] = seg[0][3];
      pk[4] = __byte_perm(seg[0][4], seg[1][0], LO_LO);            // (a8, b0)
      pk[5] = __byte_perm(seg[1][0], seg[1][1], HI_LO);            // (b1, b2)
      pk[6] = __byte_perm(seg[1][1], seg[1][2], HI_LO);
      pk[7] = __byte_perm(seg[1][2], seg[1][3], HI_LO);
      pk[8] = __byte_perm(seg[1][3], seg[1][4], HI_LO);            // (b7, b8)
      pk[9] = seg[2][0]; pk[10] = seg[2][1]; pk[11] = seg[2][2]; pk[12] = seg[2][3];
      pk[13] = (seg[2][4] & 0xffffu) | (ind[0] << 16);             // (c8, i0)            k = 26, 27
      pk[14] = ind[1] | (ind[2] << 16);                            // k = 28, 29
      pk[15] = ind[3] | (ind[4] << 16);
      pk[16] = ind[5] | (ind[6] << 16);
      pk[17] = ind[7] | (ind[8] << 16);                            // k = 34, 35
#pragma unroll
      for (int j = 18; j < 24; ++j) pk[j] = 0u;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        *reinterpret_cast<uint4*>(sA + tid * 128 + ((j ^ (tid & 7)) * 16)) =
            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      ptx::mma_f16_ss(tmem, a_desc, b_desc, IDESC, 0u);                          // hi weights, k-steps 0..2
      ptx::mma_f16_ss_acc(tmem, a_desc + 2, b_desc + 2, IDESC);
      ptx::mma_f16_ss_acc(tmem, a_desc + 4, b_desc + 4, IDESC);
      ptx::mma_f16_ss_acc(tmem, a_desc, b_desc + (8192 >> 4), IDESC);            // lo weights
      ptx::mma_f16_ss_acc(tmem, a_desc + 2, b_desc + (8192 >> 4) + 2, IDESC);
      ptx::mma_f16_ss_acc(tmem, a_desc + 4, b_desc + (8192 >> 4) + 4, IDESC);
      ptx::mma_commit(&s_bar);
    }
    ptx::mbar_wait(&s_bar, parity);
    parity ^= 1;
    ptx::tc_fence_after();
    if (tid == 0) ptx::tma_store_wait_read();      // the previous tile's stores have finished reading sOut
    __syncthreads();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float f[32];
      tmem_load_group<32>(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, f);
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(f[i], acc_scale, s_bias[c0 + i]), 0.f);
      uint32_t h[16];
      float d[32];                     // v - fp16(v): the lo plane (parity) / the scaled correction byte (compensated)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t*>(&t);
        const float2 back = __half22float2(t);
        d[2 * i] = f[2 * i] - back.x;
        d[2 * i + 1] = f[2 * i + 1] - back.y;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(sOut + tid * 128 + (((c0 >> 3) + g) ^ (tid & 7)) * 16) =
            make_uint4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
      uint8_t* row2 = sOut + 128 * 128 + tid * 128;
      if (out_mode == 1) {            // parity: the lo plane
        uint32_t l[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __half2 tl = __floats2half2_rn(d[2 * i], d[2 * i + 1]);
          l[i] = *reinterpret_cast<const uint32_t*>(&tl);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(row2 + (((c0 >> 3) + g) ^ (tid & 7)) * 16) = make_uint4(l[4 * g], l[4 * g + 1], l[4 * g + 2], l[4 * g + 3]);
      } else if (out_mode == 2) {     // compensated: [fp8(lo * 2^11) 64 B | fp8(v) 64 B]
        uint32_t xl[8], x8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xl[i] = f32x4_to_act8x4(d[4 * i] * kCompLoScale, d[4 * i + 1] * kCompLoScale, d[4 * i + 2] * kCompLoScale, d[4 * i + 3] * kCompLoScale);
          x8[i] = f32x4_to_act8x4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          *reinterpret_cast<uint4*>(row2 + (((c0 >> 4) + g) ^ (tid & 7)) * 16) = make_uint4(xl[4 * g], xl[4 * g + 1], xl[4 * g + 2], xl[4 * g + 3]);
          *reinterpret_cast<uint4*>(row2 + ((4 + (c0 >> 4) + g) ^ (tid & 7)) * 16) = make_uint4(x8[4 * g], x8[4 * g + 1], x8[4 * g + 2], x8[4 * g + 3]);
        }
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      ptx::tma_store_4d(&tmOut, sOut, 0, x0, y0, n);
      if (out_mode != 0) ptx::tma_store_4d(&tmOut, sOut + 128 * 128, 64, x0, y0, n);   // second plane: channels 64..127 (halves)
      ptx::tma_store_commit();
    }
  }
  if (tid == 0) ptx::tma_store_wait_all();
  if (warp == 0) ptx::tmem_dealloc<64>(tmem);
}

}  // namespace opb
