// conv_mlp2.cuh -- the two 1x1 convolutions that end every refinement stage, fused:
//   h = relu(Mconv6(x))  (128 -> 128)      y = Mconv7(h)  (128 -> 38 | 19 | 71 | 22, no ReLU)
// (models/CocoPoseNet.py:173-175,180-182 ...; models/FaceNet.py:114-116 ...).  Run separately they are two launches
// that stream the 128-channel intermediate through HBM (61.8 MB written and read back per stage at batch 32); here
// the intermediate never leaves the SM: GEMM 1 accumulates in TMEM, the epilogue writes relu(acc + b) as fp16
// straight into shared memory in the canonical K-major SWIZZLE_128B layout, and GEMM 2 consumes it as its A operand.
//
// One CTA = 128 threads = the 128 rows (16 x 8 pixels, row m = y*8 + x) of one tile; thread m owns TMEM lane m.
// Both weight matrices (32 KB + 12 KB) stay resident in shared memory for the CTA's lifetime; ~112 KB of shared
// memory and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's load / MMA / epilogue phases.
// Fast and compensated precision (COMP: every operand tile carries its 8-bit-float correction tile, 2 x 8 MMAs per GEMM,
// 216 KB of shared memory -> one CTA per SM); parity precision keeps the two DRAIN launches.
#pragma once
#include "conv_tcgen05.cuh"

namespace opb {

struct Mlp2Params {
  int N, H, W;
  int tiles_x, tiles_y;     // ceil(W/8), ceil(H/16)
  int n_problems;           // 1, or 2 = the L1 / L2 branches (CTA parity selects the branch)
  const float* bias1[2];    // first conv's bias [128]
  float scale1[2];          // first conv's accumulator scale (2^-S in compensated precision, else 1)
  int corr_off;             // compensated precision: channel (half) offset of the input tensor's correction plane
  int w_corr_off;           // ... and k offset of the correction rows inside one packed weight row
  ConvProblem prob[2];      // second conv: output tensor / slice, bias, cout_valid, optional planar fp32 copy
};

constexpr int kMlp2N2 = 48;                                  // padded output channels of the second conv
constexpr int kMlp2A = 2 * 16384, kMlp2W1 = 2 * 16384, kMlp2I = 2 * 16384, kMlp2W2 = 2 * kMlp2N2 * 128;
constexpr int kMlp2Smem = 1024 + kMlp2A + kMlp2W1 + kMlp2I + kMlp2W2 + 128 * 4 + kMlp2N2 * 4 + 64;
// compensated precision: every operand tile is followed by its 8-bit-float correction tile of the same size
// ([fp8(x_lo * 2^11) x 64 | fp8(x) x 64] per 64-channel chunk, conv_tcgen05.cuh) -> 216 KB, one CTA per SM
constexpr int kMlp2SmemComp = 1024 + 2 * (kMlp2A + kMlp2W1 + kMlp2I + kMlp2W2) + 128 * 4 + kMlp2N2 * 4 + 64;

template <bool COMP>
__global__ void __launch_bounds__(128)
conv_mlp2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmW1_0,
                 const __grid_constant__ CUtensorMap tmW2_0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW1_1, const __grid_constant__ CUtensorMap tmW2_1,
                 const __grid_constant__ Mlp2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int X = COMP ? 2 : 1;            // COMP: [fp16 chunks | correction chunks] per operand
  uint8_t* sA = base;                        // [2 chunks][128 rows][128 B]   input tile, by TMA
  uint8_t* sW1 = sA + X * kMlp2A;            // [2 chunks][128 rows][128 B]   Mconv6 weights
  uint8_t* sI = sW1 + X * kMlp2W1;           // [2 chunks][128 rows][128 B]   relu(Mconv6) tile, by the epilogue
  uint8_t* sW2 = sI + X * kMlp2I;            // [2 chunks][48 rows][128 B]    Mconv7 weights
  float* s_b1 = reinterpret_cast<float*>(sW2 + X * kMlp2W2);
  float* s_b2 = s_b1 + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + kMlp2N2);   // [0] weights, [1] input tile, [2] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int p = (P.n_problems > 1) ? (blockIdx.x & 1) : 0;
  const int cta = (P.n_problems > 1) ? (blockIdx.x >> 1) : blockIdx.x;
  const int n_cta = (P.n_problems > 1) ? (gridDim.x >> 1) : gridDim.x;
  const CUtensorMap* tmA = p ? &tmA1 : &tmA0;
  const CUtensorMap* tmW1 = p ? &tmW1_1 : &tmW1_0;
  const CUtensorMap* tmW2 = p ? &tmW2_1 : &tmW2_0;
  const ConvProblem& pr = P.prob[p];

  s_b1[tid] = P.bias1[p][tid];
  if (tid < kMlp2N2) s_b2[tid] = pr.bias[tid];
  if (tid == 0) {
    ptx::prefetch_tensormap(tmA);
    ptx::prefetch_tensormap(tmW1);
    ptx::prefetch_tensormap(tmW2);
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::mbar_init(&bars[2], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;            // GEMM 1 accumulator: columns 0..127, GEMM 2: columns 128..175

  if (tid == 0) {                              // both weight matrices, once
    ptx::mbar_expect_tx(&bars[0], X * (kMlp2W1 + kMlp2W2));
    ptx::tma_load_2d(sW1, tmW1, &bars[0], 0, 0);
    ptx::tma_load_2d(sW1 + 16384, tmW1, &bars[0], 64, 0);
    ptx::tma_load_2d(sW2, tmW2, &bars[0], 0, 0);
    ptx::tma_load_2d(sW2 + kMlp2N2 * 128, tmW2, &bars[0], 64, 0);
    if constexpr (COMP) {                      // the correction rows of both weight matrices
      ptx::tma_load_2d(sW1 + kMlp2W1, tmW1, &bars[0], P.w_corr_off, 0);
      ptx::tma_load_2d(sW1 + kMlp2W1 + 16384, tmW1, &bars[0], P.w_corr_off + 64, 0);
      ptx::tma_load_2d(sW2 + kMlp2W2, tmW2, &bars[0], P.w_corr_off, 0);
      ptx::tma_load_2d(sW2 + kMlp2W2 + kMlp2N2 * 128, tmW2, &bars[0], P.w_corr_off + 64, 0);
    }
  }
  const uint64_t dA = ptx::umma_desc_sw128(ptx::smem_u32(sA), 1024);
  const uint64_t dW1 = ptx::umma_desc_sw128(ptx::smem_u32(sW1), 1024);
  const uint64_t dI = ptx::umma_desc_sw128(ptx::smem_u32(sI), 1024);
  const uint64_t dW2 = ptx::umma_desc_sw128(ptx::smem_u32(sW2), 1024);
  constexpr uint32_t IDESC1 = ptx::umma_idesc_f16(128, 128);
  constexpr uint32_t IDESC2 = ptx::umma_idesc_f16(128, kMlp2N2);
  constexpr uint32_t IDESC1_8 = ptx::umma_idesc_f8(128, 128, kCompActFmt /*A: activations*/, 0 /*B: weights e4m3*/);
  constexpr uint32_t IDESC2_8 = ptx::umma_idesc_f8(128, kMlp2N2, kCompActFmt, 0);
  constexpr uint32_t CORR_A = kMlp2A >> 4, CORR_W1 = kMlp2W1 >> 4, CORR_I = kMlp2I >> 4, CORR_W2 = kMlp2W2 >> 4;
  constexpr uint32_t CHUNK_A = 16384 >> 4, CHUNK_W2 = (kMlp2N2 * 128) >> 4;   // descriptor address units (16 B)

  const int m_tiles = P.N * P.tiles_y * P.tiles_x;
  uint32_t par_a = 0, par_m = 0;
  bool weights_ready = false;
  auto load_tile = [&](int t) {                  // thread 0: input tile t -> sA (both 64-channel chunks)
    const int tn = t / (P.tiles_y * P.tiles_x);
    const int trem = t - tn * (P.tiles_y * P.tiles_x);
    const int tty = trem / P.tiles_x, ttx = trem - tty * P.tiles_x;
    ptx::mbar_expect_tx(&bars[1], X * kMlp2A);
    ptx::tma_load_4d(sA, tmA, &bars[1], 0, ttx * 8, tty * 16, tn);
    ptx::tma_load_4d(sA + 16384, tmA, &bars[1], 64, ttx * 8, tty * 16, tn);
    if constexpr (COMP) {                      // the tile's correction bytes (addressed as 64 halves per chunk)
      ptx::tma_load_4d(sA + kMlp2A, tmA, &bars[1], P.corr_off, ttx * 8, tty * 16, tn);
      ptx::tma_load_4d(sA + kMlp2A + 16384, tmA, &bars[1], P.corr_off + 64, ttx * 8, tty * 16, tn);
    }
  };
  if (tid == 0 && cta < m_tiles) load_tile(cta);
  for (int tile = cta; tile < m_tiles; tile += n_cta) {
    const int n = tile / (P.tiles_y * P.tiles_x);
    const int rem = tile - n * (P.tiles_y * P.tiles_x);
    const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
    const int y0 = ty * 16, x0 = tx * 8;
    if (warp == 0 && ptx::elect_one()) {
      if (!weights_ready) ptx::mbar_wait(&bars[0], 0);
      ptx::mbar_wait(&bars[1], par_a);
      ptx::tc_fence_after();
      // GEMM 1: [128 px x 128 ch] x W6^T -> TMEM columns 0..127
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c == 0 && k == 0) ptx::mma_f16_ss(tmem, dA, dW1, IDESC1, 0u);
          else ptx::mma_f16_ss_acc(tmem, dA + c * CHUNK_A + 2 * k, dW1 + c * CHUNK_A + 2 * k, IDESC1);
        }
        if constexpr (COMP) {                  // the chunk's first-order rounding corrections (same order as the plain kernel)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f8_ss(tmem, dA + CORR_A + c * CHUNK_A + 2 * k, dW1 + CORR_W1 + c * CHUNK_A + 2 * k, IDESC1_8, 1u);
        }
      }
      ptx::mma_commit(&bars[2]);
    }
    weights_ready = true;
    par_a ^= 1;
    ptx::mbar_wait(&bars[2], par_m);
    par_m ^= 1;
    ptx::tc_fence_after();
    // GEMM 1 has consumed sA: fetch the next tile now, behind this tile's epilogues and GEMM 2
    if (tid == 0 && tile + n_cta < m_tiles) load_tile(tile + n_cta);
    // epilogue 1: h = relu(acc + b1) -> fp16 -> sI in the K-major SWIZZLE_128B layout (row = tid)
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float f[32];
      tmem_load_group<32>(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, f);
      const float sc1 = P.scale1[p];            // 1 (fmaf(a, 1, b) == a + b), or the exact power of two 2^-S
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(f[i], sc1, s_b1[c0 + i]), 0.f);
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
      uint8_t* row = sI + (c0 >> 6) * 16384 + tid * 128;
      const int j0 = (c0 & 63) >> 3;            // first 16-byte unit of this group inside the 64-channel chunk
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(row + (((j0 + g) ^ (tid & 7)) * 16)) =
            make_uint4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
      if constexpr (COMP) {
        // the same hi / correction bytes the unfused Mconv6 launch would have stored (comp_store): per 64-channel chunk a
        // 128-byte row [fp8((v - hi) * 2^11) x 64 | fp8(v) x 64], here straight into the swizzled K-major operand tile
        uint32_t xl[8], x8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float v = f[4 * i + k];
            l[k] = (v - __half2float(__float2half_rn(v))) * kCompLoScale;
          }
          xl[i] = f32x4_to_act8x4(l[0], l[1], l[2], l[3]);
          x8[i] = f32x4_to_act8x4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
        uint8_t* crow = sI + kMlp2I + (c0 >> 6) * 16384 + tid * 128;
        const int u0 = (c0 & 63) >> 4;          // first 16-byte unit of the 32 lo bytes; the x8 bytes sit 4 units further
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          *reinterpret_cast<uint4*>(crow + (((u0 + g) ^ (tid & 7)) * 16)) = make_uint4(xl[4 * g], xl[4 * g + 1], xl[4 * g + 2], xl[4 * g + 3]);
          *reinterpret_cast<uint4*>(crow + (((4 + u0 + g) ^ (tid & 7)) * 16)) = make_uint4(x8[4 * g], x8[4 * g + 1], x8[4 * g + 2], x8[4 * g + 3]);
        }
      }
    }
    ptx::fence_proxy_async_smem();               // generic-proxy stores -> visible to the tensor core
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      // GEMM 2: h x W7^T -> TMEM columns 128..175
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c == 0 && k == 0) ptx::mma_f16_ss(tmem + 128, dI, dW2, IDESC2, 0u);
          else ptx::mma_f16_ss_acc(tmem + 128, dI + c * CHUNK_A + 2 * k, dW2 + c * CHUNK_W2 + 2 * k, IDESC2);
        }
        if constexpr (COMP) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f8_ss(tmem + 128, dI + CORR_I + c * CHUNK_A + 2 * k, dW2 + CORR_W2 + c * CHUNK_W2 + 2 * k, IDESC2_8, 1u);
        }
      }
      ptx::mma_commit(&bars[2]);
    }
    ptx::mbar_wait(&bars[2], par_m);
    par_m ^= 1;
    ptx::tc_fence_after();
    // epilogue 2: y = acc + b2 -> channel slice of the concat tensor (+ planar fp32 copy on the last stage)
    const int y = y0 + (tid >> 3), x = x0 + (tid & 7);
    const bool valid = (y < P.H) && (x < P.W);
#pragma unroll
    for (int g = 0; g < kMlp2N2 / 16; ++g) {
      float f[16];
      tmem_load_group<16>(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 128 + g * 16, f);
      epilogue_store_group<16, false>(pr, f, s_b2 + g * 16, g * 16, n, y, x, P.H, P.W, valid);
    }
    ptx::tc_fence_before();
    __syncthreads();                             // sA / sI / TMEM are free for the next tile
  }
  if (warp == 0) ptx::tmem_dealloc<256>(tmem);
}

}  // namespace opb
