// conv_tcgen05_swap7.cuh -- the role-swapped 7x7 128->128 kernel (conv_tcgen05_swap.cuh: weights = M operand,
// 16x16-pixel tile = N = 256 operand) with a LEAN single-thread MMA issue path.
//
// Why (profiles/r02_ncu_comp_swap7x7_summary.txt and the source-level samples behind it): in conv_tcgen05_swap_kernel
// the issuing thread never waits for an operand barrier (0 retries of its w_full / p_full try_wait), the tensor pipe is
// only 62-65 % active, and the thread's stall samples are spread evenly over its ~70 instructions per weight stage:
// the kernel is bound by how fast ONE thread can get through the per-stage bookkeeping (ring index / parity updates,
// descriptor rebuild from vector registers: 5 R2UR per stage, a runtime fp16/fp8 branch, a modulo per step) -- four
// 128-cycle MMAs only hide ~512 cycles of it.  This variant removes the bookkeeping instead of hiding it:
//   * the weight ring has exactly KS = 7 stages = one (chunk, filter column) step, so inside the unrolled tap loop the
//     stage index IS the tap index: every barrier address, shared-memory offset and descriptor delta is a compile-time
//     constant, and the ring parity is one bit that flips per step;
//   * the pixel ring has 2 stages (stage = step & 1); its two descriptors are precomputed per tile;
//   * fp16 / 8-bit-float correction steps alternate at compile time (template COMP), the two-level-accumulation
//     segment counter counts down instead of dividing;
//   * the CTA owns all 512 TMEM columns, so the accumulator addresses are constants (checked once).
// Shared memory: 2 x 45 056 (pixel boxes) + 7 x 16 384 (weight tiles) = 204 800 bytes.
// Results are bit-identical to conv_tcgen05_swap_kernel (same MMAs in the same order into the same accumulators).
#pragma once
#include <type_traits>
#include "conv_tcgen05_swap.cuh"

namespace opb {

struct ConvSwap7Cfg {
  static constexpr int KS = 7, NSP = 2, NSW = 7, RH = 16 + KS - 1;
  static constexpr int P_STAGE_BYTES = RH * 16 * 128;
  static constexpr int W_STAGE_BYTES = 128 * 128;
  static constexpr int SMEM_BYTES = 1024 + NSP * P_STAGE_BYTES + NSW * W_STAGE_BYTES + 512;
};

template <bool COMP, bool DRAIN>
__global__ void __launch_bounds__(DRAIN ? kSwapDrainThreads : kConvThreads, 1)
conv_tcgen05_swap7_kernel(const __grid_constant__ CUtensorMap tmP16_0, const __grid_constant__ CUtensorMap tmP8_0,
                          const __grid_constant__ CUtensorMap tmW_0, const __grid_constant__ CUtensorMap tmP16_1,
                          const __grid_constant__ CUtensorMap tmP8_1, const __grid_constant__ CUtensorMap tmW_1,
                          const __grid_constant__ ConvParams P) {
  using Cfg = ConvSwap7Cfg;
  constexpr int KS = Cfg::KS, NSP = Cfg::NSP, NSW = Cfg::NSW, PAD = 3, ACC_STAGES = 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smemP = smem;
  uint8_t* smemW = smem + NSP * Cfg::P_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemW + NSW * Cfg::W_STAGE_BYTES);
  uint64_t* p_full = bars;
  uint64_t* p_empty = p_full + NSP;
  uint64_t* w_full = p_empty + NSP;
  uint64_t* w_empty = w_full + NSW;
  uint64_t* t_full = w_empty + NSW;
  uint64_t* t_empty = t_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmP16_0);
    ptx::prefetch_tensormap(&tmP8_0);
    ptx::prefetch_tensormap(&tmW_0);
    if (P.n_problems > 1) {
      ptx::prefetch_tensormap(&tmP16_1);
      ptx::prefetch_tensormap(&tmP8_1);
      ptx::prefetch_tensormap(&tmW_1);
    }
    for (int i = 0; i < NSP; ++i) { ptx::mbar_init(&p_full[i], 1); ptx::mbar_init(&p_empty[i], 1); }
    for (int i = 0; i < NSW; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], DRAIN ? 256 : 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // one CTA per SM (204 KB of shared memory) allocating all 512 columns: the allocation starts at column 0, lane 0
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  const int m_tiles = P.N * P.tiles_y * P.tiles_x;
  const int tiles_per_problem = P.n_blocks * m_tiles;
  const int total_tiles = P.n_problems * tiles_per_problem;
  struct TileId { int p, nb, n, ty, tx; };
  auto decode = [&](int tile) {     // neighbouring CTAs work on neighbouring tiles of one image (shared halos / weights in L2)
    TileId t;
    t.p = tile / tiles_per_problem;
    int rem = tile - t.p * tiles_per_problem;
    t.nb = rem / m_tiles;
    rem -= t.nb * m_tiles;
    t.n = rem / (P.tiles_y * P.tiles_x);
    rem -= t.n * (P.tiles_y * P.tiles_x);
    t.ty = rem / P.tiles_x;
    t.tx = rem - t.ty * P.tiles_x;
    return t;
  };
  const int n_steps = P.n_pairs * KS;      // (chunk pair, filter column) steps per tile
  // Tile sequence of this CTA.  Tiles differ in cost (an 8-pixel edge tile is half a tile, the last tile row of a 46-row
  // map computes 14 of 16 rows), so `blockIdx.x + i * gridDim.x` left the SMs 15 % apart (ncu: sm__cycles_active.avg /
  // sm__cycles_elapsed.max = 0.85); the host hands every CTA a longest-processing-time-first list instead.
  const int* __restrict__ sched = P.sched ? P.sched + static_cast<size_t>(blockIdx.x) * P.sched_len : nullptr;
  auto next_tile = [&](int i) -> int {
    if (sched) return (i < P.sched_len) ? __ldg(sched + i) : -1;
    const int t = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
    return t < total_tiles ? t : -1;
  };

  if (warp == 0) {
    // ================================================================ TMA producer
    if (ptx::elect_one()) {
      uint32_t step = 0;                   // global step counter: pixel stage = step & 1, ring parities from its bits
      for (int ti = 0, tile = next_tile(0); tile >= 0; tile = next_tile(++ti)) {
        const TileId t = decode(tile);
        const int y0 = t.ty * 16, x0 = t.tx * 16;
        const bool narrow = P.pad_edge8 && (t.tx == P.tiles_x - 1);
        const CUtensorMap* tmP = narrow ? (t.p ? &tmP8_1 : &tmP8_0) : (t.p ? &tmP16_1 : &tmP16_0);
        const CUtensorMap* tmW = t.p ? &tmW_1 : &tmW_0;
        const uint32_t p_bytes = narrow ? Cfg::P_STAGE_BYTES / 2 : Cfg::P_STAGE_BYTES;
        for (int j = 0; j < P.n_pairs; ++j) {
          const int ac = P.a_off[j], bk = P.b_off[j];
          for (int s = 0; s < KS; ++s, ++step) {
            const uint32_t sp = step & 1u, pp = (step >> 1) & 1u, pw = step & 1u;
            ptx::mbar_wait(&p_empty[sp], pp ^ 1u);
            ptx::mbar_expect_tx(&p_full[sp], p_bytes);
            ptx::tma_load_4d(smemP + sp * Cfg::P_STAGE_BYTES, tmP, &p_full[sp], ac, x0 + s - PAD, y0 - PAD, t.n);
#pragma unroll
            for (int r = 0; r < KS; ++r) {
              ptx::mbar_wait(&w_empty[r], pw ^ 1u);
              ptx::mbar_expect_tx(&w_full[r], Cfg::W_STAGE_BYTES);
              ptx::tma_load_2d(smemW + r * Cfg::W_STAGE_BYTES, tmW, &w_full[r], (r * KS + s) * P.b_tap_stride + bk, t.nb * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (ptx::elect_one()) {
      uint32_t step = 0, acc = 0, pacc = 0;
      const uint64_t p_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemP), 1024);
      const uint64_t p_desc1 = p_desc0 + static_cast<uint64_t>(Cfg::P_STAGE_BYTES >> 4);
      const uint64_t w_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemW), 1024);
      for (int ti = 0, tile = next_tile(0); tile >= 0; tile = next_tile(++ti)) {
        const TileId t = decode(tile);
        const bool narrow = P.pad_edge8 && (t.tx == P.tiles_x - 1);
        const int rows = min(16, (P.H - t.ty * 16 + 1) & ~1);
        const uint32_t idesc = ptx::umma_idesc_f16(128, rows * (narrow ? 8 : 16));
        const uint32_t idesc8 = ptx::umma_idesc_f8(128, rows * (narrow ? 8 : 16), 0, kCompActFmt);
        const uint32_t row_pitch16 = narrow ? (1024 >> 4) : (2048 >> 4);
        if (!DRAIN) {
          ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
          ptx::tc_fence_after();
        }
        uint32_t d = tmem_base + acc * 256;
        uint32_t accumulate = 0;
        int seg_left = 0;                  // DRAIN: steps left in the current accumulation segment
        int steps_left = n_steps;
        // one (chunk pair, filter column) step: 7 weight stages x 4 MMAs
        auto do_step = [&](auto f8_tag) {
          constexpr bool F8 = decltype(f8_tag)::value;
          const uint32_t sp = step & 1u, pp = (step >> 1) & 1u, pw = step & 1u;
          ptx::mbar_wait(&p_full[sp], pp);
          if (DRAIN && seg_left == 0) {    // a fresh accumulator buffer per segment of drain_seg steps
            ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
            d = tmem_base + acc * 256;
            accumulate = 0;
            seg_left = P.drain_seg;
          }
          ptx::tc_fence_after();
          const uint64_t p_st = sp ? p_desc1 : p_desc0;
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            ptx::mbar_wait(&w_full[r], pw);
            ptx::tc_fence_after();
            const uint64_t w_st = w_desc0 + static_cast<uint64_t>((r * Cfg::W_STAGE_BYTES) >> 4);
            const uint64_t pd0 = p_st + static_cast<uint64_t>(r * row_pitch16);
            if constexpr (F8) {
              ptx::mma_f8_ss(d, w_st, pd0, idesc8, accumulate);
#pragma unroll
              for (int k = 1; k < 4; ++k) ptx::mma_f8_ss(d, w_st + (k * 32 >> 4), pd0 + (k * 32 >> 4), idesc8, 1u);
            } else {
              ptx::mma_f16_ss(d, w_st, pd0, idesc, accumulate);
#pragma unroll
              for (int k = 1; k < 4; ++k) ptx::mma_f16_ss_acc(d, w_st + (k * 32 >> 4), pd0 + (k * 32 >> 4), idesc);
            }
            accumulate = 1;
            ptx::mma_commit(&w_empty[r]);
          }
          ptx::mma_commit(&p_empty[sp]);
          ++step;
          --steps_left;
          if (DRAIN) {
            if (--seg_left == 0 || steps_left == 0) {
              seg_left = 0;
              ptx::mma_commit(&t_full[acc]);
              if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
            }
          }
        };
        if constexpr (COMP) {              // pairs alternate: fp16 row of a chunk, then its 8-bit-float correction row
          for (int j = 0; j < P.n_pairs; j += 2) {
            for (int s = 0; s < KS; ++s) do_step(std::false_type{});
            for (int s = 0; s < KS; ++s) do_step(std::true_type{});
          }
        } else {
          for (int i = 0; i < n_steps; ++i) do_step(std::false_type{});
        }
        if (!DRAIN) {
          ptx::mma_commit(&t_full[acc]);
          if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
        }
      }
    }
  } else {
    // ================================================================ epilogue: thread = output channel
    const int q = warp & 3;
    uint32_t acc = 0, pacc = 0;
    for (int ti = 0, tile = next_tile(0); tile >= 0; tile = next_tile(++ti)) {
      const TileId t = decode(tile);
      const int p = t.p, nb = t.nb, n = t.n;
      const int y0 = t.ty * 16, x0 = t.tx * 16;
      const bool narrow = P.pad_edge8 && (t.tx == P.tiles_x - 1);
      const int wshift = narrow ? 3 : 4;                 // pixels per tile row = 8 or 16
      const int rows = min(16, (P.H - y0 + 1) & ~1);     // rows the MMA computed for this tile
      const int n_pix = rows << wshift;
      const ConvProblem& pr = P.prob[p];
      const int ch = nb * 128 + q * 32 + lane;           // this thread's output channel
      const bool ch_ok = ch < pr.cout_valid;
      const float bias = ch_ok ? __ldg(pr.bias + ch) : 0.f;
      __half* out_c = pr.out + pr.out_coff + ch;
      const float sc = pr.acc_scale;
      uint8_t* corr_c = reinterpret_cast<uint8_t*>(pr.out + pr.out_lo_off) + comp_byte_off(pr.out_coff + ch);
      const bool comp = pr.out_lo_off != 0;
      const bool relu = pr.relu != 0;
      const size_t img_base = static_cast<size_t>(n) * P.H * P.W;
      const int cstride = pr.out_cstride;
      auto store_px = [&](int pix, float a) {
        const int y = y0 + (pix >> wshift), x = x0 + (pix & ((1 << wshift) - 1));
        float v = fmaf(a, sc, bias);
        v = relu ? fmaxf(v, 0.f) : v;
        if (ch_ok && y < P.H && x < P.W) {
          const size_t o = (img_base + static_cast<size_t>(y) * P.W + x) * cstride;
          const __half hi = __float2half_rn(v);
          out_c[o] = hi;
          if (comp) {
            const uint32_t b2 = __nv_cvt_float2_to_fp8x2(make_float2((v - __half2float(hi)) * kCompLoScale, v), __NV_SATFINITE, OPB_NV_ACT_FMT);
            corr_c[2 * o] = static_cast<uint8_t>(b2 & 0xffu);
            corr_c[2 * o + 64] = static_cast<uint8_t>(b2 >> 8);
          }
        }
      };
      if constexpr (!DRAIN) {
        ptx::mbar_wait(&t_full[acc], pacc);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < n_pix; c0 += 32) {
          float f[32];
          tmem_load_group<32>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + c0, f);
#pragma unroll
          for (int i = 0; i < 32; ++i) store_px(c0 + i, f[i]);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&t_empty[acc]);
        if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
      } else {
        // two-level accumulation: this warp's 128 pixel columns [col0, col0 + 128) of every segment's buffer
        const int col0 = ((warp - 2) >> 2) * 128;
        float sum[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) sum[i] = 0.f;
        const int n_seg = (n_steps + P.drain_seg - 1) / P.drain_seg;
        for (int seg = 0; seg < n_seg; ++seg) {
          ptx::mbar_wait(&t_full[acc], pacc);
          ptx::tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < 128; cc += 32) {
            if (col0 + cc < n_pix) {       // warp-uniform
              float f[32];
              tmem_load_group<32>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + col0 + cc, f);
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[cc + i] += f[i];
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&t_empty[acc]);
          if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
        }
#pragma unroll
        for (int cc = 0; cc < 128; cc += 32) {
          if (col0 + cc < n_pix) {
#pragma unroll
            for (int i = 0; i < 32; ++i) store_px(col0 + cc + i, sum[cc + i]);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace opb
