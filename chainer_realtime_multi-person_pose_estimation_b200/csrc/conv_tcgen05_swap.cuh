// conv_tcgen05_swap.cuh -- the implicit-GEMM conv with the operand ROLES SWAPPED, for layers whose
// output-channel count is only 128 (the 7x7 128->128 refinement layers, 50 % of all FLOPs).
//
// Measured (profiles/): a tcgen05.mma of M=128 x N=128 x K=16 (64 tensor cycles) sustains ~69 % of
// the cuBLAS peak in this pipeline, one of N=256 (128 cycles) ~83 %: the per-instruction issue /
// operand-fetch overhead is amortised over twice the work.  Cout = 128 caps N at 128 when the
// pixels are the M operand, so here the WEIGHTS are the M operand (A: 128 output channels) and a
// 16 x 16 PIXEL tile is the N operand (B: 256 rows):   D[cout, pixel] += W[cout, k] * X[pixel, k]^T.
//
//   pixel operand  one TMA box {64 ch, 16 px, 16+k-1 rows}: shared-memory row index = row*16 + px,
//                  i.e. 8-pixel groups (1024-byte swizzle atoms) follow each other with a constant
//                  1024-byte stride, so the 256 pixels of filter row r are one contiguous K-major
//                  operand starting r*2048 bytes into the box (the same trick as the A boxes of
//                  conv_tcgen05.cuh).  The last tile of an image row uses an 8-pixel box (N = 128)
//                  when <= 8 columns remain, so the column padding stays at 8-pixel granularity.
//   weight operand box {64, 128 rows} per (tap, chunk), as before.
//   accumulator    TMEM lane = output channel, column = pixel of the tile; 2 x 256 columns.
//   epilogue       thread = one output channel; a warp holds 32 consecutive channels of one pixel,
//                  so each store instruction writes 64 contiguous bytes of the NHWC row.
#pragma once
#include "conv_tcgen05.cuh"

namespace opb {

template <int KS, int NSP, int NSW>
struct ConvSwapCfg {
  static constexpr int RH = 16 + KS - 1;
  static constexpr int P_STAGE_BYTES = RH * 16 * 128;      // 16-pixel-wide box
  static constexpr int W_STAGE_BYTES = 128 * 128;
  static constexpr int SMEM_BYTES = 1024 + NSP * P_STAGE_BYTES + NSW * W_STAGE_BYTES + 512;
};

// DRAIN = two-level accumulation (compensated precision): the tensor core adds into its fp32 accumulator with
// round-toward-zero, a bias that grows with the number of chained MMAs (784 for a 7x7 128->128 layer with its correction
// rows).  Every (chunk pair, filter column) segment of KS x 4 MMAs therefore lands in a fresh TMEM buffer (the two
// buffers ping-pong WITHIN a tile) and eight epilogue warps (two per TMEM lane quarter, 128 pixel columns each) add the
// partial sums in round-to-nearest fp32 registers while the next segment's MMAs run.
constexpr int kSwapDrainThreads = 64 + 256;

// CL = 2: thread-block clusters of two CTAs that work on two pixel tiles of the same (problem, channel block) in
// lockstep and SHARE every weight tile: each CTA fetches half of it (64 of the 128 output-channel rows) with a
// multicast TMA load that lands in both CTAs' shared memory, so the L2 -> SM weight traffic (72 % of this kernel's
// operand bytes, at ~75 % of the measured L2 throughput cap) halves.  A weight stage is recycled when BOTH CTAs'
// MMAs have read it (multicast tcgen05.commit onto both w_empty barriers, count 2).
template <int KS, int NSP, int NSW, bool DRAIN = false, int CL = 1>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(DRAIN ? kSwapDrainThreads : kConvThreads, 1)
conv_tcgen05_swap_kernel(const __grid_constant__ CUtensorMap tmP16_0, const __grid_constant__ CUtensorMap tmP8_0,
                         const __grid_constant__ CUtensorMap tmW_0, const __grid_constant__ CUtensorMap tmP16_1,
                         const __grid_constant__ CUtensorMap tmP8_1, const __grid_constant__ CUtensorMap tmW_1,
                         const __grid_constant__ ConvParams P) {
  using Cfg = ConvSwapCfg<KS, NSP, NSW>;
  constexpr int PAD = (KS - 1) / 2;
  constexpr int ACC_STAGES = 2;

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
    for (int i = 0; i < NSW; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], CL); }
    for (int i = 0; i < ACC_STAGES; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], DRAIN ? 256 : 128); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  if constexpr (CL > 1) ptx::cluster_sync_all();   // the peer's barriers are initialised before anything lands on them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // P.tiles_x counts 16-wide tiles plus (if P.pad_edge8) one trailing 8-wide tile
  const int m_tiles = P.N * P.tiles_y * P.tiles_x;
  const int tiles_per_problem = P.n_blocks * m_tiles;
  const int total_tiles = P.n_problems * tiles_per_problem;
  // Tile order inside one (problem, channel block): all full-width tiles first, then the narrow edge tiles, so that the
  // two CTAs of a cluster (consecutive tile numbers) do the same amount of work per weight tile.
  const int tx_full = P.tiles_x - (P.pad_edge8 ? 1 : 0);
  const int n_full = P.N * P.tiles_y * tx_full;
  uint32_t crank = 0;
  if constexpr (CL > 1) crank = ptx::cluster_ctarank();
  const int tile_first = (static_cast<int>(blockIdx.x) / CL) * CL + static_cast<int>(crank);
  const int tile_step = static_cast<int>(gridDim.x);     // a multiple of CL; total_tiles % CL == 0 (host)
  struct TileId { int p, nb, n, ty, tx; };
  auto decode = [&](int tile) {
    TileId t;
    t.p = tile / tiles_per_problem;
    int rem = tile - t.p * tiles_per_problem;
    t.nb = rem / m_tiles;
    rem -= t.nb * m_tiles;
    if (rem < n_full) {
      t.n = rem / (P.tiles_y * tx_full);
      rem -= t.n * (P.tiles_y * tx_full);
      t.ty = rem / tx_full;
      t.tx = rem - t.ty * tx_full;
    } else {
      rem -= n_full;
      t.n = rem / P.tiles_y;
      t.ty = rem - t.n * P.tiles_y;
      t.tx = P.tiles_x - 1;
    }
    return t;
  };

  if (warp == 0) {
    // ================================================================ TMA producer
    if (ptx::elect_one()) {
      uint32_t sp = 0, pp = 0, sw = 0, pw = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const TileId tid_ = decode(tile);
        const int p = tid_.p, nb = tid_.nb, n = tid_.n, ty = tid_.ty, tx = tid_.tx;
        const int y0 = ty * 16, x0 = tx * 16;
        const bool narrow = P.pad_edge8 && (tx == P.tiles_x - 1);
        const CUtensorMap* tmP = narrow ? (p ? &tmP8_1 : &tmP8_0) : (p ? &tmP16_1 : &tmP16_0);
        const CUtensorMap* tmW = p ? &tmW_1 : &tmW_0;
        const uint32_t p_bytes = narrow ? Cfg::P_STAGE_BYTES / 2 : Cfg::P_STAGE_BYTES;
        for (int j = 0; j < P.n_pairs; ++j) {
          const int ac = P.a_off[j], bk = P.b_off[j];
          for (int s = 0; s < KS; ++s) {
            ptx::mbar_wait(&p_empty[sp], pp ^ 1);
            ptx::mbar_expect_tx(&p_full[sp], p_bytes);
            ptx::tma_load_4d(smemP + sp * Cfg::P_STAGE_BYTES, tmP, &p_full[sp], ac, x0 + s - PAD, y0 - PAD, n);
            if (++sp == NSP) { sp = 0; pp ^= 1; }
            for (int r = 0; r < KS; ++r) {
              ptx::mbar_wait(&w_empty[sw], pw ^ 1);
              ptx::mbar_expect_tx(&w_full[sw], Cfg::W_STAGE_BYTES);
              if constexpr (CL > 1)   // this CTA's 64 rows of the tile, delivered to both CTAs (tmW box = {64, 64})
                ptx::tma_load_2d_mc(smemW + sw * Cfg::W_STAGE_BYTES + crank * (Cfg::W_STAGE_BYTES / 2), tmW, &w_full[sw],
                                    (r * KS + s) * P.b_tap_stride + bk, nb * 128 + static_cast<int>(crank) * 64, 3);
              else
                ptx::tma_load_2d(smemW + sw * Cfg::W_STAGE_BYTES, tmW, &w_full[sw], (r * KS + s) * P.b_tap_stride + bk,
                                 nb * 128);
              if (++sw == NSW) { sw = 0; pw ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (ptx::elect_one()) {
      uint32_t sp = 0, pp = 0, sw = 0, pw = 0, acc = 0, pacc = 0;
      const uint64_t p_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemP), 1024);
      const uint64_t w_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemW), 1024);
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const TileId tid_ = decode(tile);
        const int ty = tid_.ty, tx = tid_.tx;
        const bool narrow = P.pad_edge8 && (tx == P.tiles_x - 1);
        // the last tile row of an image only needs its valid rows (rounded to even): N = rows x 16 (or x 8)
        const int rows = min(16, (P.H - ty * 16 + 1) & ~1);
        const uint32_t idesc = ptx::umma_idesc_f16(128, rows * (narrow ? 8 : 16));
        const uint32_t idesc8 = ptx::umma_idesc_f8(128, rows * (narrow ? 8 : 16), 0 /*A: weights e4m3*/, kCompActFmt /*B: activations*/);
        const uint32_t row_pitch16 = narrow ? (1024 >> 4) : (2048 >> 4);   // bytes per image row of the box, >> 4
        if (!DRAIN) {
          ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
          ptx::tc_fence_after();
        }
        uint32_t d = tmem_base + acc * 256;
        uint32_t accumulate = 0;
        int step = 0;
        const int n_steps = P.n_pairs * KS;
        for (int j = 0; j < P.n_pairs; ++j) {
          const bool f8 = P.comp && (j & 1);   // compensated precision: odd pairs are the 8-bit correction rows
          for (int s = 0; s < KS; ++s) {
            ptx::mbar_wait(&p_full[sp], pp);
            if (DRAIN && step % P.drain_seg == 0) {   // a fresh accumulator buffer per segment of drain_seg steps
              ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
              d = tmem_base + acc * 256;
              accumulate = 0;
            }
            ptx::tc_fence_after();
            const uint64_t p_st = p_desc0 + static_cast<uint64_t>((sp * Cfg::P_STAGE_BYTES) >> 4);
#pragma unroll
            for (int r = 0; r < KS; ++r) {
              ptx::mbar_wait(&w_full[sw], pw);
              ptx::tc_fence_after();
              const uint64_t w_st = w_desc0 + static_cast<uint64_t>((sw * Cfg::W_STAGE_BYTES) >> 4);
              const uint64_t pd0 = p_st + static_cast<uint64_t>(r * row_pitch16);
              if (f8) {
                ptx::mma_f8_ss(d, w_st, pd0, idesc8, accumulate);
#pragma unroll
                for (int k = 1; k < 4; ++k) ptx::mma_f8_ss(d, w_st + (k * 32 >> 4), pd0 + (k * 32 >> 4), idesc8, 1u);
              } else {
                ptx::mma_f16_ss(d, w_st, pd0, idesc, accumulate);
#pragma unroll
                for (int k = 1; k < 4; ++k) ptx::mma_f16_ss_acc(d, w_st + (k * 32 >> 4), pd0 + (k * 32 >> 4), idesc);
              }
              accumulate = 1;
              if constexpr (CL > 1) ptx::mma_commit_mc(&w_empty[sw], 3);
              else ptx::mma_commit(&w_empty[sw]);
              if (++sw == NSW) { sw = 0; pw ^= 1; }
            }
            ptx::mma_commit(&p_empty[sp]);
            if (++sp == NSP) { sp = 0; pp ^= 1; }
            ++step;
            if (DRAIN && (step % P.drain_seg == 0 || step == n_steps)) {
              ptx::mma_commit(&t_full[acc]);
              if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
            }
          }
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
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const TileId tid_ = decode(tile);
      const int p = tid_.p, nb = tid_.nb, n = tid_.n, ty = tid_.ty, tx = tid_.tx;
      const int y0 = ty * 16, x0 = tx * 16;
      const bool narrow = P.pad_edge8 && (tx == P.tiles_x - 1);
      const int wshift = narrow ? 3 : 4;                 // pixels per tile row = 8 or 16
      const int rows = min(16, (P.H - y0 + 1) & ~1);     // rows the MMA computed for this tile
      const int n_pix = rows << wshift;
      const ConvProblem& pr = P.prob[p];
      const int ch = nb * 128 + q * 32 + lane;           // this thread's output channel
      const bool ch_ok = ch < pr.cout_valid;
      const float bias = ch_ok ? __ldg(pr.bias + ch) : 0.f;
      __half* out_c = pr.out + pr.out_coff + ch;
      const float sc = pr.acc_scale;
      // compensated precision: this channel's two bytes in the correction plane of a pixel
      uint8_t* corr_c = reinterpret_cast<uint8_t*>(pr.out + pr.out_lo_off) + comp_byte_off(pr.out_coff + ch);
      const bool comp = pr.out_lo_off != 0;
      if constexpr (!DRAIN) {
        ptx::mbar_wait(&t_full[acc], pacc);
        ptx::tc_fence_after();
  #pragma unroll 1
        for (int c0 = 0; c0 < n_pix; c0 += 32) {
          float f[32];
          tmem_load_group<32>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + c0, f);
  #pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int pix = c0 + i;
            const int y = y0 + (pix >> wshift), x = x0 + (pix & ((1 << wshift) - 1));
            float v = fmaf(f[i], sc, bias);
            v = pr.relu ? fmaxf(v, 0.f) : v;
            if (ch_ok && y < P.H && x < P.W) {
              const size_t o = ((static_cast<size_t>(n) * P.H + y) * P.W + x) * pr.out_cstride;
              const __half hi = __float2half_rn(v);
              out_c[o] = hi;
              if (comp) {
                const uint32_t b2 = __nv_cvt_float2_to_fp8x2(make_float2((v - __half2float(hi)) * kCompLoScale, v), __NV_SATFINITE, OPB_NV_ACT_FMT);
                corr_c[2 * o] = static_cast<uint8_t>(b2 & 0xffu);
                corr_c[2 * o + 64] = static_cast<uint8_t>(b2 >> 8);
              }
            }
          }
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
        const int n_seg = (P.n_pairs * KS + P.drain_seg - 1) / P.drain_seg;
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
            for (int i = 0; i < 32; ++i) {
              const int pix = col0 + cc + i;
              const int y = y0 + (pix >> wshift), x = x0 + (pix & ((1 << wshift) - 1));
              float v = fmaf(sum[cc + i], sc, bias);
              v = pr.relu ? fmaxf(v, 0.f) : v;
              if (ch_ok && y < P.H && x < P.W) {
                const size_t o = ((static_cast<size_t>(n) * P.H + y) * P.W + x) * pr.out_cstride;
                const __half hi = __float2half_rn(v);
                out_c[o] = hi;
                if (comp) {
                  const uint32_t b2 = __nv_cvt_float2_to_fp8x2(make_float2((v - __half2float(hi)) * kCompLoScale, v), __NV_SATFINITE, OPB_NV_ACT_FMT);
                  corr_c[2 * o] = static_cast<uint8_t>(b2 & 0xffu);
                  corr_c[2 * o + 64] = static_cast<uint8_t>(b2 >> 8);
                }
              }
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  if constexpr (CL > 1) ptx::cluster_sync_all();   // nothing may still land in a CTA that has exited
  else __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace opb
