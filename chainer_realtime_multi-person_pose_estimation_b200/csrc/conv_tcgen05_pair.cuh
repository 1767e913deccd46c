// conv_tcgen05_pair.cuh -- the implicit-GEMM conv of conv_tcgen05.cuh on CTA PAIRS
// (tcgen05 cta_group::2, thread-block cluster of 2, one CTA per SM of a TPC).
//
// Why: an M=128 x N=128 SS-mode UMMA reads 8 KB of operands per 64 cycles = the whole
// 128 B/clk shared-memory port, so the single-CTA kernel is shared-memory bound (ncu: tensor
// pipe 67 % active, profiles/r01_conv7x7_ncu_full_summary.txt).  With cta_group::2 one MMA
// covers M = 256 (128 pixel rows from each CTA) and each CTA supplies only HALF of the weight
// tile (N/2 rows): operand reads drop to 6 KB per 64 cycles per SM and each weight byte is
// fetched from L2 once per pair.
//
// Protocol (names as in conv_tcgen05.cuh):
//   * both CTAs run a TMA producer; the loads are the cta_group::2 form whose transaction
//     bytes are credited to the LEADER's (cluster rank 0) full barriers; the leader's producer
//     arms them with the byte count of both CTAs.
//   * only the leader's elected thread issues tcgen05.mma.cta_group::2; tcgen05.commit with
//     .multicast::cluster frees the stage (a_empty / b_empty) and publishes the accumulator
//     (t_full) in BOTH CTAs.
//   * each CTA's epilogue warps drain their own TMEM (their 128 pixel rows) and arrive on the
//     leader's t_empty barrier (count 256) through its shared::cluster address.
// Pair tile = 16 rows x 16*MT columns: 8-column block j = 2*mt + rank belongs to CTA `rank`.
#pragma once
#include "conv_tcgen05.cuh"

namespace opb {

template <int KS, int BN, int MT, int NSA, int NSB, int ACC_STAGES>
struct ConvPairCfg {
  static constexpr int RH = 16 + KS - 1;
  static constexpr int A_SUB_BYTES = RH * 1024;
  static constexpr int A_STAGE_BYTES = MT * A_SUB_BYTES;
  static constexpr int B_STAGE_BYTES = (BN / 2) * 128;     // this CTA's half of the weight tile
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * MT * BN;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static_assert(TMEM_COLS_RAW <= 512, "accumulators do not fit TMEM");
  static constexpr int SMEM_BYTES = 1024 + NSA * A_STAGE_BYTES + NSB * B_STAGE_BYTES + 512 + 8 * BN * 4;
};

// DRAIN (compensated precision, BN = 256, MT = 1): two-level accumulation as in conv_tcgen05_swap.cuh -- every
// (chunk pair, filter column) segment lands in a fresh TMEM buffer and eight epilogue warps per CTA (128 channel columns
// each) add the partial sums in round-to-nearest fp32 registers.
constexpr int kPairDrainThreads = 64 + 256;

template <int KS, int BN, int MT, int NSA, int NSB, int ACC_STAGES, bool DRAIN = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DRAIN ? kPairDrainThreads : kConvThreads2, 1)
conv_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                         const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                         const __grid_constant__ ConvParams P) {
  using Cfg = ConvPairCfg<KS, BN, MT, NSA, NSB, ACC_STAGES>;
  constexpr int PAD = (KS - 1) / 2;
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(256, BN);
  constexpr uint32_t IDESC8 = ptx::umma_idesc_f8(256, BN, kCompActFmt /*A: activations*/, 0 /*B: weights e4m3*/);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + NSA * Cfg::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + NSB * Cfg::B_STAGE_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + NSA;
  uint64_t* b_full = a_empty + NSA;
  uint64_t* b_empty = b_full + NSB;
  uint64_t* t_full = b_empty + NSB;
  uint64_t* t_empty = t_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + ACC_STAGES);
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [8 warps][BN]

  // warp roles: 0..7 epilogue, 8 TMA producer, 9 MMA issuer.  The SMSP arbiter prefers the HIGHEST
  // warp id, so the two latency-critical single-thread roles get the top ids of their SMSPs.
  static_assert(!DRAIN || (MT == 1 && BN == 256), "drain mode: one 256-channel accumulator per buffer");
  constexpr int EPI_SETS = DRAIN ? 2 : OPB_EPI_SETS;
  constexpr int kEpiWarps = 4 * EPI_SETS;
  const int warp_raw = threadIdx.x >> 5;
#if OPB_ROLE_REORDER
  const int warp = (warp_raw >= kEpiWarps) ? warp_raw - kEpiWarps : warp_raw + 2;   // logical: 0 TMA, 1 MMA, 2.. epilogue
#else
  const int warp = warp_raw;
#endif
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA0);
    ptx::prefetch_tensormap(&tmB0);
    if (P.n_problems > 1) {
      ptx::prefetch_tensormap(&tmA1);
      ptx::prefetch_tensormap(&tmB1);
    }
    for (int i = 0; i < NSA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NSB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 256 * EPI_SETS); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // peer barriers initialised, both TMEM allocations done
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool units = (MT == 1) && P.pair_units != 0;           // see ConvParams::pair_units
  const int m_tiles = units ? (P.N * P.tiles_y * P.tiles_x) >> 1 : P.N * P.tiles_y * P.tiles_x;   // pair tiles
  struct Pos { int n, y0, x0, n_sub; };                        // x0 = first column of THIS CTA's 8-column block 0
  auto locate = [&](int rem) {                                 // rem = pair-tile index inside one (problem, channel block)
    Pos q;
    if (units) {
      const int u = 2 * rem + static_cast<int>(rank);
      q.n = u / (P.tiles_y * P.tiles_x);
      const int r2 = u - q.n * (P.tiles_y * P.tiles_x);
      const int ty = r2 / P.tiles_x;
      q.y0 = ty * 16;
      q.x0 = (r2 - ty * P.tiles_x) * 8;
      q.n_sub = 2;
    } else {
      q.n = rem / (P.tiles_y * P.tiles_x);
      const int r2 = rem - q.n * (P.tiles_y * P.tiles_x);
      const int ty = r2 / P.tiles_x, tx = r2 - ty * P.tiles_x;
      q.y0 = ty * 16;
      q.x0 = tx * (16 * MT) + 8 * static_cast<int>(rank);
      q.n_sub = min(2 * MT, (P.W - tx * (16 * MT) + 7) >> 3);    // valid 8-column blocks in this pair tile
    }
    return q;
  };
  const int tiles_per_problem = P.n_blocks * m_tiles;
  const int total_tiles = P.n_problems * tiles_per_problem;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
        const int p = tile / tiles_per_problem;
        int rem = tile - p * tiles_per_problem;
        const int nb = rem / m_tiles;
        rem -= nb * m_tiles;
        const Pos q = locate(rem);
        const int n = q.n, y0 = q.y0;
        const int n_mma = (q.n_sub + 1) >> 1;                  // M=256 MMAs per k-step
        const CUtensorMap* tmA = p ? &tmA1 : &tmA0;
        const CUtensorMap* tmB = p ? &tmB1 : &tmB0;
        for (int j = 0; j < P.n_pairs; ++j) {
          const int ac = P.a_off[j], bk = P.b_off[j];
          for (int s = 0; s < KS; ++s) {
            ptx::mbar_wait(&a_empty[sa], pa ^ 1);
            if (rank == 0) ptx::mbar_expect_tx(&a_full[sa], 2 * n_mma * Cfg::A_SUB_BYTES);
            const uint32_t afull0 = ptx::mapa_u32(ptx::smem_u32(&a_full[sa]), 0);
            for (int mt = 0; mt < n_mma; ++mt)   // an out-of-image block loads zeros (TMA OOB fill)
              ptx::tma_load_4d_pair(smemA + sa * Cfg::A_STAGE_BYTES + mt * Cfg::A_SUB_BYTES, tmA, afull0, ac,
                                    q.x0 + 16 * mt + s - PAD, y0 - PAD, n);
            if (++sa == NSA) { sa = 0; pa ^= 1; }
            for (int r = 0; r < KS; ++r) {
              ptx::mbar_wait(&b_empty[sb], pb ^ 1);
              if (rank == 0) ptx::mbar_expect_tx(&b_full[sb], 2 * Cfg::B_STAGE_BYTES);
              ptx::tma_load_2d_pair(smemB + sb * Cfg::B_STAGE_BYTES, tmB, ptx::mapa_u32(ptx::smem_u32(&b_full[sb]), 0),
                                    (r * KS + s) * P.b_tap_stride + bk, nb * BN + static_cast<int>(rank) * (BN / 2));
              if (++sb == NSB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (rank == 0 && ptx::elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 0;
      const uint64_t a_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemA), 1024);
      const uint64_t b_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemB), 1024);
      for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
        const int n_mma = (locate(tile % m_tiles).n_sub + 1) >> 1;
        if (!DRAIN) {
          ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
          ptx::tc_fence_after();
        }
        uint32_t accumulate = 0;
        int step = 0;
        const int n_steps = P.n_pairs * KS;
        for (int j = 0; j < P.n_pairs; ++j) {
          const bool f8 = P.comp && (j & 1);   // compensated precision: odd pairs are the 8-bit correction rows
          for (int s = 0; s < KS; ++s) {
            ptx::mbar_wait(&a_full[sa], pa);
            if (DRAIN && step % P.drain_seg == 0) {   // a fresh accumulator buffer per segment of drain_seg steps
              ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
              accumulate = 0;
            }
            ptx::tc_fence_after();
            const uint64_t a_st = a_desc0 + static_cast<uint64_t>((sa * Cfg::A_STAGE_BYTES) >> 4);
#pragma unroll
            for (int r = 0; r < KS; ++r) {
              ptx::mbar_wait(&b_full[sb], pb);
              ptx::tc_fence_after();
              const uint64_t b_st = b_desc0 + static_cast<uint64_t>((sb * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                if (mt < n_mma) {
                  const uint32_t d = tmem_base + (acc * MT + mt) * BN;
                  const uint64_t ad0 = a_st + static_cast<uint64_t>((mt * Cfg::A_SUB_BYTES + r * 1024) >> 4);
                  if (f8) {
                    ptx::mma_f8_ss_pair(d, ad0, b_st, IDESC8, accumulate);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                      ptx::mma_f8_ss_pair(d, ad0 + (k * 32 >> 4), b_st + (k * 32 >> 4), IDESC8, 1u);
                  } else {
                    ptx::mma_f16_ss_pair(d, ad0, b_st, IDESC, accumulate);
#pragma unroll
                    for (int k = 1; k < 4; ++k)
                      ptx::mma_f16_ss_pair_acc(d, ad0 + (k * 32 >> 4), b_st + (k * 32 >> 4), IDESC);
                  }
                }
              }
              accumulate = 1;
              ptx::mma_commit_pair(&b_empty[sb]);
              if (++sb == NSB) { sb = 0; pb ^= 1; }
            }
            ptx::mma_commit_pair(&a_empty[sa]);
            if (++sa == NSA) { sa = 0; pa ^= 1; }
            ++step;
            if (DRAIN && (step % P.drain_seg == 0 || step == n_steps)) {
              ptx::mma_commit_pair(&t_full[acc]);
              if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
            }
          }
        }
        if (!DRAIN) {
          ptx::mma_commit_pair(&t_full[acc]);
          if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5, both CTAs)
    const int q = warp_raw & 3;
    const int eset = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    float* s_bias_w = s_bias + (warp - 2) * BN;
    int bias_key = -1;
    uint32_t acc = 0, pacc = 0;
    constexpr int CW = (BN % 32 == 0) ? 32 : 16;
    for (int tile = pair_id; tile < total_tiles; tile += n_pairs) {
      const int p = tile / tiles_per_problem;
      int rem = tile - p * tiles_per_problem;
      const int nb = rem / m_tiles;
      rem -= nb * m_tiles;
      const Pos pos = locate(rem);
      const int n = pos.n;
      const int y = pos.y0 + hl;
      const int n_sub = pos.n_sub;
      const int n_mma = (n_sub + 1) >> 1;
      const ConvProblem& pr = P.prob[p];
      if (bias_key != p * 1024 + nb) {
        bias_key = p * 1024 + nb;
        epilogue_load_bias<BN>(s_bias_w, pr.bias + nb * BN, lane);
      }
      if constexpr (DRAIN) {
        // two-level accumulation: this warp's 128 channel columns [col0, col0 + 128) of every segment's buffer
        const int col0 = eset * 128;
        float sum[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) sum[i] = 0.f;
        const int n_seg = (P.n_pairs * KS + P.drain_seg - 1) / P.drain_seg;
        const bool mine = static_cast<int>(rank) < n_sub;        // this CTA's 8-column block is inside the image
        for (int seg = 0; seg < n_seg; ++seg) {
          ptx::mbar_wait(&t_full[acc], pacc);
          ptx::tc_fence_after();
          if (mine) {
#pragma unroll
            for (int cc = 0; cc < 128; cc += 32) {
              float f[32];
              tmem_load_group<32>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + col0 + cc, f);
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[cc + i] += f[i];
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&t_empty[acc]), 0));
          if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
        }
        if (mine) {
          const int x = pos.x0 + wl;
          const bool valid = (y < P.H) && (x < P.W);
#pragma unroll
          for (int cc = 0; cc < 128; cc += 32) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = sum[cc + i];
            epilogue_store_group<32, false>(pr, f, s_bias_w + col0 + cc, nb * BN + col0 + cc, n, y, x, P.H, P.W, valid);
          }
        }
      } else {
      ptx::mbar_wait(&t_full[acc], pacc);
      ptx::tc_fence_after();
      int item = 0;
      for (int mt = 0; mt < n_mma; ++mt) {
        const int jb = 2 * mt + static_cast<int>(rank);
        if (jb >= n_sub) continue;                       // this CTA's block of the last MMA is outside the image
        const int x = pos.x0 + 16 * mt + wl;
        const bool valid = (y < P.H) && (x < P.W);
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += CW, ++item) {
          if ((item & (EPI_SETS - 1)) != eset) continue;
          float f[CW];
          tmem_load_group<CW>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (acc * MT + mt) * BN + cc, f);
          epilogue_store_group<CW, false>(pr, f, s_bias_w + cc, nb * BN + cc, n, y, x, P.H, P.W, valid);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&t_empty[acc]), 0));
      if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // nobody signals into / reads from the peer after this point
  if (warp == 1) ptx::tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace opb
