// conv_tcgen05.cuh -- implicit-GEMM convolution (stride 1, "same" zero padding, ksize 1/3/7)
// on the sm_100a 5th-gen tensor cores.  Replaces Chainer's L.Convolution2D + F.relu
// (+ F.concat by writing into channel slices) of models/CocoPoseNet.py:136-260.
//
// Data layout
//   activations  NHWC fp16, channel count padded to a multiple of 64; tensor map
//                {C, W, H, N}, box {64 ch, 8 px, 16+ks-1 rows, 1}, SWIZZLE_128B.  TMA
//                zero-fills out-of-image coordinates, which *is* the conv zero padding.
//   weights      [Cout_pad][ks*ks][Cin_pad] fp16, K-major ("B" operand), tensor map
//                {Ktot, Cout_pad}, box {64, BN}, SWIZZLE_128B.
//   accumulators fp32 in TMEM, ACC_STAGES x MT x BN columns.
//
// Tiling: one CTA tile = 16 output rows x (8*MT) output columns of one image x BN output
// channels.  Sub-tile mt (16 x 8 pixels = the 128 rows of one UMMA, row m = y*8 + x) has
// its own accumulator; all MT sub-tiles share every weight tile ("B" stage), so the
// L2->SM weight traffic per FLOP drops by MT.
//
// The "A" operand for the 7 (or 3) vertical taps r of one horizontal tap s comes from ONE
// TMA box of 16+ks-1 rows: tap r simply starts r*1024 bytes (one 8-pixel row group = one
// 1024-byte swizzle atom) further into the same shared-memory buffer, so every UMMA
// descriptor stays 1024-byte aligned and activations are fetched ks (not ks*ks) times.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane issues tcgen05.mma), warps 2..5 = epilogue (tcgen05.ld -> bias -> ReLU
// -> fp16 [hi/lo] -> global).  Three mbarrier pipelines: A ring, B ring, TMEM full/empty.
// The kernel is persistent: tile = blockIdx.x + i*gridDim.x.
//
// Precision: OPB_PRECISION_FAST stores fp16 activations/weights (fp32 accumulate).
// OPB_PRECISION_PARITY stores x = hi + lo (two fp16 planes) and accumulates
// hi*Whi + lo*Whi + hi*Wlo into the same TMEM accumulator: the K loop just walks a table of
// (activation-channel-offset, weight-k-offset) chunk pairs, so both modes run this kernel.
// OPB_PRECISION_COMP ("compensated") keeps the fp16 main product and adds the two first-order rounding
// corrections as 8-bit-float MMAs at twice the fp16 rate: per 64-channel chunk one extra K = 128 row
//   [ fp8(x_lo * 2^11) | fp8(x) ]  .  [ e4m3(W * 2^kW) | e4m3(W_lo * 2^S) ],   S = kW + 11,
// accumulated into the SAME TMEM accumulator as x_hi . (W_hi * 2^S) (the fp16 weights are pre-scaled by
// the per-layer power of two 2^S, which is exact; the epilogue multiplies by 2^-S).  Cost: 2 MMAs per
// k-step instead of 3 + two-level accumulation; map error ~1e-4 (tolerance 1e-3).
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <type_traits>

#include "ptx.cuh"

namespace opb {

constexpr int kConvThreads = 192;    // 2 + 4 epilogue warps (parity / two-level accumulation kernels)
#ifndef OPB_EPI_SETS
#define OPB_EPI_SETS 1
#endif
#ifndef OPB_ROLE_REORDER
#define OPB_ROLE_REORDER 0
#endif
constexpr int kConvThreads2 = 64 + 128 * OPB_EPI_SETS;   // 2 + 8 epilogue warps (fast kernels): the epilogue is latency bound with one warp per SMSP
constexpr int kMaxPairs = 24;

struct ConvProblem {
  __half* out;          // NHWC fp16 output tensor base (or nullptr)
  float* out32;         // optional planar fp32 [N][cout_valid][H][W] (network heads), or nullptr
  const float* bias;    // [n_blocks*BN] fp32, zero padded
  int out_cstride;      // channels per pixel of the output tensor
  int out_coff;         // first channel of this problem inside the output tensor
  int out_lo_off;       // parity mode: distance (channels) from the hi plane to the lo plane, else 0
  int cout_valid;       // real number of output channels
  int relu;
  int pool;             // 1: fuse F.max_pooling_2d(2,2) -- `out` is the [N][H/2][W/2] pooled tensor
  float acc_scale;      // accumulator scale (2^-S in compensated precision, else 1)
};

struct ConvParams {
  int N, H, W;
  int tiles_x, tiles_y;  // ceil(W/(8*MT)), ceil(H/16)
  int n_blocks;          // Cout_pad / BN
  int n_pairs;           // number of 64-channel K chunk pairs
  int n_problems;        // 1 or 2 (grouped launch: the L1 / L2 branches of one stage)
  int b_tap_stride;      // K elements per filter tap in the packed weights
  int pad_edge8;         // swap kernel: the last tile of a row is 8 (not 16) pixels wide
  int comp;              // compensated precision: odd chunk pairs are the 8-bit-float correction rows
  int drain_seg;         // swap / pair DRAIN kernels: (chunk pair, filter column) steps accumulated per TMEM buffer
  int a_off[kMaxPairs];  // activation channel offset of chunk pair j
  int b_off[kMaxPairs];  // weight k offset (inside one tap) of chunk pair j
  ConvProblem prob[2];
  // swap7 kernel: host-computed tile schedule (longest-processing-time first): CTA c works through
  // sched[c * sched_len + 0 .. sched_len) until it meets -1.  nullptr = tile = blockIdx.x + i * gridDim.x.
  const int* sched;
  int sched_len;
  // pair kernel, MT = 1: 1 = tiles_x counts 8-column UNITS and the two CTAs of a pair take units 2t and 2t+1 of the
  // linear (image, tile row, unit) order -- they need not be neighbours in the image (each CTA loads its own A rows and
  // stores its own outputs), so an odd number of units per row costs no padding block: 82 columns -> 88, not 96.
  int pair_units;
};

template <int KS, int BN, int MT, int NSA, int NSB, int ACC_STAGES>
struct ConvCfg {
  static constexpr int RH = 16 + KS - 1;               // rows per activation box
  static constexpr int A_SUB_BYTES = RH * 1024;        // one sub-tile box
  static constexpr int A_STAGE_BYTES = MT * A_SUB_BYTES;
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * MT * BN;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                                   : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static_assert(TMEM_COLS_RAW <= 512, "accumulators do not fit TMEM");
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + NSA * A_STAGE_BYTES + NSB * B_STAGE_BYTES + 512 + 8 * BN * 4;
};


// ---- 8-bit float helpers (compensated precision) ----
// Activation-side correction bytes: e4m3 (4 significant bits; saturating at 448 -- a larger activation only loses part
// of its correction) by default, e5m2 (3 bits, fp16's range) with -DOPB_COMP_ACT_E5M2=1.  Weight side: always e4m3.
#ifndef OPB_COMP_ACT_E5M2
#define OPB_COMP_ACT_E5M2 0
#endif
constexpr int kCompActFmt = OPB_COMP_ACT_E5M2 ? 1 : 0;   // tcgen05 kind::f8f6f4 format code: 0 = E4M3, 1 = E5M2
#define OPB_NV_ACT_FMT (OPB_COMP_ACT_E5M2 ? __NV_E5M2 : __NV_E4M3)
constexpr float kCompLoScale = 2048.f;   // x_lo is stored as fp8(x_lo * 2^11): |x_lo| <= 2^-11 |x|, so it is never larger than |x|
__host__ __device__ __forceinline__ uint8_t f32_to_e4m3(float v) {
  return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3));
}
__host__ __device__ __forceinline__ uint8_t f32_to_act8(float v) {
  return static_cast<uint8_t>(__nv_cvt_float_to_fp8(v, __NV_SATFINITE, OPB_NV_ACT_FMT));
}
__host__ __device__ __forceinline__ uint32_t f32x4_to_act8x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, OPB_NV_ACT_FMT);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, OPB_NV_ACT_FMT);
  return lo | (hi << 16);
}
// value of one activation-side correction byte (host side of opb_test_conv)
__host__ inline float act8_to_f32(uint8_t b) {
  if (OPB_COMP_ACT_E5M2) {   // an e5m2 byte is the high byte of the fp16 with the same value
    const __half_raw hr{static_cast<unsigned short>(static_cast<unsigned short>(b) << 8)};
    return __half2float(__half(hr));
  }
  const int e = (b >> 3) & 15, m = b & 7;
  const float v = e ? ldexpf(1.f + m / 8.f, e - 7) : ldexpf(m / 8.f, -6);
  return (b & 0x80) ? -v : v;
}
// byte address of tensor channel t in the correction plane of one pixel: [chunk t/64][x_lo 64 B | x 64 B]
__host__ __device__ __forceinline__ int comp_byte_off(int t) { return ((t >> 6) << 7) + (t & 63); }

// hi (fp16) + correction bytes of CW consecutive output channels of one pixel; f = final fp32 values.
// o = address of the hi plane value of the first channel, corr = base of the pixel's correction plane, t0 = tensor
// channel of the first value.
template <int CW>
__device__ __forceinline__ void comp_store(const float (&f)[CW], __half* o, uint8_t* corr, int t0, int nvalid) {
  const bool vec = (nvalid == CW) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && ((t0 & 15) == 0);
  if (vec) {
#pragma unroll
    for (int g = 0; g < CW / 16; ++g) {
      uint32_t h[8], xl[4], x8[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __half2 t = __floats2half2_rn(f[g * 16 + 2 * i], f[g * 16 + 2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float v = f[g * 16 + 4 * i + k];
          l[k] = (v - __half2float(__float2half_rn(v))) * kCompLoScale;
        }
        xl[i] = f32x4_to_act8x4(l[0], l[1], l[2], l[3]);
        x8[i] = f32x4_to_act8x4(f[g * 16 + 4 * i], f[g * 16 + 4 * i + 1], f[g * 16 + 4 * i + 2], f[g * 16 + 4 * i + 3]);
      }
      *reinterpret_cast<uint4*>(o + g * 16) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(o + g * 16 + 8) = make_uint4(h[4], h[5], h[6], h[7]);
      uint8_t* c = corr + comp_byte_off(t0 + g * 16);
      *reinterpret_cast<uint4*>(c) = make_uint4(xl[0], xl[1], xl[2], xl[3]);
      *reinterpret_cast<uint4*>(c + 64) = make_uint4(x8[0], x8[1], x8[2], x8[3]);
    }
  } else if ((t0 & 1) == 0 && (reinterpret_cast<uintptr_t>(o) & 3) == 0 && ((t0 & 63) + CW <= 64)) {
    // unaligned slice (the heat maps start at channel 166 of the concat tensor, the PAF slice ends after 38): channel
    // PAIRS -- one 4-byte store for the two fp16 values and one 2-byte store per correction plane -- instead of three
    // scalar stores per channel.  (Each scattered store instruction of a pixel-per-lane epilogue costs 32 sectors; the
    // 1x1 head kernel is bound by them: profiles/r02_ncu_plain_comp_1_48_summary.txt.)
    uint8_t* c = corr + comp_byte_off(t0);
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) {
      if (2 * i + 1 < nvalid) {
        const float a = f[2 * i], b = f[2 * i + 1];
        const __half2 t = __floats2half2_rn(a, b);
        *reinterpret_cast<__half2*>(o + 2 * i) = t;
        const float2 back = __half22float2(t);
        *reinterpret_cast<unsigned short*>(c + 2 * i) = static_cast<unsigned short>(
            __nv_cvt_float2_to_fp8x2(make_float2((a - back.x) * kCompLoScale, (b - back.y) * kCompLoScale), __NV_SATFINITE, OPB_NV_ACT_FMT));
        *reinterpret_cast<unsigned short*>(c + 64 + 2 * i) =
            static_cast<unsigned short>(__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, OPB_NV_ACT_FMT));
      } else if (2 * i < nvalid) {
        const __half hi = __float2half_rn(f[2 * i]);
        o[2 * i] = hi;
        c[2 * i] = f32_to_act8((f[2 * i] - __half2float(hi)) * kCompLoScale);
        c[64 + 2 * i] = f32_to_act8(f[2 * i]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < CW; ++i) {
      if (i < nvalid) {
        const __half hi = __float2half_rn(f[i]);
        o[i] = hi;
        uint8_t* c = corr + comp_byte_off(t0 + i);
        c[0] = f32_to_act8((f[i] - __half2float(hi)) * kCompLoScale);
        c[64] = f32_to_act8(f[i]);
      }
    }
  }
}

// bias + ReLU (+ fused 2x2 max-pool) + fp16 (hi[/lo]) store of CW consecutive output channels of
// one pixel.  Must be called by all 32 lanes of the warp (the pool uses shuffles): lane =
// (y%4)*8 + x%8 of a 4-row x 8-column patch, so the 2x2 partners are lane^1 and lane^8.
// `bias` points at this group's CW biases (shared memory, 16-byte aligned).
// SPLIT = parity precision (hi + lo planes, pooling on the fp32 value); otherwise the values are
// packed to half2 first and pooled with HMNMX2 (rounding is monotonic, so max commutes with it).
template <int CW, bool SPLIT>
__device__ __forceinline__ void epilogue_store_group(const ConvProblem& pr, const float (&acc)[CW],
                                                     const float* __restrict__ bias, int ch0, int n, int y, int x,
                                                     int H, int W, bool valid) {
  if (ch0 >= pr.cout_valid) return;   // warp-uniform
  float f[CW];
#pragma unroll
  for (int i = 0; i < CW / 4; ++i) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * i);
    const float sc = pr.acc_scale;          // 1 (fmaf(a, 1, b) == a + b) or the exact power of two 2^-S
    f[4 * i + 0] = fmaf(acc[4 * i + 0], sc, b.x);
    f[4 * i + 1] = fmaf(acc[4 * i + 1], sc, b.y);
    f[4 * i + 2] = fmaf(acc[4 * i + 2], sc, b.z);
    f[4 * i + 3] = fmaf(acc[4 * i + 3], sc, b.w);
  }
  if (pr.relu) {
#pragma unroll
    for (int i = 0; i < CW; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  int oy = y, ox = x, oH = H, oW = W;
  const int nvalid = min(CW, pr.cout_valid - ch0);
  if (pr.pool) {
    oy = y >> 1; ox = x >> 1; oH = H >> 1; oW = W >> 1;
    valid = valid && ((y & 1) == 0) && ((x & 1) == 0);
  }
  if constexpr (SPLIT) {
    if (pr.pool) {
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        f[i] = fmaxf(f[i], __shfl_xor_sync(0xffffffffu, f[i], 1));
        f[i] = fmaxf(f[i], __shfl_xor_sync(0xffffffffu, f[i], 8));
      }
    }
    if (!valid) return;
    if (pr.out32) {
      // constant trip count + predicate: a runtime-indexed f[] would be demoted to local memory
#pragma unroll
      for (int i = 0; i < CW; ++i)
        if (i < nvalid) pr.out32[((static_cast<size_t>(n) * pr.cout_valid + ch0 + i) * oH + oy) * oW + ox] = f[i];
    }
    if (pr.out) {
      const size_t pix = (static_cast<size_t>(n) * oH + oy) * oW + ox;
      __half* o = pr.out + pix * pr.out_cstride + pr.out_coff + ch0;
      const bool vec = (nvalid == CW) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && ((pr.out_lo_off & 7) == 0);
      if (vec) {
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
          __align__(16) __half2 h[4];
          __align__(16) __half2 l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = f[g * 8 + 2 * i], b = f[g * 8 + 2 * i + 1];
            const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
            h[i] = __halves2half2(ha, hb);
            l[i] = __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
          }
          *reinterpret_cast<uint4*>(o + g * 8) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(o + pr.out_lo_off + g * 8) = *reinterpret_cast<const uint4*>(l);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          if (i < nvalid) {
            const __half hi = __float2half_rn(f[i]);
            o[i] = hi;
            o[pr.out_lo_off + i] = __float2half_rn(f[i] - __half2float(hi));
          }
        }
      }
    }
  } else {
    if (pr.out32 && valid) {   // network heads: fp32 planar maps for the post-process (never pooled)
#pragma unroll
      for (int i = 0; i < CW; ++i)
        if (i < nvalid) pr.out32[((static_cast<size_t>(n) * pr.cout_valid + ch0 + i) * oH + oy) * oW + ox] = f[i];
    }
    if (pr.out && pr.out_lo_off) {   // compensated precision (warp-uniform): pool on the fp32 values, then hi + correction bytes
      if (pr.pool) {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          f[i] = fmaxf(f[i], __shfl_xor_sync(0xffffffffu, f[i], 1));
          f[i] = fmaxf(f[i], __shfl_xor_sync(0xffffffffu, f[i], 8));
        }
      }
      if (!valid) return;
      const size_t pix = (static_cast<size_t>(n) * oH + oy) * oW + ox;
      __half* px = pr.out + pix * pr.out_cstride;
      comp_store<CW>(f, px + pr.out_coff + ch0, reinterpret_cast<uint8_t*>(px + pr.out_lo_off), pr.out_coff + ch0, nvalid);
    } else if (pr.out) {
      uint32_t h[CW / 2];        // packed half2, kept in registers (no address-taken arrays)
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        const __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
      if (pr.pool) {
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
          uint32_t o1 = __shfl_xor_sync(0xffffffffu, h[i], 1);
          __half2 m = __hmax2(*reinterpret_cast<__half2*>(&h[i]), *reinterpret_cast<__half2*>(&o1));
          h[i] = *reinterpret_cast<uint32_t*>(&m);
          uint32_t o2 = __shfl_xor_sync(0xffffffffu, h[i], 8);
          m = __hmax2(*reinterpret_cast<__half2*>(&h[i]), *reinterpret_cast<__half2*>(&o2));
          h[i] = *reinterpret_cast<uint32_t*>(&m);
        }
      }
      if (!valid) return;
      const size_t pix = (static_cast<size_t>(n) * oH + oy) * oW + ox;
      __half* o = pr.out + pix * pr.out_cstride + pr.out_coff + ch0;
      const bool vec = (nvalid == CW) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
      if (vec) {
#pragma unroll
        for (int g = 0; g < CW / 8; ++g)
          *reinterpret_cast<uint4*>(o + g * 8) = make_uint4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
      } else if ((reinterpret_cast<uintptr_t>(o) & 3) == 0) {      // 4-byte aligned slice: channel pairs
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
          if (2 * i + 1 < nvalid) reinterpret_cast<uint32_t*>(o)[i] = h[i];
          else if (2 * i < nvalid) reinterpret_cast<unsigned short*>(o)[2 * i] = static_cast<unsigned short>(h[i] & 0xffffu);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) {
          if (i < nvalid) {
            const uint32_t w = h[i >> 1];
            const unsigned short bits = (i & 1) ? static_cast<unsigned short>(w >> 16) : static_cast<unsigned short>(w & 0xffffu);
            reinterpret_cast<unsigned short*>(o)[i] = bits;
          }
        }
      }
    }
  }
}

// each epilogue warp keeps its own copy of the current (problem, n-block) bias vector in shared memory
template <int BN>
__device__ __forceinline__ void epilogue_load_bias(float* s_bias_warp, const float* __restrict__ gbias, int lane) {
  __syncwarp();
#pragma unroll
  for (int i = lane; i < BN; i += 32) s_bias_warp[i] = __ldg(gbias + i);
  __syncwarp();
}

template <int CW>
__device__ __forceinline__ void tmem_load_group(uint32_t taddr, float (&f)[CW]) {
  if constexpr (CW == 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32b_x32(taddr, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  } else {
    uint32_t v[16];
    ptx::tmem_ld_32x32b_x16(taddr, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
  }
}

template <int KS, int BN, int MT, int NSA, int NSB, int ACC_STAGES, bool DRAIN, bool BRES = false>
__global__ void __launch_bounds__(DRAIN ? kConvThreads : kConvThreads2, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                    const __grid_constant__ ConvParams P) {
  using Cfg = ConvCfg<KS, BN, MT, NSA, NSB, ACC_STAGES>;
  constexpr int PAD = (KS - 1) / 2;
  constexpr uint32_t IDESC = ptx::umma_idesc_f16(128, BN);
  constexpr uint32_t IDESC8 = ptx::umma_idesc_f8(128, BN, kCompActFmt /*A: activations*/, 0 /*B: weights e4m3*/);

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment
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
  constexpr int EPI_SETS = DRAIN ? 1 : OPB_EPI_SETS;

  // warp roles: 0..kEpiWarps-1 epilogue, then the TMA producer, then the MMA issuer.  The SMSP arbiter
  // prefers the HIGHEST warp id, so the two latency-critical single-thread roles get the top ids.
  constexpr int kEpiWarps = 4 * EPI_SETS;
  const int warp_raw = threadIdx.x >> 5;
#if OPB_ROLE_REORDER
  const int warp = (warp_raw >= kEpiWarps) ? warp_raw - kEpiWarps : warp_raw + 2;   // logical: 0 TMA, 1 MMA, 2.. epilogue
#else
  const int warp = warp_raw;
#endif
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA0);
    ptx::prefetch_tensormap(&tmB0);
    if (P.n_problems > 1) {
      ptx::prefetch_tensormap(&tmA1);
      ptx::prefetch_tensormap(&tmB1);
    }
    for (int i = 0; i < NSA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NSB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < ACC_STAGES; ++i) { ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], 128 * EPI_SETS); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = P.N * P.tiles_y * P.tiles_x;
  const int tiles_per_problem = P.n_blocks * m_tiles;
  const int total_tiles = P.n_problems * tiles_per_problem;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (ptx::elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int p = tile / tiles_per_problem;
        int rem = tile - p * tiles_per_problem;
        const int nb = rem / m_tiles;
        rem -= nb * m_tiles;
        const int n = rem / (P.tiles_y * P.tiles_x);
        rem -= n * (P.tiles_y * P.tiles_x);
        const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
        const int y0 = ty * 16, x0 = tx * (8 * MT);
        const int n_sub = min(MT, (P.W - x0 + 7) >> 3);
        const CUtensorMap* tmA = p ? &tmA1 : &tmA0;
        const CUtensorMap* tmB = p ? &tmB1 : &tmB0;
        for (int j = 0; j < P.n_pairs; ++j) {
          const int ac = P.a_off[j], bk = P.b_off[j];
          for (int s = 0; s < KS; ++s) {
            ptx::mbar_wait(&a_empty[sa], pa ^ 1);
            ptx::mbar_expect_tx(&a_full[sa], n_sub * Cfg::A_SUB_BYTES);
            for (int mt = 0; mt < n_sub; ++mt)
              ptx::tma_load_4d(smemA + sa * Cfg::A_STAGE_BYTES + mt * Cfg::A_SUB_BYTES, tmA, &a_full[sa], ac,
                               x0 + mt * 8 + s - PAD, y0 - PAD, n);
            if (++sa == NSA) { sa = 0; pa ^= 1; }
            for (int r = 0; r < KS; ++r) {
              // BRES: the whole weight matrix (NSB = taps x chunks stages) stays resident in shared
              // memory; it is fetched once, by the CTA's first tile
              if (!BRES || tile == static_cast<int>(blockIdx.x)) {
                ptx::mbar_wait(&b_empty[sb], pb ^ 1);
                ptx::mbar_expect_tx(&b_full[sb], Cfg::B_STAGE_BYTES);
                ptx::tma_load_2d(smemB + sb * Cfg::B_STAGE_BYTES, tmB, &b_full[sb],
                                 (r * KS + s) * P.b_tap_stride + bk, nb * BN);
              }
              if (++sb == NSB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (ptx::elect_one()) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, pacc = 0;
      const uint64_t a_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemA), 1024);
      const uint64_t b_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(smemB), 1024);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int rem = tile % m_tiles;
        rem %= (P.tiles_y * P.tiles_x);
        const int tx = rem % P.tiles_x;
        const int x0 = tx * (8 * MT);
        const int n_sub = min(MT, (P.W - x0 + 7) >> 3);
        if (!DRAIN) {
          ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
          ptx::tc_fence_after();
        }
        uint32_t accumulate = 0;
        // One (chunk pair, filter column) step: KS weight stages x n_sub sub-tiles x 4 MMAs.  The MMA kind is a compile-time
        // property of the step, and when the weight ring holds a whole number of steps (NSB % KS == 0) the stage index
        // inside the unrolled tap loop is `first stage of the step + r`: barrier addresses and descriptors are one base
        // value per step plus compile-time offsets, and the ring wraps per step, not per tap -- the issuing thread is
        // the bottleneck of the short-MMA (N <= 128) layers (profiles/r02_issue_path.txt).
        constexpr bool kAlignedB = (NSB % KS) == 0;
        const bool first_tile = tile == static_cast<int>(blockIdx.x);
        auto do_step = [&](auto f8_tag) {
          constexpr bool F8 = decltype(f8_tag)::value;
          ptx::mbar_wait(&a_full[sa], pa);
          if (DRAIN) {   // two-level accumulation: a fresh TMEM accumulator per A stage
            ptx::mbar_wait(&t_empty[acc], pacc ^ 1);
            accumulate = 0;
          }
          ptx::tc_fence_after();
          // descriptors differ from the stage-0 descriptor only in the 14-bit address field
          const uint64_t a_st = a_desc0 + static_cast<uint64_t>((sa * Cfg::A_STAGE_BYTES) >> 4);
          const uint32_t sb0 = sb;                                   // first weight stage of this step
          const uint64_t b_st0 = b_desc0 + static_cast<uint64_t>((sb0 * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            const uint32_t st = kAlignedB ? sb0 + r : sb;             // aligned: no wrap inside a step
            if (!BRES || first_tile) {
              ptx::mbar_wait(&b_full[st], pb);
              ptx::tc_fence_after();
            }
            const uint64_t b_st = kAlignedB ? b_st0 + static_cast<uint64_t>((r * Cfg::B_STAGE_BYTES) >> 4)
                                            : b_desc0 + static_cast<uint64_t>((sb * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              if (mt < n_sub) {
                const uint32_t d = tmem_base + (acc * MT + mt) * BN;
                const uint64_t ad0 = a_st + static_cast<uint64_t>((mt * Cfg::A_SUB_BYTES + r * 1024) >> 4);
                if constexpr (F8) {
                  ptx::mma_f8_ss(d, ad0, b_st, IDESC8, accumulate);
#pragma unroll
                  for (int k = 1; k < 4; ++k) ptx::mma_f8_ss(d, ad0 + (k * 32 >> 4), b_st + (k * 32 >> 4), IDESC8, 1u);
                } else {
                  ptx::mma_f16_ss(d, ad0, b_st, IDESC, accumulate);
#pragma unroll
                  for (int k = 1; k < 4; ++k) ptx::mma_f16_ss_acc(d, ad0 + (k * 32 >> 4), b_st + (k * 32 >> 4), IDESC);
                }
              }
            }
            accumulate = 1;
            if (!BRES) ptx::mma_commit(&b_empty[st]);
            if constexpr (!kAlignedB) { if (++sb == NSB) { sb = 0; pb ^= 1; } }
          }
          if constexpr (kAlignedB) { sb += KS; if (sb == NSB) { sb = 0; pb ^= 1; } }
          ptx::mma_commit(&a_empty[sa]);
          if (++sa == NSA) { sa = 0; pa ^= 1; }
          if (DRAIN) {
            ptx::mma_commit(&t_full[acc]);
            if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
          }
        };
        for (int j = 0; j < P.n_pairs; ++j) {
          if (P.comp && (j & 1)) {             // compensated precision: odd pairs are the 8-bit correction rows
            for (int s = 0; s < KS; ++s) do_step(std::true_type{});
          } else {
            for (int s = 0; s < KS; ++s) do_step(std::false_type{});
          }
        }
        if (!DRAIN) {
          ptx::mma_commit(&t_full[acc]);
          if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
        }
      }
    }
  } else {
    // ================================================================ epilogue (warps 2..5 [, 6..9])
    const int q = warp_raw & 3;  // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int eset = (warp - 2) >> 2;          // which set of four epilogue warps (fast kernels have two)
    const int row = q * 32 + lane;
    const int hl = row >> 3, wl = row & 7;
    float* s_bias_w = s_bias + (warp - 2) * BN;
    int bias_key = -1;
    uint32_t acc = 0, pacc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int p = tile / tiles_per_problem;
      int rem = tile - p * tiles_per_problem;
      const int nb = rem / m_tiles;
      rem -= nb * m_tiles;
      const int n = rem / (P.tiles_y * P.tiles_x);
      rem -= n * (P.tiles_y * P.tiles_x);
      const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
      const int y = ty * 16 + hl;
      const int x0 = tx * (8 * MT);
      const int n_sub = min(MT, (P.W - x0 + 7) >> 3);
      const ConvProblem& pr = P.prob[p];
      if (bias_key != p * 1024 + nb) {         // (problem, n-block) changes a handful of times per launch
        bias_key = p * 1024 + nb;
        epilogue_load_bias<BN>(s_bias_w, pr.bias + nb * BN, lane);
      }

      constexpr int CW = (BN % 32 == 0) ? 32 : 16;   // BN = 48: three groups of 16
      if constexpr (!DRAIN) {
        ptx::mbar_wait(&t_full[acc], pacc);
        ptx::tc_fence_after();
        // the (sub-tile, channel-group) work items alternate between the two sets of epilogue warps
        int item = 0;
        for (int mt = 0; mt < n_sub; ++mt) {
          const int x = x0 + mt * 8 + wl;
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
        ptx::mbar_arrive(&t_empty[acc]);
        if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
      } else {
        // two-level accumulation (parity precision): the tensor core's fp32 accumulate
        // truncates, and the bias grows linearly with the number of chained MMAs (measured:
        // 6e-5 at K = 6272).  Each A stage (ks taps x 4 k-steps) lands in a fresh TMEM
        // accumulator; the partial sums are added here in round-to-nearest fp32 registers.
        static_assert(!DRAIN || MT == 1, "drain mode uses MT = 1");
        float sum[BN];
#pragma unroll
        for (int i = 0; i < BN; ++i) sum[i] = 0.f;
        const int n_seg = P.n_pairs * KS;
        for (int seg = 0; seg < n_seg; ++seg) {
          ptx::mbar_wait(&t_full[acc], pacc);
          ptx::tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < BN; cc += CW) {
            float f[CW];
            tmem_load_group<CW>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + cc, f);
#pragma unroll
            for (int i = 0; i < CW; ++i) sum[cc + i] += f[i];
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&t_empty[acc]);
          if (++acc == ACC_STAGES) { acc = 0; pacc ^= 1; }
        }
        const int x = x0 + wl;
        const bool valid = (y < P.H) && (x < P.W);
#pragma unroll
        for (int cc = 0; cc < BN; cc += CW) {
          float f[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) f[i] = sum[cc + i];
          if (P.comp) epilogue_store_group<CW, false>(pr, f, s_bias_w + cc, nb * BN + cc, n, y, x, P.H, P.W, valid);
          else epilogue_store_group<CW, true>(pr, f, s_bias_w + cc, nb * BN + cc, n, y, x, P.H, P.W, valid);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace opb
