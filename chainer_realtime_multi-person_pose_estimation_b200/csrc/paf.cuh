// paf.cuh -- PAF line-integral scoring, greedy limb assignment and person grouping.
//   paf_candidates_kernel : compute_candidate_connections, pose_detector.py:135-157
//   limb_assign_kernel    : stable descending sort + greedy accept, pose_detector.py:158,172-177
//   group_persons_kernel  : grouping_key_points, pose_detector.py:183-250 (+ :252-265 packing)
//
// Arithmetic is float64 with explicit round-to-nearest intrinsics in the reference's order:
//   np.linspace (scalar-call semantics)  y_i = i*((b-a)/9) + a, y_9 = b
//   np.round -> rint (half-to-even), astype('i')
//   np.dot((10,2) f32, (2,) f64) evaluates as fma(p0, ux, p1*uy) (OpenBLAS dgemv on the host
//   that produced the goldens); sum of 10 in numpy's pairwise order; /10; + min(len/norm-1, 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "peaks.cuh"
#include "upsample.cuh"

namespace opb {

struct Candidate {
  double score;
  uint32_t pair;   // a_local * nB + b_local  (the reference's a-major, b-minor generation order)
  uint32_t state;  // 0 alive, 1 dead, 2 accepted
};

struct Connection {
  double score;
  int id_a, id_b;  // peak ids (rows of the peak table)
};

struct PafConsts {
  int limbs[19][2];
  double inner_product_thresh, limb_length_ratio, length_penalty_value;
  double n_subset_limbs_thresh, subset_score_thresh;
  int n_integ_points_thresh;
  int pad_;
};

constexpr int kAssignMaxType = 1024;  // peaks of one joint type per image handled by limb_assign

// PAF samplers: (Y, X) at map resolution -> the two float32 components, promoted to float64.
struct PafFull {   // full-resolution planes materialised by the upsample kernel (pose_detector.py:501)
  const float* p0;
  const float* p1;
  int W;
  __device__ __forceinline__ void operator()(int Y, int X, double& q0, double& q1) const {
    const size_t o = static_cast<size_t>(Y) * W + X;
    q0 = static_cast<double>(__ldg(p0 + o));
    q1 = static_cast<double>(__ldg(p1 + o));
  }
};
struct PafLow {    // low-resolution planes, sampled on demand with the upsample kernel's own arithmetic
  const float* p0;
  const float* p1;
  int h, w, H, W;
  double step_x, step_y;
  __device__ __forceinline__ void operator()(int Y, int X, double& q0, double& q1) const {
    const AcTap t = ac_tap(ac_axis_frac(X, w, W, step_x), ac_axis_frac(Y, h, H, step_y), w);
    q0 = static_cast<double>(ac_sample(p0, w, t));
    q1 = static_cast<double>(ac_sample(p1, w, t));
  }
};

// scores the nA x nB pairs of one (limb, image) and appends the candidates that pass (:135-157)
template <class Sampler>
__device__ __forceinline__ void paf_candidates_body(const Sampler& smp, int H, int W, const PeakD* __restrict__ pk,
                                                    const int* __restrict__ il, int a0, int nA, int b0, int nB,
                                                    const PafConsts& K, double img_len, Candidate* __restrict__ out,
                                                    int* __restrict__ cnt, int cand_cap) {
  const long long total = static_cast<long long>(nA) * nB;
  for (long long pi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pi < total;
       pi += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(pi / nB), b = static_cast<int>(pi - static_cast<long long>(a) * nB);
    const PeakD A = pk[il[a0 + a]], B = pk[il[b0 + b]];
    const double vx = __dsub_rn(B.x, A.x), vy = __dsub_rn(B.y, A.y);
    const double norm = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    if (norm == 0.0) continue;
    const double ux = __ddiv_rn(vx, norm), uy = __ddiv_rn(vy, norm);
    const double sx = __ddiv_rn(vx, 9.0), sy = __ddiv_rn(vy, 9.0);
    double ip[10];
    int nvalid = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      double ys, xs;
      if (i == 9) {
        ys = B.y; xs = B.x;
      } else {
        const double di = static_cast<double>(i);
        ys = __dadd_rn((sy == 0.0) ? __dmul_rn(__ddiv_rn(di, 9.0), vy) : __dmul_rn(di, sy), A.y);
        xs = __dadd_rn((sx == 0.0) ? __dmul_rn(__ddiv_rn(di, 9.0), vx) : __dmul_rn(di, sx), A.x);
      }
      // peaks lie inside the map, so the clamp never changes a valid input (memory safety only)
      const int Y = min(max(__double2int_rn(ys), 0), H - 1), X = min(max(__double2int_rn(xs), 0), W - 1);
      double q0, q1;
      smp(Y, X, q0, q1);
      ip[i] = __fma_rn(q0, ux, __dmul_rn(q1, uy));
      nvalid += (ip[i] > K.inner_product_thresh) ? 1 : 0;
    }
    // numpy pairwise sum for n = 10: 8-way unrolled block, then the two leftovers
    double s = __dadd_rn(__dadd_rn(__dadd_rn(ip[0], ip[1]), __dadd_rn(ip[2], ip[3])),
                         __dadd_rn(__dadd_rn(ip[4], ip[5]), __dadd_rn(ip[6], ip[7])));
    s = __dadd_rn(s, ip[8]);
    s = __dadd_rn(s, ip[9]);
    const double integ = __ddiv_rn(s, 10.0);
    double prior = __dsub_rn(__ddiv_rn(__dmul_rn(K.limb_length_ratio, img_len), norm), K.length_penalty_value);
    prior = (prior < 0.0) ? prior : 0.0;
    const double score = __dadd_rn(integ, prior);
    if (nvalid > K.n_integ_points_thresh && score > 0.0) {
      const int slot = atomicAdd(cnt, 1);
      if (slot < cand_cap) {
        Candidate c;
        c.score = score;
        c.pair = static_cast<uint32_t>(pi);
        c.state = 0;
        out[slot] = c;
      }
    }
  }
}

// grid (chunks, 19, n_img); paf [n_img][38][H][W] f32
__global__ void __launch_bounds__(128)
paf_candidates_kernel(const float* __restrict__ paf, int H, int W, const PeakD* __restrict__ peaks,
                      const int* __restrict__ idx_list, const int* __restrict__ type_start, int peaks_cap,
                      int n_types, PafConsts K, double img_len, Candidate* __restrict__ cands,
                      int* __restrict__ cand_counts, int cand_cap) {
  const int l = blockIdx.y, img = blockIdx.z;
  const int ja = K.limbs[l][0], jb = K.limbs[l][1];
  const int* ts = type_start + img * (n_types + 1);
  const int a0 = ts[ja], nA = ts[ja + 1] - a0;
  const int b0 = ts[jb], nB = ts[jb + 1] - b0;
  if (static_cast<long long>(nA) * nB == 0) return;
  PafFull smp;
  smp.p0 = paf + (static_cast<size_t>(img) * 38 + 2 * l) * H * W;
  smp.p1 = smp.p0 + static_cast<size_t>(H) * W;
  smp.W = W;
  paf_candidates_body(smp, H, W, peaks + static_cast<size_t>(img) * peaks_cap, idx_list + static_cast<size_t>(img) * peaks_cap,
                      a0, nA, b0, nB, K, img_len, cands + (static_cast<size_t>(img) * 19 + l) * cand_cap,
                      cand_counts + img * 19 + l, cand_cap);
}

// The same with the PAFs still at network resolution: paf_lo [n_img][38][h][w] f32, (H, W) = the map size the
// reference upsamples to (pose_detector.py:501).  Every sample is the value F.resize_images would have produced
// at that position, so the 38 full-resolution planes (28 MB per 320x576 image) are never written or read.
__global__ void __launch_bounds__(128)
paf_candidates_lowres_kernel(const float* __restrict__ paf_lo, int h, int w, int H, int W,
                             const PeakD* __restrict__ peaks, const int* __restrict__ idx_list,
                             const int* __restrict__ type_start, int peaks_cap, int n_types, PafConsts K, double img_len,
                             Candidate* __restrict__ cands, int* __restrict__ cand_counts, int cand_cap) {
  const int l = blockIdx.y, img = blockIdx.z;
  const int ja = K.limbs[l][0], jb = K.limbs[l][1];
  const int* ts = type_start + img * (n_types + 1);
  const int a0 = ts[ja], nA = ts[ja + 1] - a0;
  const int b0 = ts[jb], nB = ts[jb + 1] - b0;
  if (static_cast<long long>(nA) * nB == 0) return;
  PafLow smp;
  smp.p0 = paf_lo + (static_cast<size_t>(img) * 38 + 2 * l) * h * w;
  smp.p1 = smp.p0 + static_cast<size_t>(h) * w;
  smp.h = h; smp.w = w; smp.H = H; smp.W = W;
  smp.step_x = ac_step(w, W);
  smp.step_y = ac_step(h, H);
  paf_candidates_body(smp, H, W, peaks + static_cast<size_t>(img) * peaks_cap, idx_list + static_cast<size_t>(img) * peaks_cap,
                      a0, nA, b0, nB, K, img_len, cands + (static_cast<size_t>(img) * 19 + l) * cand_cap,
                      cand_counts + img * 19 + l, cand_cap);
}

// One block per (limb, image).  Exact greedy matching by descending (score, then generation
// order) without a global sort: a candidate that is the best alive one of BOTH its endpoints
// is exactly what the sequential greedy loop would accept next for those endpoints; accept
// all such candidates, kill the ones sharing an endpoint, repeat.  Every round first compacts the
// still-alive candidates into the other of two buffers (cands <-> cands_alt), so the work per
// round is proportional to what is left, not to the original list.  Accepted connections are
// finally ordered by (score desc, pair asc) = the reference's acceptance order.
constexpr int kAssignThreads = 512;
__global__ void __launch_bounds__(kAssignThreads)
limb_assign_kernel(const PeakD* __restrict__ peaks, const int* __restrict__ idx_list,
                   const int* __restrict__ type_start, int peaks_cap, int n_types, PafConsts K,
                   Candidate* __restrict__ cands, Candidate* __restrict__ cands_alt, int* __restrict__ cand_counts,
                   int cand_cap, Connection* __restrict__ conns, int* __restrict__ conn_counts, int conn_cap,
                   int* __restrict__ status) {
  __shared__ unsigned long long bestA_s[kAssignMaxType], bestB_s[kAssignMaxType];
  __shared__ unsigned int bestA_i[kAssignMaxType], bestB_i[kAssignMaxType];
  __shared__ unsigned char usedA[kAssignMaxType], usedB[kAssignMaxType];
  __shared__ int s_cnt, s_nacc;
  __shared__ double acc_score[kAssignMaxType];        // accepted connections, by value
  __shared__ unsigned int acc_pair[kAssignMaxType];

  const int l = blockIdx.x, img = blockIdx.y;
  const int ja = K.limbs[l][0], jb = K.limbs[l][1];
  const int* ts = type_start + img * (n_types + 1);
  const int a0 = ts[ja], nA = ts[ja + 1] - a0;
  const int b0 = ts[jb], nB = ts[jb + 1] - b0;
  Candidate* src = cands + (static_cast<size_t>(img) * 19 + l) * cand_cap;
  Candidate* dst = cands_alt + (static_cast<size_t>(img) * 19 + l) * cand_cap;
  Connection* out = conns + (static_cast<size_t>(img) * 19 + l) * conn_cap;
  int m = cand_counts[img * 19 + l];
  if (threadIdx.x == 0) {
    s_nacc = 0;
    if (m > cand_cap) atomicOr(&status[img], 2);
    if (nA > kAssignMaxType || nB > kAssignMaxType) atomicOr(&status[img], 4);
  }
  if (m > cand_cap) m = cand_cap;
  if (nA > kAssignMaxType || nB > kAssignMaxType || nA == 0 || nB == 0) {
    if (threadIdx.x == 0) conn_counts[img * 19 + l] = 0;
    return;
  }
  for (int i = threadIdx.x; i < nA; i += blockDim.x) usedA[i] = 0;
  for (int i = threadIdx.x; i < nB; i += blockDim.x) usedB[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;

  for (int round = 0; round < 2 * kAssignMaxType + 2 && m > 0; ++round) {
    for (int i = threadIdx.x; i < nA; i += blockDim.x) { bestA_s[i] = 0ull; bestA_i[i] = 0xffffffffu; }
    for (int i = threadIdx.x; i < nB; i += blockDim.x) { bestB_s[i] = 0ull; bestB_i[i] = 0xffffffffu; }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    // pass 1: drop candidates with a used endpoint, compact the rest, best score per endpoint
    for (int base = 0; base < m; base += blockDim.x) {
      const int i = base + threadIdx.x;
      Candidate c;
      bool alive = false;
      int a = 0, b = 0;
      if (i < m) {
        c = src[i];
        a = c.pair / nB; b = c.pair - a * nB;
        alive = !(usedA[a] || usedB[b]);
      }
      const unsigned mask = __ballot_sync(0xffffffffu, alive);
      int wbase = 0;
      if (lane == 0 && mask) wbase = atomicAdd(&s_cnt, __popc(mask));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (alive) {
        dst[wbase + __popc(mask & ((1u << lane) - 1))] = c;
        const unsigned long long sb = static_cast<unsigned long long>(__double_as_longlong(c.score));  // score > 0
        atomicMax(&bestA_s[a], sb);
        atomicMax(&bestB_s[b], sb);
      }
    }
    __syncthreads();
    m = s_cnt;
    { Candidate* t = src; src = dst; dst = t; }
    if (m == 0) break;
    // pass 2: among equal best scores the earliest generated pair wins (stable sort order)
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const Candidate c = src[i];
      const int a = c.pair / nB, b = c.pair - a * nB;
      const unsigned long long sb = static_cast<unsigned long long>(__double_as_longlong(c.score));
      if (sb == bestA_s[a]) atomicMin(&bestA_i[a], c.pair);
      if (sb == bestB_s[b]) atomicMin(&bestB_i[b], c.pair);
    }
    __syncthreads();
    // pass 3: mutual best -> accepted
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const Candidate c = src[i];
      const int a = c.pair / nB, b = c.pair - a * nB;
      const unsigned long long sb = static_cast<unsigned long long>(__double_as_longlong(c.score));
      if (sb == bestA_s[a] && c.pair == bestA_i[a] && sb == bestB_s[b] && c.pair == bestB_i[b]) {
        usedA[a] = 1;
        usedB[b] = 1;
        const int k = atomicAdd(&s_nacc, 1);
        if (k < kAssignMaxType) { acc_score[k] = c.score; acc_pair[k] = c.pair; }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int nacc = min(s_nacc, kAssignMaxType);
  if (nacc > conn_cap && threadIdx.x == 0) atomicOr(&status[img], 8);
  // rank sort by (score desc, pair asc)
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    const double si = acc_score[i];
    const unsigned int pi = acc_pair[i];
    int rank = 0;
    for (int j = 0; j < nacc; ++j) {
      const double sj = acc_score[j];
      if (sj > si || (sj == si && acc_pair[j] < pi)) ++rank;
    }
    if (rank < conn_cap) {
      const int a = pi / nB, b = pi - a * nB;
      Connection c;
      c.score = si;
      c.id_a = idx_list[static_cast<size_t>(img) * peaks_cap + a0 + a];
      c.id_b = idx_list[static_cast<size_t>(img) * peaks_cap + b0 + b];
      out[rank] = c;
    }
  }
  if (threadIdx.x == 0) conn_counts[img * 19 + l] = min(nacc, conn_cap);
}

struct PersonOut {   // == opb_person
  double score, count;
  int peak_id[18];
  int x[18];
  int y[18];
  int pad[2];
};
struct ImageHeader {  // == opb_image_header
  int n_peaks, n_persons, status, n_connections;
};

// One warp per image.  The subset table of grouping_key_points lives in shared memory for the whole kernel:
//   ids  [18][max_persons] int16  peak id per joint (-1 empty, -2 = row removed by a merge), column-major
//   tot  [max_persons]     f64    subset[-2] total score
//   cnt  [max_persons]     f64    subset[-1] joint count (non-integer after a merge, :216-217)
// (52 B per row; max_persons <= 4096 fits the 227 KB of one CTA), plus one bit per peak saying whether any row
// holds it -- a connection between two peaks no row holds starts a new subset without scanning the table; the
// scan itself reads four rows per lane and load.  Connections are fetched 32 at a time (one per
// lane, with both endpoint peak scores) and broadcast with shuffles, so the sequential merge loop never waits on
// global memory.  np.delete on a merge (:218) marks the row dead instead of shifting the rows below it; dead rows
// are skipped by the final filter, which preserves the reference's row order, and are reclaimed by one compaction
// if the table fills up.  subsets_out (optional): final kept rows [n_img][max_persons][20] float64.
inline size_t group_smem_bytes(int max_persons, int peaks_cap) {
  return static_cast<size_t>(16) * max_persons + static_cast<size_t>(36) * ((max_persons + 3) & ~3) + 4 * ((peaks_cap + 31) / 32) + 16;
}
__global__ void __launch_bounds__(32)
group_persons_kernel(const PeakD* __restrict__ peaks, const int* __restrict__ peak_counts, int peaks_cap,
                     PafConsts K, const Connection* __restrict__ conns, const int* __restrict__ conn_counts,
                     int conn_cap, int max_persons, const int* __restrict__ status_in,
                     ImageHeader* __restrict__ headers, PersonOut* __restrict__ persons,
                     double* __restrict__ subsets_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  double* tot = reinterpret_cast<double*>(s_raw);
  double* cnt = tot + max_persons;
  short* ids = reinterpret_cast<short*>(cnt + max_persons);      // ids[j * stride + k]
  const int stride = (max_persons + 3) & ~3;                      // 8-byte aligned columns
  unsigned* in_row = reinterpret_cast<unsigned*>(ids + 18 * stride);   // [peaks_cap / 32] peak held by some row
  const int img = blockIdx.x, lane = threadIdx.x;
  for (int i = lane; i < (peaks_cap + 31) / 32; i += 32) in_row[i] = 0u;
  __syncwarp();
  const PeakD* pk = peaks + static_cast<size_t>(img) * peaks_cap;
  int P = 0;               // rows in use (alive + dead)
  int n_dead = 0;
  int err = 0;
  int n_conn_total = 0;
  if (status_in[img]) err = -4;  // OPB_ERR_CAPACITY from an earlier stage

  for (int l = 0; l < 19 && !err; ++l) {
    const int ja = K.limbs[l][0], jb = K.limbs[l][1];
    const Connection* cl = conns + (static_cast<size_t>(img) * 19 + l) * conn_cap;
    const int nc = conn_counts[img * 19 + l];
    n_conn_total += nc;
    short* col_a = ids + ja * stride;
    short* col_b = ids + jb * stride;
    for (int c0 = 0; c0 < nc && !err; c0 += 32) {
      // lane i fetches connection c0 + i and the scores of its two peaks
      Connection mine;
      mine.score = 0.0; mine.id_a = 0; mine.id_b = 0;
      double my_sa = 0.0, my_sb = 0.0;
      if (c0 + lane < nc) {
        mine = cl[c0 + lane];
        my_sa = static_cast<double>(pk[mine.id_a].score);
        my_sb = static_cast<double>(pk[mine.id_b].score);
      }
      const int chunk = min(32, nc - c0);
      for (int t = 0; t < chunk && !err; ++t) {
        const int id_a = __shfl_sync(0xffffffffu, mine.id_a, t), id_b = __shfl_sync(0xffffffffu, mine.id_b, t);
        const double score = __shfl_sync(0xffffffffu, mine.score, t);
        const double sa = __shfl_sync(0xffffffffu, my_sa, t), sb = __shfl_sync(0xffffffffu, my_sb, t);
        const short ha = static_cast<short>(id_a), hb = static_cast<short>(id_b);
        // find the subsets that already hold one endpoint (first two, in row order)
        int found = 0, f0 = -1, f1 = -1;
        const bool known = ((in_row[id_a >> 5] >> (id_a & 31)) | (in_row[id_b >> 5] >> (id_b & 31))) & 1u;
        __syncwarp();   // every lane has read in_row before lane 0 may set bits for this connection below
        for (int base = 0; known && base < P; base += 128) {
          const int k0 = base + lane * 4;
          unsigned h = 0;
          if (k0 < P) {
            const uint2 va = *reinterpret_cast<const uint2*>(col_a + k0), vb = *reinterpret_cast<const uint2*>(col_b + k0);
            const short a4[4] = {static_cast<short>(va.x & 0xffff), static_cast<short>(va.x >> 16),
                                 static_cast<short>(va.y & 0xffff), static_cast<short>(va.y >> 16)};
            const short b4[4] = {static_cast<short>(vb.x & 0xffff), static_cast<short>(vb.x >> 16),
                                 static_cast<short>(vb.y & 0xffff), static_cast<short>(vb.y >> 16)};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k0 + e < P && (a4[e] == ha || b4[e] == hb)) h |= 1u << e;
          }
          unsigned any = __ballot_sync(0xffffffffu, h != 0);
          while (any) {                                    // lanes ascending, bits ascending = row order
            const int src = __ffs(any) - 1;
            any &= any - 1;
            unsigned hl = __shfl_sync(0xffffffffu, h, src);
            while (hl) {
              const int e = __ffs(hl) - 1;
              hl &= hl - 1;
              if (found == 0) f0 = base + src * 4 + e;
              else if (found == 1) f1 = base + src * 4 + e;
              ++found;
            }
          }
        }
        if (found >= 3) { err = -5; break; }   // reference: IndexError at pose_detector.py:197
        if (found == 1) {
          if (lane == 0 && col_b[f0] != hb) {            // a match through joint_b changes nothing (:200-206)
            col_b[f0] = hb;
            in_row[id_b >> 5] |= 1u << (id_b & 31);
            cnt[f0] = __dadd_rn(cnt[f0], 1.0);
            tot[f0] = __dadd_rn(tot[f0], __dadd_rn(sb, score));
          }
        } else if (found == 2) {
          const int v1 = (lane < 18) ? ids[lane * stride + f0] : -1;
          const int v2 = (lane < 18) ? ids[lane * stride + f1] : -1;
          const bool overlap = __any_sync(0xffffffffu, v1 >= 0 && v2 >= 0);
          if (!overlap) {
            if (lane < 18) {
              ids[lane * stride + f0] = static_cast<short>(v1 + v2 + 1);    // subset1[:-2] += subset2[:-2] + 1
              ids[lane * stride + f1] = -2;                                  // np.delete(subsets, f1, axis=0)
            }
            if (lane == 18) tot[f0] = __dadd_rn(__dadd_rn(tot[f0], tot[f1]), score);   // [-2:] += ... ; += score (:216-217)
            if (lane == 19) cnt[f0] = __dadd_rn(__dadd_rn(cnt[f0], cnt[f1]), score);
            ++n_dead;
          } else if (lane == 0) {
            for (int w = 0; w < 2; ++w) {
              const int row = w ? f1 : f0;
              if (col_a[row] == -1) {
                col_a[row] = ha;
                in_row[id_a >> 5] |= 1u << (id_a & 31);
                cnt[row] = __dadd_rn(cnt[row], 1.0);
                tot[row] = __dadd_rn(tot[row], __dadd_rn(sa, score));
              } else if (col_b[row] == -1) {
                col_b[row] = hb;
                in_row[id_b >> 5] |= 1u << (id_b & 31);
                cnt[row] = __dadd_rn(cnt[row], 1.0);
                tot[row] = __dadd_rn(tot[row], __dadd_rn(sb, score));
              }
            }
          }
        } else if (found == 0 && l != 9 && l != 13) {
          if (P >= max_persons && n_dead > 0) {            // reclaim removed rows before giving up
            __syncwarp();
            int w = 0;
            for (int r = 0; r < P; ++r) {
              if (ids[r] == -2) continue;
              if (w != r) {
                if (lane < 18) ids[lane * stride + w] = ids[lane * stride + r];
                if (lane == 18) tot[w] = tot[r];
                if (lane == 19) cnt[w] = cnt[r];
                __syncwarp();
              }
              ++w;
            }
            P = w;
            n_dead = 0;
          }
          if (P >= max_persons) { err = -4; break; }
          if (lane < 18) ids[lane * stride + P] = (lane == ja) ? ha : (lane == jb) ? hb : static_cast<short>(-1);
          if (lane == 18) tot[P] = __dadd_rn(__dadd_rn(sa, sb), score);
          if (lane == 19) cnt[P] = 2.0;
          if (lane == 0) { in_row[id_a >> 5] |= 1u << (id_a & 31); in_row[id_b >> 5] |= 1u << (id_b & 31); }
          ++P;
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  // final filter (:248-249) and packing (:252-265)
  int kept = 0;
  if (!err) {
    for (int k = 0; k < P; ++k) {
      if (ids[k] == -2) continue;                        // row removed by a merge
      const double c = cnt[k], sc = tot[k];
      const bool keep = (c >= K.n_subset_limbs_thresh) && (__ddiv_rn(sc, c) >= K.subset_score_thresh);
      if (keep) {
        const int id = (lane < 18) ? ids[lane * stride + k] : -1;
        if (subsets_out && lane < 20)
          subsets_out[(static_cast<size_t>(img) * max_persons + kept) * 20 + lane] =
              lane < 18 ? static_cast<double>(id) : (lane == 18 ? sc : c);
        if (persons) {
          PersonOut* po = persons + static_cast<size_t>(img) * max_persons + kept;
          if (lane < 18) {
            po->peak_id[lane] = id;
            po->x[lane] = (id >= 0) ? static_cast<int>(pk[id].x) : 0;
            po->y[lane] = (id >= 0) ? static_cast<int>(pk[id].y) : 0;
          }
          if (lane == 18) po->score = sc;
          if (lane == 19) po->count = c;
        }
        ++kept;
      }
    }
  }
  if (lane == 0) {
    ImageHeader h;
    h.n_peaks = peak_counts ? peak_counts[img] : 0;
    h.n_persons = kept;
    h.status = err;
    h.n_connections = n_conn_total;
    headers[img] = h;
  }
}

}  // namespace opb
