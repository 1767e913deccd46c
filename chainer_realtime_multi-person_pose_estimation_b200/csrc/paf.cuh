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
  const long long total = static_cast<long long>(nA) * nB;
  if (total == 0) return;
  const PeakD* pk = peaks + static_cast<size_t>(img) * peaks_cap;
  const int* il = idx_list + static_cast<size_t>(img) * peaks_cap;
  const float* p0 = paf + (static_cast<size_t>(img) * 38 + 2 * l) * H * W;
  const float* p1 = p0 + static_cast<size_t>(H) * W;
  Candidate* out = cands + (static_cast<size_t>(img) * 19 + l) * cand_cap;
  int* cnt = cand_counts + img * 19 + l;

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
      const size_t o = static_cast<size_t>(Y) * W + X;
      const double q0 = static_cast<double>(__ldg(p0 + o)), q1 = static_cast<double>(__ldg(p1 + o));
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

// One block per (limb, image).  Exact greedy matching by descending (score, then generation
// order) without a global sort: a candidate that is the best alive one of BOTH its endpoints
// is exactly what the sequential greedy loop would accept next for those endpoints; accept
// all such candidates, kill the ones sharing an endpoint, repeat.  Accepted connections are
// finally ordered by (score desc, pair asc) = the reference's acceptance order.
__global__ void __launch_bounds__(256)
limb_assign_kernel(const PeakD* __restrict__ peaks, const int* __restrict__ idx_list,
                   const int* __restrict__ type_start, int peaks_cap, int n_types, PafConsts K,
                   Candidate* __restrict__ cands, int* __restrict__ cand_counts, int cand_cap,
                   Connection* __restrict__ conns, int* __restrict__ conn_counts, int conn_cap,
                   int* __restrict__ status) {
  __shared__ unsigned long long bestA_s[kAssignMaxType], bestB_s[kAssignMaxType];
  __shared__ unsigned int bestA_i[kAssignMaxType], bestB_i[kAssignMaxType];
  __shared__ unsigned char usedA[kAssignMaxType], usedB[kAssignMaxType];
  __shared__ int s_alive, s_nacc;
  __shared__ unsigned int acc_slot[kAssignMaxType];   // candidate index of each accepted connection

  const int l = blockIdx.x, img = blockIdx.y;
  const int ja = K.limbs[l][0], jb = K.limbs[l][1];
  const int* ts = type_start + img * (n_types + 1);
  const int a0 = ts[ja], nA = ts[ja + 1] - a0;
  const int b0 = ts[jb], nB = ts[jb + 1] - b0;
  Candidate* cd = cands + (static_cast<size_t>(img) * 19 + l) * cand_cap;
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

  for (int round = 0; round < 2 * kAssignMaxType + 2; ++round) {
    for (int i = threadIdx.x; i < nA; i += blockDim.x) { bestA_s[i] = 0ull; bestA_i[i] = 0xffffffffu; }
    for (int i = threadIdx.x; i < nB; i += blockDim.x) { bestB_s[i] = 0ull; bestB_i[i] = 0xffffffffu; }
    if (threadIdx.x == 0) s_alive = 0;
    __syncthreads();
    int alive_local = 0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      if (cd[i].state != 0) continue;
      const uint32_t pr = cd[i].pair;
      const int a = pr / nB, b = pr - a * nB;
      if (usedA[a] || usedB[b]) { cd[i].state = 1; continue; }
      alive_local = 1;
      const unsigned long long sb = static_cast<unsigned long long>(__double_as_longlong(cd[i].score));  // score > 0
      atomicMax(&bestA_s[a], sb);
      atomicMax(&bestB_s[b], sb);
    }
    if (alive_local) s_alive = 1;
    __syncthreads();
    if (!s_alive) break;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      if (cd[i].state != 0) continue;
      const uint32_t pr = cd[i].pair;
      const int a = pr / nB, b = pr - a * nB;
      const unsigned long long sb = static_cast<unsigned long long>(__double_as_longlong(cd[i].score));
      if (sb == bestA_s[a]) atomicMin(&bestA_i[a], pr);
      if (sb == bestB_s[b]) atomicMin(&bestB_i[b], pr);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      if (cd[i].state != 0) continue;
      const uint32_t pr = cd[i].pair;
      const int a = pr / nB, b = pr - a * nB;
      const unsigned long long sb = static_cast<unsigned long long>(__double_as_longlong(cd[i].score));
      if (sb == bestA_s[a] && pr == bestA_i[a] && sb == bestB_s[b] && pr == bestB_i[b]) {
        cd[i].state = 2;
        usedA[a] = 1;
        usedB[b] = 1;
        const int k = atomicAdd(&s_nacc, 1);
        if (k < kAssignMaxType) acc_slot[k] = i;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int nacc = min(s_nacc, kAssignMaxType);
  if (nacc > conn_cap && threadIdx.x == 0) atomicOr(&status[img], 8);
  // rank sort by (score desc, pair asc)
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    const Candidate ci = cd[acc_slot[i]];
    int rank = 0;
    for (int j = 0; j < nacc; ++j) {
      const Candidate cj = cd[acc_slot[j]];
      if (cj.score > ci.score || (cj.score == ci.score && cj.pair < ci.pair)) ++rank;
    }
    if (rank < conn_cap) {
      const int a = ci.pair / nB, b = ci.pair - a * nB;
      Connection c;
      c.score = ci.score;
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

// One warp per image.  subsets: [n_img][max_persons][20] float64 workspace.
// subsets_out (optional): final kept rows [n_img][max_persons][20].
__global__ void __launch_bounds__(32)
group_persons_kernel(const PeakD* __restrict__ peaks, const int* __restrict__ peak_counts, int peaks_cap,
                     PafConsts K, const Connection* __restrict__ conns, const int* __restrict__ conn_counts,
                     int conn_cap, double* __restrict__ subsets, int max_persons, const int* __restrict__ status_in,
                     ImageHeader* __restrict__ headers, PersonOut* __restrict__ persons,
                     double* __restrict__ subsets_out) {
  const int img = blockIdx.x, lane = threadIdx.x;
  const PeakD* pk = peaks + static_cast<size_t>(img) * peaks_cap;
  double* S = subsets + static_cast<size_t>(img) * max_persons * 20;
  int P = 0;
  int err = 0;
  int n_conn_total = 0;
  const int st_in = status_in[img];
  if (st_in) err = -4;  // OPB_ERR_CAPACITY from an earlier stage

  for (int l = 0; l < 19 && !err; ++l) {
    const int ja = K.limbs[l][0], jb = K.limbs[l][1];
    const Connection* cl = conns + (static_cast<size_t>(img) * 19 + l) * conn_cap;
    const int nc = conn_counts[img * 19 + l];
    n_conn_total += nc;
    for (int ci = 0; ci < nc && !err; ++ci) {
      const Connection c = cl[ci];
      const double da = static_cast<double>(c.id_a), db = static_cast<double>(c.id_b);
      // find the subsets that already hold one endpoint (first two, in row order)
      int found = 0, f0 = -1, f1 = -1;
      for (int base = 0; base < P; base += 32) {
        const int k = base + lane;
        const bool hit = (k < P) && (S[k * 20 + ja] == da || S[k * 20 + jb] == db);
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          const int bit = __ffs(mask) - 1;
          mask &= mask - 1;
          if (found == 0) f0 = base + bit;
          else if (found == 1) f1 = base + bit;
          ++found;
        }
      }
      if (found >= 3) { err = -5; break; }   // reference: IndexError at pose_detector.py:197
      if (found == 1) {
        if (lane == 0) {
          double* s = S + f0 * 20;
          if (s[jb] != db) {
            s[jb] = db;
            s[19] = __dadd_rn(s[19], 1.0);
            s[18] = __dadd_rn(s[18], __dadd_rn(static_cast<double>(pk[c.id_b].score), c.score));
          }
        }
      } else if (found == 2) {
        double* s1 = S + f0 * 20;
        double* s2 = S + f1 * 20;
        const bool both = (lane < 18) && (s1[lane] >= 0.0) && (s2[lane] >= 0.0);
        const bool overlap = __any_sync(0xffffffffu, both);
        if (!overlap) {
          if (lane < 18) s1[lane] = __dadd_rn(s1[lane], __dadd_rn(s2[lane], 1.0));
          if (lane >= 18 && lane < 20) s1[lane] = __dadd_rn(__dadd_rn(s1[lane], s2[lane]), c.score);  // :216-217
          __syncwarp();
          for (int r = f1; r < P - 1; ++r) {            // np.delete(subsets, f1, axis=0)
            double v = 0.0;
            if (lane < 20) v = S[(r + 1) * 20 + lane];
            __syncwarp();
            if (lane < 20) S[r * 20 + lane] = v;
            __syncwarp();
          }
          --P;
        } else if (lane == 0) {
          for (int w = 0; w < 2; ++w) {
            double* s = w ? s2 : s1;
            if (s[ja] == -1.0) {
              s[ja] = da;
              s[19] = __dadd_rn(s[19], 1.0);
              s[18] = __dadd_rn(s[18], __dadd_rn(static_cast<double>(pk[c.id_a].score), c.score));
            } else if (s[jb] == -1.0) {
              s[jb] = db;
              s[19] = __dadd_rn(s[19], 1.0);
              s[18] = __dadd_rn(s[18], __dadd_rn(static_cast<double>(pk[c.id_b].score), c.score));
            }
          }
        }
      } else if (found == 0 && l != 9 && l != 13) {
        if (P >= max_persons) { err = -4; break; }
        if (lane < 20) {
          double v = -1.0;
          if (lane == ja) v = da;
          if (lane == jb) v = db;
          if (lane == 19) v = 2.0;
          if (lane == 18)
            v = __dadd_rn(__dadd_rn(static_cast<double>(pk[c.id_a].score), static_cast<double>(pk[c.id_b].score)),
                          c.score);
          S[P * 20 + lane] = v;
        }
        ++P;
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // final filter (:248-249) and packing (:252-265)
  int kept = 0;
  if (!err) {
    for (int k = 0; k < P; ++k) {
      const double cnt = S[k * 20 + 19], sc = S[k * 20 + 18];
      const bool keep = (cnt >= K.n_subset_limbs_thresh) && (__ddiv_rn(sc, cnt) >= K.subset_score_thresh);
      if (keep) {
        if (subsets_out && lane < 20)
          subsets_out[(static_cast<size_t>(img) * max_persons + kept) * 20 + lane] = S[k * 20 + lane];
        if (persons) {
          PersonOut* po = persons + static_cast<size_t>(img) * max_persons + kept;
          if (lane < 18) {
            const int id = static_cast<int>(S[k * 20 + lane]);
            po->peak_id[lane] = id;
            po->x[lane] = (id >= 0) ? static_cast<int>(pk[id].x) : 0;
            po->y[lane] = (id >= 0) ? static_cast<int>(pk[id].y) : 0;
          }
          if (lane == 18) po->score = sc;
          if (lane == 19) po->count = cnt;
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
