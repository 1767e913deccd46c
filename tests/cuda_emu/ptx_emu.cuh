// ptx_emu.cuh -- TEST INFRASTRUCTURE ONLY.  Replaces csrc/ptx.cuh in the emulated build (build_emu.py): the same
// opb::ptx:: wrappers, implemented as a functional model of the sm_100a features the conv kernels use, so that the
// tcgen05 / TMA kernels themselves (descriptor arithmetic, swizzled layouts, pipeline phases, TMEM addressing,
// epilogues) run on a box without a GPU and can be checked against the oracle.
//
// Model (what the kernels are entitled to assume, executed adversarially where that is cheap):
//   mbarrier   64-bit word = {phase, expected arrivals, pending arrivals, pending transaction bytes}; a phase completes
//              when both pending counts reach zero; try_wait(parity) is true once the phase of that parity has
//              completed; a failing try_wait yields to the other threads of the block.
//   TMA        cuTensorMapEncodeTiled (emu_runtime.cpp) records the map; a load copies the box element by element
//              (out-of-bounds -> zero fill), writes it with the 128-byte swizzle (16-byte chunk index ^= bits 7..9 of
//              the shared address) and completes `box bytes` on the mbarrier -- DEFERRED until a thread next polls that
//              mbarrier, so a stage read without waiting on its full barrier still holds its previous contents.
//   tcgen05    TMEM = 128 lanes x 512 fp32 columns per CTA.  tcgen05.mma is QUEUED and executed only when the issuing
//              thread commits (tcgen05.commit) -- "as late as legal": a kernel that recycles an operand stage or
//              reads an accumulator before the corresponding commit sees stale data and fails its parity test.
//              Operands are fetched through the UMMA shared-memory descriptors (K-major, SWIZZLE_128B: start address,
//              stride between 8-row groups, address-bit swizzle), fp16 x fp16 products accumulated in fp32.
//              tcgen05.ld.32x32b checks that a warp only touches its own lane quadrant.
//   cluster    CTA pairs: both CTAs of a pair are resident together (cuda_emu.h); mapa / remote mbarrier arrive /
//              cta_group::2 TMA crediting the leader's barriers / M = 256 MMA over both CTAs' operands and TMEM /
//              multicast commit are modelled.  Larger clusters and TMA multicast are not.
// Accumulation order inside an MMA is unspecified on hardware; here it is k-ascending in fp32, so results agree with the
// GPU to rounding, not bit for bit -- the conv parity tests use the same tolerances on both.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "emu_tmap.h"

#ifdef EMU_TSAN
extern "C" {
void __tsan_read_range(void* addr, unsigned long size);
void __tsan_write_range(void* addr, unsigned long size);
}
// Race detection for the asynchronous agents: an mbarrier is a happens-before token (every arrive / expect_tx /
// complete_tx releases it, a successful wait acquires it); TMA writes, tcgen05.mma operand reads and TMEM accesses
// are reported to ThreadSanitizer as accesses of the fiber that caused them, at the time the model performs them.
#define EMU_HB_RELEASE(p) __tsan_release(p)
#define EMU_HB_ACQUIRE(p) __tsan_acquire(p)
#define EMU_RACE_READ(p, n) __tsan_read_range(const_cast<void*>(static_cast<const void*>(p)), n)
#define EMU_RACE_WRITE(p, n) __tsan_write_range(static_cast<void*>(p), n)
#else
#define EMU_HB_RELEASE(p)
#define EMU_HB_ACQUIRE(p)
#define EMU_RACE_READ(p, n)
#define EMU_RACE_WRITE(p, n)
#endif

namespace emu {

// Shared-window addresses (what cvta.to.shared yields on the GPU; UMMA descriptors keep 18 bits of them): a pointer
// into the launch's dynamic shared memory maps to its offset; static __shared__ arrays (host globals here) are mapped
// behind the dynamic window, host-contiguously and 1024-byte congruent (the 128-byte swizzle depends on address bits
// 7..9), centred on the first static pointer the launch converts -- all static operands of one launch must lie within the
// remaining window (checked).
EMU_INTERNAL inline uint32_t static_window_base() { return (static_cast<uint32_t>(g_dyn_smem_bytes) + 1023u) & ~1023u; }
EMU_INTERNAL inline uint32_t smem_addr_of(const void* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p), d = reinterpret_cast<uintptr_t>(g_dyn_smem);
  if (g_dyn_smem_bytes && a >= d && a < d + g_dyn_smem_bytes) return static_cast<uint32_t>(a - d);
  const uint32_t win0 = static_window_base(), span = 0x40000u - win0;
  if (!g_static_smem_hi) g_static_smem_hi = (a & ~static_cast<uintptr_t>(1023)) - ((span / 2) & ~1023u);
  if (a < g_static_smem_hi || a - g_static_smem_hi >= span) {
    fprintf(stderr, "emu: static __shared__ operands of one launch span more than the %u KB shared window left by its dynamic "
            "shared memory (emulation limit)\n", span >> 10);
    abort();
  }
  return win0 + static_cast<uint32_t>(a - g_static_smem_hi);
}
EMU_INTERNAL inline unsigned char* smem_ptr(uint32_t addr, int cta = -1) {   // cta < 0: the running thread's CTA
  if (addr < g_dyn_smem_bytes) return (cta < 0 ? g_dyn_smem : g_cta[cta].dyn) + addr;
  if (!g_static_smem_hi || addr < static_window_base()) { fprintf(stderr, "emu: shared-window address 0x%x outside the launch's shared memory\n", addr); abort(); }
  return reinterpret_cast<unsigned char*>(g_static_smem_hi + (addr - static_window_base()));
}

inline const TensorMapRec* tmap_rec(const CUtensorMap* m) {
  const TensorMapRec* r = *reinterpret_cast<TensorMapRec* const*>(m);
  if (!r || r->magic != TMAP_MAGIC) { fprintf(stderr, "emu: TMA with a tensor map that was not encoded\n"); abort(); }
  return r;
}

struct MBar {              // layout of the 64-bit mbarrier word in the emulation
  uint32_t phase : 1;
  uint32_t expected : 15;
  uint32_t pending : 16;
  int32_t tx;
};
static_assert(sizeof(MBar) == 8, "mbarrier word");

EMU_INTERNAL inline void mbar_check(MBar* b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1u; b->pending = b->expected; }
}
EMU_INTERNAL inline void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  MBar* b = reinterpret_cast<MBar*>(bar);
  if (b->pending < n) { fprintf(stderr, "emu: more arrivals than the mbarrier expects (block %u thread %d)\n", g_bid.x, g_cur); abort(); }
  b->pending -= n;
  EMU_HB_RELEASE(bar);
  mbar_check(b);
}
EMU_INTERNAL inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  MBar* b = reinterpret_cast<MBar*>(bar);
  b->tx -= static_cast<int32_t>(bytes);
  EMU_HB_RELEASE(bar);
  mbar_check(b);
}

struct QueuedMma { uint32_t tmem_d; uint64_t adesc, bdesc; uint32_t idesc; uint32_t accumulate; int pair; int f8 = 0; };
// MMAs issued and not yet committed, per CTA (one thread per CTA issues them; plain arrays touched only inside
// EMU_INTERNAL functions so that the racecheck build does not see the emulator's own bookkeeping)
constexpr int MAX_QUEUED_MMA = 4096;
inline QueuedMma g_mma_queue[MAX_CTAS][MAX_QUEUED_MMA];
inline int g_n_mma_queued[MAX_CTAS] = {0, 0}, g_mma_issuer[MAX_CTAS] = {-1, -1};
EMU_INTERNAL inline void mma_enqueue(const QueuedMma& q) {
  const int c = g_cur_cta;
  if (g_n_mma_queued[c] && g_mma_issuer[c] != g_cur) { fprintf(stderr, "emu: tcgen05.mma issued by two threads of one CTA between commits (threads %d and %d)\n", g_mma_issuer[c], g_cur); abort(); }
  if (g_n_mma_queued[c] == MAX_QUEUED_MMA) { fprintf(stderr, "emu: %d tcgen05.mma issued without a commit\n", MAX_QUEUED_MMA); abort(); }
  g_mma_issuer[c] = g_cur;
  g_mma_queue[c][g_n_mma_queued[c]++] = q;
}
inline float g_tmem[MAX_CTAS][128][512];
inline bool g_tmem_allocated[MAX_CTAS] = {false, false};
inline long long g_mma_count = 0, g_tma_count = 0;

EMU_INTERNAL inline uint32_t swz128(uint32_t addr) { return addr ^ (((addr >> 7) & 7u) << 4); }

EMU_INTERNAL inline float half_at(uint32_t addr, int cta) {
  __half h;
  EMU_RACE_READ(smem_ptr(swz128(addr), cta), 2);
  memcpy(&h, smem_ptr(swz128(addr), cta), 2);
  return __half2float(h);
}

// kind::f8f6f4 operand element: one byte, E4M3 (fmt 0) or E5M2 (fmt 1)
EMU_INTERNAL inline float f8_at(uint32_t addr, int cta, int fmt) {
  uint8_t b;
  EMU_RACE_READ(smem_ptr(swz128(addr), cta), 1);
  memcpy(&b, smem_ptr(swz128(addr), cta), 1);
  if (fmt == 1) {   // e5m2 = the high byte of the fp16 with the same value
    const __half_raw hr{static_cast<unsigned short>(static_cast<unsigned short>(b) << 8)};
    return __half2float(__half(hr));
  }
  if (fmt != 0) { fprintf(stderr, "emu: unsupported kind::f8f6f4 operand format %d\n", fmt); abort(); }
  const int e = (b >> 3) & 15, m = b & 7;
  float v = e ? ldexpf(1.f + m / 8.f, e - 7) : ldexpf(m / 8.f, -6);
  if (e == 15 && m == 7) v = NAN;
  return (b & 0x80) ? -v : v;
}

// cta_group::1: D[128 x N] in the issuing CTA's TMEM.  cta_group::2 (pair): M = 256 -- rows 0..127 are CTA 0's A operand
// and accumulate in CTA 0's TMEM, rows 128..255 CTA 1's; the B operand's rows 0..N/2-1 come from CTA 0's shared memory and
// rows N/2..N-1 from CTA 1's, each through the same descriptor.
__attribute__((optimize("O3"))) EMU_INTERNAL inline void execute_mma(const QueuedMma& q, int issuer_cta) {
  const int N = static_cast<int>((q.idesc >> 17) & 0x3f) << 3, M = static_cast<int>((q.idesc >> 24) & 0x1f) << 4;
  if (M != (q.pair ? 256 : 128) || N < 8 || N > 256 || (q.pair && (N & 15)) || ((q.idesc >> 15) & 3u) != 0 || ((q.idesc >> 4) & 3u) != 1) {
    fprintf(stderr, "emu: unsupported tcgen05.mma instruction descriptor 0x%x (M=%d N=%d cta_group::%d)\n", q.idesc, M, N, q.pair ? 2 : 1); abort();
  }
  if ((q.adesc >> 61) != 2 || (q.bdesc >> 61) != 2) { fprintf(stderr, "emu: only SWIZZLE_128B K-major operands are modelled\n"); abort(); }
  if (q.pair && (issuer_cta != 0 || g_ncta != 2)) { fprintf(stderr, "emu: cta_group::2 MMA must be issued by the leader CTA of a pair\n"); abort(); }
  const uint32_t a0 = static_cast<uint32_t>(q.adesc & 0x3FFF) << 4, b0 = static_cast<uint32_t>(q.bdesc & 0x3FFF) << 4;
  const uint32_t a_sbo = static_cast<uint32_t>((q.adesc >> 32) & 0x3FFF) << 4, b_sbo = static_cast<uint32_t>((q.bdesc >> 32) & 0x3FFF) << 4;
  const uint32_t dcol = q.tmem_d & 0xffff, dlane = q.tmem_d >> 16;
  if (dlane != 0 || dcol + N > 512) { fprintf(stderr, "emu: tcgen05.mma accumulator outside TMEM (lane %u col %u N %d)\n", dlane, dcol, N); abort(); }
  static float A[128][32], Bt[32][256];
  const int nhalf = q.pair ? N / 2 : N;
  const int KK = q.f8 ? 32 : 16;      // K per instruction: 32 bytes of either element size
  const int afmt = static_cast<int>((q.idesc >> 7) & 7u), bfmt = static_cast<int>((q.idesc >> 10) & 7u);
  if (!q.f8 && (afmt || bfmt)) { fprintf(stderr, "emu: kind::f16 is modelled for fp16 operands only\n"); abort(); }
  for (int n = 0; n < N; ++n) {
    const int src_cta = q.pair ? n / nhalf : issuer_cta, row = q.pair ? n % nhalf : n;
    const uint32_t rb = b0 + (row >> 3) * b_sbo + (row & 7) * 128;
    for (int k = 0; k < KK; ++k) Bt[k][n] = q.f8 ? f8_at(rb + k, src_cta, bfmt) : half_at(rb + 2 * k, src_cta);
  }
  for (int half = 0; half < (q.pair ? 2 : 1); ++half) {
    const int cta = q.pair ? half : issuer_cta;
    for (int m = 0; m < 128; ++m) {
      const uint32_t ra = a0 + (m >> 3) * a_sbo + (m & 7) * 128;
      for (int k = 0; k < KK; ++k) A[m][k] = q.f8 ? f8_at(ra + k, cta, afmt) : half_at(ra + 2 * k, cta);
    }
    for (int m = 0; m < 128; ++m) {    // per element: acc = (((d + a0*b0) + a1*b1) + ...), k ascending; vectorises over n
      float* d = &g_tmem[cta][m][dcol];
      EMU_RACE_WRITE(d, 4ul * N);
      if (!q.accumulate) for (int n = 0; n < N; ++n) d[n] = 0.f;
      for (int k = 0; k < KK; ++k) {
        const float a = A[m][k];
        const float* b = Bt[k];
        for (int n = 0; n < N; ++n) d[n] += a * b[n];
      }
    }
  }
  ++g_mma_count;
}

EMU_INTERNAL inline void mma_flush() {   // the committing thread's queued MMAs run now
  const int c = g_cur_cta;
  if (g_n_mma_queued[c] && g_mma_issuer[c] != g_cur) { fprintf(stderr, "emu: tcgen05.commit by thread %d while thread %d has uncommitted MMAs\n", g_cur, g_mma_issuer[c]); abort(); }
  for (int i = 0; i < g_n_mma_queued[c]; ++i) execute_mma(g_mma_queue[c][i], c);
  g_n_mma_queued[c] = 0;
}

// shared::cluster addresses (mapa): bits 28.. = CTA rank + 1, low bits = the shared-window address inside that CTA
EMU_INTERNAL inline uint32_t cluster_addr(uint32_t local, uint32_t rank) { return ((rank + 1u) << 28) | local; }
EMU_INTERNAL inline uint64_t* cluster_bar(uint32_t caddr) {
  const int rank = static_cast<int>(caddr >> 28) - 1;
  if (rank < 0 || rank >= g_ncta) { fprintf(stderr, "emu: bad shared::cluster address 0x%x\n", caddr); abort(); }
  return reinterpret_cast<uint64_t*>(smem_ptr(caddr & 0x0FFFFFFFu, rank));
}
EMU_INTERNAL inline void need_cluster() {   // a cluster primitive in a single-CTA attempt: ask the launcher for CTA pairs
  if (g_ncta > 1) return;
  g_cluster_request = 2;
  Fiber& f = g_fibers[g_cur];
  f.state = DONE;
#ifdef EMU_TSAN
  __tsan_switch_to_fiber(g_sched_tsan, 1);
#endif
#ifdef EMU_FAST_SWITCH
  emu_switch(&f.sp, g_sched_sp);
#else
  swapcontext(&f.ctx, &g_sched);
#endif
  abort();   // never resumed
}

EMU_INTERNAL inline void yield_ready() {   // spin-wait: let the other threads of the block run
  Fiber& f = g_fibers[g_cur];
  f.state = READY;
#ifdef EMU_TSAN
  __tsan_switch_to_fiber(g_sched_tsan, 1);
#endif
#ifdef EMU_FAST_SWITCH
  emu_switch(&f.sp, g_sched_sp);
#else
  swapcontext(&f.ctx, &g_sched);
#endif
}

EMU_INTERNAL inline void tma_execute(void* smem_dst, const TensorMapRec* r, uint64_t* bar, const int* c, int rank) {
  if (static_cast<int>(r->rank) != rank) { fprintf(stderr, "emu: %dD TMA load through a %uD tensor map\n", rank, r->rank); abort(); }
  if (r->swizzle != 3 /* CU_TENSOR_MAP_SWIZZLE_128B */ || r->box[0] * r->elem_bytes != 128) {
    fprintf(stderr, "emu: only SWIZZLE_128B boxes with 128-byte inner rows are modelled (swizzle %u, inner %u B)\n", r->swizzle, r->box[0] * r->elem_bytes);
    abort();
  }
  const uint32_t dst = smem_addr_of(smem_dst);
  uint32_t rows = 1;
  for (int d = 1; d < rank; ++d) rows *= r->box[d];
  const uint32_t eb = r->elem_bytes;
  for (uint32_t row = 0; row < rows; ++row) {
    // coordinates of this row in dims 1..rank-1
    uint32_t rem = row;
    long long off = 0;
    bool oob_row = false;
    for (int d = 1; d < rank; ++d) {
      const uint32_t bi = rem % r->box[d];
      rem /= r->box[d];
      const long long g = static_cast<long long>(c[d]) + static_cast<long long>(bi) * r->estr[d];
      if (g < 0 || g >= static_cast<long long>(r->dims[d])) oob_row = true;
      off += g * static_cast<long long>(r->strides[d]);
    }
    for (uint32_t e = 0; e < r->box[0]; ++e) {
      const long long g0 = static_cast<long long>(c[0]) + static_cast<long long>(e) * r->estr[0];
      unsigned char* d8 = smem_ptr(swz128(dst + row * 128 + e * eb));
      EMU_RACE_WRITE(d8, eb);
      if (oob_row || g0 < 0 || g0 >= static_cast<long long>(r->dims[0])) memset(d8, 0, eb);
      else memcpy(d8, static_cast<const unsigned char*>(r->base) + off + g0 * static_cast<long long>(eb), eb);
    }
  }
  ++g_tma_count;
  mbar_complete_tx(bar, rows * 128);
}

// TMA loads are DEFERRED: the copy and its complete_tx happen only when some thread next polls the mbarrier the load
// signals (or at tcgen05.dealloc / kernel end) -- "as late as legal", like the queued MMAs: a consumer that reads an
// operand stage without waiting on its full barrier sees the stage's previous contents and fails its parity test.
struct PendingTma { uint64_t* bar; void* dst; const TensorMapRec* rec; int c[5]; int rank; int cta; };
// (a plain array, touched only inside EMU_INTERNAL functions: the racecheck build must not see the emulator's own
// bookkeeping, and std::vector's member functions would be instrumented)
constexpr int MAX_PENDING_TMA = 1024;
inline PendingTma g_tma_pending[MAX_PENDING_TMA];
inline int g_n_tma_pending = 0;
EMU_INTERNAL inline void tma_load(void* smem_dst, const CUtensorMap* m, uint64_t* bar, const int* c, int rank) {
  if (g_n_tma_pending == MAX_PENDING_TMA) { fprintf(stderr, "emu: %d TMA loads in flight whose mbarriers nobody polls\n", MAX_PENDING_TMA); abort(); }
  PendingTma& t = g_tma_pending[g_n_tma_pending++];
  t.bar = bar; t.dst = smem_dst; t.rec = tmap_rec(m); t.rank = rank; t.cta = g_cur_cta;
  for (int d = 0; d < 5; ++d) t.c[d] = d < rank ? c[d] : 0;
}
// TMA store (executed at once: the kernels make their shared-memory writes visible and synchronise before issuing it):
// de-swizzle the box and write the in-bounds part to global memory
EMU_INTERNAL inline void tma_store(const CUtensorMap* m, const void* smem_src, const int* c, int rank) {
  const TensorMapRec* r = tmap_rec(m);
  if (static_cast<int>(r->rank) != rank || r->swizzle != 3 || r->box[0] * r->elem_bytes != 128) {
    fprintf(stderr, "emu: unsupported TMA store (rank %u, swizzle %u, inner %u B)\n", r->rank, r->swizzle, r->box[0] * r->elem_bytes); abort();
  }
  const uint32_t src = smem_addr_of(smem_src);
  uint32_t rows = 1;
  for (int d = 1; d < rank; ++d) rows *= r->box[d];
  const uint32_t eb = r->elem_bytes;
  for (uint32_t row = 0; row < rows; ++row) {
    uint32_t rem = row;
    long long off = 0;
    bool oob_row = false;
    for (int d = 1; d < rank; ++d) {
      const uint32_t bi = rem % r->box[d];
      rem /= r->box[d];
      const long long g = static_cast<long long>(c[d]) + static_cast<long long>(bi) * r->estr[d];
      if (g < 0 || g >= static_cast<long long>(r->dims[d])) oob_row = true;
      off += g * static_cast<long long>(r->strides[d]);
    }
    if (oob_row) continue;
    for (uint32_t e = 0; e < r->box[0]; ++e) {
      const long long g0 = static_cast<long long>(c[0]) + static_cast<long long>(e) * r->estr[0];
      if (g0 < 0 || g0 >= static_cast<long long>(r->dims[0])) continue;
      const unsigned char* s8 = smem_ptr(swz128(src + row * 128 + e * eb));
      EMU_RACE_READ(s8, eb);
      memcpy(static_cast<unsigned char*>(const_cast<void*>(r->base)) + off + g0 * static_cast<long long>(eb), s8, eb);
    }
  }
  ++g_tma_count;
}

// multicast: one pending copy per destination CTA (same offsets of destination and barrier in each)
EMU_INTERNAL inline void tma_load_to(int cta, void* smem_dst, const CUtensorMap* m, uint64_t* bar, const int* c, int rank) {
  if (g_n_tma_pending == MAX_PENDING_TMA) { fprintf(stderr, "emu: %d TMA loads in flight whose mbarriers nobody polls\n", MAX_PENDING_TMA); abort(); }
  PendingTma& t = g_tma_pending[g_n_tma_pending++];
  t.bar = bar; t.dst = smem_dst; t.rec = tmap_rec(m); t.rank = rank; t.cta = cta;
  for (int d = 0; d < 5; ++d) t.c[d] = d < rank ? c[d] : 0;
}
EMU_INTERNAL inline void tma_flush(uint64_t* bar) {   // bar == nullptr: everything
  if (g_n_tma_pending == 0) return;
  unsigned char* const saved_dyn = g_dyn_smem;
  const int saved_cta = g_cur_cta;
  int kept = 0;
  const int n = g_n_tma_pending;
  for (int i = 0; i < n; ++i) {
    const PendingTma t = g_tma_pending[i];
    if (bar != nullptr && t.bar != bar) { g_tma_pending[kept++] = t; continue; }
    g_cur_cta = t.cta; g_dyn_smem = g_cta[t.cta].dyn;     // the destination is in the ISSUING CTA's shared memory
    tma_execute(t.dst, t.rec, t.bar, t.c, t.rank);
  }
  g_n_tma_pending = kept;
  g_cur_cta = saved_cta; g_dyn_smem = saved_dyn;
}

}  // namespace emu

namespace opb {
namespace ptx {

EMU_INTERNAL inline uint32_t smem_u32(const void* p) { return emu::smem_addr_of(p); }

// ---------------------------------------------------------------- mbarrier
EMU_INTERNAL inline void mbar_init(uint64_t* bar, uint32_t count) {
  emu::MBar b;
  b.phase = 0; b.expected = count; b.pending = count; b.tx = 0;
  memcpy(bar, &b, 8);
}
inline void fence_barrier_init() {}
inline void fence_proxy_async_smem() {}
EMU_INTERNAL inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  reinterpret_cast<emu::MBar*>(bar)->tx += static_cast<int32_t>(bytes);
  emu::mbar_arrive_n(bar, 1);
}
EMU_INTERNAL inline void mbar_arrive(uint64_t* bar) { emu::mbar_arrive_n(bar, 1); }
EMU_INTERNAL inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  emu::tma_flush(bar);
  if (reinterpret_cast<emu::MBar*>(bar)->phase == (parity & 1u)) emu::yield_ready();
  const bool done = reinterpret_cast<emu::MBar*>(bar)->phase != (parity & 1u);
  if (done) EMU_HB_ACQUIRE(bar);
  return done;
}
EMU_INTERNAL inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  long long spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 2000000) {   // every other thread of the block had two million turns: a protocol deadlock
      fprintf(stderr, "emu: mbarrier wait never satisfied: block %u thread %d bar +%u parity %u\n", emu::g_bid.x, emu::g_cur,
              smem_u32(bar), parity);
      abort();
    }
  }
}

// ---------------------------------------------------------------- TMA
inline void prefetch_tensormap(const CUtensorMap*) {}
EMU_INTERNAL inline void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const int c[2] = {c0, c1};
  emu::tma_load(smem_dst, m, bar, c, 2);
}
EMU_INTERNAL inline void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  const int c[4] = {c0, c1, c2, c3};
  emu::tma_store(m, smem_src, c, 4);
}
inline void tma_store_commit() {}
inline void tma_store_wait_read() {}
inline void tma_store_wait_all() {}
EMU_INTERNAL inline void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const int c[4] = {c0, c1, c2, c3};
  emu::tma_load(smem_dst, m, bar, c, 4);
}

// ---------------------------------------------------------------- tcgen05
template <int kCols>
EMU_INTERNAL inline void tmem_alloc(uint32_t* smem_result) {
  if (emu::lane_id() == 0) {   // .sync.aligned: one allocation per warp-wide call
    if (emu::g_tmem_allocated[emu::g_cur_cta]) { fprintf(stderr, "emu: second tcgen05.alloc in one CTA\n"); abort(); }
    emu::g_tmem_allocated[emu::g_cur_cta] = true;
    for (auto& row : emu::g_tmem[emu::g_cur_cta]) for (float& v : row) v = __builtin_nanf("");   // fresh TMEM holds garbage
    *smem_result = 0;
  }
}
template <int kCols>
EMU_INTERNAL inline void tmem_dealloc(uint32_t) {
  if (emu::lane_id() == 0) {
    emu::tma_flush(nullptr);
    if (emu::g_n_mma_queued[emu::g_cur_cta]) { fprintf(stderr, "emu: %d tcgen05.mma issued by thread %d were never committed\n", emu::g_n_mma_queued[emu::g_cur_cta], emu::g_mma_issuer[emu::g_cur_cta]); abort(); }
    emu::g_tmem_allocated[emu::g_cur_cta] = false;
  }
}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
EMU_INTERNAL inline void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  emu::mma_enqueue(emu::QueuedMma{tmem_d, adesc, bdesc, idesc, accumulate, 0});
}
EMU_INTERNAL inline void mma_f16_ss_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  emu::mma_enqueue(emu::QueuedMma{tmem_d, adesc, bdesc, idesc, 1u, 0});
}
EMU_INTERNAL inline void mma_f8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  emu::mma_enqueue(emu::QueuedMma{tmem_d, adesc, bdesc, idesc, accumulate, 0, 1});
}
EMU_INTERNAL inline void mma_commit(uint64_t* bar) {
  emu::mma_flush();
  emu::mbar_arrive_n(bar, 1);
}
template <int NCOL>
EMU_INTERNAL inline void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
  const uint32_t lane0 = taddr >> 16, col = taddr & 0xffff;
  if (lane0 != 32u * (static_cast<uint32_t>(emu::warp_id()) & 3u) || col + NCOL > 512) {
    fprintf(stderr, "emu: tcgen05.ld outside the warp's lane quadrant / TMEM (warp %d lane base %u col %u)\n", emu::warp_id(), lane0, col);
    abort();
  }
  const float* src = &emu::g_tmem[emu::g_cur_cta][lane0 + static_cast<uint32_t>(emu::lane_id())][col];
  EMU_RACE_READ(src, 4ul * NCOL);
  memcpy(r, src, 4 * NCOL);
}
EMU_INTERNAL inline void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_cols<32>(taddr, r); }
EMU_INTERNAL inline void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_cols<16>(taddr, r); }
inline void tmem_ld_wait() {}

// ---------------------------------------------------------------- cluster multicast (cta_group::1 kernels)
EMU_INTERNAL inline void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  emu::need_cluster();
  const int c[2] = {c0, c1};
  const uint32_t off_d = emu::smem_addr_of(smem_dst), off_b = emu::smem_addr_of(bar);
  for (int k = 0; k < emu::g_ncta; ++k)
    if ((cta_mask >> k) & 1)
      emu::tma_load_to(k, emu::smem_ptr(off_d, k), m, reinterpret_cast<uint64_t*>(emu::smem_ptr(off_b, k)), c, 2);
}
EMU_INTERNAL inline void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  emu::need_cluster();
  emu::mma_flush();
  const uint32_t off = emu::smem_addr_of(bar);
  for (int k = 0; k < emu::g_ncta; ++k)
    if ((cta_mask >> k) & 1) emu::mbar_arrive_n(reinterpret_cast<uint64_t*>(emu::smem_ptr(off, k)), 1);
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2): both CTAs resident
// elect.sync: one lane of the (converged) warp -- the emulator elects lane 0
EMU_INTERNAL inline bool elect_one() { return (threadIdx.x & 31u) == 0u; }
EMU_INTERNAL inline uint32_t cluster_ctarank() { emu::need_cluster(); return static_cast<uint32_t>(emu::g_cur_cta); }
EMU_INTERNAL inline void cluster_sync_all() { emu::need_cluster(); emu::yield_wait(emu::WAIT_CLUSTER); }
EMU_INTERNAL inline uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) { emu::need_cluster(); return emu::cluster_addr(local_addr, rank); }
EMU_INTERNAL inline void mbar_arrive_cluster(uint32_t cluster_addr) { emu::mbar_arrive_n(emu::cluster_bar(cluster_addr), 1); }
// data lands in the ISSUING CTA's shared memory, the bytes are credited to the barrier at the cluster address
EMU_INTERNAL inline void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1) {
  const int c[2] = {c0, c1};
  emu::tma_load(smem_dst, m, emu::cluster_bar(mbar_cluster_addr), c, 2);
}
EMU_INTERNAL inline void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0, int c1, int c2,
                                          int c3) {
  const int c[4] = {c0, c1, c2, c3};
  emu::tma_load(smem_dst, m, emu::cluster_bar(mbar_cluster_addr), c, 4);
}
template <int kCols>
EMU_INTERNAL inline void tmem_alloc_pair(uint32_t* smem_result) { emu::need_cluster(); tmem_alloc<kCols>(smem_result); }
template <int kCols>
EMU_INTERNAL inline void tmem_dealloc_pair(uint32_t taddr) { tmem_dealloc<kCols>(taddr); }
EMU_INTERNAL inline void mma_f16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  emu::mma_enqueue(emu::QueuedMma{tmem_d, adesc, bdesc, idesc, accumulate, 1});
}
EMU_INTERNAL inline void mma_f16_ss_pair_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  emu::mma_enqueue(emu::QueuedMma{tmem_d, adesc, bdesc, idesc, 1u, 1});
}
// multicast commit: arrive (count 1) on the barrier at this offset in BOTH CTAs once the queued MMAs have run
EMU_INTERNAL inline void mma_f8_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  emu::mma_enqueue(emu::QueuedMma{tmem_d, adesc, bdesc, idesc, accumulate, 1, 1});
}
EMU_INTERNAL inline void mma_commit_pair(uint64_t* bar) {
  emu::mma_flush();
  const uint32_t off = emu::smem_addr_of(bar);
  for (int c = 0; c < emu::g_ncta; ++c) emu::mbar_arrive_n(reinterpret_cast<uint64_t*>(emu::smem_ptr(off, c)), 1);
}

// descriptor encodings: the kernels' own (pasted from csrc/ptx.cuh by build_emu.py)
// @@DESCRIPTORS@@

}  // namespace ptx
}  // namespace opb

// counters for the tests: [0] kernel launches, [1] of them as CTA pairs, [2] tcgen05.mma executed, [3] TMA loads
extern "C" void opb_emu_stats(long long* out) {
  out[0] = emu::g_launches; out[1] = emu::g_cluster_launches; out[2] = emu::g_mma_count; out[3] = emu::g_tma_count;
}
