// cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never loaded by the package).
//
// A single-OS-thread emulation of the CUDA execution model, large enough to run the library's *unmodified* kernel
// sources (post-process, resize, keypoints, and -- with ptx_emu.cuh -- the tcgen05 / TMA convolution kernels) on a box
// without a GPU, so that `pytest -m "not gpu"` can check device code against the oracle:
//
//   * every CUDA thread is a fiber (hand-written x86-64 context switch, ucontext elsewhere); the CTAs of a grid run one
//     after another, the two CTAs of a CTA pair together;
//   * __syncthreads* / __syncwarp / shuffles / votes / the cluster barrier are cooperative yields to a scheduler that
//     releases a barrier when all live participants wait on it, checks that the lanes of a warp wait in the SAME
//     primitive (code that only works under lockstep execution aborts), and detects deadlocks;
//   * __shared__ variables become function statics (one CTA per cluster slot is resident at a time), the dynamic shared
//     memory window of a launch is an exact-size heap block;
//   * atomics are relaxed atomics, __ldg/__stcs plain accesses;
//   * the _rn arithmetic intrinsics are single IEEE operations fenced against contraction, so the emulated build may be
//     compiled with -ffp-contract=fast to expose arithmetic that would depend on nvcc's FMA contraction;
//   * OPB_EMU_SANITIZE=address|thread (build_emu.py) turn AddressSanitizer / ThreadSanitizer into memcheck / racecheck
//     analogues: under TSan every fiber is a TSan fiber and barriers are the only happens-before edges.
//
// Environment: OPB_EMU_ORDER=reverse (highest thread first), OPB_EMU_SMS=n (emulated SM count), OPB_EMU_BACKTRACE=1.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#undef __shared__
#define __shared__ static
#undef __grid_constant__
#define __grid_constant__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __cluster_dims__
#define __cluster_dims__(...)

// Race detection (OPB_EMU_SANITIZE=thread): every fiber is a ThreadSanitizer fiber without implicit
// synchronisation at switches; barriers are the only happens-before edges inside a block, so two CUDA threads
// touching the same shared / global location without a barrier (or atomic) in between are reported as a data race --
// a racecheck analogue that also flags warp-synchronous code relying on lockstep execution.
#if defined(__SANITIZE_THREAD__)
#define EMU_TSAN 1
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
}
#endif

#ifdef EMU_TSAN
#define EMU_INTERNAL __attribute__((no_sanitize("thread"), noinline))   // emulator bookkeeping is not kernel code
#else
#define EMU_INTERNAL
#endif

namespace emu {

enum { READY = 0, WAIT_BLOCK = 1, WAIT_WARP = 2, DONE = 3, WAIT_CLUSTER = 4 };
constexpr int MAX_CTAS = 2;               // CTAs of one thread-block cluster resident together (CTA pairs)
constexpr int MAX_THREADS = 1024 * MAX_CTAS;
constexpr size_t STACK_BYTES = 256 * 1024;

struct Fiber {
  ucontext_t ctx;      // portable fallback
  void* sp = nullptr;  // x86-64 fast path: saved stack pointer (emu_switch in emu_runtime.cpp)
  void* tsan = nullptr;  // ThreadSanitizer fiber handle (race-detection build)
  char* stack = nullptr;
  int state = DONE;
  unsigned mask = 0;
  int pred = 0;
  int cta = 0;       // CTA of the cluster this thread belongs to
  int site = 0;      // which warp-level primitive (and which half of it) the fiber waits in
  long long nwaits = 0;
  long long nblock = 0;   // block-level barriers passed
  long long ncluster = 0; // cluster-level barriers passed
  uint3 tid;
};

inline Fiber g_fibers[MAX_THREADS];
inline int g_nfib = 0, g_cur = 0;
inline ucontext_t g_sched;
inline void* g_sched_sp = nullptr;
inline void* g_sched_tsan = nullptr;
// happens-before tokens: one per barrier GENERATION (a fiber that is resumed late must not pick up what faster
// fibers released when they arrived at the NEXT barrier); participants are at most one generation apart
inline char g_hb_block[MAX_CTAS][4], g_hb_warp[MAX_THREADS / 32][4], g_hb_cluster[4], g_hb_done;
inline const std::function<void()>* g_body = nullptr;
inline int g_red_or[MAX_CTAS], g_red_and[MAX_CTAS], g_red_cnt[MAX_CTAS];
struct CtaCtx { uint3 bid; unsigned char* dyn = nullptr; };
inline CtaCtx g_cta[MAX_CTAS];
inline int g_ncta = 1, g_cur_cta = 0, g_cta_threads = 0;
inline long long g_cluster_launches = 0;
inline int g_cluster_request = 0;   // set by a cluster primitive reached in a non-cluster launch: rerun as CTA pairs
inline uint64_t g_warp_buf[MAX_THREADS / 32][32];
inline int g_warp_pred[MAX_THREADS / 32][32];
inline unsigned char* g_dyn_smem = nullptr;   // exactly the launch's dynamic shared memory (a heap block: ASan sees overruns)
constexpr size_t MAX_DYN_SMEM = 232448;
inline size_t g_dyn_smem_bytes = 0;
inline uintptr_t g_static_smem_hi = 0;   // host address that maps to the first shared-window address behind the dynamic window (static operands)
inline uint3 g_tid, g_bid;
inline dim3 g_bdim, g_gdim;
inline long long g_launches = 0;
inline bool g_reverse = false;

#if defined(__x86_64__)
#define EMU_FAST_SWITCH 1
// saves the callee-saved registers and the stack pointer of the running fiber, resumes `to`
// (no signal-mask system call, unlike swapcontext: ~20x cheaper per synchronisation point)
extern "C" void emu_switch(void** save_sp, void* to_sp);
#endif

EMU_INTERNAL inline void fiber_entry() {
  (*g_body)();
  g_fibers[g_cur].state = DONE;   // uc_link returns to the scheduler
}

EMU_INTERNAL inline void yield_wait(int state) {
  Fiber& f = g_fibers[g_cur];
  f.state = state;
#ifdef EMU_TSAN
  void* hb = (state == WAIT_BLOCK)     ? static_cast<void*>(&g_hb_block[f.cta][f.nblock++ & 3])
             : (state == WAIT_CLUSTER) ? static_cast<void*>(&g_hb_cluster[f.ncluster++ & 3])
                                       : static_cast<void*>(&g_hb_warp[g_cur >> 5][f.nwaits & 3]);   // (a warp never spans CTAs: threads per CTA % 32 == 0 in cluster launches)
  __tsan_release(hb);                       // everything before the barrier ...
  __tsan_switch_to_fiber(g_sched_tsan, 1);  // (1 = no implicit synchronisation at the switch)
#endif
#ifdef EMU_FAST_SWITCH
  emu_switch(&f.sp, g_sched_sp);
#else
  swapcontext(&f.ctx, &g_sched);
#endif
#ifdef EMU_TSAN
  __tsan_acquire(hb);                       // ... happens before everything after it, for all participants
#endif
}

#ifdef EMU_FAST_SWITCH
EMU_INTERNAL inline void fiber_trampoline() {
  fiber_entry();
#ifdef EMU_TSAN
  __tsan_release(&g_hb_done);
  __tsan_switch_to_fiber(g_sched_tsan, 1);
#endif
  emu_switch(&g_fibers[g_cur].sp, g_sched_sp);   // never resumed
  abort();
}
#endif

// Runs the g_ncta CTAs of one cluster (1, or 2 for CTA pairs) to completion: nthreads threads each.  Returns false
// when a thread reached a cluster primitive in a non-cluster launch (the launch is then repeated as CTA pairs).
EMU_INTERNAL inline bool run_block(int nthreads, const std::function<void()>& body) {
  const int total = nthreads * g_ncta;
  if (total > MAX_THREADS || (g_ncta > 1 && (nthreads & 31))) { fprintf(stderr, "emu: %d CTA(s) of %d threads\n", g_ncta, nthreads); abort(); }
  g_nfib = total;
  g_cta_threads = nthreads;
  g_body = &body;
#ifdef EMU_TSAN
  g_sched_tsan = __tsan_get_current_fiber();
#endif
  for (int i = 0; i < total; ++i) {
    Fiber& f = g_fibers[i];
    if (!f.stack) f.stack = static_cast<char*>(malloc(STACK_BYTES));
#ifdef EMU_FAST_SWITCH
    {  // initial frame: six callee-saved registers, then the address emu_switch's `ret` jumps to
      uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + STACK_BYTES) & ~static_cast<uintptr_t>(15);
      void** sp = reinterpret_cast<void**>(top);
      *--sp = nullptr;                                            // fake return address of the trampoline
      *--sp = reinterpret_cast<void*>(&fiber_trampoline);
      for (int r = 0; r < 6; ++r) *--sp = nullptr;
      f.sp = sp;
    }
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK_BYTES;
    f.ctx.uc_link = &g_sched;
    makecontext(&f.ctx, fiber_entry, 0);
#endif
    f.state = READY;
    f.nwaits = 0;
    f.nblock = 0;
    f.ncluster = 0;
    f.cta = i / nthreads;
#ifdef EMU_TSAN
    f.tsan = __tsan_create_fiber(0);
#endif
    const int t = i % nthreads;
    f.tid.x = t % g_bdim.x;
    f.tid.y = (t / g_bdim.x) % g_bdim.y;
    f.tid.z = t / (g_bdim.x * g_bdim.y);
  }
  for (;;) {
    bool ran = false;
    for (int k = 0; k < total; ++k) {
      // OPB_EMU_ORDER=reverse: highest thread first.  Code that is correct under independent thread scheduling
      // gives the same result for any order; lane-0-first and lane-31-first bracket the usual hazards.
      const int i = g_reverse ? total - 1 - k : k;
      if (g_fibers[i].state != READY) continue;
      g_cur = i;
      g_tid = g_fibers[i].tid;
      g_cur_cta = g_fibers[i].cta;
      g_bid = g_cta[g_cur_cta].bid;
      g_dyn_smem = g_cta[g_cur_cta].dyn;
#ifdef EMU_TSAN
      __tsan_switch_to_fiber(g_fibers[i].tsan, 1);
#endif
#ifdef EMU_FAST_SWITCH
      emu_switch(&g_sched_sp, g_fibers[i].sp);
#else
      swapcontext(&g_sched, &g_fibers[i].ctx);
#endif
      ran = true;
      if (g_cluster_request && g_ncta == 1) {   // abandon this attempt (nothing but shared memory was touched yet)
#ifdef EMU_TSAN
        for (int q = 0; q < total; ++q) { __tsan_destroy_fiber(g_fibers[q].tsan); g_fibers[q].tsan = nullptr; }
#endif
        return false;
      }
    }
    // warp-level barriers: release a group when every live lane of its mask waits
    bool released = false;
    for (int base = 0; base < total; base += 32) {
      const int wend = std::min(base + 32, (base / nthreads + 1) * nthreads);   // a warp never spans two CTAs
      for (int l = 0; base + l < wend; ++l) {
        Fiber& f = g_fibers[base + l];
        if (f.state != WAIT_WARP) continue;
        bool all = true;
        for (int j = 0; base + j < wend; ++j) {
          if (!((f.mask >> j) & 1u)) continue;
          const int st = g_fibers[base + j].state;
          if (st == DONE) continue;
          if (st != WAIT_WARP || g_fibers[base + j].mask != f.mask) { all = false; break; }
        }
        if (!all) continue;
        for (int j = 0; base + j < wend; ++j) {   // convergence check: same primitive, same count
          const Fiber& o = g_fibers[base + j];
          if (!((f.mask >> j) & 1u) || o.state != WAIT_WARP) continue;
          if (o.site != f.site || o.nwaits != f.nwaits) {
            fprintf(stderr, "emu: divergent warp-synchronous code in block (%u,%u,%u): lane %d waits in primitive %d (#%lld), "
                    "lane %d in primitive %d (#%lld)\n", g_bid.x, g_bid.y, g_bid.z, l, f.site, f.nwaits, j, o.site, o.nwaits);
            abort();
          }
        }
        const unsigned m = f.mask;
        for (int j = 0; base + j < wend; ++j)
          if (((m >> j) & 1u) && g_fibers[base + j].state == WAIT_WARP) g_fibers[base + j].state = READY;
        released = true;
      }
    }
    if (released) continue;
    // block-level barriers, per CTA; the cluster barrier over all CTAs
    int live_all = 0, at_cluster = 0, at_block_all = 0;
    for (int c = 0; c < g_ncta; ++c) {
      int live = 0, waiting = 0, r_or = 0, r_and = 1, r_cnt = 0;
      for (int i = c * nthreads; i < (c + 1) * nthreads; ++i) {
        const Fiber& f = g_fibers[i];
        if (f.state == DONE) continue;
        ++live;
        if (f.state == WAIT_CLUSTER) ++at_cluster;
        if (f.state == WAIT_BLOCK) { ++waiting; r_or |= (f.pred != 0); r_and &= (f.pred != 0); r_cnt += (f.pred != 0); }
      }
      live_all += live;
      at_block_all += waiting;
      if (live && waiting == live) {
        g_red_or[c] = r_or; g_red_and[c] = r_and; g_red_cnt[c] = r_cnt;
        for (int i = c * nthreads; i < (c + 1) * nthreads; ++i) if (g_fibers[i].state == WAIT_BLOCK) g_fibers[i].state = READY;
        released = true;
      }
    }
    if (live_all == 0) {
#ifdef EMU_TSAN
      __tsan_acquire(&g_hb_done);             // the block's writes are visible to later blocks / the host
      for (int i = 0; i < total; ++i) { __tsan_destroy_fiber(g_fibers[i].tsan); g_fibers[i].tsan = nullptr; }
#endif
      return true;
    }
    if (released) continue;
    if (at_cluster == live_all) {
      for (int i = 0; i < total; ++i) if (g_fibers[i].state == WAIT_CLUSTER) g_fibers[i].state = READY;
      continue;
    }
    if (!ran) {
      fprintf(stderr, "emu: deadlock in block (%u,%u,%u): %d live threads, %d at __syncthreads, %d at the cluster barrier, rest at "
              "divergent warp barriers\n", g_bid.x, g_bid.y, g_bid.z, live_all, at_block_all, at_cluster);
      abort();
    }
  }
}

struct Launcher {
  dim3 g, b;
  size_t smem;
  Launcher(dim3 grid, dim3 block, size_t smem_bytes = 0, cudaStream_t = nullptr) : g(grid), b(block), smem(smem_bytes) {}
  template <class... KA, class... A>
  void run(void (*k)(KA...), A&&... a) {
    if (smem > MAX_DYN_SMEM) { fprintf(stderr, "emu: %zu B of dynamic shared memory\n", smem); abort(); }
    std::tuple<std::decay_t<KA>...> args(std::forward<A>(a)...);
    const std::function<void()> body = [&] { std::apply(k, args); };
    g_bdim = b; g_gdim = g;
    ++g_launches;
    { const char* o = getenv("OPB_EMU_ORDER"); g_reverse = o && o[0] == 'r'; }
    const int nthreads = static_cast<int>(b.x * b.y * b.z);
    g_cluster_request = 0;
    // A kernel declared with __cluster_dims__(2, 1, 1) reveals itself by reaching a cluster primitive: the first
    // attempt is then abandoned and the grid is run as CTA pairs (both CTAs of a pair resident together).
    for (int ncta = 1; ncta <= MAX_CTAS; ++ncta) {
      g_ncta = ncta;
      if (g.x % ncta) { fprintf(stderr, "emu: grid of %u blocks launched in clusters of %d\n", g.x, ncta); abort(); }
      void* dyn[MAX_CTAS] = {nullptr, nullptr};   // 256 KB-aligned: the low 18 bits of a pointer into a block are its shared-window address
      for (int c = 0; c < ncta; ++c) {
        if (posix_memalign(&dyn[c], 262144, smem ? smem : 1)) abort();
        g_cta[c].dyn = static_cast<unsigned char*>(dyn[c]);
      }
      g_dyn_smem_bytes = smem;
      g_static_smem_hi = 0;
      bool ok = true;
      for (unsigned z = 0; z < g.z && ok; ++z)
        for (unsigned y = 0; y < g.y && ok; ++y)
          for (unsigned x = 0; x < g.x && ok; x += ncta) {
            for (int c = 0; c < ncta; ++c) { g_cta[c].bid.x = x + c; g_cta[c].bid.y = y; g_cta[c].bid.z = z; }
            g_bid = g_cta[0].bid;
            g_dyn_smem = g_cta[0].dyn;
            ok = run_block(nthreads, body);
          }
      g_dyn_smem = nullptr;
      for (int c = 0; c < ncta; ++c) free(dyn[c]);
      if (ok) { if (ncta > 1) ++g_cluster_launches; break; }
      if (ncta == MAX_CTAS) { fprintf(stderr, "emu: cluster launch could not be completed\n"); abort(); }
    }
    g_ncta = 1;
  }
};

[[noreturn]] inline void unsupported_ptx() {
  fprintf(stderr, "emu: inline PTX (tcgen05 / TMA / mbarrier) cannot be emulated -- this kernel needs a B200\n");
  abort();
}

EMU_INTERNAL inline int lane_id() { return (g_cur % g_cta_threads) & 31; }
EMU_INTERNAL inline int warp_id() { return (g_cur % g_cta_threads) >> 5; }          // within the CTA
EMU_INTERNAL inline int warp_slot() { return g_cur / g_cta_threads * ((g_cta_threads + 31) >> 5) + warp_id(); }   // exchange-buffer row
EMU_INTERNAL inline void warp_wait(unsigned mask, int site) {
  Fiber& f = g_fibers[g_cur];
  f.mask = mask; f.site = site; ++f.nwaits;
  yield_wait(WAIT_WARP);
}

template <class T>
EMU_INTERNAL inline T shfl_from(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  g_warp_buf[warp_slot()][lane_id()] = bits;
  warp_wait(mask, 10 + static_cast<int>(sizeof(T)));
  const uint64_t r = g_warp_buf[warp_slot()][src_lane & 31];
  warp_wait(mask, 20 + static_cast<int>(sizeof(T)));
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}

EMU_INTERNAL inline unsigned ballot(unsigned mask, int pred) {
  g_warp_pred[warp_slot()][lane_id()] = pred != 0;
  warp_wait(mask, 3);
  unsigned r = 0;
  const int base = g_cur - lane_id(), wend = (g_cur / g_cta_threads + 1) * g_cta_threads;
  for (int j = 0; j < 32 && base + j < wend; ++j)
    if (((mask >> j) & 1u) && g_fibers[base + j].state != DONE && g_warp_pred[warp_slot()][j]) r |= 1u << j;
  warp_wait(mask, 4);
  return r;
}

EMU_INTERNAL inline unsigned live_mask(unsigned mask) {
  unsigned r = 0;
  const int base = g_cur - lane_id(), wend = (g_cur / g_cta_threads + 1) * g_cta_threads;
  for (int j = 0; j < 32 && base + j < wend; ++j)
    if (((mask >> j) & 1u) && g_fibers[base + j].state != DONE) r |= 1u << j;
  return r;
}

}  // namespace emu

namespace emu {
EMU_INTERNAL inline uint3 tid() { return g_tid; }
EMU_INTERNAL inline unsigned char* dyn_smem() { return g_dyn_smem; }   // what `extern __shared__` arrays are rewritten to
}
#define threadIdx emu::tid()
#define blockIdx emu::g_bid
#define blockDim emu::g_bdim
#define gridDim emu::g_gdim
constexpr int warpSize = 32;

// ---- synchronisation ---------------------------------------------------------------------------
EMU_INTERNAL inline void __syncthreads() { emu::g_fibers[emu::g_cur].pred = 0; emu::yield_wait(emu::WAIT_BLOCK); }
EMU_INTERNAL inline int __syncthreads_or(int p) { emu::g_fibers[emu::g_cur].pred = p; emu::yield_wait(emu::WAIT_BLOCK); return emu::g_red_or[emu::g_cur_cta]; }
EMU_INTERNAL inline int __syncthreads_and(int p) { emu::g_fibers[emu::g_cur].pred = p; emu::yield_wait(emu::WAIT_BLOCK); return emu::g_red_and[emu::g_cur_cta]; }
EMU_INTERNAL inline int __syncthreads_count(int p) { emu::g_fibers[emu::g_cur].pred = p; emu::yield_wait(emu::WAIT_BLOCK); return emu::g_red_cnt[emu::g_cur_cta]; }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_wait(mask, 1); }
template <class T> inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  const int l = emu::lane_id();
  return emu::shfl_from(m, v, (l & ~(width - 1)) | (src & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int lane_mask, int width = 32) {
  const int l = emu::lane_id(), s = l ^ lane_mask;
  return emu::shfl_from(m, v, ((s & ~(width - 1)) == (l & ~(width - 1))) ? s : l);
}
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  const int l = emu::lane_id(), s = l + static_cast<int>(d);
  return emu::shfl_from(m, v, ((s & ~(width - 1)) == (l & ~(width - 1))) ? s : l);
}
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  const int l = emu::lane_id(), s = l - static_cast<int>(d);
  return emu::shfl_from(m, v, (s >= (l & ~(width - 1))) ? s : l);
}
inline unsigned __ballot_sync(unsigned m, int p) { return emu::ballot(m, p); }
inline int __any_sync(unsigned m, int p) { return emu::ballot(m, p) != 0; }
inline int __all_sync(unsigned m, int p) { const unsigned b = emu::ballot(m, p); return b == emu::live_mask(m); }
[[noreturn]] inline void __trap() { fprintf(stderr, "emu: __trap()\n"); abort(); }

// ---- memory ------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T, class U> inline void __stcs(T* p, U v) { *p = v; }
// relaxed atomics (one OS thread, but the race-detection build must see them as atomic accesses)
template <class T, class U> inline T atomicAdd(T* p, U v) {
  static_assert(std::is_integral<T>::value, "only integer atomicAdd is used by the emulated kernels");
  return __atomic_fetch_add(p, static_cast<T>(v), __ATOMIC_RELAXED);
}
template <class T, class U> inline T atomicMax(T* p, U v) { T o = __atomic_load_n(p, __ATOMIC_RELAXED); if (static_cast<T>(v) > o) __atomic_store_n(p, static_cast<T>(v), __ATOMIC_RELAXED); return o; }
template <class T, class U> inline T atomicMin(T* p, U v) { T o = __atomic_load_n(p, __ATOMIC_RELAXED); if (static_cast<T>(v) < o) __atomic_store_n(p, static_cast<T>(v), __ATOMIC_RELAXED); return o; }
template <class T, class U> inline T atomicOr(T* p, U v) { return __atomic_fetch_or(p, static_cast<T>(v), __ATOMIC_RELAXED); }
template <class T, class U> inline T atomicAnd(T* p, U v) { return __atomic_fetch_and(p, static_cast<T>(v), __ATOMIC_RELAXED); }
template <class T, class U> inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, static_cast<T>(v), __ATOMIC_RELAXED); }
template <class T, class U, class V> inline T atomicCAS(T* p, U c, V v) { T e = static_cast<T>(c); __atomic_compare_exchange_n(p, &e, static_cast<T>(v), false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return e; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

// ---- single IEEE operations, fenced against contraction ---------------------------------------------
#define EMU_FENCE(x) asm volatile("" : "+x"(x))
inline double __dadd_rn(double a, double b) { double r = a + b; EMU_FENCE(r); return r; }
inline double __dsub_rn(double a, double b) { double r = a - b; EMU_FENCE(r); return r; }
inline double __dmul_rn(double a, double b) { double r = a * b; EMU_FENCE(r); return r; }
inline double __ddiv_rn(double a, double b) { double r = a / b; EMU_FENCE(r); return r; }
inline double __dsqrt_rn(double a) { double r = std::sqrt(a); EMU_FENCE(r); return r; }
inline double __fma_rn(double a, double b, double c) { double r = std::fma(a, b, c); EMU_FENCE(r); return r; }
inline float __fadd_rn(float a, float b) { float r = a + b; EMU_FENCE(r); return r; }
inline float __fsub_rn(float a, float b) { float r = a - b; EMU_FENCE(r); return r; }
inline float __fmul_rn(float a, float b) { float r = a * b; EMU_FENCE(r); return r; }
inline float __fdiv_rn(float a, float b) { float r = a / b; EMU_FENCE(r); return r; }
inline float __fmaf_rn(float a, float b, float c) { float r = std::fmaf(a, b, c); EMU_FENCE(r); return r; }
inline float __double2float_rn(double a) { float r = static_cast<float>(a); EMU_FENCE(r); return r; }
inline int __double2int_rn(double a) { return static_cast<int>(std::nearbyint(a)); }
inline long long __double2ll_rn(double a) { return static_cast<long long>(std::nearbyint(a)); }
inline int __float2int_rn(float a) { return static_cast<int>(std::nearbyintf(a)); }
inline int __float2int_rd(float a) { return static_cast<int>(std::floor(a)); }
inline int __double2int_rd(double a) { return static_cast<int>(std::floor(a)); }

// ---- bit tricks ----------------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz(static_cast<unsigned>(x)) : 32; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const uint64_t src = (static_cast<uint64_t>(y) << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned byte = static_cast<unsigned>((src >> (8 * (sel & 7))) & 0xff);
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
  const uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
  return static_cast<unsigned>(v >> (shift & 31));
}

// CUDA's global min / max overloads
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> max(A a, B b) { using T = std::common_type_t<A, B>; return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(b) : static_cast<T>(a); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> min(A a, B b) { using T = std::common_type_t<A, B>; return static_cast<T>(b) < static_cast<T>(a) ? static_cast<T>(b) : static_cast<T>(a); }

// C++ conveniences that cuda_runtime.h only declares under nvcc
template <class R, class... A>
inline cudaError_t cudaFuncSetAttribute(R (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }
inline long long clock64() { static long long t = 0; return t += 1000; }
