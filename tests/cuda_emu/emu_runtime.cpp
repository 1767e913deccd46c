// emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY: the handful of CUDA runtime entry points that
// csrc/opb_api.cu calls, implemented over host memory for the emulated build (cuda_emu.h).
// "Device" memory is malloc'ed host memory, streams and events are dummies (everything is
// synchronous), graphs / cluster launches / tensor-map encoding report "not supported" or
// do nothing: the tensor-core paths cannot run here.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "emu_tmap.h"

extern "C" {

static int g_dummy_handles = 0;
static void* new_handle() { ++g_dummy_handles; return malloc(16); }

cudaError_t cudaMalloc(void** p, size_t n) {
  *p = nullptr;
  if (posix_memalign(p, 1024, n ? n : 1)) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, n);                     // uninitialised device memory is not zero
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }

cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {   // macro-renamed to _v2 by the header
  memset(p, 0, sizeof(*p));
  strcpy(p->name, "cuda_emu (CPU fibers, tests only)");
  p->major = 10; p->minor = 0;
  // OPB_EMU_SMS: a small SM count makes every persistent CTA loop over many tiles (pipeline wrap-around, phase parity)
  p->multiProcessorCount = getenv("OPB_EMU_SMS") ? atoi(getenv("OPB_EMU_SMS")) : 148;
  p->sharedMemPerBlockOptin = 232448; p->warpSize = 32;
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emu: operation not supported under emulation"; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  memset(a, 0, sizeof(*a));
  a->type = cudaMemoryTypeUnregistered;
  a->hostPointer = const_cast<void*>(p);
  return cudaSuccess;
}

cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = static_cast<cudaStream_t>(new_handle()); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { if (g) *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = static_cast<cudaEvent_t>(new_handle()); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { if (e) *e = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t*, const void*, void**) { return cudaErrorNotSupported; }
cudaError_t cudaLaunchKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return cudaErrorNotSupported; }

// cuTensorMapEncodeTiled: record the map for the emulated TMA (ptx_emu.cuh); the 128-byte CUtensorMap holds a
// pointer to the record (never freed: a few hundred maps per test process)
static CUresult emu_cuTensorMapEncodeTiled(CUtensorMap* tm, CUtensorMapDataType dtype, cuuint32_t rank, void* base,
                                           const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                           const cuuint32_t* estr, CUtensorMapInterleave, CUtensorMapSwizzle swizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  static const uint32_t kBytes[] = {1, 2, 4, 4, 8, 8, 2, 4, 8, 2, 4, 4, 4};
  if (rank < 1 || rank > 5 || static_cast<unsigned>(dtype) >= sizeof(kBytes) / sizeof(kBytes[0])) return CUDA_ERROR_INVALID_VALUE;
  emu::TensorMapRec* r = new emu::TensorMapRec();
  r->magic = emu::TMAP_MAGIC;
  r->base = base; r->rank = rank; r->elem_bytes = kBytes[dtype]; r->swizzle = static_cast<uint32_t>(swizzle);
  for (cuuint32_t d = 0; d < rank; ++d) {
    r->dims[d] = dims[d];
    r->strides[d] = d ? strides[d - 1] : r->elem_bytes;
    r->box[d] = box[d];
    r->estr[d] = estr[d];
    // the documented constraints of the real encoder (cuda.h): box <= 256, traversal stride 1..8, strides multiples
    // of 16 bytes and < 2^40, non-empty dimensions <= 2^32
    if (box[d] == 0 || box[d] > 256 || estr[d] == 0 || estr[d] > 8 || dims[d] == 0 || dims[d] > (1ull << 32) ||
        (d && ((strides[d - 1] % 16) || strides[d - 1] >= (1ull << 40)))) { delete r; return CUDA_ERROR_INVALID_VALUE; }
  }
  static const uint32_t kSwizzleSpan[] = {0, 32, 64, 128};       // NONE, 32B, 64B, 128B: inner box bytes must fit the span
  const uint32_t inner = box[0] * r->elem_bytes;
  if (reinterpret_cast<uintptr_t>(base) % 16 || inner % 16 ||
      (r->swizzle >= 1 && r->swizzle <= 3 && inner > kSwizzleSpan[r->swizzle])) { delete r; return CUDA_ERROR_INVALID_VALUE; }
  memset(tm, 0, sizeof(*tm));
  memcpy(tm, &r, sizeof(r));
  return CUDA_SUCCESS;
}
static CUresult emu_driver_stub(...) { return CUDA_SUCCESS; }
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
  *fn = !strcmp(name, "cuTensorMapEncodeTiled") ? reinterpret_cast<void*>(&emu_cuTensorMapEncodeTiled)
                                                : reinterpret_cast<void*>(&emu_driver_stub);
  if (q) *q = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}

}  // extern "C"

// Debug aid: OPB_EMU_BACKTRACE=1 prints a symbolised backtrace on SIGSEGV (fiber stacks included).
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
namespace {
void emu_segv(int sig) {
  void* frames[48];
  const int n = backtrace(frames, 48);
  const char msg[] = "emu: fatal signal, backtrace:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}
struct EmuSegvInstaller {
  EmuSegvInstaller() {
    if (!getenv("OPB_EMU_BACKTRACE")) return;
    static char alt[1 << 16];
    stack_t ss; ss.ss_sp = alt; ss.ss_size = sizeof(alt); ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa; memset(&sa, 0, sizeof(sa));
    sa.sa_handler = emu_segv; sa.sa_flags = SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
  }
} g_emu_segv_installer;
}  // namespace

#if defined(__x86_64__)
// void emu_switch(void** save_sp, void* to_sp): cooperative fiber switch used by cuda_emu.h
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
#endif
