// selftest_kernels.cpp -- TEST INFRASTRUCTURE ONLY: small kernels that pin down the emulator itself
// (tests/test_emu_selftest.py).  Built as an executable on top of cuda_emu.h; `selftest <mode>` prints one line.
//   func          warp / block primitives against hand-computed values
//   race_block    thread 0 writes shared memory, everybody reads it with no barrier        (racecheck must report)
//   sync_block    the same with __syncthreads                                              (racecheck must stay silent)
//   race_warp     lanes read a shared word that lane 0 updates in the same step            (racecheck must report)
//   sync_warp     the same with __syncwarp between the read and the update                 (racecheck must stay silent)
//   diverge       lanes of one warp wait in different primitives                           (emulator must abort)
#include <string>

__global__ void k_func(int* out) {
  __shared__ int s[64];
  const int t = threadIdx.x, lane = t & 31;
  s[t] = t * 3;
  __syncthreads();
  int v = s[63 - t];                                                   // block-level exchange
  v += __shfl_sync(0xffffffffu, lane, (lane + 1) & 31);                // rotate
  v += __shfl_xor_sync(0xffffffffu, lane, 16);
  v += __shfl_down_sync(0xffffffffu, lane, 4);                         // lanes 28..31 keep their own value
  const unsigned b = __ballot_sync(0xffffffffu, (lane & 3) == 0);
  v += __popc(b) + __any_sync(0xffffffffu, lane == 7) + __all_sync(0xffffffffu, lane < 32);
  v += __syncthreads_or(t == 5) + __syncthreads_count(t < 10);
  atomicAdd(&out[64], 1);
  out[t] = v;
}
__global__ void k_block(int* out, int use_barrier) {
  __shared__ int s;
  if (threadIdx.x == 0) s = 42;
  if (use_barrier) __syncthreads();
  out[threadIdx.x] = s;
}
__global__ void k_warp(int* out, int fix) {
  __shared__ unsigned bits[2];
  const int lane = threadIdx.x;
  if (lane < 2) bits[lane] = 0;
  __syncwarp();
  for (int t = 0; t < 3; ++t) {
    const unsigned seen = bits[0];
    if (fix) __syncwarp();
    if (lane == 0) bits[0] |= 1u << t;
    __syncwarp();
    out[lane] += seen;
  }
}
__global__ void k_diverge(int* out) {
  const int lane = threadIdx.x;
  int v = lane;
  if (lane & 1) v = __shfl_sync(0xffffffffu, v, 0);   // half of the warp shuffles ...
  else __syncwarp();                                  // ... the other half waits in a different primitive
  out[lane] = v;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "func";
  static int out[80];
  if (mode == "func") {
    emu::Launcher(1, 64, 0, 0).run(k_func, out);
    long long sum = 0;
    for (int t = 0; t < 64; ++t) {
      const int lane = t & 31;
      const int expect = (63 - t) * 3 + ((lane + 1) & 31) + (lane ^ 16) + (lane + 4 < 32 ? lane + 4 : lane) + 8 + 1 + 1 + 1 + 10;
      if (out[t] != expect) { printf("MISMATCH thread %d: %d != %d\n", t, out[t], expect); return 1; }
      sum += out[t];
    }
    printf("func ok %lld atomics %d\n", sum, out[64]);
    return out[64] == 64 ? 0 : 1;
  }
  if (mode == "race_block") emu::Launcher(1, 64, 0, 0).run(k_block, out, 0);
  else if (mode == "sync_block") emu::Launcher(1, 64, 0, 0).run(k_block, out, 1);
  else if (mode == "race_warp") emu::Launcher(1, 32, 0, 0).run(k_warp, out, 0);
  else if (mode == "sync_warp") emu::Launcher(1, 32, 0, 0).run(k_warp, out, 1);
  else if (mode == "diverge") emu::Launcher(1, 32, 0, 0).run(k_diverge, out);
  else return 2;
  printf("%s done out[5]=%d\n", mode.c_str(), out[5]);
  return 0;
}
