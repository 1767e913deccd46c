// emu_tmap.h -- TEST INFRASTRUCTURE ONLY: what the emulated cuTensorMapEncodeTiled records for the emulated TMA.
#pragma once
#include <stdint.h>

namespace emu {
struct TensorMapRec {
  uint64_t magic;
  void* base;
  uint32_t rank, elem_bytes, swizzle;
  uint64_t dims[5], strides[5];   // strides in bytes, strides[0] = elem_bytes
  uint32_t box[5], estr[5];
};
constexpr uint64_t TMAP_MAGIC = 0x6f70625f746d6170ull;   // "opb_tmap"
}  // namespace emu
