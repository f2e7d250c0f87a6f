// hostdev.cuh — the rule functions are plain integer logic: they are compiled for the device (the product) and for the
// host (tests/host_core runs them on the CPU against the oracle, no GPU needed).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define DS_FN __host__ __device__ __forceinline__

namespace dogstep {
DS_FN int ds_popcll(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}
DS_FN uint32_t ds_min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }
}  // namespace dogstep
