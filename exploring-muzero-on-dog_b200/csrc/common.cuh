// common.cuh — launch-error plumbing shared by every translation unit of libdogstep.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/dogstep.h"

namespace dogstep {
// thread-local copy of the last CUDA error string (dogstep_last_error())
void set_last_error(const char* msg);

// after a <<<>>> launch: peek (do not clear sticky state), never synchronise
inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(cudaGetErrorString(e));
    return DOGSTEP_ERR_CUDA;
  }
  return DOGSTEP_OK;
}

#ifdef __CUDACC__
// Cooperative copy of `bytes` bytes by `nthr` threads (thread index `tid`): 16-byte vectors with four loads in flight per
// thread when both pointers are 16-byte aligned, bytes otherwise (and for the tail).  The rows these kernels move
// (observations, policy rows) are gathers from HBM: a byte-per-thread loop exposes one DRAM latency per element.
__device__ __forceinline__ void coop_copy_bytes(void* __restrict__ dst, const void* __restrict__ src, int64_t bytes, int tid, int nthr) {
  unsigned char* d = (unsigned char*)dst;
  const unsigned char* s = (const unsigned char*)src;
  int64_t done = 0;
  if ((((uintptr_t)d | (uintptr_t)s) & 15u) == 0) {
    const int64_t nv = bytes >> 4;
    const uint4* sv = (const uint4*)s;
    uint4* dv = (uint4*)d;
    int64_t k = tid;
    for (; k + 3 * (int64_t)nthr < nv; k += 4 * (int64_t)nthr) {
      const uint4 a = sv[k], b = sv[k + nthr], c = sv[k + 2 * (int64_t)nthr], e = sv[k + 3 * (int64_t)nthr];
      dv[k] = a; dv[k + nthr] = b; dv[k + 2 * (int64_t)nthr] = c; dv[k + 3 * (int64_t)nthr] = e;
    }
    for (; k < nv; k += nthr) dv[k] = sv[k];
    done = nv << 4;
  }
  for (int64_t k = done + tid; k < bytes; k += nthr) d[k] = s[k];
}
__device__ __forceinline__ void coop_zero_bytes(void* __restrict__ dst, int64_t bytes, int tid, int nthr) {
  unsigned char* d = (unsigned char*)dst;
  int64_t done = 0;
  if (((uintptr_t)d & 15u) == 0) {
    const int64_t nv = bytes >> 4;
    uint4* dv = (uint4*)d;
    for (int64_t k = tid; k < nv; k += nthr) dv[k] = make_uint4(0u, 0u, 0u, 0u);
    done = nv << 4;
  }
  for (int64_t k = done + tid; k < bytes; k += nthr) d[k] = 0;
}
// int8 -> float32 widening copy of n elements (4 at a time when aligned)
__device__ __forceinline__ void coop_widen_i8_f32(float* __restrict__ dst, const int8_t* __restrict__ src, int64_t n, int tid, int nthr) {
  int64_t done = 0;
  if ((((uintptr_t)src) & 3u) == 0 && (((uintptr_t)dst) & 15u) == 0) {
    const int64_t nv = n >> 2;
    const uint32_t* sv = (const uint32_t*)src;
    float4* dv = (float4*)dst;
    for (int64_t k = tid; k < nv; k += nthr) {
      const uint32_t w = sv[k];
      dv[k] = make_float4((float)(int8_t)(w & 0xFF), (float)(int8_t)((w >> 8) & 0xFF), (float)(int8_t)((w >> 16) & 0xFF), (float)(int8_t)(w >> 24));
    }
    done = nv << 2;
  }
  for (int64_t k = done + tid; k < n; k += nthr) dst[k] = (float)src[k];
}
#endif
}  // namespace dogstep
