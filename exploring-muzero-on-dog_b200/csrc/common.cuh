// common.cuh — launch-error plumbing shared by every translation unit of libdogstep.so.
#pragma once
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
}  // namespace dogstep
