// Shared helpers of libsp3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/sp3d.h"

namespace sp3d {

void set_last_error(cudaError_t e);

// Returns SP3D_OK or SP3D_ERR_LAUNCH after a kernel launch (no synchronisation).
inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error(e);
    return SP3D_ERR_LAUNCH;
  }
  return SP3D_OK;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace sp3d
