// K2/K5 (tensor-core form) -- placeholder dispatcher until the tcgen05 kernels land.
#include "sp3d_common.cuh"

namespace sp3d {
int conv_tc(const sp3d_conv_args*, cudaStream_t) { return SP3D_ERR_UNSUPPORTED; }
}  // namespace sp3d
