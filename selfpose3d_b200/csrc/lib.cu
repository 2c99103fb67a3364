// libsp3d entry points that are not tied to one kernel file: version, error strings, and the
// convolution dispatcher.
#include "sp3d_common.cuh"
#include <string.h>

namespace sp3d {

static thread_local char g_last_error[256] = "";

void set_last_error(cudaError_t e) {
  strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

int conv_simt_f32(const sp3d_conv_args* a, cudaStream_t st);
int conv_tc(const sp3d_conv_args* a, cudaStream_t st);
void set_conv_profile(void* dev);
void set_conv_pair(int on);

}  // namespace sp3d

extern "C" int sp3d_abi_version(void) { return SP3D_ABI_VERSION; }

extern "C" const char* sp3d_strerror(int status) {
  switch (status) {
    case SP3D_OK: return "ok";
    case SP3D_ERR_INVALID_ARG: return "invalid argument";
    case SP3D_ERR_UNSUPPORTED: return "unsupported configuration";
    case SP3D_ERR_LAUNCH: return "CUDA launch failed (see sp3d_last_cuda_error)";
    case SP3D_ERR_WORKSPACE: return "workspace missing, too small or misaligned";
    default: return "unknown status";
  }
}

extern "C" const char* sp3d_last_cuda_error(void) { return sp3d::g_last_error; }

extern "C" void sp3d_debug_conv_profile(void* dev_u64_buffer) { sp3d::set_conv_profile(dev_u64_buffer); }
extern "C" void sp3d_debug_conv_pair(int enable) { sp3d::set_conv_pair(enable); }

extern "C" int64_t sp3d_conv_head_workspace(const sp3d_conv_args* a) {
  if (a == nullptr || a->head_softargmax == nullptr || a->N < 0) return 0;
  int dev = 0, n_sm = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  return (int64_t)a->N * (2 * n_sm) * a->head_softargmax->C * 5 * (int64_t)sizeof(double);
}

extern "C" int sp3d_conv_fwd(const sp3d_conv_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->in == nullptr || a->weight == nullptr || (a->out == nullptr && a->head_softargmax == nullptr) ||
      a->N < 0 || a->cin < 1 ||
      a->cout < 1 || a->cout_pitch < a->cout || a->cout_pitch_w < 1 || a->OD < 0 || a->OH < 0 || a->OW < 0)
    return SP3D_ERR_INVALID_ARG;
  for (int d = 0; d < 3; ++d)
    if (a->ksize[d] < 1 || a->stride[d] < 1 || a->ostride[d] < 1 || a->ooffset[d] < 0) return SP3D_ERR_INVALID_ARG;
  if ((int64_t)a->N * a->OD * a->OH * a->OW == 0) return SP3D_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->algo) {
    case SP3D_CONV_SIMT_F32:
      if (a->cout_pitch_w < a->cout) return SP3D_ERR_INVALID_ARG;
      if (a->head_softargmax != nullptr) return SP3D_ERR_UNSUPPORTED;
      return conv_simt_f32(a, st);
    case SP3D_CONV_TC_BF16:
    case SP3D_CONV_TC_BF16X3: return conv_tc(a, st);
    default: return SP3D_ERR_UNSUPPORTED;
  }
}
