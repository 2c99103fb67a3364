// Training targets rendered on the device -- replaces the per-item numpy work of the
// reference's DataLoader workers: generate_target_heatmap / generate_3d_target, lib/dataset/JointsDataset.py:237-341.
#include "sp3d_common.cuh"

namespace sp3d {

// One CTA per (joint, item): people's truncated centres and visibility in shared memory, one thread per pixel (strided).
__global__ void __launch_bounds__(256) target_heatmaps_kernel(const sp3d_target_heatmaps_args a) {
  __shared__ int s_mx[64], s_my[64], s_on[64];
  __shared__ int s_any;
  const int j = blockIdx.x, item = blockIdx.y;
  const int n = min(a.n_people[item], a.P);
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    // a person takes part when any of its joints is visible (compute_human_scale != 0, :260-262) ...
    const double* vis = a.joints_vis + (((int64_t)item * a.P + p) * a.J) * a.vstride;
    bool any = false;
    for (int q = 0; q < a.J; ++q) any |= vis[(int64_t)q * a.vstride] == 1.0;
    const double* jt = a.joints + (((int64_t)item * a.P + p) * a.J + j) * a.jstride;
    // ... and this joint of it when it is visible itself (:271).  int() truncates toward zero (:268-269)
    const bool on = any && vis[(int64_t)j * a.vstride] != 0.0;
    s_mx[p] = (int)(jt[0] / a.stride_x);
    s_my[p] = (int)(jt[1] / a.stride_y);
    s_on[p] = on ? 1 : 0;
    if (vis[(int64_t)j * a.vstride] == 1.0) s_any = 1;      // target_weight (:245-249): any person shows joint j
  }
  __syncthreads();
  const int r = a.radius, size = 2 * r + 1;
  float* out = a.target + ((int64_t)item * a.J + j) * a.h * a.w;
  for (int i = threadIdx.x; i < a.h * a.w; i += blockDim.x) {
    const int y = i / a.w, x = i % a.w;
    float m = 0.0f;
    for (int p = 0; p < n; ++p) {
      if (!s_on[p]) continue;
      const int dx = x - s_mx[p] + r, dy = y - s_my[p] + r;
      if (dx >= 0 && dx < size && dy >= 0 && dy < size) m = fmaxf(m, __ldg(a.window + dy * size + dx));
    }
    out[i] = fminf(fmaxf(m, 0.0f), 1.0f);
  }
  if (threadIdx.x == 0) a.target_weight[(int64_t)item * a.J + j] = s_any ? 1.0f : 0.0f;
}

// One thread per voxel (z fastest); float64 Gaussian as numpy evaluates it, stored as float32.
__global__ void __launch_bounds__(256) target_volume_kernel(const sp3d_target_volume_args a) {
  const int64_t n_vox = (int64_t)a.X * a.Y * a.Z;
  const int64_t total = n_vox * a.n_items;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int item = (int)(i / n_vox);
    const int64_t v = i % n_vox;
    const int iz = (int)(v % a.Z), iy = (int)((v / a.Z) % a.Y), ix = (int)(v / ((int64_t)a.Z * a.Y));
    const double gx = a.grid_x[ix], gy = a.grid_y[iy], gz = a.grid_z[iz];
    const int n = min(a.n_people[item], a.P);
    const double lim = 3.0 * a.sigma;
    float m = 0.0f;
    for (int p = 0; p < n; ++p) {
      const double* mu = a.roots + ((int64_t)item * a.P + p) * 3;
      // np.searchsorted(grid, mu - 3 sigma) .. searchsorted(grid, mu + 3 sigma, 'right'): mu - 3s <= g <= mu + 3s
      if (gx < mu[0] - lim || gx > mu[0] + lim || gy < mu[1] - lim || gy > mu[1] + lim || gz < mu[2] - lim || gz > mu[2] + lim)
        continue;
      const double dx = gx - mu[0], dy = gy - mu[1], dz = gz - mu[2];
      const double g = exp(-(dx * dx + dy * dy + dz * dz) / (2.0 * a.sigma * a.sigma));
      m = fmaxf(m, (float)g);
    }
    a.target[i] = fminf(fmaxf(m, 0.0f), 1.0f);
  }
}

}  // namespace sp3d

extern "C" int sp3d_target_heatmaps(const sp3d_target_heatmaps_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->n_items < 0) return SP3D_ERR_INVALID_ARG;
  if (a->n_items == 0) return SP3D_OK;
  if (a->joints == nullptr || a->joints_vis == nullptr || a->n_people == nullptr || a->window == nullptr ||
      a->target == nullptr || a->target_weight == nullptr || a->P < 1 || a->P > 64 || a->J < 1 || a->jstride < 2 ||
      a->vstride < 1 || a->h < 1 || a->w < 1 || a->radius < 0 || !(a->stride_x > 0.0) || !(a->stride_y > 0.0) ||
      a->n_items > 65535)
    return SP3D_ERR_INVALID_ARG;
  target_heatmaps_kernel<<<dim3(a->J, a->n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}

extern "C" int sp3d_target_volume(const sp3d_target_volume_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->n_items < 0) return SP3D_ERR_INVALID_ARG;
  if (a->n_items == 0) return SP3D_OK;
  if (a->roots == nullptr || a->n_people == nullptr || a->grid_x == nullptr || a->grid_y == nullptr ||
      a->grid_z == nullptr || a->target == nullptr || a->P < 1 || a->X < 1 || a->Y < 1 || a->Z < 1 || !(a->sigma > 0.0))
    return SP3D_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->X * a->Y * a->Z * a->n_items;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  target_volume_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}
