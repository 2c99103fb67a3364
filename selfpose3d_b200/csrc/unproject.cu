// K1 -- fused multi-view un-projection (voxel fill).
//
// One thread per voxel (z fastest, so a warp covers 32 consecutive voxels and its stores are
// contiguous).  For every view the voxel centre is projected with the radial/tangential camera
// model, tested against the image bounds, mapped through the input affine to heat-map
// coordinates and bilinearly sampled for all channels; numerators and the view count stay in
// registers; one store per channel at the end.  The coordinate arithmetic follows the
// reference's float32 operation order with explicit round-to-nearest intrinsics (no FMA
// contraction), so the in-image mask agrees with the oracle bit for bit.
//
// Reference semantics: lib/models/project_layer.py:42-102, lib/utils/cameras.py:27-55,
// lib/utils/transforms.py:119-123.
#include "sp3d_common.cuh"

namespace sp3d {

constexpr int kUnprojThreads = 256;
constexpr int kChanGroup = 16;

struct ViewSample {
  float wx0, wx1, wy0, wy1;  // bilinear weights
  int x0, y0;                // top-left tap
  float m;                   // in-image mask (0/1)
};

// World point -> heat-map sampling position of one view.  `cam` points at SP3D_CAM_FLOATS floats.
__device__ __forceinline__ ViewSample project_view(const float* __restrict__ cam, float gx, float gy, float gz,
                                                   float img_w, float img_h, float cfg_w, float cfg_h,
                                                   float hm_w, float hm_h) {
  const float dx = __fsub_rn(gx, cam[9]);
  const float dy = __fsub_rn(gy, cam[10]);
  const float dz = __fsub_rn(gz, cam[11]);
  const float xc = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam[0]), __fmul_rn(dy, cam[1])), __fmul_rn(dz, cam[2]));
  const float yc = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam[3]), __fmul_rn(dy, cam[4])), __fmul_rn(dz, cam[5]));
  float zc = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam[6]), __fmul_rn(dy, cam[7])), __fmul_rn(dz, cam[8]));
  zc = __fadd_rn(zc, 1e-5f);
  const float y0 = __fdiv_rn(xc, zc);
  const float y1 = __fdiv_rn(yc, zc);
  const float r2 = fminf(__fadd_rn(__fmul_rn(y0, y0), __fmul_rn(y1, y1)), 1e10f);
  const float r4 = __fmul_rn(r2, r2);
  const float r6 = __fmul_rn(r4, r2);
  const float radial = __fadd_rn(
      1.0f, __fadd_rn(__fadd_rn(__fmul_rn(cam[16], r2), __fmul_rn(cam[17], r4)), __fmul_rn(cam[18], r6)));
  const float tan = __fadd_rn(__fmul_rn(cam[19], y1), __fmul_rn(cam[20], y0));
  const float corr = __fadd_rn(radial, __fmul_rn(2.0f, tan));
  const float u = __fadd_rn(__fmul_rn(y0, corr), __fmul_rn(cam[20], r2));
  const float v = __fadd_rn(__fmul_rn(y1, corr), __fmul_rn(cam[19], r2));
  float px = __fadd_rn(__fmul_rn(cam[12], u), cam[14]);
  float py = __fadd_rn(__fmul_rn(cam[13], v), cam[15]);

  const float width = cam[27], height = cam[28];
  ViewSample s;
  s.m = (px >= 0.0f && py >= 0.0f && px < width && py < height) ? 1.0f : 0.0f;  // mask on un-clamped pixels
  const float hi = fmaxf(width, height);
  // torch.clamp semantics: NaN propagates; fminf/fmaxf would drop it, so keep NaN explicitly
  px = (px != px) ? px : fminf(fmaxf(px, -1.0f), hi);
  py = (py != py) ? py : fminf(fmaxf(py, -1.0f), hi);
  float qx = __fadd_rn(__fadd_rn(__fmul_rn(cam[21], px), __fmul_rn(cam[22], py)), cam[23]);
  const float qy = __fadd_rn(__fadd_rn(__fmul_rn(cam[24], px), __fmul_rn(cam[25], py)), cam[26]);
  if (cam[29] != 0.0f) qx = __fsub_rn(img_w, qx);
  const float uu = __fdiv_rn(__fmul_rn(qx, cfg_w), img_w);
  const float vv = __fdiv_rn(__fmul_rn(qy, cfg_h), img_h);
  float sx = __fsub_rn(__fmul_rn(__fdiv_rn(uu, __fsub_rn(cfg_w, 1.0f)), 2.0f), 1.0f);
  float sy = __fsub_rn(__fmul_rn(__fdiv_rn(vv, __fsub_rn(cfg_h, 1.0f)), 2.0f), 1.0f);
  sx = (sx != sx) ? sx : fminf(fmaxf(sx, -1.1f), 1.1f);
  sy = (sy != sy) ? sy : fminf(fmaxf(sy, -1.1f), 1.1f);
  // grid_sample(align_corners=True) un-normalisation
  const float fx = __fmul_rn(__fdiv_rn(__fadd_rn(sx, 1.0f), 2.0f), __fsub_rn(hm_w, 1.0f));
  const float fy = __fmul_rn(__fdiv_rn(__fadd_rn(sy, 1.0f), 2.0f), __fsub_rn(hm_h, 1.0f));
  const float x0f = floorf(fx), y0f = floorf(fy);
  s.wx1 = __fsub_rn(fx, x0f);
  s.wy1 = __fsub_rn(fy, y0f);
  s.wx0 = __fsub_rn(1.0f, s.wx1);
  s.wy0 = __fsub_rn(1.0f, s.wy1);
  // NaN coordinates (never produced by finite cameras) sample nothing
  s.x0 = (fx == fx) ? (int)x0f : -4;
  s.y0 = (fy == fy) ? (int)y0f : -4;
  return s;
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// HM_CL: heat-maps are channel-last (stride_c == 1) with 16-byte aligned pixels -> float4 taps.
template <bool HM_CL, typename OutT>
__global__ void __launch_bounds__(kUnprojThreads) unproject_kernel(const sp3d_unproject_args a) {
  __shared__ float s_cam[SP3D_MAX_VIEWS * SP3D_CAM_FLOATS];
  __shared__ float s_center[4];
  const int cube = blockIdx.y;
  const int sample = a.cube_sample ? a.cube_sample[cube] : cube / a.cubes_per_sample;
  const int N = a.X * a.Y * a.Z;
  const int tid = threadIdx.x;
  for (int i = tid; i < a.V * SP3D_CAM_FLOATS; i += kUnprojThreads)
    s_cam[i] = a.cams[(int64_t)sample * a.V * SP3D_CAM_FLOATS + i];
  if (tid < 4) s_center[tid] = (tid < 3 || a.center_stride > 3) ? a.centers[(int64_t)cube * a.center_stride + tid] : 0.0f;
  __syncthreads();

  const int vox = blockIdx.x * kUnprojThreads + tid;
  if (vox >= N) return;
  const bool skip = a.check_flag && !(s_center[3] >= 0.0f);
  const int iz = vox % a.Z;
  const int iy = (vox / a.Z) % a.Y;
  const int ix = vox / (a.Z * a.Y);
  const float gx = __fadd_rn(a.lin_x[ix], s_center[0]);
  const float gy = __fadd_rn(a.lin_y[iy], s_center[1]);
  const float gz = __fadd_rn(a.lin_z[iz], s_center[2]);
  if (a.grids != nullptr) {
    float* g = a.grids + ((int64_t)cube * N + vox) * 3;
    g[0] = skip ? 0.0f : gx;
    g[1] = skip ? 0.0f : gy;
    g[2] = skip ? 0.0f : gz;
  }
  OutT* out = reinterpret_cast<OutT*>(a.cubes) + (int64_t)cube * a.out_stride_cube + (int64_t)vox * a.out_stride_vox;
  const int c_store = a.partial ? a.C + 1 : a.C;
  const int c_total = a.out_c_pad > c_store ? a.out_c_pad : c_store;
  if (skip) {
    for (int c = 0; c < c_total; ++c) out[(int64_t)c * a.out_stride_c] = to_out<OutT>(0.0f);
    return;
  }
  const float hm_w = (float)a.w, hm_h = (float)a.h;

  for (int c0 = 0; c0 < a.C; c0 += kChanGroup) {
    const int cn = min(kChanGroup, a.C - c0);
    float num[kChanGroup];
#pragma unroll
    for (int j = 0; j < kChanGroup; ++j) num[j] = 0.0f;
    float den = 0.0f;
    for (int v = a.view_begin; v < a.view_end; ++v) {
      const ViewSample s = project_view(s_cam + v * SP3D_CAM_FLOATS, gx, gy, gz, a.img_w, a.img_h, a.hm_cfg_w,
                                          a.hm_cfg_h, hm_w, hm_h);
      den = __fadd_rn(den, s.m);
      if (s.m == 0.0f) continue;  // masked views contribute exact zeros
      const float* hm = a.heatmaps[v] + (int64_t)sample * a.hm_stride_b + (int64_t)c0 * a.hm_stride_c;
      float acc[kChanGroup];
#pragma unroll
      for (int j = 0; j < kChanGroup; ++j) acc[j] = 0.0f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xi = s.x0 + (t & 1);
        const int yi = s.y0 + (t >> 1);
        const float wgt = __fmul_rn((t & 1) ? s.wx1 : s.wx0, (t >> 1) ? s.wy1 : s.wy0);
        if (xi < 0 || xi >= a.w || yi < 0 || yi >= a.h) continue;  // zero padding
        const float* p = hm + (int64_t)yi * a.hm_stride_h + (int64_t)xi * a.hm_stride_w;
        if (HM_CL) {
#pragma unroll
          for (int j = 0; j < kChanGroup; j += 4) {
            if (j < cn) {
              const float4 q = ldg4(p + j);
              acc[j + 0] = __fadd_rn(acc[j + 0], __fmul_rn(q.x, wgt));
              acc[j + 1] = __fadd_rn(acc[j + 1], __fmul_rn(q.y, wgt));
              acc[j + 2] = __fadd_rn(acc[j + 2], __fmul_rn(q.z, wgt));
              acc[j + 3] = __fadd_rn(acc[j + 3], __fmul_rn(q.w, wgt));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < kChanGroup; ++j)
            if (j < cn) acc[j] = __fadd_rn(acc[j], __fmul_rn(__ldg(p + (int64_t)j * a.hm_stride_c), wgt));
        }
      }
#pragma unroll
      for (int j = 0; j < kChanGroup; ++j) num[j] = __fadd_rn(num[j], acc[j]);
    }
    if (a.partial) {
#pragma unroll
      for (int j = 0; j < kChanGroup; ++j)
        if (j < cn) out[(int64_t)(c0 + j) * a.out_stride_c] = to_out<OutT>(num[j]);
      if (c0 + kChanGroup >= a.C) out[(int64_t)a.C * a.out_stride_c] = to_out<OutT>(den);
    } else {
      const float d = __fadd_rn(den, 1e-6f);
#pragma unroll
      for (int j = 0; j < kChanGroup; ++j) {
        if (j < cn) {
          float r = __fdiv_rn(num[j], d);
          r = (r != r) ? 0.0f : fminf(fmaxf(r, 0.0f), 1.0f);
          out[(int64_t)(c0 + j) * a.out_stride_c] = to_out<OutT>(r);
        }
      }
    }
  }
  for (int c = c_store; c < c_total; ++c) out[(int64_t)c * a.out_stride_c] = to_out<OutT>(0.0f);
}

__global__ void unproject_finalize_kernel(const sp3d_unproject_finalize_args a) {
  const int64_t total = a.n_cubes * a.N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cube = i / a.N, vox = i % a.N;
    float* p = a.buf + cube * a.stride_cube + vox * a.stride_vox;
    const float d = __fadd_rn(p[a.C * a.stride_c], 1e-6f);
    for (int64_t c = 0; c < a.C; ++c) {
      float r = __fdiv_rn(p[c * a.stride_c], d);
      p[c * a.stride_c] = (r != r) ? 0.0f : fminf(fmaxf(r, 0.0f), 1.0f);
    }
  }
}

}  // namespace sp3d

namespace sp3d {
int unproject_fast(const sp3d_unproject_args* a, cudaStream_t st);
}

extern "C" int sp3d_unproject_fwd(const sp3d_unproject_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->n_cubes < 0) return SP3D_ERR_INVALID_ARG;
  if (a->n_cubes == 0) return SP3D_OK;  // nothing to fill (no valid proposal): pointers may be null
  if (a->V < 1 || a->V > SP3D_MAX_VIEWS || a->C < 1 || a->cubes == nullptr ||
      a->cams == nullptr || a->centers == nullptr || a->lin_x == nullptr || a->lin_y == nullptr ||
      a->lin_z == nullptr || a->center_stride < 3 || a->cubes_per_sample < 1 || a->view_begin < 0 ||
      a->view_end > a->V || a->view_begin > a->view_end || (a->check_flag && a->center_stride < 4))
    return SP3D_ERR_INVALID_ARG;
  for (int v = a->view_begin; v < a->view_end; ++v)
    if (a->heatmaps[v] == nullptr) return SP3D_ERR_INVALID_ARG;
  if (a->out_dtype != SP3D_F32 && a->out_dtype != SP3D_BF16) return SP3D_ERR_UNSUPPORTED;
  const int N = a->X * a->Y * a->Z;
  if (N <= 0 || a->n_cubes > 65535) return SP3D_ERR_INVALID_ARG;
  if (a->math_mode == 1) return unproject_fast(a, static_cast<cudaStream_t>(stream));
  if (a->math_mode != 0 || a->hm_dtype != SP3D_F32) return SP3D_ERR_UNSUPPORTED;
  bool cl = a->hm_stride_c == 1 && (a->hm_stride_w % 4) == 0 && (a->hm_stride_h % 4) == 0 && (a->hm_stride_b % 4) == 0;
  for (int v = a->view_begin; v < a->view_end && cl; ++v) cl = (reinterpret_cast<uintptr_t>(a->heatmaps[v]) % 16) == 0;
  // float4 taps read whole groups of 4 channels: the pitch must cover them
  if (cl && a->hm_stride_w < ((a->C + 3) / 4) * 4) cl = false;
  dim3 grid(ceil_div(N, kUnprojThreads), a->n_cubes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->out_dtype == SP3D_F32) {
    if (cl) unproject_kernel<true, float><<<grid, kUnprojThreads, 0, st>>>(*a);
    else unproject_kernel<false, float><<<grid, kUnprojThreads, 0, st>>>(*a);
  } else {
    if (cl) unproject_kernel<true, __nv_bfloat16><<<grid, kUnprojThreads, 0, st>>>(*a);
    else unproject_kernel<false, __nv_bfloat16><<<grid, kUnprojThreads, 0, st>>>(*a);
  }
  return check_launch();
}

extern "C" int sp3d_unproject_finalize(const sp3d_unproject_finalize_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->buf == nullptr || a->C < 1 || a->N < 1) return SP3D_ERR_INVALID_ARG;
  if (a->n_cubes == 0) return SP3D_OK;
  const int64_t total = a->n_cubes * a->N;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  unproject_finalize_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}
