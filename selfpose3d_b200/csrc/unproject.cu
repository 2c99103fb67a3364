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
#include "unproject_geom.cuh"

namespace sp3d {

constexpr int kUnprojThreads = 256;
constexpr int kChanGroup = 16;

template <typename OutT>
__device__ __forceinline__ OutT to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Output element `i` of a voxel.  PAIR: the float32 value goes out as two bf16 terms (SP3D_BF16X2), hi = bf16(v) into
// plane 0 and lo = bf16(v - hi) into plane 1 (`plane` elements further).
template <typename OutT, bool PAIR>
__device__ __forceinline__ void put(OutT* out, int64_t i, int64_t plane, float v) {
  if (PAIR) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16*>(out)[i] = hi;
    reinterpret_cast<__nv_bfloat16*>(out)[plane + i] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
  } else {
    out[i] = to_out<OutT>(v);
  }
}

// Channel-last fast path: the kChanGroup = 16 values of one voxel as full 16-byte stores.
__device__ __forceinline__ void store16(float* out, const float* r) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
}
__device__ __forceinline__ void store16(__nv_bfloat16* out, const float* r) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(out) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(out + 8) = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void store16_pair(__nv_bfloat16* out, int64_t plane, const float* r) {
  float lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) lo[j] = __fsub_rn(r[j], __bfloat162float(__float2bfloat16_rn(r[j])));
  store16(out, r);
  store16(out + plane, lo);
}

// HM_CL: heat-maps are channel-last (stride_c == 1) with 16-byte aligned pixels -> float4 taps.
template <bool HM_CL, typename OutT, bool PAIR>
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
  const int64_t plane = (int64_t)a.n_cubes * a.out_stride_cube;      // PAIR: distance between the two term planes
  // channel-last voxels of exactly one channel group: one vectorised store of all 16 channels (padding included)
  const bool vec16 = !a.partial && a.out_stride_c == 1 && c_total == kChanGroup && (a.out_stride_vox % kChanGroup) == 0 &&
                     (a.out_stride_cube % 8) == 0 && (reinterpret_cast<uintptr_t>(a.cubes) % 16) == 0;
  if (skip) {
    for (int c = 0; c < c_total; ++c) put<OutT, PAIR>(out, (int64_t)c * a.out_stride_c, plane, 0.0f);
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
        if (j < cn) put<OutT, PAIR>(out, (int64_t)(c0 + j) * a.out_stride_c, plane, num[j]);
      if (c0 + kChanGroup >= a.C) put<OutT, PAIR>(out, (int64_t)a.C * a.out_stride_c, plane, den);
    } else {
      const float d = __fadd_rn(den, 1e-6f);
      float r[kChanGroup];
#pragma unroll
      for (int j = 0; j < kChanGroup; ++j) {
        r[j] = 0.0f;
        if (j < cn) {
          const float q = __fdiv_rn(num[j], d);
          r[j] = (q != q) ? 0.0f : fminf(fmaxf(q, 0.0f), 1.0f);
        }
      }
      if (vec16) {
        if (PAIR) store16_pair(reinterpret_cast<__nv_bfloat16*>(out), plane, r);
        else store16(out, r);
        return;
      }
#pragma unroll
      for (int j = 0; j < kChanGroup; ++j)
        if (j < cn) put<OutT, PAIR>(out, (int64_t)(c0 + j) * a.out_stride_c, plane, r[j]);
    }
  }
  for (int c = c_store; c < c_total; ++c) put<OutT, PAIR>(out, (int64_t)c * a.out_stride_c, plane, 0.0f);
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
  if (a->out_dtype != SP3D_F32 && a->out_dtype != SP3D_BF16 && a->out_dtype != SP3D_BF16X2) return SP3D_ERR_UNSUPPORTED;
  if (a->out_dtype == SP3D_BF16X2 && (a->math_mode != 0 || a->partial)) return SP3D_ERR_UNSUPPORTED;
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
    if (cl) unproject_kernel<true, float, false><<<grid, kUnprojThreads, 0, st>>>(*a);
    else unproject_kernel<false, float, false><<<grid, kUnprojThreads, 0, st>>>(*a);
  } else if (a->out_dtype == SP3D_BF16) {
    if (cl) unproject_kernel<true, __nv_bfloat16, false><<<grid, kUnprojThreads, 0, st>>>(*a);
    else unproject_kernel<false, __nv_bfloat16, false><<<grid, kUnprojThreads, 0, st>>>(*a);
  } else {
    if (cl) unproject_kernel<true, __nv_bfloat16, true><<<grid, kUnprojThreads, 0, st>>>(*a);
    else unproject_kernel<false, __nv_bfloat16, true><<<grid, kUnprojThreads, 0, st>>>(*a);
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
