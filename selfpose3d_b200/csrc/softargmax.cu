// K4 -- soft-argmax over a voxel cube (online softmax, single pass over the cube).
//
// grid = (splits, n_cubes).  Each CTA streams a contiguous voxel range of one cube; a thread owns
// one voxel per step and keeps, per channel, the running (max, sum e, sum e*(g - centre)) of the
// online softmax in registers; warps and CTAs merge those states in float64; a second tiny kernel
// merges the splits and adds the centre back.  Voxel coordinates g are rebuilt as
// fl(lin + centre), i.e. exactly the `grids` tensor of ProjectLayer, and accumulated relative to
// the cube centre to keep magnitudes (and float32 rounding) small.
//
// Reference semantics: SoftArgmaxLayer.forward, lib/models/pose_regression_net.py:19-28.
#include "sp3d_common.cuh"
#include <math.h>

namespace sp3d {

constexpr int kSaThreads = 256;
constexpr int kSaGroup = 16;
constexpr int kSaState = 5;  // m, s, wx, wy, wz (float64 in the workspace)

struct SaState {
  double m, s, wx, wy, wz;
};

__device__ __forceinline__ SaState sa_merge(const SaState& a, const SaState& b) {
  if (b.s == 0.0) return a;
  if (a.s == 0.0) return b;
  const double M = a.m > b.m ? a.m : b.m;
  const double fa = exp(a.m - M), fb = exp(b.m - M);
  SaState r;
  r.m = M;
  r.s = a.s * fa + b.s * fb;
  r.wx = a.wx * fa + b.wx * fb;
  r.wy = a.wy * fa + b.wy * fb;
  r.wz = a.wz * fa + b.wz * fb;
  return r;
}

__device__ __forceinline__ SaState sa_shfl_down(const SaState& a, int off) {
  SaState r;
  r.m = __shfl_down_sync(0xffffffffu, a.m, off);
  r.s = __shfl_down_sync(0xffffffffu, a.s, off);
  r.wx = __shfl_down_sync(0xffffffffu, a.wx, off);
  r.wy = __shfl_down_sync(0xffffffffu, a.wy, off);
  r.wz = __shfl_down_sync(0xffffffffu, a.wz, off);
  return r;
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// VEC: channel-last float32 input with 16-byte aligned voxels -> float4 loads of 4 channels.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kSaThreads) softargmax_partial_kernel(const sp3d_softargmax_args a, int splits,
                                                                         int vox_per_split) {
  __shared__ SaState s_red[kSaThreads / 32];
  const int cube = blockIdx.y;
  const int split = blockIdx.x;
  const int tid = threadIdx.x;
  const int N = a.X * a.Y * a.Z;
  const float* cen = a.centers + (int64_t)cube * a.center_stride;
  double* ws = reinterpret_cast<double*>(a.workspace) + ((int64_t)cube * splits + split) * a.C * kSaState;
  if (a.check_flag && !(cen[3] >= 0.0f)) {
    for (int i = tid; i < a.C * kSaState; i += kSaThreads) ws[i] = 0.0;
    return;
  }
  const float cx = cen[0], cy = cen[1], cz = cen[2];
  const int v_begin = split * vox_per_split;
  const int v_end = min(N, v_begin + vox_per_split);
  const T* xb = reinterpret_cast<const T*>(a.x) + (int64_t)cube * a.stride_cube;

  for (int c0 = 0; c0 < a.C; c0 += kSaGroup) {
    const int cn = min(kSaGroup, a.C - c0);
    float m[kSaGroup], s[kSaGroup], wx[kSaGroup], wy[kSaGroup], wz[kSaGroup];
#pragma unroll
    for (int j = 0; j < kSaGroup; ++j) { m[j] = -INFINITY; s[j] = 0.f; wx[j] = 0.f; wy[j] = 0.f; wz[j] = 0.f; }
    for (int vox = v_begin + tid; vox < v_end; vox += kSaThreads) {
      const int iz = vox % a.Z, iy = (vox / a.Z) % a.Y, ix = vox / (a.Z * a.Y);
      // g = fl(lin + centre) as in ProjectLayer.compute_grid; accumulate g - centre
      const float gx = __fsub_rn(__fadd_rn(a.lin_x[ix], cx), cx);
      const float gy = __fsub_rn(__fadd_rn(a.lin_y[iy], cy), cy);
      const float gz = __fsub_rn(__fadd_rn(a.lin_z[iz], cz), cz);
      const T* p = xb + (int64_t)vox * a.stride_vox + (int64_t)c0 * a.stride_c;
      float xv[kSaGroup];
      if (VEC) {
#pragma unroll
        for (int j = 0; j < kSaGroup; j += 4) {
          if (j < cn) {
            const float4 q = ldg4(reinterpret_cast<const float*>(p) + j);
            xv[j] = q.x; xv[j + 1] = q.y; xv[j + 2] = q.z; xv[j + 3] = q.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kSaGroup; ++j)
          if (j < cn) xv[j] = load_as_float<T>(p + (int64_t)j * a.stride_c);
      }
#pragma unroll
      for (int j = 0; j < kSaGroup; ++j) {
        if (j < cn) {
          const float z = __fmul_rn(a.beta, xv[j]);
          if (z > m[j]) {
            const float sc = expf(m[j] - z);  // exp(-inf) = 0 on the first element
            s[j] = fmaf(s[j], sc, 1.0f);
            wx[j] = fmaf(wx[j], sc, gx);
            wy[j] = fmaf(wy[j], sc, gy);
            wz[j] = fmaf(wz[j], sc, gz);
            m[j] = z;
          } else {
            const float e = expf(z - m[j]);
            s[j] += e;
            wx[j] = fmaf(e, gx, wx[j]);
            wy[j] = fmaf(e, gy, wy[j]);
            wz[j] = fmaf(e, gz, wz[j]);
          }
        }
      }
    }
    // merge threads -> CTA, one channel at a time, in float64
    for (int j = 0; j < cn; ++j) {
      SaState st{(double)m[j], (double)s[j], (double)wx[j], (double)wy[j], (double)wz[j]};
      for (int off = 16; off > 0; off >>= 1) st = sa_merge(st, sa_shfl_down(st, off));
      if ((tid & 31) == 0) s_red[tid >> 5] = st;
      __syncthreads();
      if (tid == 0) {
        SaState r = s_red[0];
        for (int w = 1; w < kSaThreads / 32; ++w) r = sa_merge(r, s_red[w]);
        double* o = ws + (int64_t)(c0 + j) * kSaState;
        o[0] = r.m; o[1] = r.s; o[2] = r.wx; o[3] = r.wy; o[4] = r.wz;
      }
      __syncthreads();
    }
  }
}


// Streaming form for channel-last float32 volumes (the layout V2VNet's head writes): `LPV` lanes share a voxel,
// each owning 4 consecutive channels, so one 16-byte load per lane is a fully coalesced 512-byte warp request
// and the online-softmax state is 20 registers per lane.  The CTA merge is parallel: shared per-channel maximum
// first, then every lane rescales its state once and plain sums are reduced (float64 from here on).
// 2^x on the special-function unit (2 ulp); the argument is a small non-positive difference wherever the
// weight matters, so the result is as accurate as expf for this use
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int LPV>
__global__ void __launch_bounds__(kSaThreads) softargmax_stream_kernel(const sp3d_softargmax_args a, int splits,
                                                                        int vox_per_split) {
  constexpr int kWarps = kSaThreads / 32;
  __shared__ float s_g[3][256];
  __shared__ float s_max[kWarps][4 * LPV];
  __shared__ double s_sum[kWarps][4 * LPV][4];
  const int cube = blockIdx.y, split = blockIdx.x, tid = threadIdx.x;
  const int N = a.X * a.Y * a.Z;
  const float* cen = a.centers + (int64_t)cube * a.center_stride;
  double* ws = reinterpret_cast<double*>(a.workspace) + ((int64_t)cube * splits + split) * a.C * kSaState;
  if (a.check_flag && !(cen[3] >= 0.0f)) {
    for (int i = tid; i < a.C * kSaState; i += kSaThreads) ws[i] = 0.0;
    return;
  }
  const float cx = cen[0], cy = cen[1], cz = cen[2];
  // g - centre with g = fl(lin + centre), the `grids` values of ProjectLayer.compute_grid
  for (int i = tid; i < a.X; i += kSaThreads) s_g[0][i] = __fsub_rn(__fadd_rn(a.lin_x[i], cx), cx);
  for (int i = tid; i < a.Y; i += kSaThreads) s_g[1][i] = __fsub_rn(__fadd_rn(a.lin_y[i], cy), cy);
  for (int i = tid; i < a.Z; i += kSaThreads) s_g[2][i] = __fsub_rn(__fadd_rn(a.lin_z[i], cz), cz);
  __syncthreads();
  const int v_begin = split * vox_per_split;
  const int v_end = min(N, v_begin + vox_per_split);
  const float* xb = reinterpret_cast<const float*>(a.x) + (int64_t)cube * a.stride_cube;
  const int cq = tid % LPV;                      // channel quad of this lane
  const int nch = min(4, a.C - 4 * cq);          // live channels in the quad (<= 0: none)
  float m[4], s[4], wx[4], wy[4], wz[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; s[k] = 0.f; wx[k] = 0.f; wy[k] = 0.f; wz[k] = 0.f; }
  // voxel (ix, iy, iz) advances by a constant step: carried incrementally instead of two integer divisions per voxel
  constexpr int kStep = kSaThreads / LPV;
  const int step_z = kStep % a.Z, step_yq = kStep / a.Z;
  const int step_y = step_yq % a.Y, step_x = step_yq / a.Y;
  int vox = v_begin + tid / LPV;
  int iz = vox % a.Z, iy = (vox / a.Z) % a.Y, ix = vox / (a.Z * a.Y);
  const float kLog2e = 1.4426950408889634f;
  constexpr int kUnroll = 4;   // loads of 4 voxels in flight per lane before the (serially dependent) updates
  while (vox < v_end) {
    float4 q[kUnroll];
    float g[kUnroll][3];
    bool live[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      live[j] = vox < v_end;
      if (live[j]) {
        q[j] = ldg4(xb + (int64_t)vox * a.stride_vox + 4 * cq);
        g[j][0] = s_g[0][ix]; g[j][1] = s_g[1][iy]; g[j][2] = s_g[2][iz];
      }
      vox += kStep;
      iz += step_z;
      if (iz >= a.Z) { iz -= a.Z; ++iy; }
      iy += step_y;
      if (iy >= a.Y) { iy -= a.Y; ++ix; }
      ix += step_x;
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      if (!live[j]) continue;
      const float gx = g[j][0], gy = g[j][1], gz = g[j][2];
      const float xv[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nch) {
          const float z = __fmul_rn(a.beta, xv[k]);        // the reference's float32 beta * x
          if (z > m[k]) {
            const float sc = ex2_approx((m[k] - z) * kLog2e);  // 2^-inf = 0 on the first element
            s[k] = fmaf(s[k], sc, 1.0f);
            wx[k] = fmaf(wx[k], sc, gx);
            wy[k] = fmaf(wy[k], sc, gy);
            wz[k] = fmaf(wz[k], sc, gz);
            m[k] = z;
          } else {
            const float e = ex2_approx((z - m[k]) * kLog2e);
            s[k] += e;
            wx[k] = fmaf(e, gx, wx[k]);
            wy[k] = fmaf(e, gy, wy[k]);
            wz[k] = fmaf(e, gz, wz[k]);
          }
        }
      }
    }
  }
  // CTA-wide maximum per channel
  const int lane = tid & 31, warp = tid >> 5;
  float mm[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mm[k] = m[k];
    for (int off = LPV; off < 32; off <<= 1) mm[k] = fmaxf(mm[k], __shfl_xor_sync(0xffffffffu, mm[k], off));
  }
  if (lane < LPV) {
#pragma unroll
    for (int k = 0; k < 4; ++k) s_max[warp][4 * lane + k] = mm[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float M = s_max[0][4 * cq + k];
    for (int w = 1; w < kWarps; ++w) M = fmaxf(M, s_max[w][4 * cq + k]);
    mm[k] = M;
    // this lane's state on the common maximum (an empty state has s = 0 and m = -inf: factor forced to 0)
    const double f = (s[k] > 0.0f) ? exp((double)m[k] - (double)M) : 0.0;
    double v4[4] = {(double)s[k] * f, (double)wx[k] * f, (double)wy[k] * f, (double)wz[k] * f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      for (int off = LPV; off < 32; off <<= 1) v4[i] += __shfl_xor_sync(0xffffffffu, v4[i], off);
    if (lane < LPV) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s_sum[warp][4 * lane + k][i] = v4[i];
    }
  }
  __syncthreads();
  if (tid < 4 * LPV && tid < a.C) {
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    float M = s_max[0][tid];
    for (int w = 0; w < kWarps; ++w) {
      M = fmaxf(M, s_max[w][tid]);
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] += s_sum[w][tid][i];
    }
    double* o = ws + (int64_t)tid * kSaState;
    o[0] = (double)M; o[1] = r[0]; o[2] = r[1]; o[3] = r[2]; o[4] = r[3];
  }
}

// one warp per (cube, channel): lanes take the splits, shared maximum first, then plain float64 sums
__global__ void softargmax_merge_kernel(const sp3d_softargmax_args a, int splits) {
  const int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);  // (cube, channel)
  const int lane = threadIdx.x & 31;
  if (i >= a.n_cubes * a.C) return;
  const int cube = i / a.C, c = i % a.C;
  const float* cen = a.centers + (int64_t)cube * a.center_stride;
  float* o = a.out + (int64_t)i * 3;
  if (a.check_flag && !(cen[3] >= 0.0f)) {
    if (lane < 3) o[lane] = 0.0f;
    return;
  }
  const double* ws = reinterpret_cast<const double*>(a.workspace);
  double M = -INFINITY;
  for (int sp = lane; sp < splits; sp += 32) {
    const double* p = ws + (((int64_t)cube * splits + sp) * a.C + c) * kSaState;
    if (p[1] > 0.0) M = fmax(M, p[0]);
  }
  for (int off = 16; off > 0; off >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, off));
  double r[4] = {0.0, 0.0, 0.0, 0.0};
  for (int sp = lane; sp < splits; sp += 32) {
    const double* p = ws + (((int64_t)cube * splits + sp) * a.C + c) * kSaState;
    if (p[1] > 0.0) {
      const double f = exp(p[0] - M);
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] += p[1 + k] * f;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    for (int off = 16; off > 0; off >>= 1) r[k] += __shfl_xor_sync(0xffffffffu, r[k], off);
  if (lane == 0) {
    o[0] = (float)((double)cen[0] + r[1] / r[0]);
    o[1] = (float)((double)cen[1] + r[2] / r[0]);
    o[2] = (float)((double)cen[2] + r[3] / r[0]);
  }
}

static int sa_splits(const sp3d_softargmax_args* a) {
  const int64_t N = (int64_t)a->X * a->Y * a->Z;
  int splits = (int)((148 * 8 + a->n_cubes - 1) / (a->n_cubes > 0 ? a->n_cubes : 1));
  const int max_splits = (int)((N + 1023) / 1024);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return splits;
}

}  // namespace sp3d

extern "C" int64_t sp3d_softargmax3d_workspace(const sp3d_softargmax_args* a) {
  if (a == nullptr || a->n_cubes < 0 || a->C < 1) return 0;
  return (int64_t)a->n_cubes * sp3d::sa_splits(a) * a->C * sp3d::kSaState * (int64_t)sizeof(double);
}

extern "C" int sp3d_softargmax3d_fwd(const sp3d_softargmax_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->x == nullptr || a->out == nullptr || a->centers == nullptr || a->lin_x == nullptr ||
      a->lin_y == nullptr || a->lin_z == nullptr || a->C < 1 || a->n_cubes < 0 || a->X < 1 || a->Y < 1 || a->Z < 1 ||
      a->center_stride < 3 || (a->check_flag && a->center_stride < 4))
    return SP3D_ERR_INVALID_ARG;
  if (a->x_dtype != SP3D_F32 && a->x_dtype != SP3D_BF16) return SP3D_ERR_UNSUPPORTED;
  if (a->n_cubes == 0) return SP3D_OK;
  if (a->n_cubes > 65535) return SP3D_ERR_INVALID_ARG;
  if (a->workspace == nullptr || a->workspace_bytes < sp3d_softargmax3d_workspace(a) ||
      (reinterpret_cast<uintptr_t>(a->workspace) % 8) != 0)
    return SP3D_ERR_WORKSPACE;
  const int splits = sa_splits(a);
  const int N = a->X * a->Y * a->Z;
  const int vps = (N + splits - 1) / splits;
  dim3 grid(splits, a->n_cubes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->x_dtype == SP3D_F32) {
    const bool vec = a->stride_c == 1 && (a->stride_vox % 4) == 0 && (a->stride_cube % 4) == 0 &&
                     (reinterpret_cast<uintptr_t>(a->x) % 16) == 0 && a->stride_vox >= ((a->C + 3) / 4) * 4;
    const bool stream = vec && a->C <= 16 && a->X <= 256 && a->Y <= 256 && a->Z <= 256;
    if (stream && a->C > 8) softargmax_stream_kernel<4><<<grid, kSaThreads, 0, st>>>(*a, splits, vps);
    else if (stream && a->C > 4) softargmax_stream_kernel<2><<<grid, kSaThreads, 0, st>>>(*a, splits, vps);
    else if (stream) softargmax_stream_kernel<1><<<grid, kSaThreads, 0, st>>>(*a, splits, vps);
    else if (vec) softargmax_partial_kernel<float, true><<<grid, kSaThreads, 0, st>>>(*a, splits, vps);
    else softargmax_partial_kernel<float, false><<<grid, kSaThreads, 0, st>>>(*a, splits, vps);
  } else {
    softargmax_partial_kernel<__nv_bfloat16, false><<<grid, kSaThreads, 0, st>>>(*a, splits, vps);
  }
  int rc = check_launch();
  if (rc != SP3D_OK) return rc;
  const int total = a->n_cubes * a->C;
  softargmax_merge_kernel<<<(total + 3) / 4, 128, 0, st>>>(*a, splits);
  return check_launch();
}
