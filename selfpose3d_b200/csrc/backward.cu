// Backward operators of the voxel pose path (float32): un-projection (adjoint of the bilinear gather), soft-argmax,
// max pooling, convolution weight gradient, training-mode BatchNorm (statistics / apply / backward).
//
// These are the first, correctness-first forms: one thread per voxel / window / element, float atomics for the
// scatters, double atomics for the per-channel reductions.  The input gradient of a convolution needs no kernel of
// its own: it is sp3d_conv_fwd on the flipped / transposed weight (selfpose3d_b200/ops.py, conv_dgrad).
//
// Reference semantics: what torch.autograd runs for lib/models/project_layer.py:93-99 (grid_sample + masked mean +
// clamp), lib/models/pose_regression_net.py:22-27 (softmax-weighted sum), F.max_pool3d / nn.MaxPool2d,
// nn.Conv{2,3}d / nn.ConvTranspose{2,3}d (backward-filter) and nn.BatchNorm{2,3}d in training mode.
#include "sp3d_common.cuh"
#include "unproject_geom.cuh"
#include <math.h>

namespace sp3d {

// ------------------------------------------------------------------------------------------ un-projection
constexpr int kUbThreads = 256;
constexpr int kUbGroup = 16;

__global__ void __launch_bounds__(kUbThreads) unproject_bwd_kernel(const sp3d_unproject_bwd_args b) {
  const sp3d_unproject_args& a = b.fwd;
  __shared__ float s_cam[SP3D_MAX_VIEWS * SP3D_CAM_FLOATS];
  __shared__ float s_center[4];
  const int cube = blockIdx.y;
  const int sample = a.cube_sample ? a.cube_sample[cube] : cube / a.cubes_per_sample;
  const int N = a.X * a.Y * a.Z;
  const int tid = threadIdx.x;
  for (int i = tid; i < a.V * SP3D_CAM_FLOATS; i += kUbThreads)
    s_cam[i] = a.cams[(int64_t)sample * a.V * SP3D_CAM_FLOATS + i];
  if (tid < 4) s_center[tid] = (tid < 3 || a.center_stride > 3) ? a.centers[(int64_t)cube * a.center_stride + tid] : 0.0f;
  __syncthreads();
  const int vox = blockIdx.x * kUbThreads + tid;
  if (vox >= N) return;
  if (a.check_flag && !(s_center[3] >= 0.0f)) return;   // skipped cubes are constant zeros: no gradient
  const int iz = vox % a.Z;
  const int iy = (vox / a.Z) % a.Y;
  const int ix = vox / (a.Z * a.Y);
  const float gx = __fadd_rn(a.lin_x[ix], s_center[0]);
  const float gy = __fadd_rn(a.lin_y[iy], s_center[1]);
  const float gz = __fadd_rn(a.lin_z[iz], s_center[2]);
  const float hm_w = (float)a.w, hm_h = (float)a.h;
  const float* go = b.grad_cubes + (int64_t)cube * a.out_stride_cube + (int64_t)vox * a.out_stride_vox;

  for (int c0 = 0; c0 < a.C; c0 += kUbGroup) {
    const int cn = min(kUbGroup, a.C - c0);
    // pass 1: the forward value (same operation order as unproject_kernel) -> clamp gate and denominator
    float num[kUbGroup];
#pragma unroll
    for (int j = 0; j < kUbGroup; ++j) num[j] = 0.0f;
    float den = 0.0f;
    for (int v = a.view_begin; v < a.view_end; ++v) {
      const ViewSample s = project_view(s_cam + v * SP3D_CAM_FLOATS, gx, gy, gz, a.img_w, a.img_h, a.hm_cfg_w,
                                          a.hm_cfg_h, hm_w, hm_h);
      den = __fadd_rn(den, s.m);
      if (s.m == 0.0f) continue;
      const float* hm = a.heatmaps[v] + (int64_t)sample * a.hm_stride_b + (int64_t)c0 * a.hm_stride_c;
      float acc[kUbGroup];
#pragma unroll
      for (int j = 0; j < kUbGroup; ++j) acc[j] = 0.0f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xi = s.x0 + (t & 1), yi = s.y0 + (t >> 1);
        const float wgt = __fmul_rn((t & 1) ? s.wx1 : s.wx0, (t >> 1) ? s.wy1 : s.wy0);
        if (xi < 0 || xi >= a.w || yi < 0 || yi >= a.h) continue;
        const float* p = hm + (int64_t)yi * a.hm_stride_h + (int64_t)xi * a.hm_stride_w;
#pragma unroll
        for (int j = 0; j < kUbGroup; ++j)
          if (j < cn) acc[j] = __fadd_rn(acc[j], __fmul_rn(__ldg(p + (int64_t)j * a.hm_stride_c), wgt));
      }
#pragma unroll
      for (int j = 0; j < kUbGroup; ++j) num[j] = __fadd_rn(num[j], acc[j]);
    }
    const float d = __fadd_rn(den, 1e-6f);
    float coef[kUbGroup];
    bool any = false;
#pragma unroll
    for (int j = 0; j < kUbGroup; ++j) {
      coef[j] = 0.0f;
      if (j < cn) {
        const float r = __fdiv_rn(num[j], d);
        const bool gate = (r == r) && r >= 0.0f && r <= 1.0f;   // NaN -> 0 assignment and clamp(0, 1) pass no gradient outside
        if (gate) coef[j] = __fdiv_rn(go[(int64_t)(c0 + j) * a.out_stride_c], d);
        any = any || coef[j] != 0.0f;
      }
    }
    if (!any) continue;
    // pass 2: scatter coef * bilinear weight into the taps of every view that sees the voxel
    for (int v = a.view_begin; v < a.view_end; ++v) {
      const ViewSample s = project_view(s_cam + v * SP3D_CAM_FLOATS, gx, gy, gz, a.img_w, a.img_h, a.hm_cfg_w,
                                          a.hm_cfg_h, hm_w, hm_h);
      if (s.m == 0.0f) continue;
      float* ghm = b.grad_heatmaps[v] + (int64_t)sample * a.hm_stride_b + (int64_t)c0 * a.hm_stride_c;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xi = s.x0 + (t & 1), yi = s.y0 + (t >> 1);
        const float wgt = __fmul_rn((t & 1) ? s.wx1 : s.wx0, (t >> 1) ? s.wy1 : s.wy0);
        if (xi < 0 || xi >= a.w || yi < 0 || yi >= a.h) continue;
        float* p = ghm + (int64_t)yi * a.hm_stride_h + (int64_t)xi * a.hm_stride_w;
#pragma unroll
        for (int j = 0; j < kUbGroup; ++j)
          if (j < cn && coef[j] != 0.0f) atomicAdd(p + (int64_t)j * a.hm_stride_c, __fmul_rn(coef[j], wgt));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ block reductions
__device__ __forceinline__ float block_max(float v, float* s_red) {
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  float r = s_red[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, s_red[i]);
  return r;
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < nw; ++i) r += s_red[i];
  return r;
}

// ------------------------------------------------------------------------------------------ soft-argmax
// grid = (C, n_cubes): one CTA per (cube, channel) makes three passes over the channel's voxels (max, sum, gradient);
// a 64^3 x 16-channel float32 cube is 16.8 MB, so passes two and three are served by L2.
__global__ void __launch_bounds__(256) softargmax_bwd_kernel(const sp3d_softargmax_bwd_args b) {
  const sp3d_softargmax_args& a = b.fwd;
  __shared__ float s_redf[8];
  __shared__ double s_redd[8];
  const int c = blockIdx.x, cube = blockIdx.y;
  const int N = a.X * a.Y * a.Z;
  const float* x = reinterpret_cast<const float*>(a.x) + (int64_t)cube * a.stride_cube + (int64_t)c * a.stride_c;
  float* gxp = b.grad_x + (int64_t)cube * a.stride_cube + (int64_t)c * a.stride_c;
  const float* cen = a.centers + (int64_t)cube * a.center_stride;
  if (a.check_flag && !(cen[3] >= 0.0f)) {
    for (int v = threadIdx.x; v < N; v += blockDim.x) gxp[(int64_t)v * a.stride_vox] = 0.0f;
    return;
  }
  float m = -INFINITY;
  for (int v = threadIdx.x; v < N; v += blockDim.x) m = fmaxf(m, __fmul_rn(a.beta, __ldg(x + (int64_t)v * a.stride_vox)));
  m = block_max(m, s_redf);
  double s = 0.0;
  for (int v = threadIdx.x; v < N; v += blockDim.x)
    s += (double)expf(__fsub_rn(__fmul_rn(a.beta, __ldg(x + (int64_t)v * a.stride_vox)), m));
  s = block_sum(s, s_redd);
  const float inv_s = (float)(1.0 / s);
  const float* o = a.out + ((int64_t)cube * a.C + c) * 3;
  const float* go = b.grad_out + ((int64_t)cube * a.C + c) * 3;
  const float ox = o[0], oy = o[1], oz = o[2], g0 = go[0], g1 = go[1], g2 = go[2];
  const float cx = cen[0], cy = cen[1], cz = cen[2];
  for (int v = threadIdx.x; v < N; v += blockDim.x) {
    const int iz = v % a.Z, iy = (v / a.Z) % a.Y, ix = v / (a.Z * a.Y);
    const float px = __fadd_rn(a.lin_x[ix], cx), py = __fadd_rn(a.lin_y[iy], cy), pz = __fadd_rn(a.lin_z[iz], cz);
    const float p = expf(__fsub_rn(__fmul_rn(a.beta, __ldg(x + (int64_t)v * a.stride_vox)), m)) * inv_s;
    const float dot = (px - ox) * g0 + (py - oy) * g1 + (pz - oz) * g2;
    gxp[(int64_t)v * a.stride_vox] = a.beta * p * dot;
  }
}

// Streaming form for channel-last volumes (stride_c = 1, a pitch that is a multiple of 4): all channels of a voxel are
// one or a few float4; grid = (splits, cubes).  Pass 1 keeps a running (max, sum) per channel and thread, merges the
// CTA's threads and writes one partial per (cube, split, channel); pass 2 merges the partials (a few dozen values) and
// streams the volume once more: dx = beta p (g - out) . dout.  The one-CTA-per-(cube, channel) form above re-reads
// 4-byte elements at 64-byte stride three times (353 GB/s in the training step).
constexpr int kSabThreads = 256;
constexpr int kSabMaxSplits = 64;

__global__ void __launch_bounds__(kSabThreads) softargmax_bwd_stats_kernel(const sp3d_softargmax_bwd_args b, int c4n, int splits,
                                                                           float* __restrict__ ws) {
  const sp3d_softargmax_args& a = b.fwd;
  __shared__ float s_m[kSabThreads][4];
  __shared__ double s_s[kSabThreads][4];
  const int split = blockIdx.x, cube = blockIdx.y, tid = threadIdx.x;
  const int N = a.X * a.Y * a.Z;
  const int rows = kSabThreads / c4n;                   // voxels per CTA step
  const int col = tid % c4n, prow = tid / c4n;
  const float* x = reinterpret_cast<const float*>(a.x) + (int64_t)cube * a.stride_cube;
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  if (prow < rows) {
    for (int v = split * rows + prow; v < N; v += splits * rows) {
      const float4 q = ldg4(x + (int64_t)v * a.stride_vox + 4 * col);
      const float z[4] = {__fmul_rn(a.beta, q.x), __fmul_rn(a.beta, q.y), __fmul_rn(a.beta, q.z), __fmul_rn(a.beta, q.w)};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (z[j] > m[j]) {                               // new maximum: rescale the running sum
          s[j] = s[j] * (double)expf(m[j] - z[j]) + 1.0;
          m[j] = z[j];
        } else {
          s[j] += (double)expf(z[j] - m[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { s_m[tid][j] = m[j]; s_s[tid][j] = s[j]; }
  __syncthreads();
  if (tid < c4n) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float M = -INFINITY;
      for (int r = 0; r < rows; ++r) M = fmaxf(M, s_m[r * c4n + tid][j]);
      double S = 0.0;
      for (int r = 0; r < rows; ++r) {
        const float mr = s_m[r * c4n + tid][j];
        if (mr > -INFINITY) S += s_s[r * c4n + tid][j] * (double)expf(mr - M);
      }
      const int c = 4 * tid + j;
      if (c < a.C) {
        float* o = ws + (((int64_t)cube * splits + split) * a.C + c) * 2;
        o[0] = M;
        o[1] = (float)S;      // (a partial sum of at most N terms <= 1: float is exact enough for the merge below)
      }
    }
  }
}

__global__ void __launch_bounds__(kSabThreads) softargmax_bwd_apply_kernel(const sp3d_softargmax_bwd_args b, int c4n, int splits,
                                                                           int stat_splits, const float* __restrict__ ws) {
  const sp3d_softargmax_args& a = b.fwd;
  const int split = blockIdx.x, cube = blockIdx.y, tid = threadIdx.x;
  const int N = a.X * a.Y * a.Z;
  const int rows = kSabThreads / c4n;
  const int col = tid % c4n, prow = tid / c4n;
  if (prow >= rows) return;
  const float* x = reinterpret_cast<const float*>(a.x) + (int64_t)cube * a.stride_cube;
  float* gx = b.grad_x + (int64_t)cube * a.stride_cube;
  const float* cen = a.centers + (int64_t)cube * a.center_stride;
  const bool skip = a.check_flag && !(cen[3] >= 0.0f);
  float M[4], inv_s[4], o3[4][3], g3[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * col + j;
    M[j] = 0.0f; inv_s[j] = 0.0f;
#pragma unroll
    for (int d = 0; d < 3; ++d) { o3[j][d] = 0.0f; g3[j][d] = 0.0f; }
    if (c < a.C && !skip) {
      const float* w = ws + ((int64_t)cube * stat_splits * a.C + c) * 2;
      float mm = -INFINITY;
      for (int i = 0; i < stat_splits; ++i) mm = fmaxf(mm, w[(int64_t)i * a.C * 2]);
      double S = 0.0;
      for (int i = 0; i < stat_splits; ++i) {
        const float mi = w[(int64_t)i * a.C * 2];
        if (mi > -INFINITY) S += (double)w[(int64_t)i * a.C * 2 + 1] * (double)expf(mi - mm);
      }
      M[j] = mm;
      inv_s[j] = (float)(1.0 / S);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        o3[j][d] = a.out[((int64_t)cube * a.C + c) * 3 + d];
        g3[j][d] = b.grad_out[((int64_t)cube * a.C + c) * 3 + d];
      }
    }
  }
  const float cx = cen[0], cy = cen[1], cz = cen[2];
  for (int v = split * rows + prow; v < N; v += splits * rows) {
    float r[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (!skip) {
      const int iz = v % a.Z, iy = (v / a.Z) % a.Y, ix = v / (a.Z * a.Y);
      const float px = __fadd_rn(a.lin_x[ix], cx), py = __fadd_rn(a.lin_y[iy], cy), pz = __fadd_rn(a.lin_z[iz], cz);
      const float4 q = ldg4(x + (int64_t)v * a.stride_vox + 4 * col);
      const float xs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (4 * col + j < a.C) {
          const float p = expf(__fsub_rn(__fmul_rn(a.beta, xs[j]), M[j])) * inv_s[j];
          const float dot = (px - o3[j][0]) * g3[j][0] + (py - o3[j][1]) * g3[j][1] + (pz - o3[j][2]) * g3[j][2];
          r[j] = a.beta * p * dot;
        }
      }
    }
    *reinterpret_cast<float4*>(gx + (int64_t)v * a.stride_vox + 4 * col) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// ------------------------------------------------------------------------------------------ max pool
// one thread per (window, 4 channels): the first maximum in (d, h, w) scan order receives the window's gradient
__global__ void maxpool_bwd_kernel(const sp3d_maxpool_bwd_args b) {
  const sp3d_maxpool_args& a = b.fwd;
  const int cvec = a.c_pitch / 4;
  const int64_t total = (int64_t)a.N * a.OD * a.OH * a.OW * cvec;
  const float* in = reinterpret_cast<const float*>(a.in);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t pos = i / cvec;
    const int ow = (int)(pos % a.OW); pos /= a.OW;
    const int oh = (int)(pos % a.OH); pos /= a.OH;
    const int od = (int)(pos % a.OD);
    const int n = (int)(pos / a.OD);
    float mv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int64_t mi[4] = {-1, -1, -1, -1};
    for (int kd = 0; kd < a.k[0]; ++kd) {
      const int id = od * a.s[0] - a.p[0] + kd;
      if (id < 0 || id >= a.D) continue;
      for (int kh = 0; kh < a.k[1]; ++kh) {
        const int ih = oh * a.s[1] - a.p[1] + kh;
        if (ih < 0 || ih >= a.H) continue;
        for (int kw = 0; kw < a.k[2]; ++kw) {
          const int iw = ow * a.s[2] - a.p[2] + kw;
          if (iw < 0 || iw >= a.W) continue;
          const int64_t off = ((((int64_t)n * a.D + id) * a.H + ih) * a.W + iw) * a.c_pitch + cv * 4;
          const float4 q = ldg4(in + off);
          const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (qv[j] > mv[j] || mi[j] < 0) { mv[j] = qv[j]; mi[j] = off + j; }
        }
      }
    }
    const float4 g = ldg4(b.grad_out + ((((int64_t)n * a.OD + od) * a.OH + oh) * a.OW + ow) * a.c_pitch + cv * 4);
    const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (mi[j] >= 0 && cv * 4 + j < a.C) atomicAdd(b.grad_in + mi[j], gv[j]);
  }
}

// ------------------------------------------------------------------------------------------ convolution wgrad
// CTA = (chunk of output positions, group of TL taps x CT input channels = 64 GEMM rows, tile of 64 output channels);
// 256 threads each own a 4 x 4 block of grad_weight, positions stream through shared memory 32 at a time.
constexpr int kWgThreads = 256, kWgPos = 32, kWgTile = 64;

__global__ void __launch_bounds__(kWgThreads) conv_wgrad_kernel(const sp3d_conv_wgrad_args b, int CT, int TL, int ppc) {
  const sp3d_conv_args& a = b.fwd;
  __shared__ __align__(16) float sx[kWgPos][kWgTile];
  __shared__ __align__(16) float sdy[kWgPos][kWgTile];
  __shared__ int s_pos[kWgPos][4];   // n, od, oh, ow of the staged positions (n = -1: past the end)
  const int ntaps = a.ksize[0] * a.ksize[1] * a.ksize[2];
  const int ci_tiles = (a.cin + CT - 1) / CT;
  const int tap0 = ((int)blockIdx.y / ci_tiles) * TL, ci0 = ((int)blockIdx.y % ci_tiles) * CT, co0 = (int)blockIdx.z * kWgTile;
  const int64_t P = (int64_t)a.N * a.OD * a.OH * a.OW;
  const int64_t p_begin = (int64_t)blockIdx.x * ppc;
  const int64_t p_end = p_begin + ppc < P ? p_begin + ppc : P;
  const int tid = threadIdx.x, ri = tid >> 4, cj = tid & 15;
  const bool do_bias = b.grad_bias != nullptr && blockIdx.y == 0 && ri == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bacc[4] = {0.0f, 0.0f, 0.0f, 0.0f};

  for (int64_t p0 = p_begin; p0 < p_end; p0 += kWgPos) {
    if (tid < kWgPos) {
      int64_t p = p0 + tid;
      int n = -1, od = 0, oh = 0, ow = 0;
      if (p < p_end) {
        ow = (int)(p % a.OW); p /= a.OW;
        oh = (int)(p % a.OH); p /= a.OH;
        od = (int)(p % a.OD);
        n = (int)(p / a.OD);
      }
      s_pos[tid][0] = n; s_pos[tid][1] = od; s_pos[tid][2] = oh; s_pos[tid][3] = ow;
    }
    __syncthreads();
    for (int e = tid; e < kWgPos * kWgTile; e += kWgThreads) {
      const int pl = e / kWgTile, c = e % kWgTile;
      const int n = s_pos[pl][0], od = s_pos[pl][1], oh = s_pos[pl][2], ow = s_pos[pl][3];
      // grad_out column c of position pl
      float v = 0.0f;
      if (n >= 0 && co0 + c < a.cout) {
        const int64_t opos = (((int64_t)n * a.TD + (od * a.ostride[0] + a.ooffset[0])) * a.TH + (oh * a.ostride[1] + a.ooffset[1])) *
                                 a.TW + (ow * a.ostride[2] + a.ooffset[2]);
        v = __ldg(b.grad_out + opos * a.cout_pitch + co0 + c);
      }
      sdy[pl][c] = v;
      // input row r = (tap within the group, channel) of position pl
      const int tl = c / CT, ci = ci0 + c % CT, tap = tap0 + tl;
      float xv = 0.0f;
      if (n >= 0 && tl < TL && tap < ntaps && ci < a.cin) {
        const int tw = tap % a.ksize[2], th = (tap / a.ksize[2]) % a.ksize[1], td = tap / (a.ksize[2] * a.ksize[1]);
        const int id = od * a.stride[0] + a.tap_off0[0] + td * a.tap_step[0];
        const int ih = oh * a.stride[1] + a.tap_off0[1] + th * a.tap_step[1];
        const int iw = ow * a.stride[2] + a.tap_off0[2] + tw * a.tap_step[2];
        if (id >= 0 && id < a.D && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
          xv = __ldg(reinterpret_cast<const float*>(a.in) + ((((int64_t)n * a.D + id) * a.H + ih) * a.W + iw) * a.cin_pitch + ci);
      }
      sx[pl][c] = xv;
    }
    __syncthreads();
#pragma unroll 8
    for (int pl = 0; pl < kWgPos; ++pl) {
      const float4 xv = *reinterpret_cast<const float4*>(&sx[pl][4 * ri]);
      const float4 dv = *reinterpret_cast<const float4*>(&sdy[pl][4 * cj]);
      const float xr[4] = {xv.x, xv.y, xv.z, xv.w}, dr[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], dr[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bacc[j] += dr[j];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 4 * ri + i, tl = r / CT, ci = ci0 + r % CT, tap = tap0 + tl;
    if (tl >= TL || tap >= ntaps || ci >= a.cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + 4 * cj + j;
      if (co < a.cout && acc[i][j] != 0.0f) atomicAdd(b.grad_weight + ((int64_t)tap * a.cin + ci) * a.cout_pitch_w + co, acc[i][j]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + 4 * cj + j;
      if (co < a.cout) atomicAdd(b.grad_bias + co, bacc[j]);
    }
  }
}

__global__ void relu_bwd_kernel(const sp3d_relu_bwd_args a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x)
    a.grad_x[i] = __ldg(a.y + i) > 0.0f ? __ldg(a.grad_y + i) : 0.0f;
}

// ------------------------------------------------------------------------------------------ Gaussian joint rendering
// grid = (J, V * B): one CTA per (view, sample, joint) heat-map; the people's joint pixels sit in shared memory.
constexpr int kGrMaxPeople = 32;

__device__ __forceinline__ float gauss_term(float xx, float yy, float kx, float ky, float inv_sigma) {
  const float a = (xx - kx) * inv_sigma, b = (yy - ky) * inv_sigma;
  return expf(-(a * a) * 0.5f - (b * b) * 0.5f);
}

__global__ void __launch_bounds__(256) gauss_render_fwd_kernel(const sp3d_gauss_render_args a) {
  __shared__ float s_k[kGrMaxPeople][2];
  const int j = blockIdx.x, item = blockIdx.y, b = item % a.B;
  const int n = min(a.n_people[b], a.P);
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const float* k = a.kps + (((int64_t)item * a.P + p) * a.J + j) * 2;
    s_k[p][0] = k[0] * a.inv_scale;
    s_k[p][1] = k[1] * a.inv_scale;
  }
  __syncthreads();
  const float inv_sigma = 1.0f / a.sigma;
  float* out = a.heatmaps + ((int64_t)item * a.J + j) * a.h * a.w;
  for (int i = threadIdx.x; i < a.h * a.w; i += blockDim.x) {
    const float xx = (float)(i % a.w), yy = (float)(i / a.w);
    float s = 0.0f;
    for (int p = 0; p < n; ++p) s += gauss_term(xx, yy, s_k[p][0], s_k[p][1], inv_sigma);
    out[i] = fminf(fmaxf(s, 0.0f), 1.0f);
  }
}

__global__ void __launch_bounds__(256) gauss_render_bwd_kernel(const sp3d_gauss_render_bwd_args bw) {
  const sp3d_gauss_render_args& a = bw.fwd;
  __shared__ float s_k[kGrMaxPeople][2];
  __shared__ double s_red[8];
  const int j = blockIdx.x, item = blockIdx.y, b = item % a.B;
  const int n = min(a.n_people[b], a.P);
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const float* k = a.kps + (((int64_t)item * a.P + p) * a.J + j) * 2;
    s_k[p][0] = k[0] * a.inv_scale;
    s_k[p][1] = k[1] * a.inv_scale;
  }
  __syncthreads();
  const float inv_sigma = 1.0f / a.sigma;
  const float* g = bw.grad_heatmaps + ((int64_t)item * a.J + j) * a.h * a.w;
  for (int p = 0; p < a.P; ++p) {
    double gx = 0.0, gy = 0.0;
    if (p < n) {
      for (int i = threadIdx.x; i < a.h * a.w; i += blockDim.x) {
        const float xx = (float)(i % a.w), yy = (float)(i / a.w);
        float s = 0.0f;
        for (int q = 0; q < n; ++q) s += gauss_term(xx, yy, s_k[q][0], s_k[q][1], inv_sigma);
        if (s > 1.0f) continue;                       // clip(., 0, 1): no gradient above 1 (the sum is never below 0)
        const float e = gauss_term(xx, yy, s_k[p][0], s_k[p][1], inv_sigma) * __ldg(g + i);
        // d/dk of exp(-((x - k s)/sigma)^2 / 2) = e * (x - k s)/sigma * s/sigma with s = inv_scale
        gx += (double)(e * (xx - s_k[p][0]));
        gy += (double)(e * (yy - s_k[p][1]));
      }
    }
    gx = block_sum(gx, s_red);
    gy = block_sum(gy, s_red);
    if (threadIdx.x == 0) {
      float* o = bw.grad_kps + (((int64_t)item * a.P + p) * a.J + j) * 2;
      const float c = inv_sigma * inv_sigma * a.inv_scale;
      o[0] = (float)gx * c;
      o[1] = (float)gy * c;
    }
  }
}

static int grid_for(int64_t total, int threads, int cap) {
  const int64_t want = (total + threads - 1) / threads;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace sp3d

using namespace sp3d;

extern "C" int sp3d_unproject_bwd(const sp3d_unproject_bwd_args* b, void* stream) {
  if (b == nullptr || b->fwd.n_cubes < 0) return SP3D_ERR_INVALID_ARG;
  const sp3d_unproject_args* a = &b->fwd;
  if (a->n_cubes == 0) return SP3D_OK;
  if (a->V < 1 || a->V > SP3D_MAX_VIEWS || a->C < 1 || b->grad_cubes == nullptr || a->cams == nullptr ||
      a->centers == nullptr || a->lin_x == nullptr || a->lin_y == nullptr || a->lin_z == nullptr || a->center_stride < 3 ||
      a->cubes_per_sample < 1 || a->view_begin < 0 || a->view_end > a->V || a->view_begin > a->view_end ||
      (a->check_flag && a->center_stride < 4))
    return SP3D_ERR_INVALID_ARG;
  for (int v = a->view_begin; v < a->view_end; ++v)
    if (a->heatmaps[v] == nullptr || b->grad_heatmaps[v] == nullptr) return SP3D_ERR_INVALID_ARG;
  if (a->math_mode != 0 || a->hm_dtype != SP3D_F32 || a->partial) return SP3D_ERR_UNSUPPORTED;
  const int N = a->X * a->Y * a->Z;
  if (N <= 0 || a->n_cubes > 65535) return SP3D_ERR_INVALID_ARG;
  dim3 grid(ceil_div(N, kUbThreads), a->n_cubes);
  unproject_bwd_kernel<<<grid, kUbThreads, 0, static_cast<cudaStream_t>(stream)>>>(*b);
  return check_launch();
}

extern "C" int sp3d_softargmax3d_bwd(const sp3d_softargmax_bwd_args* b, void* stream) {
  if (b == nullptr || b->fwd.n_cubes < 0 || b->fwd.C < 1) return SP3D_ERR_INVALID_ARG;
  const sp3d_softargmax_args* a = &b->fwd;
  if (a->n_cubes == 0) return SP3D_OK;
  if (a->x == nullptr || a->out == nullptr || b->grad_out == nullptr || b->grad_x == nullptr || a->centers == nullptr ||
      a->lin_x == nullptr || a->lin_y == nullptr || a->lin_z == nullptr || a->X < 1 || a->Y < 1 || a->Z < 1 ||
      a->center_stride < 3 || (a->check_flag && a->center_stride < 4) || a->n_cubes > 65535)
    return SP3D_ERR_INVALID_ARG;
  if (a->x_dtype != SP3D_F32) return SP3D_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // channel-last volumes with a workspace: the streaming form (2 launches, every byte read twice as float4)
  const int64_t N = (int64_t)a->X * a->Y * a->Z;
  const int c4n = (int)(a->stride_vox / 4);
  int splits = (int)((N + 4095) / 4096);                                   // ~4096 voxels per CTA ...
  if (splits * a->n_cubes < 148 && N >= 1024) splits = (148 + a->n_cubes - 1) / a->n_cubes;   // ... but fill the GPU
  if (splits > kSabMaxSplits) splits = kSabMaxSplits;
  if (splits < 1) splits = 1;
  const int64_t need = (int64_t)a->n_cubes * splits * a->C * 2 * (int64_t)sizeof(float);
  if (a->stride_c == 1 && a->stride_vox % 4 == 0 && c4n >= 1 && c4n <= 64 && a->stride_vox >= a->C && a->stride_cube % 4 == 0 &&
      reinterpret_cast<uintptr_t>(a->x) % 16 == 0 && reinterpret_cast<uintptr_t>(b->grad_x) % 16 == 0 &&
      a->workspace != nullptr && a->workspace_bytes >= need) {
    softargmax_bwd_stats_kernel<<<dim3(splits, a->n_cubes), kSabThreads, 0, st>>>(*b, c4n, splits, a->workspace);
    int rc = check_launch();
    if (rc != SP3D_OK) return rc;
    int asplits = (int)((N + 2047) / 2048);
    if (asplits > 256) asplits = 256;
    softargmax_bwd_apply_kernel<<<dim3(asplits, a->n_cubes), kSabThreads, 0, st>>>(*b, c4n, asplits, splits, a->workspace);
    return check_launch();
  }
  dim3 grid(a->C, a->n_cubes);
  softargmax_bwd_kernel<<<grid, 256, 0, st>>>(*b);
  return check_launch();
}

extern "C" int sp3d_maxpool_bwd(const sp3d_maxpool_bwd_args* b, void* stream) {
  if (b == nullptr) return SP3D_ERR_INVALID_ARG;
  const sp3d_maxpool_args* a = &b->fwd;
  if (a->in == nullptr || b->grad_out == nullptr || b->grad_in == nullptr || a->N < 0 || a->C < 1 || a->c_pitch < a->C ||
      (a->c_pitch % 4))
    return SP3D_ERR_INVALID_ARG;
  if (a->dtype != SP3D_F32) return SP3D_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t in_elems = (int64_t)a->N * a->D * a->H * a->W * a->c_pitch;
  if (in_elems > 0) {
    cudaError_t e = cudaMemsetAsync(b->grad_in, 0, (size_t)in_elems * sizeof(float), st);
    if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
  }
  const int64_t total = (int64_t)a->N * a->OD * a->OH * a->OW * (a->c_pitch / 4);
  if (total == 0) return SP3D_OK;
  maxpool_bwd_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(*b);
  return check_launch();
}

extern "C" int sp3d_conv_wgrad(const sp3d_conv_wgrad_args* b, void* stream) {
  if (b == nullptr) return SP3D_ERR_INVALID_ARG;
  const sp3d_conv_args* a = &b->fwd;
  if (a->in == nullptr || b->grad_out == nullptr || b->grad_weight == nullptr || a->N < 0 || a->cin < 1 || a->cout < 1 ||
      a->cin_pitch < a->cin || a->cout_pitch < a->cout || a->cout_pitch_w < a->cout || a->OD < 0 || a->OH < 0 || a->OW < 0)
    return SP3D_ERR_INVALID_ARG;
  for (int d = 0; d < 3; ++d)
    if (a->ksize[d] < 1 || a->stride[d] < 1 || a->ostride[d] < 1 || a->ooffset[d] < 0) return SP3D_ERR_INVALID_ARG;
  if (a->in_dtype != SP3D_F32) return SP3D_ERR_UNSUPPORTED;
  const int64_t P = (int64_t)a->N * a->OD * a->OH * a->OW;
  if (P == 0) return SP3D_OK;
  int CT = 64;
  if (a->cin < 64) {
    CT = 4;
    while (CT < a->cin) CT *= 2;
  }
  const int ntaps = a->ksize[0] * a->ksize[1] * a->ksize[2];
  int TL = kWgTile / CT;
  if (TL > ntaps) TL = ntaps;
  const int ci_tiles = (a->cin + CT - 1) / CT;
  // positions per CTA: enough CTAs to fill the GPU, few enough that the atomics stay a small share
  int ppc = 2048;
  while (ppc > 128 && (P + ppc - 1) / ppc * ((ntaps + TL - 1) / TL) * ci_tiles < 4 * 148) ppc /= 2;
  const int64_t gx = (P + ppc - 1) / ppc;
  const int64_t gy = (int64_t)((ntaps + TL - 1) / TL) * ci_tiles;
  const int gz = (a->cout + kWgTile - 1) / kWgTile;
  if (gx > 2147483647LL || gy > 65535 || gz > 65535) return SP3D_ERR_UNSUPPORTED;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  conv_wgrad_kernel<<<grid, kWgThreads, 0, static_cast<cudaStream_t>(stream)>>>(*b, CT, TL, ppc);
  return check_launch();
}

extern "C" int sp3d_relu_bwd(const sp3d_relu_bwd_args* a, void* stream) {
  if (a == nullptr || a->n < 0 || (a->n > 0 && (a->grad_y == nullptr || a->y == nullptr || a->grad_x == nullptr)))
    return SP3D_ERR_INVALID_ARG;
  if (a->n == 0) return SP3D_OK;
  relu_bwd_kernel<<<grid_for(a->n, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}

static int gauss_render_check(const sp3d_gauss_render_args* a) {
  if (a == nullptr || a->kps == nullptr || a->n_people == nullptr || a->V < 1 || a->B < 1 || a->P < 0 || a->J < 1 ||
      a->h < 1 || a->w < 1 || !(a->sigma > 0.0f) || !(a->inv_scale > 0.0f))
    return SP3D_ERR_INVALID_ARG;
  if (a->P > kGrMaxPeople || a->J > 65535 || (int64_t)a->V * a->B > 65535) return SP3D_ERR_UNSUPPORTED;
  return SP3D_OK;
}

extern "C" int sp3d_gauss_render_fwd(const sp3d_gauss_render_args* a, void* stream) {
  int rc = gauss_render_check(a);
  if (rc != SP3D_OK) return rc;
  if (a->heatmaps == nullptr) return SP3D_ERR_INVALID_ARG;
  gauss_render_fwd_kernel<<<dim3(a->J, a->V * a->B), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}

extern "C" int sp3d_gauss_render_bwd(const sp3d_gauss_render_bwd_args* b, void* stream) {
  if (b == nullptr) return SP3D_ERR_INVALID_ARG;
  int rc = gauss_render_check(&b->fwd);
  if (rc != SP3D_OK) return rc;
  if (b->grad_heatmaps == nullptr || b->grad_kps == nullptr) return SP3D_ERR_INVALID_ARG;
  if (b->fwd.P == 0) return SP3D_OK;
  gauss_render_bwd_kernel<<<dim3(b->fwd.J, b->fwd.V * b->fwd.B), 256, 0, static_cast<cudaStream_t>(stream)>>>(*b);
  return check_launch();
}
