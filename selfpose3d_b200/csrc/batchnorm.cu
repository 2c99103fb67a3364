// Training-mode BatchNorm on channel-last float32 activations: batch statistics, apply (+ residual / ReLU), backward.
//
// x is [n_items][item_positions][pitch] (pitch a multiple of 4).  With groups (item_group / group_items given) every
// item belongs to one of n_groups statistic groups -- the pose net runs ALL proposal slots of a training step in one
// launch set, while the reference calls it once per slot (lib/models/multi_person_posenet.py:88-99,
// multi_person_posenet_ssv.py:330-407), i.e. normalises every slot's cubes with that slot's own batch statistics.
// mean / var / scale / shift are then [n_groups][C]; grad_gamma / grad_beta are summed over the groups (shared
// parameters).  Without groups: one item of P positions.
//
// HBM-bound passes: float4 loads with several independent loads in flight per thread; per-thread double accumulators
// (the variance E[x^2] - mean^2 is cancellation-prone), shared-memory then global double atomics per CTA.
//
// Reference semantics: F.batch_norm(training=True) of nn.BatchNorm{2,3}d and its autograd
// (lib/models/v2v_net.py:14,27,30, lib/models/pose_resnet.py:49-...).
#include "sp3d_common.cuh"
#include <math.h>

namespace sp3d {

constexpr int kBnThreads = 256;
constexpr int kBnUnroll = 4;

struct BnGeom {
  int c4n;                 // float4 columns per position (pitch / 4)
  int64_t item_positions;
  const int32_t* item_group;    // or NULL
  const int32_t* group_items;   // items per group (or NULL: one group of n_items)
  int n_items, n_groups, C;
};

__device__ __forceinline__ int bn_group_of(const BnGeom& g, int item) { return g.item_group ? g.item_group[item] : 0; }
__device__ __forceinline__ double bn_group_positions(const BnGeom& g, int grp) {
  return (double)(g.group_items ? g.group_items[grp] : g.n_items) * (double)g.item_positions;
}

// grid = (chunks of positions, items).  MODE 0: sum x, sum x^2.  MODE 1: sum dz, sum dz * xhat (dz = grad_y masked by
// y > 0 when y is given).  ws: [n_groups][2][C] doubles, zeroed by the caller.
template <int MODE>
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               const float* __restrict__ y, const float* __restrict__ mean,
                                                               const float* __restrict__ var, float eps, BnGeom g,
                                                               double* __restrict__ ws) {
  __shared__ double s_acc[kBnThreads][8];
  const int item = blockIdx.y;
  const int grp = bn_group_of(g, item);
  const int tid = threadIdx.x;
  const int cw = g.c4n < kBnThreads ? g.c4n : kBnThreads;      // float4 columns handled per pass
  const int rows = kBnThreads / cw;                            // positions per CTA step
  const int col_in = tid % cw, prow = tid / cw;
  const bool active = prow < rows;
  const int64_t base = (int64_t)item * g.item_positions;
  for (int cb = 0; cb < g.c4n; cb += cw) {
    const int col = cb + col_in;
    const bool on = active && col < g.c4n;
    double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
    float m[4] = {0, 0, 0, 0}, is[4] = {0, 0, 0, 0};
    if (MODE == 1 && on) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * col + j;
        if (c < g.C) {
          m[j] = mean[grp * g.C + c];
          is[j] = 1.0f / sqrtf(var[grp * g.C + c] + eps);
        }
      }
    }
    if (on) {
      const int64_t step = (int64_t)gridDim.x * rows;
      for (int64_t p0 = (int64_t)blockIdx.x * rows + prow; p0 < g.item_positions; p0 += step * kBnUnroll) {
        float4 xv[kBnUnroll], dv[kBnUnroll], yv[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const int64_t p = p0 + u * step;
          if (p < g.item_positions) {
            const int64_t off = ((base + p) * g.c4n + col) * 4;
            xv[u] = ldg4(x + off);
            if (MODE == 1) {
              dv[u] = ldg4(dy + off);
              if (y != nullptr) yv[u] = ldg4(y + off);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
          const int64_t p = p0 + u * step;
          if (p < g.item_positions) {
            const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
            if (MODE == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                a0[j] += (double)xs[j];
                a1[j] += (double)xs[j] * (double)xs[j];
              }
            } else {
              float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
              if (y != nullptr) {
                const float ys[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (!(ys[j] > 0.0f)) ds[j] = 0.0f;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                a0[j] += (double)ds[j];
                a1[j] += (double)ds[j] * (double)((xs[j] - m[j]) * is[j]);
              }
            }
          }
        }
      }
    }
    // CTA reduction over the position rows of each column, then one global atomic per (column, value)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_acc[tid][j] = a0[j];
      s_acc[tid][4 + j] = a1[j];
    }
    __syncthreads();
    if (tid < cw && cb + tid < g.c4n) {
      double t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = 0.0;
      for (int r = 0; r < rows; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] += s_acc[r * cw + tid][j];
      double* o = ws + (int64_t)grp * 2 * g.C;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * (cb + tid) + j;
        if (c < g.C) {
          atomicAdd(o + c, t[j]);
          atomicAdd(o + g.C + c, t[4 + j]);
        }
      }
    }
    __syncthreads();
  }
}

// one thread per channel: mean / biased variance of every group, then (optionally) the running-average update of the
// module exactly as n_groups successive F.batch_norm calls make it: r <- (1 - m) r + m stat_g, the variance unbiased
__global__ void bn_stats_finish_kernel(const double* ws, BnGeom g, float* mean, float* var, float* running_mean,
                                       float* running_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.C) return;
  float rm = running_mean ? running_mean[c] : 0.0f, rv = running_var ? running_var[c] : 0.0f;
  for (int grp = 0; grp < g.n_groups; ++grp) {
    const double P = bn_group_positions(g, grp);
    float m = 0.0f, v = 0.0f;
    if (P > 0.0) {
      const double md = ws[(int64_t)grp * 2 * g.C + c] / P;
      double vd = ws[(int64_t)grp * 2 * g.C + g.C + c] / P - md * md;
      if (vd < 0.0) vd = 0.0;
      m = (float)md;
      v = (float)vd;
      // (float32 arithmetic in torch's order: running.mul_(1 - m).add_(stat, alpha = m))
      rm = rm * (1.0f - momentum) + momentum * m;
      rv = rv * (1.0f - momentum) + momentum * (v * (float)(P / (P > 1.0 ? P - 1.0 : 1.0)));
    }
    mean[grp * g.C + c] = m;
    var[grp * g.C + c] = v;
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// grid = (chunks, items): y = act(x * scale[g][c] + shift[g][c] (+ residual)); padding channels are written as zeros
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const sp3d_bn_apply_args a, BnGeom g) {
  const int item = blockIdx.y;
  const int grp = bn_group_of(g, item);
  const int tid = threadIdx.x;
  const int cw = g.c4n < kBnThreads ? g.c4n : kBnThreads;
  const int rows = kBnThreads / cw;
  const int col_in = tid % cw, prow = tid / cw;
  if (prow >= rows) return;
  const int64_t base = (int64_t)item * g.item_positions;
  const int64_t step = (int64_t)gridDim.x * rows;
  for (int cb = 0; cb < g.c4n; cb += cw) {
    const int col = cb + col_in;
    if (col >= g.c4n) continue;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * col + j;
      sc[j] = c < g.C ? a.scale[grp * g.C + c] : 0.0f;
      sh[j] = c < g.C ? a.shift[grp * g.C + c] : 0.0f;
    }
    for (int64_t p0 = (int64_t)blockIdx.x * rows + prow; p0 < g.item_positions; p0 += step * kBnUnroll) {
      float4 xv[kBnUnroll], rv[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int64_t p = p0 + u * step;
        if (p < g.item_positions) {
          const int64_t off = ((base + p) * g.c4n + col) * 4;
          xv[u] = ldg4(a.x + off);
          if (a.residual != nullptr) rv[u] = ldg4(a.residual + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int64_t p = p0 + u * step;
        if (p < g.item_positions) {
          const int64_t off = ((base + p) * g.c4n + col) * 4;
          const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
          float r[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            r[j] = xs[j] * sc[j] + sh[j];
            if (a.relu == 2) r[j] = fmaxf(r[j], 0.0f);
          }
          if (a.residual != nullptr) {
            r[0] += rv[u].x; r[1] += rv[u].y; r[2] += rv[u].z; r[3] += rv[u].w;
          }
          if (a.relu == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = fmaxf(r[j], 0.0f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (4 * col + j >= g.C) r[j] = 0.0f;
          *reinterpret_cast<float4*>(a.y + off) = make_float4(r[0], r[1], r[2], r[3]);
        }
      }
    }
  }
}

// dx = gamma * rsqrt(var + eps) * (dz - dbeta_g / P_g - xhat * dgamma_g / P_g) with the group's sums in ws
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const sp3d_bn_bwd_args a, BnGeom g) {
  const int item = blockIdx.y;
  const int grp = bn_group_of(g, item);
  const int tid = threadIdx.x;
  const int cw = g.c4n < kBnThreads ? g.c4n : kBnThreads;
  const int rows = kBnThreads / cw;
  const int col_in = tid % cw, prow = tid / cw;
  if (prow >= rows) return;
  const int64_t base = (int64_t)item * g.item_positions;
  const int64_t step = (int64_t)gridDim.x * rows;
  const double invP = 1.0 / bn_group_positions(g, grp);
  for (int cb = 0; cb < g.c4n; cb += cw) {
    const int col = cb + col_in;
    if (col >= g.c4n) continue;
    float m[4], is[4], k[4], db[4], dg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 4 * col + j;
      m[j] = is[j] = k[j] = db[j] = dg[j] = 0.0f;
      if (c < g.C) {
        m[j] = a.mean[grp * g.C + c];
        is[j] = 1.0f / sqrtf(a.var[grp * g.C + c] + a.eps);
        k[j] = (a.gamma != nullptr ? a.gamma[c] : 1.0f) * is[j];
        db[j] = (float)(a.workspace[(int64_t)grp * 2 * g.C + c] * invP);
        dg[j] = (float)(a.workspace[(int64_t)grp * 2 * g.C + g.C + c] * invP);
      }
    }
    for (int64_t p0 = (int64_t)blockIdx.x * rows + prow; p0 < g.item_positions; p0 += step * kBnUnroll) {
      float4 xv[kBnUnroll], dv[kBnUnroll], yv[kBnUnroll];
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int64_t p = p0 + u * step;
        if (p < g.item_positions) {
          const int64_t off = ((base + p) * g.c4n + col) * 4;
          xv[u] = ldg4(a.x + off);
          dv[u] = ldg4(a.grad_y + off);
          if (a.y != nullptr) yv[u] = ldg4(a.y + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kBnUnroll; ++u) {
        const int64_t p = p0 + u * step;
        if (p < g.item_positions) {
          const int64_t off = ((base + p) * g.c4n + col) * 4;
          const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
          float ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
          if (a.y != nullptr) {
            const float ys[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (!(ys[j] > 0.0f)) ds[j] = 0.0f;
          }
          float r[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xh = (xs[j] - m[j]) * is[j];
            r[j] = k[j] * (ds[j] - db[j] - xh * dg[j]);       // padding channels: k = 0
          }
          *reinterpret_cast<float4*>(a.grad_x + off) = make_float4(r[0], r[1], r[2], r[3]);
        }
      }
    }
  }
}

// shared parameters: gradients summed over the groups
__global__ void bn_bwd_finish_kernel(const double* ws, int n_groups, int C, float* grad_gamma, float* grad_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double b = 0.0, gm = 0.0;
  for (int grp = 0; grp < n_groups; ++grp) {
    b += ws[(int64_t)grp * 2 * C + c];
    gm += ws[(int64_t)grp * 2 * C + C + c];
  }
  if (grad_beta != nullptr) grad_beta[c] = (float)b;
  if (grad_gamma != nullptr) grad_gamma[c] = (float)gm;
}

// items / groups of a launch; returns false on inconsistent arguments
static bool bn_geometry(int64_t P, int C, int pitch, int n_groups, const int32_t* item_group, const int32_t* group_items,
                        bool need_counts, int n_items, BnGeom* g) {
  if (P < 0 || C < 1 || pitch < C || (pitch % 4)) return false;
  const bool grouped = n_groups > 1 || item_group != nullptr;
  if (grouped && (item_group == nullptr || (need_counts && group_items == nullptr) || n_groups < 1 || n_items < 1 ||
                  n_items > 65535 || (P % n_items)))
    return false;
  g->c4n = pitch / 4;
  g->n_items = grouped ? n_items : 1;
  g->item_positions = grouped ? P / n_items : P;
  g->item_group = grouped ? item_group : nullptr;
  g->group_items = grouped ? group_items : nullptr;
  g->n_groups = grouped ? n_groups : 1;
  g->C = C;
  return true;
}

// CTAs along the positions of an item: enough to fill the GPU a few times over, at most one per CTA step
static unsigned bn_grid_x(const BnGeom& g, int waves) {
  const int cw = g.c4n < kBnThreads ? g.c4n : kBnThreads;
  const int rows = kBnThreads / cw;
  const int64_t need = (g.item_positions + (int64_t)rows * kBnUnroll - 1) / ((int64_t)rows * kBnUnroll);
  int64_t cap = (148 * waves + g.n_items - 1) / g.n_items;
  if (cap < 1) cap = 1;
  const int64_t n = need < cap ? need : cap;
  return (unsigned)(n < 1 ? 1 : n);
}

}  // namespace sp3d

extern "C" int sp3d_bn_stats(const sp3d_bn_stats_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->x == nullptr || a->mean == nullptr || a->var == nullptr || a->P < 1) return SP3D_ERR_INVALID_ARG;
  BnGeom g;
  if (!bn_geometry(a->P, a->C, a->pitch, a->n_groups, a->item_group, a->group_items, true, a->n_items, &g) ||
      (reinterpret_cast<uintptr_t>(a->x) % 16))
    return SP3D_ERR_INVALID_ARG;
  const int64_t ws_bytes = (int64_t)g.n_groups * 2 * a->C * (int64_t)sizeof(double);
  if (a->workspace == nullptr || a->workspace_bytes < ws_bytes || (reinterpret_cast<uintptr_t>(a->workspace) % 8))
    return SP3D_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(a->workspace, 0, (size_t)ws_bytes, st);
  if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
  bn_reduce_kernel<0><<<dim3(bn_grid_x(g, 8), g.n_items), kBnThreads, 0, st>>>(a->x, nullptr, nullptr, nullptr, nullptr, 0.0f, g,
                                                                               a->workspace);
  int rc = check_launch();
  if (rc != SP3D_OK) return rc;
  if ((a->running_mean == nullptr) != (a->running_var == nullptr)) return SP3D_ERR_INVALID_ARG;
  bn_stats_finish_kernel<<<(a->C + 127) / 128, 128, 0, st>>>(a->workspace, g, a->mean, a->var, a->running_mean, a->running_var,
                                                             a->momentum);
  return check_launch();
}

extern "C" int sp3d_bn_apply(const sp3d_bn_apply_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->x == nullptr || a->y == nullptr || a->scale == nullptr || a->shift == nullptr || a->relu < 0 ||
      a->relu > 2)
    return SP3D_ERR_INVALID_ARG;
  BnGeom g;
  if (!bn_geometry(a->P, a->C, a->pitch, a->n_groups, a->item_group, nullptr, false, a->n_items, &g))
    return SP3D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(a->x) % 16) || (reinterpret_cast<uintptr_t>(a->y) % 16) ||
      (a->residual != nullptr && (reinterpret_cast<uintptr_t>(a->residual) % 16)))
    return SP3D_ERR_INVALID_ARG;
  if (a->P == 0) return SP3D_OK;
  bn_apply_kernel<<<dim3(bn_grid_x(g, 16), g.n_items), kBnThreads, 0, static_cast<cudaStream_t>(stream)>>>(*a, g);
  return check_launch();
}

extern "C" int sp3d_bn_bwd(const sp3d_bn_bwd_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->x == nullptr || a->grad_y == nullptr || a->mean == nullptr || a->var == nullptr ||
      a->grad_x == nullptr || a->P < 1)
    return SP3D_ERR_INVALID_ARG;
  BnGeom g;
  if (!bn_geometry(a->P, a->C, a->pitch, a->n_groups, a->item_group, a->group_items, true, a->n_items, &g) ||
      (reinterpret_cast<uintptr_t>(a->x) % 16) || (reinterpret_cast<uintptr_t>(a->grad_y) % 16) ||
      (reinterpret_cast<uintptr_t>(a->grad_x) % 16) || (a->y != nullptr && (reinterpret_cast<uintptr_t>(a->y) % 16)))
    return SP3D_ERR_INVALID_ARG;
  const int64_t ws_bytes = (int64_t)g.n_groups * 2 * a->C * (int64_t)sizeof(double);
  if (a->workspace == nullptr || a->workspace_bytes < ws_bytes || (reinterpret_cast<uintptr_t>(a->workspace) % 8))
    return SP3D_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(a->workspace, 0, (size_t)ws_bytes, st);
  if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
  bn_reduce_kernel<1><<<dim3(bn_grid_x(g, 8), g.n_items), kBnThreads, 0, st>>>(a->x, a->grad_y, a->y, a->mean, a->var, a->eps, g,
                                                                               a->workspace);
  int rc = check_launch();
  if (rc != SP3D_OK) return rc;
  bn_bwd_apply_kernel<<<dim3(bn_grid_x(g, 16), g.n_items), kBnThreads, 0, st>>>(*a, g);
  rc = check_launch();
  if (rc != SP3D_OK) return rc;
  bn_bwd_finish_kernel<<<(a->C + 127) / 128, 128, 0, st>>>(a->workspace, g.n_groups, a->C, a->grad_gamma, a->grad_beta);
  return check_launch();
}
