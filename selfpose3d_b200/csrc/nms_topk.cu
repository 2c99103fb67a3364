// K3 -- 3-D non-maximum suppression + top-K proposals + index -> world location.
//
// One CTA per sample.  Pass 1: every thread walks its voxels in ascending flat index, evaluates
// v' = (x == max over the 3x3x3 neighbourhood) ? x : 0 (the reference zeroes non-maxima rather
// than removing them) and keeps its K best (value desc, index asc) in shared memory.  Pass 2: K
// rounds of a block-wide arg-max over the list heads.  Tie policy: highest value, lowest index.
//
// Reference semantics: lib/core/proposal.py:18-48, lib/models/cuboid_proposal_net_soft.py:46-68.
#include "sp3d_common.cuh"
#include <math.h>

namespace sp3d {

constexpr int kNmsMaxThreads = 1024;   // one CTA per sample: as many threads as the candidate lists leave room for
constexpr int kNmsMaxK = 32;
constexpr int kNmsSlabs = 16;          // CTAs per sample when the caller provides a workspace

__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

// proposal k of sample b: flat index -> world location (float32 or float64 as torch's type promotion does), flag, score
__device__ __forceinline__ void write_proposal(const sp3d_nms_topk_args& a, int b, int k, float bv, int bi) {
  const int X = a.X, Y = a.Y, Z = a.Z, K = a.K;
  const int iz = bi % Z, iy = (bi / Z) % Y, ix = bi / (Z * Y);
  const int idx[3] = {ix, iy, iz};
  const int dims[3] = {X, Y, Z};
  float* gc = a.grid_centers + ((int64_t)b * K + k) * 5;
  for (int d = 0; d < 3; ++d) {
    const float q = __fdiv_rn((float)idx[d], __fsub_rn((float)dims[d], 1.0f));
    float loc;
    if (a.loc_f64) {
      const double s = a.space_size[d];
      loc = (float)(__dsub_rn(__dadd_rn(__dmul_rn((double)q, s), a.space_center[d]), s / 2.0));
    } else {
      const float s = (float)a.space_size[d];
      loc = __fsub_rn(__fadd_rn(__fmul_rn(q, s), (float)a.space_center[d]), __fdiv_rn(s, 2.0f));
    }
    gc[d] = loc;
  }
  gc[3] = (bv > a.threshold) ? 0.0f : -1.0f;
  gc[4] = bv;
  if (a.topk_index != nullptr) a.topk_index[(int64_t)b * K + k] = bi;
}

__global__ void __launch_bounds__(kNmsMaxThreads) nms_topk_kernel(const sp3d_nms_topk_args a) {
  const int kNmsThreads = blockDim.x;
  extern __shared__ unsigned char smem_raw[];
  // per-thread candidate lists, interleaved so that slot s of thread t is at [s * kNmsThreads + t]
  float* cand_v = reinterpret_cast<float*>(smem_raw);
  int* cand_i = reinterpret_cast<int*>(cand_v + a.K * kNmsThreads);
  __shared__ float red_v[kNmsMaxThreads / 32];
  __shared__ int red_i[kNmsMaxThreads / 32];
  __shared__ int red_t[kNmsMaxThreads / 32];
  __shared__ int win_t;

  const int b = blockIdx.y;                       // sample; blockIdx.x = slab of x-slices (1 slab without a workspace)
  const int tid = threadIdx.x;
  const int K = a.K;
  const int X = a.X, Y = a.Y, Z = a.Z;
  const int N = X * Y * Z;
  const float* x = a.root_cubes + (int64_t)b * N;
  const float ninf = -INFINITY;

  for (int s = 0; s < K; ++s) {
    cand_v[s * kNmsThreads + tid] = ninf;
    cand_i[s * kNmsThreads + tid] = 0x7fffffff;
  }
  const int xs = (X + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_begin = (int)blockIdx.x * xs * Y * Z, n_end = min(N, ((int)blockIdx.x + 1) * xs * Y * Z);
  for (int n = n_begin + tid; n < n_end; n += kNmsThreads) {
    const int iz = n % Z, iy = (n / Z) % Y, ix = n / (Z * Y);
    const float c = x[n];
    float mx = ninf;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = ix + dx;
      if (xx < 0 || xx >= X) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = iy + dy;
        if (yy < 0 || yy >= Y) continue;
        for (int dz = -1; dz <= 1; ++dz) {
          const int zz = iz + dz;
          if (zz < 0 || zz >= Z) continue;
          mx = fmaxf(mx, x[(xx * Y + yy) * Z + zz]);
        }
      }
    }
    const float v = (c == mx) ? c : 0.0f;
    // insertion into this thread's sorted list; indices arrive ascending, so strict '>' keeps ties ordered
    if (v > cand_v[(K - 1) * kNmsThreads + tid]) {
      int s = K - 1;
      while (s > 0 && v > cand_v[(s - 1) * kNmsThreads + tid]) {
        cand_v[s * kNmsThreads + tid] = cand_v[(s - 1) * kNmsThreads + tid];
        cand_i[s * kNmsThreads + tid] = cand_i[(s - 1) * kNmsThreads + tid];
        --s;
      }
      cand_v[s * kNmsThreads + tid] = v;
      cand_i[s * kNmsThreads + tid] = n;
    }
  }
  __syncthreads();

  int head = 0;  // this thread's next unconsumed candidate
  for (int k = 0; k < K; ++k) {
    float v = head < K ? cand_v[head * kNmsThreads + tid] : ninf;
    int i = head < K ? cand_i[head * kNmsThreads + tid] : 0x7fffffff;
    int t = tid;
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_down_sync(0xffffffffu, v, off);
      const int oi = __shfl_down_sync(0xffffffffu, i, off);
      const int ot = __shfl_down_sync(0xffffffffu, t, off);
      if (better(ov, oi, v, i)) { v = ov; i = oi; t = ot; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = v; red_i[tid >> 5] = i; red_t[tid >> 5] = t; }
    __syncthreads();
    if (tid == 0) {
      float bv = red_v[0]; int bi = red_i[0]; int bt = red_t[0];
      for (int w = 1; w < kNmsThreads / 32; ++w)
        if (better(red_v[w], red_i[w], bv, bi)) { bv = red_v[w]; bi = red_i[w]; bt = red_t[w]; }
      win_t = bt;
      if (a.workspace != nullptr && gridDim.x > 1) {   // slab mode: this CTA's k-th candidate, merged by nms_merge_kernel
        float* cv = reinterpret_cast<float*>(a.workspace) + (((int64_t)b * gridDim.x + blockIdx.x) * K + k) * 2;
        cv[0] = bv;
        reinterpret_cast<int*>(cv)[1] = bi;
      } else {
        write_proposal(a, b, k, bv, bi);
      }
    }
    __syncthreads();
    if (tid == win_t) ++head;
  }
}

// slab mode, stage 2: one warp per sample picks the K best of the S * K slab candidates (same order relation)
__global__ void nms_merge_kernel(const sp3d_nms_topk_args a, int slabs) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int K = a.K, total = slabs * K;
  const float* cand = reinterpret_cast<const float*>(a.workspace) + (int64_t)b * total * 2;
  unsigned long long taken_lo = 0, taken_hi = 0;        // candidates of this lane already emitted (<= 128 per lane)
  for (int k = 0; k < K; ++k) {
    float v = -INFINITY;
    int i = 0x7fffffff, src = -1;
    for (int c = lane, j = 0; c < total; c += 32, ++j) {
      const bool used = j < 64 ? ((taken_lo >> j) & 1ull) : ((taken_hi >> (j - 64)) & 1ull);
      if (used) continue;
      const float cv = cand[2 * c];
      const int ci = reinterpret_cast<const int*>(cand)[2 * c + 1];
      if (better(cv, ci, v, i)) { v = cv; i = ci; src = j; }
    }
    float bv = v;
    int bi = i, bl = lane;
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
      if (better(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) { bv = ov; bi = oi; bl = ol; }
    }
    if (lane == bl && src >= 0) {
      if (src < 64) taken_lo |= 1ull << src;
      else taken_hi |= 1ull << (src - 64);
    }
    if (lane == 0) write_proposal(a, b, k, bv, bi);
  }
}

}  // namespace sp3d

extern "C" int64_t sp3d_nms_topk3d_workspace(const sp3d_nms_topk_args* a) {
  if (a == nullptr || a->B < 1 || a->K < 1) return 0;
  return (int64_t)a->B * sp3d::kNmsSlabs * a->K * 2 * (int64_t)sizeof(float);
}

extern "C" int sp3d_nms_topk3d(const sp3d_nms_topk_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->root_cubes == nullptr || a->grid_centers == nullptr || a->K < 1 || a->K > kNmsMaxK ||
      a->B < 0 || a->X < 1 || a->Y < 1 || a->Z < 1)
    return SP3D_ERR_INVALID_ARG;
  if ((int64_t)a->X * a->Y * a->Z < a->K) return SP3D_ERR_INVALID_ARG;  // torch.topk would raise
  if (a->B == 0) return SP3D_OK;
  // with a workspace: kNmsSlabs CTAs of 256 threads per sample (x slabs) + a one-warp merge; else one big CTA.
  const bool slabs = a->workspace != nullptr && a->workspace_bytes >= sp3d_nms_topk3d_workspace(a) &&
                     a->X >= kNmsSlabs && (int64_t)((a->X + kNmsSlabs - 1) / kNmsSlabs) * a->Y * a->Z >= a->K &&
                     kNmsSlabs * a->K <= 32 * 128 && (reinterpret_cast<uintptr_t>(a->workspace) % 8) == 0;
  // per-thread candidate lists of K (value, index) pairs live in shared memory: 1024 threads up to K = 25
  int threads = slabs ? 256 : kNmsMaxThreads;
  while ((size_t)a->K * threads * 8 > 200 * 1024) threads /= 2;
  const size_t smem = (size_t)a->K * threads * (sizeof(float) + sizeof(int));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sp3d_nms_topk_args k = *a;
  if (!slabs) k.workspace = nullptr;
  nms_topk_kernel<<<dim3(slabs ? kNmsSlabs : 1, a->B), threads, smem, st>>>(k);
  int rc = check_launch();
  if (rc != SP3D_OK || !slabs) return rc;
  nms_merge_kernel<<<a->B, 32, 0, st>>>(k, kNmsSlabs);
  return check_launch();
}
