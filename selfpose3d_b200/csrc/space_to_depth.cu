// 2 x 2 space-to-depth of an image batch into the channel-last bf16 layout the tensor-core convolution reads:
//   dst[n, y', x', (py * 2 + px) * C + c] = src[n, c, 2 y' + py, 2 x' + px]          (channels >= 4 C are zero)
// A stride-2 convolution on src becomes a stride-1 convolution (kernel ceil-halved) on dst, which is how the 7x7/s2
// stem and the 3x3/s2 convolutions of PoseResNet (lib/models/pose_resnet.py:102-105, 58-93) reach the tcgen05 path.
#include "sp3d_common.cuh"

namespace sp3d {

template <typename T>
__device__ __forceinline__ float s2d_load(const T* p);
template <>
__device__ __forceinline__ float s2d_load<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float s2d_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// one thread per (output position, group of 8 output channels) -> one 16-byte store
template <typename T, bool PAIR>
__global__ void space_to_depth_kernel(const sp3d_s2d_args a) {
  const int groups = a.dst_pitch / 8;
  const int OH = a.H / 2, OW = a.W / 2;
  const int64_t total = (int64_t)a.N * OH * OW * groups;
  const T* src = reinterpret_cast<const T*>(a.src);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int64_t pos = i / groups;
    const int x = (int)(pos % OW), y = (int)((pos / OW) % OH);
    const int64_t n = pos / ((int64_t)OW * OH);
    __align__(16) __nv_bfloat16 o[8], lo[8];
    if (sizeof(T) == 2 && a.stride_c == 1 && (a.C % 8) == 0 && 8 * g < 4 * a.C) {
      // channel-last bf16 source: 8 consecutive channels of one source pixel = one 16-byte copy
      const int q = (8 * g) / a.C, c0 = (8 * g) % a.C;
      const T* p = src + n * a.stride_n + (int64_t)(2 * y + (q >> 1)) * a.stride_y + (int64_t)(2 * x + (q & 1)) * a.stride_x + c0;
      *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(p);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = 8 * g + j;
        float v = 0.0f;
        if (ch < 4 * a.C) {
          const int q = ch / a.C, c = ch % a.C;
          v = s2d_load<T>(src + n * a.stride_n + (int64_t)c * a.stride_c + (int64_t)(2 * y + (q >> 1)) * a.stride_y +
                          (int64_t)(2 * x + (q & 1)) * a.stride_x);
        }
        o[j] = __float2bfloat16_rn(v);
        lo[j] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(o[j])));
      }
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dst) + pos * a.dst_pitch + 8 * g) =
        *reinterpret_cast<const uint4*>(o);
    if (PAIR)   // SP3D_BF16X2: second term plane
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dst) + ((int64_t)a.N * OH * OW + pos) * a.dst_pitch + 8 * g) =
          *reinterpret_cast<const uint4*>(lo);
  }
}

// dst[n, x, y, z, j] = src[n, x + j - pad, y, z, 0] for j < taps (zero outside the volume), channels >= taps zero:
// a 1-channel volume with its x-neighbourhood stacked into the channel dimension, so that a k^3 convolution on one
// input channel becomes a 1 x k x k convolution on `taps` channels (K = 16 of the tensor-core MMA is then used for
// taps instead of 15 padding zeros).  One thread per voxel, one 32-byte store.
__global__ void stack_x_shifts_kernel(const sp3d_stack_args a) {
  const int64_t yz = (int64_t)a.Y * a.Z;
  const int64_t total = (int64_t)a.N * a.X * yz;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.src);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)((i / yz) % a.X);
    __align__(16) __nv_bfloat16 o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int xs = x + j - a.pad;
      const bool ok = j < a.taps && xs >= 0 && xs < a.X;
      o[j] = ok ? src[(i + (int64_t)(j - a.pad) * yz) * a.src_pitch] : __float2bfloat16_rn(0.0f);
    }
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dst) + i * 16);
    d[0] = *reinterpret_cast<const uint4*>(o);
    d[1] = *reinterpret_cast<const uint4*>(o + 8);
  }
}

// float32 -> sum of S bf16 terms, term s into plane s (the operand layout of the split-operand tensor-core
// convolution).  One thread per (position, group of 8 channels): two 16-byte loads, S 16-byte stores.
template <int S>
__global__ void split_bf16_kernel(const sp3d_split_args a) {
  const int groups = a.c_block / 8;
  const int64_t total = a.P * groups;
  const bool vec = (a.src_pitch % 4) == 0 && (reinterpret_cast<uintptr_t>(a.src) % 16) == 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int64_t pos = i / groups;
    const float* sp = a.src + pos * a.src_pitch + 8 * g;
    float v[8];
    if (vec && 8 * g + 8 <= a.C) {
      const float4 lo = ldg4(sp), hi = ldg4(sp + 4);
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
      v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (8 * g + j < a.C) ? __ldg(sp + j) : 0.0f;
    }
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(a.dst) + pos * a.c_block + 8 * g;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      __align__(16) __nv_bfloat16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = __float2bfloat16_rn(v[j]);
        v[j] = __fsub_rn(v[j], __bfloat162float(o[j]));   // exact: the remainder of a bf16 rounding fits float32
      }
      *reinterpret_cast<uint4*>(dp + (int64_t)s * a.P * a.c_block) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// bf16 term planes -> float32 (plane 0 + plane 1), 8 channels per thread
__global__ void merge_bf16_kernel(const sp3d_split_args a) {
  const int groups = a.c_block / 8;
  const int64_t total = a.P * groups;
  float* dst = const_cast<float*>(a.src);
  const __nv_bfloat16* planes = reinterpret_cast<const __nv_bfloat16*>(a.dst);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const int64_t pos = i / groups;
    const uint4 qh = __ldg(reinterpret_cast<const uint4*>(planes + pos * a.c_block + 8 * g));
    const uint4 ql = __ldg(reinterpret_cast<const uint4*>(planes + (a.P + pos) * a.c_block + 8 * g));
    const uint32_t h4[4] = {qh.x, qh.y, qh.z, qh.w}, l4[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t h = (j & 1) ? (h4[j >> 1] & 0xffff0000u) : (h4[j >> 1] << 16);
      const uint32_t l = (j & 1) ? (l4[j >> 1] & 0xffff0000u) : (l4[j >> 1] << 16);
      if (8 * g + j < a.src_pitch) dst[pos * a.src_pitch + 8 * g + j] = (8 * g + j < a.C) ? __uint_as_float(h) + __uint_as_float(l) : 0.0f;
    }
  }
}

}  // namespace sp3d

extern "C" int sp3d_merge_bf16(const sp3d_split_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->src == nullptr || a->dst == nullptr || a->P < 0 || a->C < 1 || a->src_pitch < a->C ||
      a->c_block < a->C || (a->c_block % 8) || (reinterpret_cast<uintptr_t>(a->dst) % 16) || a->S != 2)
    return SP3D_ERR_INVALID_ARG;
  const int64_t total = a->P * (a->c_block / 8);
  if (total == 0) return SP3D_OK;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  merge_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}

extern "C" int sp3d_split_bf16(const sp3d_split_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->src == nullptr || a->dst == nullptr || a->P < 0 || a->C < 1 || a->src_pitch < a->C ||
      a->c_block < a->C || (a->c_block % 8) || (reinterpret_cast<uintptr_t>(a->dst) % 16))
    return SP3D_ERR_INVALID_ARG;
  if (a->S != 2 && a->S != 3) return SP3D_ERR_UNSUPPORTED;
  const int64_t total = a->P * (a->c_block / 8);
  if (total == 0) return SP3D_OK;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->S == 2) split_bf16_kernel<2><<<blocks, 256, 0, st>>>(*a);
  else split_bf16_kernel<3><<<blocks, 256, 0, st>>>(*a);
  return check_launch();
}

extern "C" int sp3d_stack_x_shifts(const sp3d_stack_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->src == nullptr || a->dst == nullptr || a->N < 0 || a->X < 1 || a->Y < 1 || a->Z < 1 ||
      a->taps < 1 || a->taps > 16 || a->pad < 0 || a->src_pitch < 1 || (reinterpret_cast<uintptr_t>(a->dst) % 16))
    return SP3D_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->N * a->X * a->Y * a->Z;
  if (total == 0) return SP3D_OK;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  stack_x_shifts_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}

extern "C" int sp3d_space_to_depth(const sp3d_s2d_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->src == nullptr || a->dst == nullptr || a->N < 0 || a->C < 1 || a->H < 2 || a->W < 2 ||
      (a->H & 1) || (a->W & 1) || a->dst_pitch < 4 * a->C || (a->dst_pitch % 8) ||
      (reinterpret_cast<uintptr_t>(a->dst) % 16))
    return SP3D_ERR_INVALID_ARG;
  if (a->src_dtype != SP3D_F32 && a->src_dtype != SP3D_BF16) return SP3D_ERR_UNSUPPORTED;
  const int64_t total = (int64_t)a->N * (a->H / 2) * (a->W / 2) * (a->dst_pitch / 8);
  if (total == 0) return SP3D_OK;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool pair = a->dst_dtype == SP3D_BF16X2;
  if (a->dst_dtype != 0 && a->dst_dtype != SP3D_BF16 && !pair) return SP3D_ERR_UNSUPPORTED;
  if (pair && a->src_dtype != SP3D_F32) return SP3D_ERR_UNSUPPORTED;   // (bf16 term planes are re-arranged plane by plane)
  if (a->src_dtype == SP3D_F32) {
    if (pair) space_to_depth_kernel<float, true><<<blocks, 256, 0, st>>>(*a);
    else space_to_depth_kernel<float, false><<<blocks, 256, 0, st>>>(*a);
  } else {
    // the 16-byte copy path needs aligned source pixels
    if (a->stride_c == 1 && (a->C % 8) == 0 &&
        ((reinterpret_cast<uintptr_t>(a->src) % 16) || (a->stride_x % 8) || (a->stride_y % 8) || (a->stride_n % 8)))
      return SP3D_ERR_INVALID_ARG;
    space_to_depth_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>(*a);
  }
  return check_launch();
}
