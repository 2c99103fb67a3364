// K2/K5 (float32 SIMT form) -- convolution family as an implicit GEMM on channel-last tensors,
// with the folded-BatchNorm / bias / residual / ReLU epilogue fused.  This is the float32-exact
// path used for parity and for the shapes the tensor-core path does not take (see conv_tc.cu).
//
//   M = N*OD*OH*OW output positions (128 per CTA), N = cout (16..128 per CTA), K = taps*cin.
//   A tile gathered from the activation tensor with zero padding, B tile from the packed weights;
//   both staged through shared memory with register prefetch of the next K step.
//
// Reference semantics: cudnn conv3d / conv_transpose3d / conv2d / conv_transpose2d + batch_norm +
// relu (+ residual add) as wired in lib/models/v2v_net.py:10-69,124 and
// lib/models/pose_resnet.py:58-93,102-124,161-207.
#include "sp3d_common.cuh"

namespace sp3d {

constexpr int kBM = 128;
constexpr int kBK = 16;
constexpr int kConvThreads = 256;
constexpr int kMaxTaps = 512;

template <int BN>
struct ConvTile {
  static constexpr int TN = BN >= 64 ? BN / 16 : (BN == 32 ? 4 : 2);  // columns per thread
  static constexpr int TX = BN / TN;                                  // threads along N
  static constexpr int TY = kConvThreads / TX;                        // threads along M
  static constexpr int TM = kBM / TY;                                 // rows per thread
};

// 4 consecutive channels of an activation tensor as float4 (float32 or bf16 storage)
__device__ __forceinline__ float4 load4(const float* p) { return ldg4(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
  return make_float4(__bfloat162float(lo.x), __bfloat162float(lo.y), __bfloat162float(hi.x), __bfloat162float(hi.y));
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store1(float* p, float v) { *p = v; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int BN, typename InT, typename OutT>
__global__ void __launch_bounds__(kConvThreads) conv_igemm_f32_kernel(const sp3d_conv_args a) {
  using T = ConvTile<BN>;
  constexpr int TM = T::TM, TN = T::TN, TX = T::TX;
  __shared__ __align__(16) float As[kBK][kBM + 4];
  __shared__ __align__(16) float Bs[kBK][BN + 4];
  __shared__ int4 s_row[kBM];        // n, input-origin d/h/w of each tile row (n < 0: row out of range)
  __shared__ int s_tap[kMaxTaps];    // packed tap offsets (10 bits each, biased by 512)

  const int tid = threadIdx.x;
  const int64_t M_total = (int64_t)a.N * a.OD * a.OH * a.OW;
  const int64_t m0 = (int64_t)blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int ntaps = a.ksize[0] * a.ksize[1] * a.ksize[2];
  const int K_total = ntaps * a.cin;

  for (int r = tid; r < kBM; r += kConvThreads) {
    const int64_t m = m0 + r;
    int4 info = make_int4(-1, 0, 0, 0);
    if (m < M_total) {
      const int ow = (int)(m % a.OW);
      const int oh = (int)((m / a.OW) % a.OH);
      const int od = (int)((m / ((int64_t)a.OW * a.OH)) % a.OD);
      const int n = (int)(m / ((int64_t)a.OW * a.OH * a.OD));
      info = make_int4(n, od * a.stride[0] + a.tap_off0[0], oh * a.stride[1] + a.tap_off0[1],
                       ow * a.stride[2] + a.tap_off0[2]);
    }
    s_row[r] = info;
  }
  for (int t = tid; t < ntaps; t += kConvThreads) {
    const int tw = t % a.ksize[2];
    const int th = (t / a.ksize[2]) % a.ksize[1];
    const int td = t / (a.ksize[2] * a.ksize[1]);
    s_tap[t] = ((td * a.tap_step[0] + 512) << 20) | ((th * a.tap_step[1] + 512) << 10) | (tw * a.tap_step[2] + 512);
  }
  __syncthreads();

  // A loader: 128 rows x 4 float4 per K step -> 2 per thread
  const int a_kv = tid & 3;
  const int a_r0 = tid >> 2;
  // B loader: 16 rows x BN/4 float4 per K step
  constexpr int kBVecPerRow = BN / 4;
  constexpr int kBVec = kBK * kBVecPerRow;
  constexpr int kBIter = (kBVec + kConvThreads - 1) / kConvThreads;

  const InT* in = reinterpret_cast<const InT*>(a.in);
  const float* wgt = reinterpret_cast<const float*>(a.weight);

  float4 a_reg[2];
  float4 b_reg[kBIter];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = a_r0 + i * 64;
      const int kk = k0 + a_kv * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int4 info = s_row[r];
      if (info.x >= 0 && kk < K_total) {
        const int t = kk / a.cin;
        const int c = kk - t * a.cin;
        const int pk = s_tap[t];
        const int id = info.y + ((pk >> 20) & 1023) - 512;
        const int ih = info.z + ((pk >> 10) & 1023) - 512;
        const int iw = info.w + (pk & 1023) - 512;
        if (id >= 0 && id < a.D && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W)
          v = load4(in + ((((int64_t)info.x * a.D + id) * a.H + ih) * a.W + iw) * a.cin_pitch + c);
      }
      a_reg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < kBIter; ++i) {
      const int idx = tid + i * kConvThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < kBVec) {
        const int kr = idx / kBVecPerRow;
        const int nc = (idx - kr * kBVecPerRow) * 4;
        const int kk = k0 + kr;
        if (kk < K_total && n0 + nc < a.cout_pitch_w) v = ldg4(wgt + (int64_t)kk * a.cout_pitch_w + n0 + nc);
      }
      b_reg[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = a_r0 + i * 64;
      As[a_kv * 4 + 0][r] = a_reg[i].x;
      As[a_kv * 4 + 1][r] = a_reg[i].y;
      As[a_kv * 4 + 2][r] = a_reg[i].z;
      As[a_kv * 4 + 3][r] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < kBIter; ++i) {
      const int idx = tid + i * kConvThreads;
      if (idx < kBVec) {
        const int kr = idx / kBVecPerRow;
        const int nc = (idx - kr * kBVecPerRow) * 4;
        *reinterpret_cast<float4*>(&Bs[kr][nc]) = b_reg[i];
      }
    }
  };

  const int tx = tid % TX;
  const int ty = tid / TX;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  for (int k0 = 0; k0 < K_total; k0 += kBK) {
    store_tiles();
    __syncthreads();
    if (k0 + kBK < K_total) load_tiles(k0 + kBK);
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        av[i] = q.x; av[i + 1] = q.y; av[i + 2] = q.z; av[i + 3] = q.w;
      }
      if (TN >= 4) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 q = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
          bv[j] = q.x; bv[j + 1] = q.y; bv[j + 2] = q.z; bv[j + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: scale/shift (+ residual) (+ ReLU), channel-last store
  OutT* out = reinterpret_cast<OutT*>(a.out);
  const OutT* res = reinterpret_cast<const OutT*>(a.residual);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = ty * TM + i;
    const int64_t m = m0 + r;
    if (m >= M_total) continue;
    const int ow = (int)(m % a.OW);
    const int oh = (int)((m / a.OW) % a.OH);
    const int od = (int)((m / ((int64_t)a.OW * a.OH)) % a.OD);
    const int n = (int)(m / ((int64_t)a.OW * a.OH * a.OD));
    const int64_t pos = (((int64_t)n * a.TD + od * a.ostride[0] + a.ooffset[0]) * a.TH + oh * a.ostride[1] + a.ooffset[1]) *
                            a.TW + ow * a.ostride[2] + a.ooffset[2];
    OutT* o = out + pos * a.cout_pitch;
    const OutT* rp = res ? res + pos * a.cout_pitch : nullptr;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co < a.cout) {
        float v = acc[i][j];
        if (a.scale) v *= a.scale[co];
        if (a.shift) v += a.shift[co];
        if (a.relu == 2) v = fmaxf(v, 0.f);
        if (rp) v += to_f32(rp[co]);
        if (a.relu == 1) v = fmaxf(v, 0.f);
        store1(o + co, v);
      } else if (co < a.cout_pitch) {
        store1(o + co, 0.f);
      }
    }
  }
}

template <int BN>
static int launch_conv(const sp3d_conv_args* a, cudaStream_t st) {
  const int64_t M_total = (int64_t)a->N * a->OD * a->OH * a->OW;
  dim3 grid(ceil_div(M_total, kBM), ceil_div(a->cout, BN));
  if (a->in_dtype == SP3D_F32 && a->out_dtype == SP3D_F32)
    conv_igemm_f32_kernel<BN, float, float><<<grid, kConvThreads, 0, st>>>(*a);
  else if (a->in_dtype == SP3D_F32 && a->out_dtype == SP3D_BF16)
    conv_igemm_f32_kernel<BN, float, __nv_bfloat16><<<grid, kConvThreads, 0, st>>>(*a);
  else if (a->in_dtype == SP3D_BF16 && a->out_dtype == SP3D_BF16)
    conv_igemm_f32_kernel<BN, __nv_bfloat16, __nv_bfloat16><<<grid, kConvThreads, 0, st>>>(*a);
  else
    return SP3D_ERR_UNSUPPORTED;
  return check_launch();
}

int conv_simt_f32(const sp3d_conv_args* a, cudaStream_t st) {
  const int ntaps = a->ksize[0] * a->ksize[1] * a->ksize[2];
  if (ntaps < 1 || ntaps > kMaxTaps) return SP3D_ERR_UNSUPPORTED;
  if ((a->cin % 4) != 0 || (a->cin_pitch % 4) != 0 || (a->cout_pitch_w % 4) != 0 || a->cin > a->cin_pitch)
    return SP3D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(a->in) % 16) != 0 || (reinterpret_cast<uintptr_t>(a->weight) % 16) != 0)
    return SP3D_ERR_INVALID_ARG;
  // float32 math on float32 or bf16 storage (the bf16 forms serve the strided 2-D convolutions of the bf16 backbone)
  for (int d = 0; d < 3; ++d) {
    const int lo = a->tap_off0[d] + (a->tap_step[d] < 0 ? (a->ksize[d] - 1) * a->tap_step[d] : 0);
    const int hi = a->tap_off0[d] + (a->tap_step[d] > 0 ? (a->ksize[d] - 1) * a->tap_step[d] : 0);
    if (lo < -500 || hi > 500) return SP3D_ERR_UNSUPPORTED;
  }
  if (a->cout > 64) return launch_conv<128>(a, st);
  if (a->cout > 32) return launch_conv<64>(a, st);
  if (a->cout > 16) return launch_conv<32>(a, st);
  return launch_conv<16>(a, st);
}

// ------------------------------------------------------------------------------------------ max pool
__global__ void maxpool_f32_kernel(const sp3d_maxpool_args a) {
  const int cvec = a.c_pitch / 4;
  const int64_t total = (int64_t)a.N * a.OD * a.OH * a.OW * cvec;
  const float* in = reinterpret_cast<const float*>(a.in);
  float* out = reinterpret_cast<float*>(a.out);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t pos = i / cvec;
    const int ow = (int)(pos % a.OW); pos /= a.OW;
    const int oh = (int)(pos % a.OH); pos /= a.OH;
    const int od = (int)(pos % a.OD);
    const int n = (int)(pos / a.OD);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int kd = 0; kd < a.k[0]; ++kd) {
      const int id = od * a.s[0] - a.p[0] + kd;
      if (id < 0 || id >= a.D) continue;
      for (int kh = 0; kh < a.k[1]; ++kh) {
        const int ih = oh * a.s[1] - a.p[1] + kh;
        if (ih < 0 || ih >= a.H) continue;
        for (int kw = 0; kw < a.k[2]; ++kw) {
          const int iw = ow * a.s[2] - a.p[2] + kw;
          if (iw < 0 || iw >= a.W) continue;
          const float4 q = ldg4(in + ((((int64_t)n * a.D + id) * a.H + ih) * a.W + iw) * a.c_pitch + cv * 4);
          m.x = fmaxf(m.x, q.x); m.y = fmaxf(m.y, q.y); m.z = fmaxf(m.z, q.z); m.w = fmaxf(m.w, q.w);
        }
      }
    }
    *reinterpret_cast<float4*>(out + ((((int64_t)n * a.OD + od) * a.OH + oh) * a.OW + ow) * a.c_pitch + cv * 4) = m;
  }
}

// bf16 variant: 8 channels per 16-byte vector
__global__ void maxpool_bf16_kernel(const sp3d_maxpool_args a) {
  const int cvec = a.c_pitch / 8;
  const int64_t total = (int64_t)a.N * a.OD * a.OH * a.OW * cvec;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(a.in);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t pos = i / cvec;
    const int ow = (int)(pos % a.OW); pos /= a.OW;
    const int oh = (int)(pos % a.OH); pos /= a.OH;
    const int od = (int)(pos % a.OD);
    const int n = (int)(pos / a.OD);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kd = 0; kd < a.k[0]; ++kd) {
      const int id = od * a.s[0] - a.p[0] + kd;
      if (id < 0 || id >= a.D) continue;
      for (int kh = 0; kh < a.k[1]; ++kh) {
        const int ih = oh * a.s[1] - a.p[1] + kh;
        if (ih < 0 || ih >= a.H) continue;
        for (int kw = 0; kw < a.k[2]; ++kw) {
          const int iw = ow * a.s[2] - a.p[2] + kw;
          if (iw < 0 || iw >= a.W) continue;
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(
              in + ((((int64_t)n * a.D + id) * a.H + ih) * a.W + iw) * a.c_pitch + cv * 8));
          const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&q);
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], __bfloat162float(h[j]));
        }
      }
    }
    __align__(16) __nv_bfloat16 o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(m[j]);
    *reinterpret_cast<uint4*>(out + ((((int64_t)n * a.OD + od) * a.OH + oh) * a.OW + ow) * a.c_pitch + cv * 8) =
        *reinterpret_cast<const uint4*>(o);
  }
}

// SP3D_BF16X2: float32 values as two bf16 term planes.  The maximum is taken over plane 0 + plane 1 (exact in float32:
// the pair came from splitting a float32 value) and re-split, 8 channels per thread.
__global__ void maxpool_pair_kernel(const sp3d_maxpool_args a) {
  const int cvec = a.c_pitch / 8;
  const int64_t total = (int64_t)a.N * a.OD * a.OH * a.OW * cvec;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(a.in);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
  const int64_t in_plane = (int64_t)a.N * a.D * a.H * a.W * a.c_pitch;
  const int64_t out_plane = (int64_t)a.N * a.OD * a.OH * a.OW * a.c_pitch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t pos = i / cvec;
    const int ow = (int)(pos % a.OW); pos /= a.OW;
    const int oh = (int)(pos % a.OH); pos /= a.OH;
    const int od = (int)(pos % a.OD);
    const int n = (int)(pos / a.OD);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int kd = 0; kd < a.k[0]; ++kd) {
      const int id = od * a.s[0] - a.p[0] + kd;
      if (id < 0 || id >= a.D) continue;
      for (int kh = 0; kh < a.k[1]; ++kh) {
        const int ih = oh * a.s[1] - a.p[1] + kh;
        if (ih < 0 || ih >= a.H) continue;
#pragma unroll 2
        for (int kw = 0; kw < a.k[2]; ++kw) {
          const int iw = ow * a.s[2] - a.p[2] + kw;
          if (iw < 0 || iw >= a.W) continue;
          const __nv_bfloat16* p = in + ((((int64_t)n * a.D + id) * a.H + ih) * a.W + iw) * a.c_pitch + cv * 8;
          const uint4 qh = __ldg(reinterpret_cast<const uint4*>(p));
          const uint4 ql = __ldg(reinterpret_cast<const uint4*>(p + in_plane));
          const uint32_t h4[4] = {qh.x, qh.y, qh.z, qh.w}, l4[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            m[2 * j] = fmaxf(m[2 * j], __uint_as_float(h4[j] << 16) + __uint_as_float(l4[j] << 16));
            m[2 * j + 1] = fmaxf(m[2 * j + 1], __uint_as_float(h4[j] & 0xffff0000u) + __uint_as_float(l4[j] & 0xffff0000u));
          }
        }
      }
    }
    uint32_t w[4], u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * j], m[2 * j + 1]);
      const __nv_bfloat162 l = __floats2bfloat162_rn(m[2 * j] - __low2float(h), m[2 * j + 1] - __high2float(h));
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
      u[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    __nv_bfloat16* o = out + ((((int64_t)n * a.OD + od) * a.OH + oh) * a.OW + ow) * a.c_pitch + cv * 8;
    *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(o + out_plane) = make_uint4(u[0], u[1], u[2], u[3]);
  }
}

// bf16, window 2 stride 2 no padding (V2VNet's pools): the 8 taps are independent 16-byte loads issued together and
// reduced with packed bf16x2 max
__global__ void maxpool_bf16_k2s2_kernel(const sp3d_maxpool_args a) {
  const int cvec = a.c_pitch / 8;
  const int64_t total = (int64_t)a.N * a.OD * a.OH * a.OW * cvec;
  const uint4* in = reinterpret_cast<const uint4*>(a.in);
  uint4* out = reinterpret_cast<uint4*>(a.out);
  const int64_t sw = cvec, sh = (int64_t)a.W * cvec, sd = (int64_t)a.H * a.W * cvec;   // input strides in 16-byte units
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t pos = i / cvec;
    const int ow = (int)(pos % a.OW); pos /= a.OW;
    const int oh = (int)(pos % a.OH); pos /= a.OH;
    const int od = (int)(pos % a.OD);
    const int64_t n = pos / a.OD;
    const uint4* p = in + n * a.D * sd + (int64_t)(2 * od) * sd + (int64_t)(2 * oh) * sh + (int64_t)(2 * ow) * sw + cv;
    uint4 q[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) q[t] = __ldg(p + (t >> 2) * sd + ((t >> 1) & 1) * sh + (t & 1) * sw);
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(&(reinterpret_cast<const uint32_t*>(&q[0])[j]));
#pragma unroll
      for (int t = 1; t < 8; ++t)
        m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(&(reinterpret_cast<const uint32_t*>(&q[t])[j])));
      r[j] = *reinterpret_cast<const uint32_t*>(&m);
    }
    out[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// ------------------------------------------------------------------------------------------ layout
// [N, C, S] <-> [N, S, c_pitch] through a 32x32 shared-memory transpose tile.
template <typename SrcT, typename DstT>
__global__ void layout_to_cl_kernel(const SrcT* __restrict__ src, DstT* __restrict__ dst, int64_t C, int64_t S,
                                    int64_t c_pitch) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + j, s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && s < S) ? (float)src[(n * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t s = s0 + j, c = c0 + threadIdx.x;
    if (s < S && c < c_pitch) dst[(n * S + s) * c_pitch + c] = (DstT)tile[threadIdx.x][j];
  }
}

template <typename SrcT, typename DstT>
__global__ void layout_to_cf_kernel(const SrcT* __restrict__ src, DstT* __restrict__ dst, int64_t C, int64_t S,
                                    int64_t c_pitch) {
  __shared__ float tile[32][33];
  const int64_t n = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t s = s0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (s < S && c < C) ? (float)src[(n * S + s) * c_pitch + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + j, s = s0 + threadIdx.x;
    if (c < C && s < S) dst[(n * C + c) * S + s] = (DstT)tile[threadIdx.x][j];
  }
}

template <typename SrcT, typename DstT>
static int launch_layout(const sp3d_layout_args* a, cudaStream_t st) {
  const int64_t cspan = a->to_channel_last ? a->c_pitch : a->C;
  dim3 block(32, 8);
  dim3 grid(ceil_div(a->S, 32), ceil_div(cspan, 32), (unsigned)a->N);
  if (a->to_channel_last)
    layout_to_cl_kernel<SrcT, DstT><<<grid, block, 0, st>>>(reinterpret_cast<const SrcT*>(a->src),
                                                           reinterpret_cast<DstT*>(a->dst), a->C, a->S, a->c_pitch);
  else
    layout_to_cf_kernel<SrcT, DstT><<<grid, block, 0, st>>>(reinterpret_cast<const SrcT*>(a->src),
                                                           reinterpret_cast<DstT*>(a->dst), a->C, a->S, a->c_pitch);
  return check_launch();
}

}  // namespace sp3d

extern "C" int sp3d_maxpool_fwd(const sp3d_maxpool_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->in == nullptr || a->out == nullptr || a->N < 0 || a->C < 1 || a->c_pitch < a->C)
    return SP3D_ERR_INVALID_ARG;
  if (a->dtype != SP3D_F32 && a->dtype != SP3D_BF16 && a->dtype != SP3D_BF16X2) return SP3D_ERR_UNSUPPORTED;
  const int vec = a->dtype == SP3D_F32 ? 4 : 8;
  if (a->dtype == SP3D_BF16X2 && ((reinterpret_cast<uintptr_t>(a->in) % 16) || (reinterpret_cast<uintptr_t>(a->out) % 16)))
    return SP3D_ERR_INVALID_ARG;
  if ((a->c_pitch % vec) != 0) return SP3D_ERR_INVALID_ARG;
  const int64_t total = (int64_t)a->N * a->OD * a->OH * a->OW * (a->c_pitch / vec);
  if (total == 0) return SP3D_OK;
  const int64_t want = (total + 255) / 256;
  const int blocks = (int)(want < 148 * 32 ? want : 148 * 32);
  const bool k2s2 = a->dtype == SP3D_BF16 && a->k[0] == 2 && a->k[1] == 2 && a->k[2] == 2 && a->s[0] == 2 &&
                    a->s[1] == 2 && a->s[2] == 2 && a->p[0] == 0 && a->p[1] == 0 && a->p[2] == 0 &&
                    2 * a->OD <= a->D && 2 * a->OH <= a->H && 2 * a->OW <= a->W &&
                    (reinterpret_cast<uintptr_t>(a->in) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->out) % 16) == 0;
  if (a->dtype == SP3D_F32) maxpool_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  else if (a->dtype == SP3D_BF16X2) maxpool_pair_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  else if (k2s2) maxpool_bf16_k2s2_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  else maxpool_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
  return check_launch();
}

extern "C" int sp3d_layout_convert(const sp3d_layout_args* a, void* stream) {
  using namespace sp3d;
  if (a == nullptr || a->src == nullptr || a->dst == nullptr || a->C < 1 || a->c_pitch < a->C || a->N < 0 || a->S < 0)
    return SP3D_ERR_INVALID_ARG;
  if (a->N == 0 || a->S == 0) return SP3D_OK;
  if (a->N > 65535) return SP3D_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->src_dtype == SP3D_F32 && a->dst_dtype == SP3D_F32) return launch_layout<float, float>(a, st);
  if (a->src_dtype == SP3D_F32 && a->dst_dtype == SP3D_BF16) return launch_layout<float, __nv_bfloat16>(a, st);
  if (a->src_dtype == SP3D_BF16 && a->dst_dtype == SP3D_F32) return launch_layout<__nv_bfloat16, float>(a, st);
  return SP3D_ERR_UNSUPPORTED;
}
