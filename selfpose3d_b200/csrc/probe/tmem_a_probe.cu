// tmem_a_probe -- can the convolution's A operand live in TMEM?  (development tool, round 2)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_a_probe tmem_a_probe.cu
//
// Questions (answers recorded in profiles/r02_tmem_a_probe.log and DESIGN.md):
//   1. does `tcgen05.cp.128x256b` with a row-shifted, halo-pitched, swizzled K-major descriptor (the descriptor the
//      conv kernel builds for a tap) put a 128 x 16 bf16 A tile into TMEM such that `tcgen05.mma [d], [a], b-desc`
//      computes the same product as the smem-sourced form?
//   2. cycles per MMA with A in TMEM for N = 16 .. 128 (the smem-sourced form costs 32 + N/4: A is re-read per MMA)
//   3. cycles per tcgen05.cp, and per group "1 cp + m MMAs" (m = 1, 2, 3, 7): do copies overlap the MMAs?
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"

using namespace sp3d::tc;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]^T
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

static uint8_t* d_img;
static float* d_out;
static long long* d_cyc;

struct Params {
  uint32_t smem_bytes, a_off, b_off, a_sbo, b_sbo, layout, idesc, N, n_k, kstep;
  int mode;        // 0 correctness (cp + ts-mma), 1 ts-mma rate, 2 cp rate, 3 cp + m mma groups, 4 ss-mma rate (control)
  int m, repeat;
};

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ image, Params t, float* __restrict__ out,
                                                    long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid * 16; i < t.smem_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t tmA = tm + 256;          // A staging: 8 columns per 128 x 16 bf16 tile
  const uint32_t base = smem_u32(smem);
  if (tid == 0) {
    const long long t0 = clock64();
    if (t.mode == 0) {
      for (uint32_t k = 0; k < t.n_k; ++k) {
        const uint64_t da = make_smem_desc(base + t.a_off + k * t.kstep, 0, t.a_sbo, t.layout);
        const uint64_t db = make_smem_desc(base + t.b_off + k * t.kstep, 0, t.b_sbo, t.layout);
        tmem_cp_128x256b(tmA + (k & 1) * 8, da);
        mma_f16_ts(tm, tmA + (k & 1) * 8, db, t.idesc, k ? 1u : 0u);
      }
    } else {
      const uint64_t da = make_smem_desc(base + t.a_off, 0, t.a_sbo, t.layout);
      const uint64_t db = make_smem_desc(base + t.b_off, 0, t.b_sbo, t.layout);
      tmem_cp_128x256b(tmA, da);
      tmem_cp_128x256b(tmA + 8, da);
      for (int r = 0; r < t.repeat; ++r) {
        if (t.mode == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) mma_f16_ts(tm + (j & 1) * t.N, tmA, db, t.idesc, 1u);
        } else if (t.mode == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) tmem_cp_128x256b(tmA + (j & 3) * 8, da + (uint64_t)(2 * (j & 1)));
        } else if (t.mode == 3) {
          tmem_cp_128x256b(tmA + (r & 1) * 8, da + (uint64_t)(2 * (r & 1)));
          for (int j = 0; j < t.m; ++j) mma_f16_ts(tm + (j & 1) * t.N, tmA + (r & 1) * 8, db, t.idesc, 1u);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) mma_f16_ss(tm + (j & 1) * t.N, da + (uint64_t)(2 * (j & 1)), db, t.idesc, 1u);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[0] = clock64() - t0;
  }
  __syncthreads();
  tc_fence_after();
  if (t.mode == 0) {
    for (uint32_t n0 = 0; n0 < t.N; n0 += 8) {
      uint32_t v[8];
      tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + n0, v);
      tmem_ld_wait();
      for (int j = 0; j < 8; ++j) out[(size_t)tid * t.N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}


// ---------------------------------------------------------------------------------------------- multi-issuer rates
// `issuers` warps run the loop warp-uniformly (elect.sync-predicated issue, as the conv kernel does); each owns
// accumulator columns and four 8-column A slots.  mode 0: smem-sourced MMAs (control), 1: TMEM-sourced MMAs,
// 2: copies only, 3: software-pipelined groups "cp(slot g+D) ; m MMAs reading slot g".
template <int N>
__global__ void __launch_bounds__(128) rate_kernel(uint32_t layout, uint32_t row_bytes, uint32_t pitch_rows, int issuers, int mode,
                                                   int m, int dist, int repeat, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = (16 * pitch_rows + 32) * row_bytes, b_bytes = 8 * N * row_bytes;
  for (uint32_t i = tid * 16; i < a_bytes + b_bytes + 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmD = tmem_base + (uint32_t)warp * 96;            // 2 accumulators of <= 48 columns... N <= 32: 2 x N
  const uint32_t tmA = tmem_base + 400 + (uint32_t)warp * 28;      // 3 A slots x 8 columns (+ slack)
  const uint32_t base = smem_u32(smem);
  const uint32_t b_base = base + ((a_bytes + 1023) & ~1023u);
  const uint32_t idesc = make_idesc(kFmtBF16, 128, N);
  const uint64_t da = make_smem_desc(base + 3 * row_bytes, 0, pitch_rows * row_bytes, layout);
  const uint64_t db = make_smem_desc(b_base, 0, 8 * row_bytes, layout);
  const uint32_t nacc = N <= 48 ? 2 : 1;
  if (warp < issuers) {
    if (elect_one_sync()) {
      tmem_cp_128x256b(tmA, da);
      tmem_cp_128x256b(tmA + 8, da);
      tmem_cp_128x256b(tmA + 16, da);
    }
    __syncwarp();
    const long long t0 = clock64();
    for (int r = 0; r < repeat; ++r) {
      if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (elect_one_sync()) mma_f16_ss(tmD + (j % nacc) * N, da + (uint64_t)((j + 1) * (row_bytes >> 4)), db + (uint64_t)(j * (N * row_bytes >> 4)), idesc, 1u);
      } else if (mode == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (elect_one_sync()) mma_f16_ts(tmD + (j % nacc) * N, tmA + (j % 3) * 8, db + (uint64_t)(j * (N * row_bytes >> 4)), idesc, 1u);
      } else if (mode == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (elect_one_sync()) tmem_cp_128x256b(tmA + (j % 3) * 8, da + (uint64_t)((j + 1) * (row_bytes >> 4)));
      } else {
        if (elect_one_sync()) tmem_cp_128x256b(tmA + ((r + dist) % 3) * 8, da + (uint64_t)((r & 7) * (row_bytes >> 4)));
        for (int j = 0; j < m; ++j)
          if (elect_one_sync()) mma_f16_ts(tmD + (j % nacc) * N, tmA + (r % 3) * 8, db + (uint64_t)((j & 7) * (N * row_bytes >> 4)), idesc, 1u);
      }
    }
    if (elect_one_sync()) mma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    if (lane == 0) cycles[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N>
static void rates(uint32_t layout, int row_bytes, int pitch_rows) {
  CK(cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int smem = (16 * pitch_rows + 32) * row_bytes + 8 * N * row_bytes + 4096;
  const int repeat = 512;
  auto run = [&](int issuers, int mode, int m, int dist) {
    CK(cudaMemset(d_cyc, 0, 32));
    rate_kernel<N><<<1, 128, smem, 0>>>(layout, row_bytes, pitch_rows, issuers, mode, m, dist, repeat, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
    long long cyc[4];
    CK(cudaMemcpy(cyc, d_cyc, 32, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < issuers; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    return (double)mx;
  };
  for (int issuers : {1, 2, 4}) {
    const double ss = run(issuers, 0, 0, 0) / (repeat * 8) / issuers, ts = run(issuers, 1, 0, 0) / (repeat * 8) / issuers,
                 cp = run(issuers, 2, 0, 0) / (repeat * 8) / issuers;
    printf("rate3 row=%3dB N=%3d issuers=%d : SS mma %6.1f   TS mma %6.1f   cp %6.1f  cycles per op per SM (math floor %d)\n", row_bytes, N,
           issuers, ss, ts, cp, N / 2);
  }
  for (int issuers : {1, 2, 4})
    for (int dist : {1, 2})
      for (int m : {1, 2, 3, 4, 7}) {
        const double g = run(issuers, 3, m, dist) / repeat / issuers;
        printf("rate3 row=%3dB N=%3d issuers=%d dist=%d : 1 cp + %d TS mma = %6.1f cycles per group per SM = %5.1f per MMA (SS form: %d)\n",
               row_bytes, N, issuers, dist, m, g, g / m, 32 + N / 4);
      }
}

// ---------------------------------------------------------------------------------------------- mixed-layout SS rate
// The z-folded stem reads A from 64-byte rows (SWIZZLE_64B, halo pitch 12) and B from 32-byte rows (SWIZZLE_32B).
// Cycles per smem-sourced MMA for A fixed as in the stem and B in 32 / 64 / 128-byte rows, N = 32 .. 128, 2 issuers.
template <int N>
__global__ void __launch_bounds__(128) mixed_rate_kernel(uint32_t b_layout, uint32_t b_row_bytes, int issuers, int repeat,
                                                         long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_row = 64, a_pitch = 12;
  const uint32_t a_bytes = (16 * a_pitch + 40) * a_row * 4, b_bytes = 8 * N * b_row_bytes;
  for (uint32_t i = tid * 16; i < a_bytes + b_bytes + 2048; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmD = tmem_base + (uint32_t)warp * 128;
  const uint32_t base = smem_u32(smem);
  const uint32_t b_base = base + ((a_bytes + 1023) & ~1023u);
  const uint32_t idesc = make_idesc(kFmtBF16, 128, N);
  const uint64_t da = make_smem_desc(base + 3 * a_row, 0, a_pitch * a_row, kSwizzle64);
  const uint64_t db = make_smem_desc(b_base, 0, 8 * b_row_bytes, b_layout);
  if (warp < issuers) {
    const long long t0 = clock64();
    for (int r = 0; r < repeat; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j)     // A: another x-slice / tap window per MMA (as the conv kernel), B: the next tap
        if (elect_one_sync())
          mma_f16_ss(tmD, da + (uint64_t)((j * 37 + warp * 264 * 4 + 1) * (32 >> 4)), db + (uint64_t)(j * (N * b_row_bytes >> 4)), idesc, 1u);
    }
    if (elect_one_sync()) mma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    if (lane == 0) cycles[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N>
static void mixed_rates() {
  CK(cudaFuncSetAttribute(mixed_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  struct L { const char* name; uint32_t layout, row; } ls[] = {{"SW32/32B", kSwizzle32, 32}, {"SW64/64B", kSwizzle64, 64}, {"SW128/128B", kSwizzle128, 128}};
  for (auto& l : ls) {
    const int smem = (16 * 12 + 40) * 64 * 4 + 8 * N * (int)l.row + 4096;
    if (smem > 220 * 1024) continue;
    for (int issuers : {1, 2}) {
      CK(cudaMemset(d_cyc, 0, 32));
      mixed_rate_kernel<N><<<1, 128, smem, 0>>>(l.layout, l.row, issuers, 512, d_cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mixed rate kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
      long long cyc[4];
      CK(cudaMemcpy(cyc, d_cyc, 32, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < issuers; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      printf("rate4 A=SW64/64B rows (halo pitch 12), B=%-10s N=%3d issuers=%d : %6.1f cycles per MMA per SM (model 32 + N/4 = %d)\n",
             l.name, N, issuers, (double)mx / (512 * 8) / issuers, 32 + N / 4);
    }
  }
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static uint32_t swz(uint32_t off, uint32_t layout) {
  switch (layout) {
    case kSwizzle128: return off ^ (((off >> 7) & 7) << 4);
    case kSwizzle64: return off ^ (((off >> 7) & 3) << 4);
    case kSwizzle32: return off ^ (((off >> 7) & 1) << 4);
    default: return off;
  }
}


static int correctness(const char* name, uint32_t layout, int K, int N, int shift_rows, int pitch_rows) {
  const int M = 128, row_bytes = K * 2;
  const int a_rows = 16 * pitch_rows + shift_rows + 8;
  std::vector<float> A((size_t)a_rows * K), B((size_t)N * K);
  srand(99 + N + K + shift_rows);
  for (auto& v : A) v = (float)((rand() % 7) - 3);
  for (auto& v : B) v = (float)((rand() % 5) - 2);
  const uint32_t a_region = (uint32_t)((a_rows * row_bytes + 1023) / 1024 * 1024);
  const uint32_t b_region = (uint32_t)((N * row_bytes + 1023) / 1024 * 1024);
  std::vector<uint8_t> img(a_region + b_region, 0);
  for (int r = 0; r < a_rows; ++r)
    for (int k = 0; k < K; ++k) {
      uint16_t h = f2bf(A[(size_t)r * K + k]);
      memcpy(&img[swz((uint32_t)(r * row_bytes + k * 2), layout)], &h, 2);
    }
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < K; ++k) {
      uint16_t h = f2bf(B[(size_t)r * K + k]);
      memcpy(&img[a_region + swz((uint32_t)(r * row_bytes + k * 2), layout)], &h, 2);
    }
  Params t{};
  t.smem_bytes = (uint32_t)img.size();
  t.a_off = (uint32_t)(shift_rows * row_bytes);
  t.b_off = a_region;
  t.a_sbo = (uint32_t)(pitch_rows * row_bytes);
  t.b_sbo = (uint32_t)(8 * row_bytes);
  t.layout = layout;
  t.idesc = make_idesc(kFmtBF16, M, N);
  t.N = N;
  t.n_k = K / 16;
  t.kstep = 32;
  t.mode = 0;
  CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, (size_t)M * 256 * 4));
  probe_kernel<<<1, 128, img.size(), 0>>>(d_img, t, d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-40s LAUNCH FAILED: %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> out((size_t)M * N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double max_err = 0;
  for (int m = 0; m < M; ++m) {
    const int r = (m / 8) * pitch_rows + (m % 8) + shift_rows;
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)r * K + k] * B[(size_t)n * K + k];
      const double err = fabs(ref - out[(size_t)m * N + n]);
      max_err = err > max_err ? err : max_err;
      if (err > 1e-3) ++bad;
    }
  }
  printf("cp+ts %-32s K=%3d N=%3d shift=%2d pitch=%2d : %s max_err=%.3g bad=%d/%d\n", name, K, N, shift_rows, pitch_rows,
         bad == 0 ? "OK  " : "FAIL", max_err, bad, M * N);
  return bad != 0;
}

static void rate(const char* what, int mode, uint32_t layout, int row_bytes, int N, int m, int pitch_rows) {
  Params t{};
  const uint32_t a_region = (uint32_t)(((16 * pitch_rows + 24) * row_bytes + 1023) / 1024 * 1024);
  t.smem_bytes = a_region + (uint32_t)((N * row_bytes + 1023) / 1024 * 1024);
  t.a_off = (uint32_t)(pitch_rows == 8 ? 0 : 3 * row_bytes);
  t.b_off = a_region;
  t.a_sbo = (uint32_t)(pitch_rows * row_bytes);
  t.b_sbo = (uint32_t)(8 * row_bytes);
  t.layout = layout;
  t.idesc = make_idesc(kFmtBF16, 128, N);
  t.N = N;
  t.mode = mode;
  t.m = m;
  t.repeat = 512;
  CK(cudaMemset(d_img, 0, t.smem_bytes));
  probe_kernel<<<1, 128, t.smem_bytes, 0>>>(d_img, t, d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s LAUNCH FAILED: %s\n", what, cudaGetErrorString(e));
    exit(3);
  }
  long long cyc = 0;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  if (mode == 3)
    printf("rate %-22s row=%3dB pitch=%2d N=%3d : %7.1f cycles per group (1 cp + %d MMAs) = %.1f per MMA (math floor %d)\n", what,
           row_bytes, pitch_rows, N, (double)cyc / t.repeat, m, (double)cyc / t.repeat / m, N / 2);
  else
    printf("rate %-22s row=%3dB pitch=%2d N=%3d : %7.1f cycles per op (math floor %d)\n", what, row_bytes, pitch_rows, N,
           (double)cyc / (t.repeat * 8), N / 2);
}

int main() {
  CK(cudaMalloc(&d_img, 256 * 1024));
  CK(cudaMalloc(&d_out, 128 * 256 * 4));
  CK(cudaMalloc(&d_cyc, 32));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (getenv("PROBE_MIXED")) {
    mixed_rates<32>();
    mixed_rates<64>();
    mixed_rates<128>();
    return 0;
  }
  int fails = 0;
  fails += correctness("SW128 dense", kSwizzle128, 64, 32, 0, 8);
  fails += correctness("SW64 dense", kSwizzle64, 32, 32, 0, 8);
  fails += correctness("SW32 dense", kSwizzle32, 16, 16, 0, 8);
  for (int shift : {1, 3, 11}) {
    fails += correctness("SW128 shift + halo pitch", kSwizzle128, 64, 64, shift, 10);
    fails += correctness("SW64 shift + halo pitch", kSwizzle64, 32, 32, shift, 12);
    fails += correctness("SW32 shift + halo pitch", kSwizzle32, 16, 16, shift, 14);
  }
  for (int N : {16, 32, 64, 128}) {
    rate("mma A=smem (control)", 4, kSwizzle64, 64, N, 0, 12);
    rate("mma A=tmem", 1, kSwizzle64, 64, N, 0, 12);
  }
  rate("cp 128x256b", 2, kSwizzle64, 64, 32, 0, 12);
  rate("cp 128x256b", 2, kSwizzle128, 128, 32, 0, 10);
  rate("cp 128x256b", 2, kSwizzle32, 32, 32, 0, 14);
  for (int N : {32, 64})
    for (int m : {1, 2, 3, 4, 7}) rate("cp + m mma", 3, kSwizzle64, 64, N, m, 12);
  rates<32>(kSwizzle64, 64, 12);
  rates<64>(kSwizzle128, 128, 10);
  rates<16>(kSwizzle32, 32, 14);
  printf("tmem_a_probe finished, %d failing cases\n", fails);
  return 0;
}
