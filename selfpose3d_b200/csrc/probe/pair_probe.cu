// Probe: tcgen05.mma.cta_group::2 (one MMA over the two SMs of a CTA pair) with shared-memory operands.
//   * correctness: each CTA of a 2-CTA cluster holds ITS 128 rows of A and ITS half (N/2 rows) of B at the same
//     shared-memory offsets; the leader issues M = 256 MMAs; each CTA reads its 128 accumulator rows from its own TMEM;
//   * rate: cycles per MMA per SM for N = 32 .. 256 against the one-CTA form (32 + N/4 cycles per K = 16 step: the
//     shared-memory read of A (4 KB) plus B (N x 32 B) at 128 B/clk).  In the pair form an SM reads A plus HALF of B.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pair_probe pair_probe.cu     (run on a B200)
#include "../tc_common.cuh"
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

using namespace sp3d::tc;
namespace cg = cooperative_groups;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of the mask when the pair's MMAs have completed
__device__ __forceinline__ void mma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

struct Params {
  uint32_t a_bytes, b_bytes;   // per CTA
  int N, n_k, repeat, mode;    // mode 0: correctness, 1: rate (pair), 2: rate (one CTA, control)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
pair_kernel(const uint8_t* __restrict__ image, Params t, float* __restrict__ out, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_region = (t.a_bytes + 1023) & ~1023u;
  const uint32_t total = a_region + t.b_bytes;
  const uint8_t* src = image + (size_t)rank * total;
  for (uint32_t i = tid * 16; i < total; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(src + i);
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (t.mode == 2) { tmem_alloc(&tmem_base, 512); tmem_relinquish(); }
    else { tmem_alloc2(&tmem_base, 512); tmem_relinquish2(); }
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint64_t da = make_smem_desc(smem_u32(smem), 0, 1024, kSwizzle128);
  const uint64_t db = make_smem_desc(smem_u32(smem) + a_region, 0, 1024, kSwizzle128);
  if (t.mode == 0) {
    if (rank == 0 && warp == 1) {
      const uint32_t idesc = make_idesc(kFmtBF16, 256, t.N);
      for (int k = 0; k < t.n_k; ++k)
        if (elect_one_sync()) mma2_f16_ss(tm, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
      if (elect_one_sync()) mma2_commit_mc(&bar[0], 3);
      __syncwarp();
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    for (int c0 = 0; c0 < t.N; c0 += 8) {
      uint32_t v[8];
      tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 8; ++j) out[((size_t)rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(v[j]);
    }
  } else {
    const bool pair = t.mode == 1;
    if ((pair ? rank == 0 : true) && warp == 1) {
      const uint32_t idesc = make_idesc(kFmtBF16, pair ? 256 : 128, t.N);
      const long long t0 = clock64();
      for (int r = 0; r < t.repeat; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (elect_one_sync()) {
            if (pair) mma2_f16_ss(tm + (j & 1) * 256, da + 2 * (j & 3), db + 2 * (j & 3), idesc, 1u);
            else mma_f16_ss(tm + (j & 1) * 256, da + 2 * (j & 3), db + 2 * (j & 3), idesc, 1u);
          }
        }
      }
      if (elect_one_sync()) {
        if (pair) mma2_commit_mc(&bar[1], 3);
        else mma_commit(&bar[1]);
      }
      __syncwarp();
      mbar_wait(&bar[1], 0);
      if (lane == 0) cycles[rank] = clock64() - t0;
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    if (t.mode == 2) tmem_dealloc(tm, 512);
    else tmem_dealloc2(tm, 512);
  }
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7) << 4); }

static uint8_t* d_img;
static float* d_out;
static long long* d_cyc;

static int correctness(int N) {
  const int K = 64, row_bytes = 128, half = N / 2;
  const uint32_t a_bytes = 128 * row_bytes, b_bytes = (uint32_t)((half * row_bytes + 1023) / 1024 * 1024);
  const uint32_t total = a_bytes + b_bytes;
  std::vector<float> A(2 * 128 * K), B((size_t)N * K);
  srand(7 + N);
  for (auto& v : A) v = (float)((rand() % 7) - 3);
  for (auto& v : B) v = (float)((rand() % 5) - 2);
  std::vector<uint8_t> img(2 * total, 0);
  for (int r = 0; r < 2; ++r) {
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < K; ++k) {
        uint16_t h = f2bf(A[((size_t)r * 128 + m) * K + k]);
        memcpy(&img[(size_t)r * total + swz128((uint32_t)(m * row_bytes + k * 2))], &h, 2);
      }
    for (int n = 0; n < half; ++n)
      for (int k = 0; k < K; ++k) {
        uint16_t h = f2bf(B[((size_t)r * half + n) * K + k]);
        memcpy(&img[(size_t)r * total + a_bytes + swz128((uint32_t)(n * row_bytes + k * 2))], &h, 2);
      }
  }
  Params t{a_bytes, b_bytes, N, K / 16, 0, 0};
  CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, 2 * 128 * 256 * 4));
  pair_kernel<<<2, 128, total + 1024>>>(d_img, t, d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("pair N=%d LAUNCH FAILED: %s\n", N, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> out(2 * 128 * 256);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double max_err = 0;
  for (int r = 0; r < 2; ++r)
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[((size_t)r * 128 + m) * K + k] * B[(size_t)n * K + k];
        const double err = fabs(ref - out[((size_t)r * 128 + m) * 256 + n]);
        max_err = err > max_err ? err : max_err;
        if (err > 1e-3) ++bad;
      }
  printf("pair MMA M=256 K=64 N=%3d (B halves of %3d rows per CTA): %s max_err=%.3g bad=%d/%d\n", N, half, bad == 0 ? "OK  " : "FAIL",
         max_err, bad, 2 * 128 * N);
  return bad != 0;
}

static void rate(int N, int mode) {
  const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)(((mode == 1 ? N / 2 : N) * 128 + 1023) / 1024 * 1024);
  Params t{a_bytes, b_bytes, N, 4, 512, mode};
  CK(cudaMemset(d_img, 0, 2 * (a_bytes + b_bytes)));
  CK(cudaMemset(d_cyc, 0, 32));
  pair_kernel<<<2, 128, a_bytes + b_bytes + 1024>>>(d_img, t, d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("rate N=%d mode=%d LAUNCH FAILED: %s\n", N, mode, cudaGetErrorString(e));
    exit(3);
  }
  long long cyc[4];
  CK(cudaMemcpy(cyc, d_cyc, 32, cudaMemcpyDeviceToHost));
  const double per = (double)cyc[0] / (512 * 8);
  if (mode == 1)
    printf("rate pair (cta_group::2, M=256) N=%3d : %6.1f cycles per MMA (one-CTA model 32 + N/4 = %d, A + B/2 model %d, math floor %d)\n", N, per,
           32 + N / 4, 32 + N / 8, N / 4);
  else
    printf("rate one CTA (control, both SMs busy) N=%3d : %6.1f cycles per MMA (model %d)\n", N, per, 32 + N / 4);
}

int main() {
  CK(cudaMalloc(&d_img, 512 * 1024));
  CK(cudaMalloc(&d_out, 2 * 128 * 256 * 4));
  CK(cudaMalloc(&d_cyc, 32));
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  int fails = 0;
  for (int N : {32, 64, 128, 256}) fails += correctness(N);
  for (int N : {32, 64, 128, 256}) {
    rate(N, 2);
    rate(N, 1);
  }
  printf("pair_probe finished, %d failing cases\n", fails);
  return 0;
}
