// tc_probe -- hardware fact-finding for the tcgen05 convolution kernels (development tool).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tc_probe.cu -lcuda
//
// The host builds a byte image of shared memory (operand tiles laid out with a software model of
// the swizzle), the kernel copies it verbatim into shared memory, one thread issues tcgen05.mma
// with host-supplied descriptors, and the accumulator is read back and compared with a CPU GEMM.
// Questions answered (results recorded in DESIGN.md):
//   1. descriptor encodings for K-major bf16/tf32 tiles in SWIZZLE_{NONE,32,64,128}B
//   2. may the A start address be shifted by whole rows (im2col-free tap shifts inside a smem halo brick)
//      and may SBO be an arbitrary row pitch (8-row groups of a brick with halo)?
//   3. tcgen05.mma issue rate versus N for smem-sourced operands
//   4. TMA 5-D tiled loads with swizzle and out-of-bounds zero fill produce the layout the model assumes
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../tc_common.cuh"

using namespace sp3d::tc;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

struct MmaTest {
  uint32_t smem_bytes;        // image size
  uint32_t a_off, b_off;      // operand start offsets inside the image (bytes)
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t layout_type, a_base_off, b_base_off;
  uint32_t idesc;
  uint32_t n_k;               // MMAs per accumulation chain
  uint32_t a_kstep, b_kstep;  // start-address advance per MMA (bytes)
  uint32_t N;
  uint32_t tf32;
  uint32_t repeat;            // timing: repeat the chain this many times
};

__global__ void __launch_bounds__(128) mma_probe_kernel(const uint8_t* __restrict__ image, MmaTest t,
                                                        float* __restrict__ out, long long* __restrict__ cycles) {
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
    tmem_alloc(&tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t base = smem_u32(smem);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (uint32_t r = 0; r < t.repeat; ++r) {
      for (uint32_t k = 0; k < t.n_k; ++k) {
        const uint64_t da = make_smem_desc(base + t.a_off + k * t.a_kstep, t.a_lbo, t.a_sbo, t.layout_type, t.a_base_off);
        const uint64_t db = make_smem_desc(base + t.b_off + k * t.b_kstep, t.b_lbo, t.b_sbo, t.layout_type, t.b_base_off);
        if (t.tf32) mma_tf32_ss(tm, da, db, t.idesc, (r | k) ? 1u : 0u);
        else mma_f16_ss(tm, da, db, t.idesc, (r | k) ? 1u : 0u);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    cycles[0] = (base & 1023u) ? -(long long)(base & 1023u) : t1 - t0;   // negative: dynamic smem base not 1024-aligned
  }
  tc_fence_after();
  // accumulator row m = TMEM lane m; this warp owns lanes 32*warp .. 32*warp+31
  for (uint32_t n0 = 0; n0 < t.N; n0 += 8) {
    uint32_t v[8];
    tmem_ld_x8(tm + ((uint32_t)(warp * 32) << 16) + n0, v);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) out[(size_t)tid * t.N + n0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ---------------------------------------------------------------------------------------------- host model
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}

// Address-based swizzle model: 16-byte chunk index bits XORed with higher address bits.
//   SW128: bits[4:6] ^= bits[7:9]   SW64: bits[4:5] ^= bits[7:8]   SW32: bit[4] ^= bit[7]
static uint32_t swz(uint32_t off, uint32_t layout_type) {
  switch (layout_type) {
    case kSwizzle128: return off ^ (((off >> 7) & 7) << 4);
    case kSwizzle64: return off ^ (((off >> 7) & 3) << 4);
    case kSwizzle32: return off ^ (((off >> 7) & 1) << 4);
    default: return off;
  }
}

struct Case {
  const char* name;
  uint32_t layout_type;
  int elem_bytes;       // 2 = bf16, 4 = tf32
  int K;                // K extent of the tile = one swizzle row (or any for none)
  int N;
  int shift_rows;       // A start shifted by this many rows
  int group_pitch_rows; // rows between consecutive 8-row groups of A in the buffer (8 = dense)
  int base_off_mode;    // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
  int repeat;
};

static int run_case(const Case& c, bool verbose) {
  const int M = 128, K = c.K, N = c.N, eb = c.elem_bytes;
  const int row_bytes = (c.layout_type == kSwizzleNone) ? 16 : K * eb;   // none: rows of one 16-byte core row
  const int kchunks = (c.layout_type == kSwizzleNone) ? (K * eb / 16) : 1;
  // logical operands (small integers: exact in bf16 / tf32)
  const int a_rows_buf = 16 * c.group_pitch_rows + c.shift_rows + 8;
  std::vector<float> A((size_t)a_rows_buf * K), B((size_t)N * K);
  srand(1234 + N + K);
  for (auto& v : A) v = (float)((rand() % 7) - 3);
  for (auto& v : B) v = (float)((rand() % 5) - 2);
  // image: A region at 0 (1024-aligned), B region after it (1024-aligned)
  const uint32_t a_region = (uint32_t)((a_rows_buf * row_bytes * kchunks + 1023) / 1024 * 1024);
  const uint32_t b_region = (uint32_t)((N * row_bytes * kchunks + 1023) / 1024 * 1024);
  std::vector<uint8_t> img(a_region + b_region, 0);
  auto put = [&](uint32_t region_off, int rows_total, int r, int k, float val) {
    uint32_t off;
    if (c.layout_type == kSwizzleNone) {
      const int chunk = (k * eb) / 16, within = (k * eb) % 16;
      off = (uint32_t)(chunk * rows_total * 16 + r * 16 + within);   // [kchunk][row][16 B]
    } else {
      off = swz((uint32_t)(r * row_bytes + k * eb), c.layout_type);
    }
    if (eb == 2) {
      uint16_t h = f2bf(val);
      memcpy(&img[region_off + off], &h, 2);
    } else {
      memcpy(&img[region_off + off], &val, 4);
    }
  };
  for (int r = 0; r < a_rows_buf; ++r)
    for (int k = 0; k < K; ++k) put(0, a_rows_buf, r, k, A[(size_t)r * K + k]);
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < K; ++k) put(a_region, N, r, k, B[(size_t)r * K + k]);

  MmaTest t{};
  t.smem_bytes = (uint32_t)img.size();
  t.layout_type = c.layout_type;
  t.N = N;
  t.tf32 = (eb == 4);
  t.repeat = c.repeat;
  const int umma_k = 32 / eb;                    // 16 for bf16, 8 for tf32
  t.n_k = K / umma_k;
  t.idesc = make_idesc(eb == 2 ? kFmtBF16 : kFmtTF32, M, N);
  t.a_off = (uint32_t)(c.shift_rows * row_bytes);
  t.b_off = a_region;
  if (c.layout_type == kSwizzleNone) {
    t.a_lbo = (uint32_t)(a_rows_buf * 16);       // next 16-byte K chunk
    t.b_lbo = (uint32_t)(N * 16);
    t.a_sbo = (uint32_t)(c.group_pitch_rows * 16);
    t.b_sbo = 8 * 16;
    t.a_kstep = 2 * t.a_lbo;                     // one MMA consumes two 16-byte K chunks (32 bytes of K)
    t.b_kstep = 2 * t.b_lbo;
  } else {
    t.a_lbo = t.b_lbo = 0;
    t.a_sbo = (uint32_t)(c.group_pitch_rows * row_bytes);
    t.b_sbo = (uint32_t)(8 * row_bytes);
    t.a_kstep = t.b_kstep = 32;                  // advance 32 bytes of K inside the swizzled row
  }
  t.a_base_off = c.base_off_mode ? ((t.a_off >> 7) & 7) : 0;
  t.b_base_off = 0;

  uint8_t* d_img;
  float* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_img, img.size()));
  CK(cudaMalloc(&d_out, (size_t)M * N * 4));
  CK(cudaMalloc(&d_cyc, 8));
  CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, (size_t)M * N * 4));
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  mma_probe_kernel<<<1, 128, img.size(), 0>>>(d_img, t, d_out, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s LAUNCH FAILED: %s\n", c.name, cudaGetErrorString(e));
    exit(3);  // sticky context error: stop here
  }
  std::vector<float> out((size_t)M * N);
  long long cyc = 0;
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
  double max_err = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m) {
    const int r = (m / 8) * c.group_pitch_rows + (m % 8) + c.shift_rows;
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)r * K + k] * B[(size_t)n * K + k];
      ref *= c.repeat;
      const double err = fabs(ref - out[(size_t)m * N + n]);
      if (err > max_err) max_err = err;
      if (err > 1e-3) ++bad;
    }
  }
  printf("%-44s N=%3d K=%3d shift=%d pitch=%2d bo=%d : %s max_err=%.3g bad=%d/%d  cycles=%lld (%.1f per MMA over %u)\n",
         c.name, N, K, c.shift_rows, c.group_pitch_rows, c.base_off_mode, bad == 0 ? "OK  " : "FAIL", max_err, bad,
         M * N, cyc, (double)cyc / (t.n_k * t.repeat), t.n_k * t.repeat);
  (void)verbose;
  cudaFree(d_img);
  cudaFree(d_out);
  cudaFree(d_cyc);
  return bad == 0 ? 0 : 1;
}


// ---------------------------------------------------------------------------------------------- issue-rate probe
// `issuers` warps each issue repeat*8 MMAs (compile-time unrolled descriptor offsets) into their own accumulator.
template <int N, bool TF32>
__global__ void __launch_bounds__(128) rate_kernel(uint32_t layout_type, uint32_t row_bytes, int issuers, int repeat,
                                                   int pitch_rows, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = (16 * pitch_rows + 16) * row_bytes, b_bytes = N * row_bytes;
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
  const uint32_t tm = tmem_base + (uint32_t)warp * (N >= 128 ? 128 : N);
  const uint32_t base = smem_u32(smem);
  const uint32_t b_base = base + ((a_bytes + 1023) & ~1023u);
  const uint32_t idesc = make_idesc(TF32 ? kFmtTF32 : kFmtBF16, 128, N);
  const uint64_t da = make_smem_desc(base, 0, pitch_rows * row_bytes, layout_type);
  const uint64_t db = make_smem_desc(b_base, 0, 8 * row_bytes, layout_type);
  const int ksteps = row_bytes / 32;   // MMAs per swizzled row
  long long t0 = clock64();
  if (warp < issuers && lane == 0) {
    for (int r = 0; r < repeat; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint64_t off = (uint64_t)((j % ksteps) * 2);
        // pitch_rows != 8 mimics the conv kernel: every MMA starts at a different (unaligned) halo row
        const uint64_t shift = pitch_rows == 8 ? 0 : (uint64_t)((j + 1) * (row_bytes >> 4));
        if (TF32) mma_tf32_ss(tm, da + off + shift, db + off, idesc, 1u);
        else mma_f16_ss(tm, da + off + shift, db + off, idesc, 1u);
      }
    }
    mma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    cycles[warp] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, bool TF32>
static void run_rate(const char* name, uint32_t layout_type, uint32_t row_bytes, int pitch_rows = 8) {
  long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 32));
  CK(cudaFuncSetAttribute(rate_kernel<N, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int issuers : {1, 2, 4}) {
    if (issuers * (N >= 128 ? 128 : N) > 512) continue;
    const int repeat = 256;
    CK(cudaMemset(d_cyc, 0, 32));
    rate_kernel<N, TF32><<<1, 128, (16 * pitch_rows + 16) * row_bytes + N * row_bytes + 4096, 0>>>(layout_type, row_bytes, issuers, repeat,
                                                                                                   pitch_rows, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate kernel failed: %s\n", cudaGetErrorString(e)); exit(3); }
    long long cyc[4];
    CK(cudaMemcpy(cyc, d_cyc, 32, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < issuers; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    const double per = (double)mx / (repeat * 8);
    const int kk = TF32 ? 8 : 16;
    printf("rate2 %-18s pitch=%2d N=%3d issuers=%d : %7.1f cycles per MMA per issuer, %6.1f MACs/cycle/SM (floor %d cyc)\n", name, pitch_rows, N,
           issuers, per, issuers * 128.0 * N * kk / per, 128 * N / 256 * (TF32 ? 1 : 1));
  }
  cudaFree(d_cyc);
}

// ---------------------------------------------------------------------------------------------- TMA probe
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) tma_probe_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2,
                                                        int c3, int c4, uint32_t bytes, uint8_t* __restrict__ dump) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (uint32_t i = tid * 4; i < bytes; i += 128 * 4) *reinterpret_cast<uint32_t*>(smem + i) = 0xDEADBEEFu;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_5d(smem, &map, &bar, c0, c1, c2, c3, c4);
  }
  mbar_wait(&bar, 0);
  for (uint32_t i = tid; i < bytes; i += 128) dump[i] = smem[i];
}

static int run_tma(uint32_t layout_type, int C, bool overlap_probe) {
  // tensor [N=2][X=6][Y=9][Z=12][C] bf16, box = {C, 10, 7, 4, 1} at (0, -1, -1, -1, 1): halo with OOB on the low side
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int Nn = 2, X = 6, Y = 9, Z = 12;
  std::vector<uint16_t> h((size_t)Nn * X * Y * Z * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = f2bf((float)((i * 7 + 3) % 251));
  uint16_t* d;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap map;
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)Nn};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Z, (cuuint64_t)C * 2 * Z * Y, (cuuint64_t)C * 2 * Z * Y * X};
  cuuint32_t box[5] = {(cuuint32_t)C, 10, 7, 4, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (overlap_probe) {  // dims {C, dz = 3, Z, Y, X} with the dz stride equal to the Z stride (overlapping window)
    gdim[1] = 3; gdim[2] = Z - 2; gdim[3] = Y; gdim[4] = X;
    gstr[0] = (cuuint64_t)C * 2; gstr[1] = (cuuint64_t)C * 2; gstr[2] = (cuuint64_t)C * 2 * Z; gstr[3] = (cuuint64_t)C * 2 * Z * Y;
    box[1] = 3; box[2] = 4; box[3] = 2; box[4] = 1;
  }
  CUtensorMapSwizzle sw = layout_type == kSwizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : layout_type == kSwizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : layout_type == kSwizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("TMA encode (layout %u, C=%d, overlap=%d): cuTensorMapEncodeTiled failed with %d\n", layout_type, C, overlap_probe, (int)r);
    cudaFree(d);
    return overlap_probe ? 0 : 1;
  }
  if (overlap_probe) {
    printf("TMA encode with overlapping strides (dz window): accepted by the driver\n");
    cudaFree(d);
    return 0;
  }
  const uint32_t rows = 10 * 7 * 4, row_bytes = (uint32_t)C * 2, bytes = rows * row_bytes;
  uint8_t* dump;
  CK(cudaMalloc(&dump, bytes));
  CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tma_probe_kernel<<<1, 128, bytes, 0>>>(map, 0, -1, -1, -1, 1, bytes, dump);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("TMA probe launch failed: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<uint8_t> got(bytes);
  CK(cudaMemcpy(got.data(), dump, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int bx = 0; bx < 4; ++bx)
    for (int by = 0; by < 7; ++by)
      for (int bz = 0; bz < 10; ++bz)
        for (int c = 0; c < C; ++c) {
          const int x = bx - 1, y = by - 1, z = bz - 1;
          uint16_t want = 0;
          if (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z)
            want = h[((((size_t)1 * X + x) * Y + y) * Z + z) * C + c];
          const uint32_t row = (uint32_t)((bx * 7 + by) * 10 + bz);
          const uint32_t off = swz(row * row_bytes + (uint32_t)c * 2, layout_type);
          uint16_t g;
          memcpy(&g, &got[off], 2);
          if (g != want) ++bad;
        }
  printf("TMA 5-D halo load, layout %u, C=%d (row %u B), OOB zero fill: %s (%d mismatches of %u)\n", layout_type, C,
         row_bytes, bad == 0 ? "OK" : "FAIL", bad, rows * C);
  cudaFree(d);
  cudaFree(dump);
  return bad != 0;
}

int main(int argc, char** argv) {
  int fails = 0;
  const bool timing = argc > 1 && !strcmp(argv[1], "timing");
  if (argc > 1 && !strcmp(argv[1], "rate2")) {
    run_rate<16, false>("bf16 SW128", kSwizzle128, 128);
    run_rate<32, false>("bf16 SW128", kSwizzle128, 128);
    run_rate<64, false>("bf16 SW128", kSwizzle128, 128);
    run_rate<128, false>("bf16 SW128", kSwizzle128, 128);
    run_rate<16, false>("bf16 SW32", kSwizzle32, 32, 14);
    run_rate<32, false>("bf16 SW32", kSwizzle32, 32, 10);
    run_rate<32, false>("bf16 SW64", kSwizzle64, 64, 10);
    run_rate<64, false>("bf16 SW64", kSwizzle64, 64, 10);
    run_rate<64, false>("bf16 SW128", kSwizzle128, 128, 10);
    run_rate<128, false>("bf16 SW128", kSwizzle128, 128, 10);
    run_rate<16, false>("bf16 SW32", kSwizzle32, 32);
    run_rate<32, false>("bf16 SW64", kSwizzle64, 64);
    run_rate<16, true>("tf32 SW128", kSwizzle128, 128);
    run_rate<32, true>("tf32 SW128", kSwizzle128, 128);
    run_rate<64, true>("tf32 SW128", kSwizzle128, 128);
    run_rate<128, true>("tf32 SW128", kSwizzle128, 128);
    return 0;
  }
  if (!timing) {
    // 1. plain tiles
    Case basic[] = {
        {"bf16 SW128 dense", kSwizzle128, 2, 64, 64, 0, 8, 0, 1},
        {"bf16 SW128 dense N=16", kSwizzle128, 2, 64, 16, 0, 8, 0, 1},
        {"bf16 SW128 dense N=128", kSwizzle128, 2, 64, 128, 0, 8, 0, 1},
        {"bf16 SW64 dense", kSwizzle64, 2, 32, 32, 0, 8, 0, 1},
        {"bf16 SW32 dense", kSwizzle32, 2, 16, 16, 0, 8, 0, 1},
        {"bf16 NONE dense", kSwizzleNone, 2, 32, 32, 0, 8, 0, 1},
        {"tf32 SW128 dense", kSwizzle128, 4, 32, 32, 0, 8, 0, 1},
        {"tf32 SW64 dense", kSwizzle64, 4, 16, 16, 0, 8, 0, 1},
    };
    for (auto& c : basic) fails += run_case(c, false);
    // 2. row-shifted A starts and non-dense group pitch (halo brick addressing)
    for (int bo = 0; bo < 2; ++bo)
      for (int shift : {1, 3, 8, 11}) {
        Case c1{"bf16 SW128 row shift", kSwizzle128, 2, 64, 32, shift, 8, bo, 1};
        fails += run_case(c1, false);
      }
    for (int bo = 0; bo < 2; ++bo)
      for (int shift : {0, 1, 2, 11}) {
        Case c2{"bf16 SW128 shift + pitch 10", kSwizzle128, 2, 64, 32, shift, 10, bo, 1};
        fails += run_case(c2, false);
      }
    for (int shift : {0, 1, 2, 3, 11}) {
      Case c3{"bf16 SW64 shift + pitch 10", kSwizzle64, 2, 32, 32, shift, 10, 0, 1};
      fails += run_case(c3, false);
      Case c4{"bf16 SW32 shift + pitch 10", kSwizzle32, 2, 16, 16, shift, 10, 0, 1};
      fails += run_case(c4, false);
      Case c5{"bf16 NONE shift + pitch 10", kSwizzleNone, 2, 32, 32, shift, 10, 0, 1};
      fails += run_case(c5, false);
      Case c6{"tf32 SW128 shift + pitch 10", kSwizzle128, 4, 32, 32, shift, 10, 0, 1};
      fails += run_case(c6, false);
    }
    Case c7{"bf16 SW128 shift + pitch 14 (7^3 halo)", kSwizzle128, 2, 64, 16, 5, 14, 0, 1};
    fails += run_case(c7, false);
    // 4. TMA
    fails += run_tma(kSwizzle128, 64, false);
    fails += run_tma(kSwizzle64, 32, false);
    fails += run_tma(kSwizzle32, 16, false);
    fails += run_tma(kSwizzleNone, 8, false);
    run_tma(kSwizzle32, 16, true);
  } else {
    // 3. issue rate: 256 chained MMAs per configuration
    for (int N : {16, 32, 64, 128, 256}) {
      Case a{"rate bf16 SW128 K=64", kSwizzle128, 2, 64, N, 0, 8, 0, 64};
      run_case(a, false);
      if (N <= 128) {
        Case b{"rate bf16 SW64  K=32", kSwizzle64, 2, 32, N, 0, 8, 0, 128};
        run_case(b, false);
        Case c{"rate bf16 SW32  K=16", kSwizzle32, 2, 16, N, 0, 8, 0, 256};
        run_case(c, false);
        Case d{"rate tf32 SW128 K=32", kSwizzle128, 4, 32, N, 0, 8, 0, 64};
        run_case(d, false);
        Case e{"rate bf16 NONE  K=32", kSwizzleNone, 2, 32, N, 0, 8, 0, 128};
        run_case(e, false);
      }
    }
  }
  printf("probe finished, %d failing cases\n", fails);
  return 0;
}
