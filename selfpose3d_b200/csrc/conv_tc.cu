// K2 / K5 (tensor-core form) -- convolutions as an im2col-free implicit GEMM on tcgen05.
//
// One persistent CTA per SM walks work items = (output brick of TX x 16 x 8 positions (x, y, z; z fastest),
// tile of N output channels).  3-D tensors map (X, Y, Z) to (x, y, z); 2-D image batches map
// (image, row, column) to (x, y, z) with a kernel extent of 1 along x.
//   * warp 0 (1 thread)  : TMA-loads the brick's input HALO ((TX+kx-1) x (16+k-1) x (8+k-1) positions x <=64
//                          channels, bf16, channel-last) as ONE 5-D tiled box with hardware swizzle and
//                          out-of-bounds zero fill (= the convolution's zero padding), double buffered.
//                          Stride-2 convolutions use the tensor map's element strides;
//   * warp 1 (1 thread)  : streams the packed weights, G taps per stage, through a small TMA ring;
//   * warps 2,3          : issue tcgen05.mma (warp-uniform loop, elect.sync-predicated).  A tap (dx,dy,dz) is
//                          NOT a new load: it is the same shared-memory halo with the A-descriptor start
//                          address moved by ((dx*HY + dy)*HZ + dz) rows and SBO = HZ rows (8 consecutive z
//                          positions form a core-matrix group, 16 y values form the 128 rows of M).  Verified on
//                          B200 by csrc/probe/tc_probe.cu: the swizzle is a function of the absolute smem
//                          address, so row-shifted starts with base_offset = 0 address the right data.
//                          Accumulators (TX tiles x N fp32 columns, double buffered) live in TMEM;
//   * warps 4..7         : epilogue.  tcgen05.ld the accumulator row of "their" position, apply folded
//                          BatchNorm scale/shift, optional residual and ReLU, convert, store channel-last
//                          (strided positions for the phases of a transposed convolution).
//
// Reference semantics: cudnn conv3d / conv_transpose3d(k2,s2) + batch_norm(eval) + relu (+ add) as wired in
// lib/models/v2v_net.py:10-69,124, and conv2d / conv_transpose2d(k4,s2,p1) + batch_norm + relu (+ add) of
// lib/models/pose_resnet.py:58-93,118-124,161-207.
#include "sp3d_common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <string.h>
#include <type_traits>

namespace sp3d {

using namespace tc;

constexpr int kTcBaseThreads = 128;   // warps 0-3: halo producer, weight producer, MMA issuers; then EG x 4 epilogue warps
constexpr int kBY = 16, kBZ = 8;   // brick extent in y and z: 128 rows of M = 16 groups of 8 z-positions

constexpr int largest_divisor_le(int n, int cap) {
  int best = 1;
  for (int d = 1; d <= n && d <= cap; ++d)
    if (n % d == 0) best = d;
  return best;
}

// 2^x on the special-function unit (2 ulp)
__device__ __forceinline__ float ex2_sfu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct SplitPair {};   // epilogue tag: float32 result stored as the bf16 pair (hi, lo), hi = bf16(r), lo = bf16(r - hi)

struct alignas(64) TcStoreMaps {
  CUtensorMap m[4];              // [0] for a regular launch; one per (px, py) phase pair for the fused transposed form
};

struct TcConvParams {
  int X, Y, Z;                   // output grid of this launch (bricks enumerate it)
  int n_outer;                   // leading (cube) dimension
  int bricks_x, bricks_y, bricks_z, n_bricks;
  int n_tiles;                   // tiles of N output channels
  int n_chunks;                  // K chunks of (RB / 2) input channels (all K blocks of the split-operand mode)
  int nc_block;                  // chunks per K block (= n_chunks unless the operands are split into bf16 terms)
  int w_rows_tile;               // packed-weight rows of one channel tile (the weight producer streams them linearly)
  uint32_t block_act;            // split-operand mode: nibble b = activation term plane read by K block b (plane t of
                                 // cube n is outer index t * n_outer + n of the activation tensor map), else 0
  int istride[3];                // input step per output step (x, y, z)
  int origin[3];                 // input offset of tap 0 (x, y, z)
  int cout, cout_pitch;          // channels computed / distance between output positions (elements)
  int TD, TH, TW;                // full output tensor extent
  int ostride[3], ooff[3];       // output position = o * ostride + ooff (transposed-conv phases)
  int relu;
  int out_mode;                  // 0: store bf16, 1: float32, 2: float32 values as two bf16 term planes (SP3D_BF16X2:
                                 //    plane t of cube n = outer index t * n_outer + n of the output / residual maps)
  const float* scale;
  const float* shift;
  const void* residual;          // same dtype / addressing as out
  void* out;
  int fused_cols;                // fused k2/s2 transposed convolution: columns per (px, py) output phase pair
                                 // (= 2 * cout: (pz, co) is contiguous in the output), else 0
  int blocked;                   // 1: contiguous item blocks per CTA (fused soft-argmax head)
  // fused soft-argmax head (N = 16 kernels): the output is not stored; per (CTA, epilogue group, cube) online-softmax
  // partials [m, s, wx, wy, wz] per channel go to sa_ws (float64) and softargmax_merge_kernel finishes them
  double* sa_ws;
  const float* sa_lin_x;
  const float* sa_lin_y;
  const float* sa_lin_z;
  const float* sa_centers;
  int sa_center_stride, sa_C, sa_slots;
  float sa_beta;
  int col_mod;                   // > 0: GEMM column -> channel is col % col_mod (fused phases, z-fold)
  int tma_store;                 // 1: epilogue stages rows in smem and stores (and pre-loads the residual) by TMA
  int has_res;
  unsigned long long* prof;      // optional per-CTA wait-cycle counters (sp3d_debug_conv_profile), else NULL
};
constexpr int kProfSlots = 16;    // per CTA: mma total, wait halo, wait weights, wait acc_empty, epi total, epi wait, items, -

// KSX / KS kernel extent along x / along y and z, RB bytes per smem row (= channels per K chunk * 2),
// N = MMA N (output-channel tile), TX = x-slices (M tiles) per brick, G = taps per weight stage, S = weight stages.
// HB = halo buffers (2 for the compute-heavy kernels, deeper for 1x1 where a work item is a few MMAs).
// WD = 2 (split-operand mode, split_terms 2): the accumulator of an x-slice is 2 N columns wide -- the K block on the
// activation term x0 multiplies the weight rows [w0 | w1] in ONE MMA of 2 N columns (the A operand, whose shared-memory
// read bounds narrow tiles, is read once for two products), the K block on x1 multiplies w0 into the upper N columns;
// the epilogue adds the two halves.
template <int KSX, int KS, int RB, int N, int TX, int G, int S, int HB = 2, int SB = 2, int SR = 128, int EG = 2, int F = 1,
          int WD = 1>
struct TcCfg {
  static constexpr int kThreads = kTcBaseThreads + 128 * EG;
  // epilogue staging ring: SB slots of one x-slice (128 positions) x kStageRow bytes (one swizzled TMA-store box)
  static constexpr int kStageRow = N * 2 < SR ? N * 2 : SR;   // SR caps the chunk where smem is short
  static constexpr int kStageBytes = 128 * kStageRow;
  static_assert(SB >= 2, "the staging ring needs two slots");
  static_assert(EG == 1 || (EG == 2 && TX % 2 == 0), "x-slices must split evenly over the epilogue groups");
  // z-fold F: F consecutive z positions form ONE smem row (RB = F * bytes per position) and F outputs share one GEMM
  // row (N = F * channel tile); a z "tap" becomes a window starting at any position inside the row sequence, so
  // KS + F - 1 windows (with zero weight blocks where window and output do not meet) replace F * KS narrow MMAs.
  static constexpr int kPad = (KS - 1) / 2;
  static constexpr int kPosBytes = RB / F;                          // bytes of one position (K of one window)
  static constexpr int kE0 = F * ((kPad + F - 1) / F) - kPad;       // first window, in positions from the aligned row start
  static constexpr int KZ = KS + F - 1;                             // windows along z
  static constexpr int HX = TX + KSX - 1, HY = kBY + KS - 1, HZ = kBZ + (kE0 + KZ - 1) / F;
  static constexpr int kHaloRows = HX * HY * HZ;
  static constexpr int kHaloBytes = kHaloRows * RB;
  // Single-halo-buffer kernels with taps along x load and release the halo PER X-SLICE: slice d is last read by the taps
  // with dx = d, so the next brick's slice d streams in while the taps dx > d still run (the wait on a monolithic
  // single buffer was 5 - 22 % of the MMA warps' time).  Slices are padded to 1024 bytes (swizzle atom alignment).
  // (3^3 kernels; for the F = 2 7^3 stem the ten extra boxes per fill cost more weight-stream stalls than the 4.7 % halo
  // wait they remove -- 22.9 against 22.5 Mclk per CTA, profiles/r02_conv_stalls*.log)
  // The slices live in a RING of kSlots slots addressed by a running slice counter: where all HX slices of a brick fit,
  // kSlots = HX (every slice has its own slot); the F = 4 stem (HX = 10 slices of 28 KB) keeps TX + 1 -- the TX slices the
  // current x tap reads plus the one the next tap adds -- which is what makes a 4-fold z-fold fit in shared memory.
  static constexpr bool kSliced = HB == 1 && KSX > 1 && (KSX <= 3 || F == 4);
  static constexpr int kSliceBytes = HY * HZ * RB;
  static constexpr int kSliceStride = kSliced ? (kSliceBytes + 1023) / 1024 * 1024 : kSliceBytes;
  static constexpr int kSlots = !kSliced ? HB : (HX * kSliceStride > 160 * 1024 ? TX + 1 : HX);
  static constexpr int kHBar = kSlots;                 // halo barriers (full / empty each)
  static constexpr int kHaloStride = kSliced ? kSlots * kSliceStride : (kHaloBytes + 1023) / 1024 * 1024;
  static constexpr int kTaps = KSX * KS * KZ;
  static constexpr int kGroups = kTaps / G;             // weight stages consumed per (brick, chunk)
  static constexpr int kTapBytes = N * kPosBytes;
  static constexpr int kWBytes = G * kTapBytes * WD;    // one stage = G consecutive taps (WD = 2: of up to 2 N rows each)
  static constexpr int kAcc = N * WD;                   // accumulator columns of one x-slice
  static constexpr int kWStride = (kWBytes + 1023) / 1024 * 1024;
  static constexpr int kTapsPerLoad = largest_divisor_le(G, 256 / N);   // TMA box rows <= 256
  static constexpr int kLoads = G / kTapsPerLoad;
  static constexpr int kHaloRegion = kSliced ? kHaloStride : HB * kHaloStride;
  static constexpr int kSmemBytes = kHaloRegion + S * kWStride + EG * SB * kStageBytes + 1024;   // + alignment slack
  static_assert(kTaps % G == 0, "taps per stage must divide the tap count");
  static constexpr int kGroupsPerDx = (KS * KZ) / G;    // weight stages per x tap (sliced halo: slice hand-over points)
  static_assert(!kSliced || (KS * KZ) % G == 0, "sliced halo: weight stages must not straddle x taps");
  static constexpr int kKSteps = kPosBytes / 32;   // tcgen05.mma K = 16 bf16 = 32 bytes
  static constexpr int kIssuers = TX >= 2 ? 2 : 1;
  static constexpr uint32_t kLayout = RB == 128 ? kSwizzle128 : (RB == 64 ? kSwizzle64 : kSwizzle32);
  static constexpr uint32_t kLayoutW = kPosBytes == 128 ? kSwizzle128 : (kPosBytes == 64 ? kSwizzle64 : kSwizzle32);
  static_assert(F == 1 || ((KSX == KS || KSX == 1) && kPosBytes >= 32), "z-fold: same convolutions, one K chunk");
  // accumulator sets in TMEM: two (the epilogue drains one brick while the next one accumulates) where they fit the 512
  // columns, else one (the MMA warps wait for the drain: a few percent, against the MMAs a wider tile saves)
  static constexpr int kAB = 2 * TX * N * WD <= 512 ? 2 : 1;
  static_assert(kAB * TX * N * WD <= 512, "accumulators exceed TMEM");
  static_assert(WD == 1 || WD == 2, "accumulator width factor");
  static_assert(kSmemBytes + 3072 <= 227 * 1024, "shared memory budget (dynamic + ~3 KB static)");
};

// CL = 2: CTA pairs (clusters of two) share the weight stream.  Both CTAs of a pair walk the same stage sequence; each
// loads HALF of every stage and multicasts it into both shared memories, so the L2 serves every weight byte once per pair
// (tried against the 13 - 33 % of the MMA warps' time spent waiting on weights in the 3^3 layers; see g_conv_pair below).
// A stage is recycled when the MMA warps of BOTH CTAs have released it (multicast tcgen05.commit).
template <int KSX, int KS, int RB, int N, int TX, int G, int S, int HB, int SB, int SR, int EG, int F, int WD, int CL>
__global__ void __launch_bounds__(kTcBaseThreads + 128 * EG, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ TcStoreMaps maps_out, const __grid_constant__ TcStoreMaps maps_res,
               const TcConvParams p) {
  using C = TcCfg<KSX, KS, RB, N, TX, G, S, HB, SB, SR, EG, F, WD>;
  constexpr int kWStages = S;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t halo_full[C::kHBar], halo_empty[C::kHBar], w_full[kWStages], w_empty[kWStages], acc_full[2], acc_empty[2];
  __shared__ uint64_t res_full[EG][SB];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_scale[2][N], s_shift[2][N];      // per accumulator buffer (the channel tile may change per item)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* halo = smem;                               // [HB][kHaloStride]
  uint8_t* wbuf = smem + C::kHaloRegion;              // [kWStages][kWStride]
  uint8_t* stage = wbuf + kWStages * C::kWStride;     // [SB][kStageBytes]

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  if (tid == 0) {
    for (int i = 0; i < C::kHBar; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_empty[i], C::kIssuers);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], C::kIssuers);
      mbar_init(&acc_empty[i], 128 * EG);
    }
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], C::kIssuers * CL);
    }
    for (int i = 0; i < EG * SB; ++i) mbar_init(&res_full[0][0] + i, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * N; i += C::kThreads) {     // channel tile 0 (the only one unless n_tiles > 1)
    const int co = i % N;
    const int cs = p.col_mod ? co % p.col_mod : co;
    s_scale[i / N][co] = (p.scale != nullptr && cs < p.cout) ? p.scale[cs] : 1.0f;
    s_shift[i / N][co] = (p.shift != nullptr && cs < p.cout) ? p.shift[cs] : 0.0f;
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&map_in);
    tma_prefetch_desc(&map_w);
    if (p.tma_store) tma_prefetch_desc(&maps_out.m[0]);
    if (p.has_res) tma_prefetch_desc(&maps_res.m[0]);
  }
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);   // warp-uniform for the compiler
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0u;

  // work item -> (outer index, brick origin in output coordinates, channel tile); the channel tile is the fastest
  // index so that concurrently running CTAs share one input halo through L2
  const int n_items = p.n_bricks * p.n_tiles;
  // round-robin over CTAs, or (fused soft-argmax head) one contiguous block of items per CTA so that a CTA meets at
  // most a few cubes
  const int items_per_cta = (n_items + (int)gridDim.x - 1) / (int)gridDim.x;
  const int wi_begin = p.blocked ? (int)blockIdx.x * items_per_cta : (int)blockIdx.x;
  const int wi_end = p.blocked ? min(n_items, wi_begin + items_per_cta) : n_items;
  const int wi_step = p.blocked ? 1 : (int)gridDim.x;
  // CL = 2 (round-robin items, one channel tile): the odd CTA may own one item less than the even one; it still walks the
  // weight stages of that item (producer and MMA warps, no MMAs) so that the pair's stage sequences stay identical
  const int pair_items = CL > 1 ? (n_items - ((int)blockIdx.x & ~1) + wi_step - 1) / wi_step : 0;
  const int my_items = CL > 1 ? (n_items - (int)blockIdx.x + wi_step - 1) / wi_step : 0;
  auto item_coords = [&](int wi, int& n, int& x0, int& y0, int& z0, int& nt) {
    nt = wi % p.n_tiles;
    const int b = wi / p.n_tiles;
    const int bz = b % p.bricks_z;
    const int by = (b / p.bricks_z) % p.bricks_y;
    const int bx = (b / (p.bricks_z * p.bricks_y)) % p.bricks_x;
    n = b / (p.bricks_z * p.bricks_y * p.bricks_x);
    x0 = bx * TX; y0 = by * kBY; z0 = bz * kBZ;
  };

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ halo producer
    uint32_t u = 0;
    for (int wi = wi_begin; wi < wi_end; wi += wi_step) {
      int n, x0, y0, z0, nt;
      item_coords(wi, n, x0, y0, z0, nt);
      for (int c = 0; c < p.n_chunks; ++c, ++u) {
        // split-operand mode: K block kb = c / nc_block reads activation term plane (block_act >> 4 kb) & 15
        const int kb = c / p.nc_block;
        const int plane = (int)((p.block_act >> (4 * kb)) & 15u);
        if constexpr (C::kSliced) {
          for (int xs = 0; xs < C::HX; ++xs) {          // one box per x-slice (the tensor map's box is one slice thick)
            const uint32_t q = u * C::HX + xs, sl = q % C::kSlots;      // running slice counter -> ring slot
            mbar_wait(&halo_empty[sl], ((q / C::kSlots) & 1) ^ 1);
            mbar_arrive_expect_tx(&halo_full[sl], C::kSliceBytes);
            tma_load_5d(halo + sl * C::kSliceStride, &map_in, &halo_full[sl], (c - kb * p.nc_block) * (RB / 2),
                        z0 * p.istride[2] + p.origin[2], y0 * p.istride[1] + p.origin[1],
                        x0 * p.istride[0] + p.origin[0] + xs, plane * p.n_outer + n);
          }
        } else {
          const uint32_t buf = u % HB;
          mbar_wait(&halo_empty[buf], ((u / HB) & 1) ^ 1);
          mbar_arrive_expect_tx(&halo_full[buf], C::kHaloBytes);
          tma_load_5d(halo + buf * C::kHaloStride, &map_in, &halo_full[buf], (c - kb * p.nc_block) * (RB / 2),
                      z0 * p.istride[2] + p.origin[2], y0 * p.istride[1] + p.origin[1],
                      x0 * p.istride[0] + p.origin[0], plane * p.n_outer + n);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ weight producer (G taps per stage)
    // packed weights: rows ordered [n_tile][chunk][tap][N]; a stage holds taps g*G .. g*G+G-1 of one chunk
    uint32_t w = 0;
    const int wi_end_w = CL > 1 ? wi_begin + pair_items * wi_step : wi_end;
    for (int wi = wi_begin; wi < wi_end_w; wi += wi_step) {
      const int nt = wi % p.n_tiles;
      int row = nt * p.w_rows_tile;            // rows are consumed in storage order: [chunk][tap][N (or 2 N) rows]
      for (int c = 0; c < p.n_chunks; ++c) {
        const bool wide = WD == 2 && c < p.nc_block;       // K block 0 of the 2-block split mode: taps of 2 N rows
        const int loads = wide ? 2 * C::kLoads : C::kLoads;
        for (int g = 0; g < C::kGroups; ++g, ++w) {
          const uint32_t st = w % kWStages;
          mbar_wait(&w_empty[st], ((w / kWStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&w_full[st], (uint32_t)loads * (C::kTapsPerLoad * C::kTapBytes));
          if constexpr (CL > 1) {
            // this CTA's half of every box, delivered to both CTAs (data and complete_tx at the same offsets)
            constexpr int kHalfRows = C::kTapsPerLoad * N / 2, kHalfBytes = C::kTapsPerLoad * C::kTapBytes / 2;
#pragma unroll 1
            for (int l = 0; l < loads; ++l, row += C::kTapsPerLoad * N)
              tma_load_2d_mc(wbuf + st * C::kWStride + l * (C::kTapsPerLoad * C::kTapBytes) + cta_rank * kHalfBytes, &map_w,
                             &w_full[st], 0, row + (int)cta_rank * kHalfRows, (uint16_t)3);
          } else {
#pragma unroll 1
            for (int l = 0; l < loads; ++l, row += C::kTapsPerLoad * N)
              tma_load_2d(wbuf + st * C::kWStride + l * (C::kTapsPerLoad * C::kTapBytes), &map_w, &w_full[st], 0, row);
          }
        }
      }
    }
  } else if ((warp == 2 || warp == 3) && (warp - 2) < C::kIssuers) {
    // ------------------------------------------------------------------ MMA issuers
    // The whole warp runs this loop (warp-uniform control flow and values); only the tcgen05.mma / commit
    // instructions are predicated on the elected lane.
    const int q = warp - 2;
    const uint32_t idesc_n = make_idesc(kFmtBF16, 128, N), idesc_w = make_idesc(kFmtBF16, 128, N * WD);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(halo), 0, C::HZ * RB, C::kLayout);
    const uint64_t b_desc0 = make_smem_desc(smem_u32(wbuf), 0, 8 * C::kPosBytes, C::kLayoutW);
    uint32_t u = 0, w = 0, it = 0;
    long long t_begin = 0, t_halo = 0, t_w = 0, t_acc = 0, t0 = 0;
    const bool prof = p.prof != nullptr;
    if (prof) t_begin = clock64();
    for (int wi = wi_begin; wi < wi_end; wi += wi_step, ++it) {
      const uint32_t accbuf = it % C::kAB;
      if (prof) t0 = clock64();
      mbar_wait(&acc_empty[accbuf], ((it / C::kAB) & 1) ^ 1);
      if (prof) t_acc += clock64() - t0;
      tc_fence_after();
      for (int c = 0; c < p.n_chunks; ++c, ++u) {
        const uint32_t buf = C::kSliced ? 0u : u % HB;
        if (prof) t0 = clock64();
        const uint32_t q0 = u * C::HX;                     // sliced halo: this chunk's first slice in the running count
        if constexpr (C::kSliced) {
#pragma unroll 1
          for (int xs = 0; xs < TX; ++xs)                  // the slices the taps dx = 0 read
            mbar_wait(&halo_full[(q0 + xs) % C::kSlots], ((q0 + xs) / C::kSlots) & 1);
        } else {
          mbar_wait(&halo_full[buf], (u / HB) & 1);
        }
        if (prof) t_halo += clock64() - t0;
        // descriptor low words advance in 16-byte units; the high words (SBO, version, layout) never change
        const uint32_t a_lo0 = (uint32_t)a_desc0 + (uint32_t)((buf * C::kHaloStride) >> 4);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        constexpr uint32_t kRow16 = RB / 16;                       // one halo row in 16-byte units
        constexpr uint32_t kPos16 = C::kPosBytes / 16;             // one position (= row unless z-folded)
        uint32_t accum = c ? 1u : 0u;                              // first tap of the first chunk overwrites
        int dz = 0, dy = 0;
        uint32_t a_tap = a_lo0 + C::kE0 * kPos16;                  // start of the current tap's shifted window
        // sliced halo: a_tap is the window's offset INSIDE a slice; soff[t] = ring slot of slice dx + t, in 16-byte units
        uint32_t soff[TX];
        if constexpr (C::kSliced) {
#pragma unroll
          for (int t = 0; t < TX; ++t) soff[t] = ((q0 + t) % C::kSlots) * (uint32_t)(C::kSliceStride >> 4);
        }
        // One K chunk.  WIDE (WD = 2 only, K block 0 = activation term x0): MMAs of 2 N columns over the weight rows
        // [w0 | w1]; the other block (x1) adds its N columns onto the upper half, so that the small products
        // x0 w1 + x1 w0 share one accumulator.  Compile-time so that descriptor steps stay immediates in the issue loop.
        auto issue_chunk = [&](auto wide_tag) {
          constexpr bool WIDE = decltype(wide_tag)::value;
          const uint32_t idesc = WIDE ? idesc_w : idesc_n;
          constexpr uint32_t d_off = (WD == 2 && !WIDE) ? (uint32_t)N : 0u;
          constexpr uint32_t b_step = (uint32_t)((WIDE ? 2 * C::kTapBytes : C::kTapBytes) >> 4);
          for (int g = 0; g < C::kGroups; ++g, ++w) {
            if constexpr (C::kSliced) {
              if (g > 0 && g % C::kGroupsPerDx == 0) {      // taps move on to dx = g / kGroupsPerDx
                const uint32_t dxn = (uint32_t)(g / C::kGroupsPerDx);
                if (elect_one_sync()) mma_commit(&halo_empty[(q0 + dxn - 1) % C::kSlots]);   // slice dx - 1: read for the last time
                if (prof) t0 = clock64();
                const uint32_t qn = q0 + dxn + TX - 1;                         // the one new slice these taps touch
                mbar_wait(&halo_full[qn % C::kSlots], (qn / C::kSlots) & 1);
                if (prof) t_halo += clock64() - t0;
                a_tap = a_lo0 + C::kE0 * kPos16;
#pragma unroll
                for (int t = 0; t < TX; ++t) soff[t] = ((q0 + dxn + t) % C::kSlots) * (uint32_t)(C::kSliceStride >> 4);
              }
            }
            const uint32_t st = w % kWStages;
            if (prof) t0 = clock64();
            mbar_wait(&w_full[st], (w / kWStages) & 1);
            if (prof) t_w += clock64() - t0;
            tc_fence_after();
            uint32_t b_lo = (uint32_t)b_desc0 + (uint32_t)((st * C::kWStride) >> 4);
#pragma unroll(G % C::KZ == 0 ? C::KZ : (G <= 5 ? G : 1))
            for (int j = 0; j < G; ++j) {
#pragma unroll
              for (int t = q; t < TX; t += C::kIssuers) {
                const uint32_t a_lo = C::kSliced ? a_tap + soff[t] : a_tap + (uint32_t)t * (uint32_t)(C::kSliceStride >> 4);
                const uint32_t d_tmem = tmem_base + (accbuf * TX + t) * C::kAcc + d_off;
#pragma unroll
                for (int k = 0; k < C::kKSteps; ++k)
                  if (elect_one_sync()) mma_f16_ss_lohi(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k ? 1u : accum);
              }
              accum = 1u;
              b_lo += b_step;
              // next tap: z fastest, then y, then x
              a_tap += kPos16;
              if (++dz == C::KZ) {
                dz = 0;
                a_tap += (uint32_t)C::HZ * kRow16 - (uint32_t)C::KZ * kPos16;
                if (++dy == KS) {       // next x tap (sliced halo: re-based at the slice hand-over above)
                  dy = 0;
                  a_tap += (uint32_t)((C::HY - KS) * C::HZ) * kRow16 + (uint32_t)((C::kSliceStride - C::kSliceBytes) >> 4);
                }
              }
            }
            if (elect_one_sync()) {
              if constexpr (CL > 1) mma_commit_mc(&w_empty[st], (uint16_t)3);
              else mma_commit(&w_empty[st]);
            }
          }
        };
        if (WD == 2 && c < p.nc_block) issue_chunk(std::integral_constant<bool, WD == 2>{});
        else issue_chunk(std::false_type{});
        if constexpr (C::kSliced) {
#pragma unroll 1
          for (int xs = KSX - 1; xs < C::HX; ++xs)
            if (elect_one_sync()) mma_commit(&halo_empty[(q0 + xs) % C::kSlots]);
        } else {
          if (elect_one_sync()) mma_commit(&halo_empty[buf]);
        }
      }
      if (elect_one_sync()) mma_commit(&acc_full[accbuf]);
    }
    if constexpr (CL > 1) {
      // the pair's last item has no twin here: release its weight stages unread
      for (int i = my_items; i < pair_items; ++i)
        for (int c = 0; c < p.n_chunks; ++c)
          for (int g = 0; g < C::kGroups; ++g, ++w) {
            const uint32_t st = w % kWStages;
            mbar_wait(&w_full[st], (w / kWStages) & 1);
            if (elect_one_sync()) mma_commit_mc(&w_empty[st], (uint16_t)3);
          }
    }
    if (prof && q == 0 && lane == 0) {
      unsigned long long* o = p.prof + (size_t)blockIdx.x * kProfSlots;
      o[0] = clock64() - t_begin; o[1] = t_halo; o[2] = t_w; o[3] = t_acc; o[6] = it;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (thread <-> accumulator row)
    const int row = (tid - 128) & 127;         // 0..127 = TMEM lane
    const int grp = (tid - 128) >> 7;          // epilogue group: x-slices grp, grp + EG, ... of every brick
    uint8_t* gstage = stage + grp * SB * C::kStageBytes;
    uint64_t* gres_full = res_full[grp];
    const int bar_id = 1 + grp;
    const int ly = row >> 3, lz = row & 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t it = 0;
    long long e_begin = 0, e_wait = 0, e0 = 0;
    const bool prof = p.prof != nullptr;
    if (prof) e_begin = clock64();
    // ---- staged form: TMEM -> registers -> swizzled smem rows (residual pre-loaded there by TMA) -> TMA store.
    // A unit = one x-slice (128 positions) x one chunk of kStageRow bytes of channels; units walk a ring of SB slots.
    auto staged = [&](auto tag) {
      using T = decltype(tag);
      // split-pair output: a slot holds two half-width boxes (hi terms, then lo terms) of kStageRow / 2 bytes per row
      constexpr bool kSplit = std::is_same<T, SplitPair>::value;
      constexpr int ELT = kSplit ? 4 : (int)sizeof(T);      // staged bytes per accumulator column
      constexpr int COLS = C::kStageRow / ELT;              // accumulator columns per unit
      constexpr int CPU = kSplit ? 8 : 16 / ELT;            // columns per 16-byte smem unit (per box)
      constexpr int BOXROW = kSplit ? C::kStageRow / 2 : C::kStageRow;   // bytes per row of one TMA box
      constexpr int UNITS = BOXROW / 16;
      constexpr int NSC = N / COLS;                         // units per x-slice
      constexpr uint32_t SWZ = BOXROW == 128 ? 7u : (BOXROW == 64 ? 3u : (BOXROW == 32 ? 1u : 0u));
      constexpr int LO_OFF = C::kStageBytes / 2;            // split: byte offset of the lo box inside a slot
      const bool has_res = p.has_res != 0;
      const bool leader = row == 0;
      int pf_wi = wi_begin, pf_t = grp, pf_sc = 0;          // residual prefetch cursor (leader)
      uint32_t pf_u = 0;
      auto issue_res = [&]() {
        int n, x0, y0, z0, nt;
        item_coords(pf_wi, n, x0, y0, z0, nt);
        const uint32_t b = pf_u % SB;
        mbar_arrive_expect_tx(&gres_full[b], C::kStageBytes);
        int gc = nt * N + pf_sc * COLS, mi = 0;
        if (p.fused_cols) {
          mi = gc / p.fused_cols;
          gc -= mi * p.fused_cols;
        }
        tma_load_5d(gstage + b * C::kStageBytes, &maps_res.m[mi], &gres_full[b], gc, z0, y0, x0 + pf_t, n);
        if constexpr (kSplit)
          tma_load_5d(gstage + b * C::kStageBytes + LO_OFF, &maps_res.m[mi], &gres_full[b], gc, z0, y0, x0 + pf_t,
                      p.n_outer + n);
        ++pf_u;
        if (++pf_sc == NSC) {
          pf_sc = 0;
          if ((pf_t += EG) >= TX) {
            pf_t = grp;
            pf_wi += wi_step;
          }
        }
      };
      if (has_res && leader)
        for (int i = 0; i < SB - 1 && pf_wi < wi_end; ++i) issue_res();
      // N <= 32: the folded-BatchNorm scale / shift stay in registers (shared-memory loads queue behind the tensor
      // core's operand reads while MMAs run); wider tiles read them as float4 from shared memory
      constexpr bool kRegSS = N <= 32;
      float reg_sc[kRegSS ? N : 1], reg_sh[kRegSS ? N : 1];
      if constexpr (kRegSS) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
          reg_sc[c] = s_scale[0][c];
          reg_sh[c] = s_shift[0][c];
        }
      }
      uint32_t u = 0;
      long long pe[5] = {0, 0, 0, 0, 0};   // tmem load, residual wait, math + smem, fence + barrier, store + ring wait
      for (int wi = wi_begin; wi < wi_end; wi += wi_step, ++it) {
        int n, x0, y0, z0, nt;
        item_coords(wi, n, x0, y0, z0, nt);
        const uint32_t accbuf = it % C::kAB;
        if (prof) e0 = clock64();
        mbar_wait(&acc_full[accbuf], (it / C::kAB) & 1);
        if (prof) e_wait += clock64() - e0;
        tc_fence_after();
        const int ch0 = nt * N;
        if (p.n_tiles > 1) {
          if (row < N) {
            const int co = p.col_mod ? (ch0 + row) % p.col_mod : ch0 + row;
            s_scale[accbuf][row] = (p.scale != nullptr && co < p.cout) ? __ldg(p.scale + co) : 1.0f;
            s_shift[accbuf][row] = (p.shift != nullptr && co < p.cout) ? __ldg(p.shift + co) : 0.0f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if constexpr (kRegSS) {
#pragma unroll
            for (int c = 0; c < N; ++c) {
              reg_sc[c] = s_scale[accbuf][c];
              reg_sh[c] = s_shift[accbuf][c];
            }
          }
        }
        const float* sc_s = s_scale[p.n_tiles > 1 ? accbuf : 0];
        const float* sh_s = s_shift[p.n_tiles > 1 ? accbuf : 0];
#pragma unroll 1
        for (int t = grp; t < TX; t += EG) {
#pragma unroll(NSC <= 2 ? NSC : 1)
          for (int sc = 0; sc < NSC; ++sc, ++u) {
            const uint32_t b = u % SB;
            uint8_t* sbuf = gstage + b * C::kStageBytes;
            const uint32_t taddr = tmem_base + lane_base + (accbuf * TX + t) * C::kAcc + sc * COLS;
            uint32_t v[COLS];
            if constexpr (COLS >= 16) {
#pragma unroll
              for (int c = 0; c < COLS; c += 16) tmem_ld_x16(taddr + c, v + c);
            } else {
              tmem_ld_x8(taddr, v);
            }
            long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
            if (prof) c0 = clock64();
            tmem_ld_wait();
            if constexpr (WD == 2) {      // add the small-product accumulator (upper N columns)
              uint32_t v2[COLS];
              if constexpr (COLS >= 16) {
#pragma unroll
                for (int c = 0; c < COLS; c += 16) tmem_ld_x16(taddr + N + c, v2 + c);
              } else {
                tmem_ld_x8(taddr + N, v2);
              }
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < COLS; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(v2[c]));
            }
            if (prof) c1 = clock64();
            if (has_res) mbar_wait(&gres_full[b], (u / SB) & 1);
            if (prof) c2 = clock64();
            const uint32_t sbuf_s = smem_u32(sbuf);
#pragma unroll
            for (int j = 0; j < UNITS; ++j) {
              const uint32_t off = (uint32_t)row * BOXROW + j * 16;
              const uint32_t q = sbuf_s + (off ^ (((off >> 7) & SWZ) << 4));
              const int col0 = sc * COLS + j * CPU;
              float r[CPU];
#pragma unroll
              for (int i = 0; i < CPU; i += 4) {
                float4 s4, h4;
                if constexpr (kRegSS) {
                  s4 = make_float4(reg_sc[col0 + i], reg_sc[col0 + i + 1], reg_sc[col0 + i + 2], reg_sc[col0 + i + 3]);
                  h4 = make_float4(reg_sh[col0 + i], reg_sh[col0 + i + 1], reg_sh[col0 + i + 2], reg_sh[col0 + i + 3]);
                } else {
                  s4 = *reinterpret_cast<const float4*>(sc_s + col0 + i);
                  h4 = *reinterpret_cast<const float4*>(sh_s + col0 + i);
                }
                r[i + 0] = __uint_as_float(v[j * CPU + i + 0]) * s4.x + h4.x;
                r[i + 1] = __uint_as_float(v[j * CPU + i + 1]) * s4.y + h4.y;
                r[i + 2] = __uint_as_float(v[j * CPU + i + 2]) * s4.z + h4.z;
                r[i + 3] = __uint_as_float(v[j * CPU + i + 3]) * s4.w + h4.w;
              }
              if (p.relu == 2) {
#pragma unroll
                for (int i = 0; i < CPU; ++i) r[i] = fmaxf(r[i], 0.0f);
              }
              uint4 o;
              if constexpr (kSplit) {
                if (has_res) {
                  const uint4 rh = lds128(q), rl = lds128(q + LO_OFF);
                  const uint32_t h4[4] = {rh.x, rh.y, rh.z, rh.w}, l4[4] = {rl.x, rl.y, rl.z, rl.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    r[2 * i] += __uint_as_float(h4[i] << 16) + __uint_as_float(l4[i] << 16);
                    r[2 * i + 1] += __uint_as_float(h4[i] & 0xffff0000u) + __uint_as_float(l4[i] & 0xffff0000u);
                  }
                }
                if (p.relu == 1) {
#pragma unroll
                  for (int i = 0; i < CPU; ++i) r[i] = fmaxf(r[i], 0.0f);
                }
                uint32_t w4[4], u4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 h = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
                  // the remainder of a bf16 rounding is exact in float32
                  const __nv_bfloat162 l = __floats2bfloat162_rn(r[2 * i] - __low2float(h), r[2 * i + 1] - __high2float(h));
                  w4[i] = *reinterpret_cast<const uint32_t*>(&h);
                  u4[i] = *reinterpret_cast<const uint32_t*>(&l);
                }
                o = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                sts128(q + LO_OFF, make_uint4(u4[0], u4[1], u4[2], u4[3]));
              } else if constexpr (sizeof(T) == 4) {
                if (has_res) {
                  const uint4 rv = lds128(q);
                  r[0] += __uint_as_float(rv.x); r[1] += __uint_as_float(rv.y);
                  r[2] += __uint_as_float(rv.z); r[3] += __uint_as_float(rv.w);
                }
                if (p.relu == 1) {
#pragma unroll
                  for (int i = 0; i < CPU; ++i) r[i] = fmaxf(r[i], 0.0f);
                }
                o = make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3]));
              } else {
                if (has_res) {
                  const uint4 rv = lds128(q);
                  const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    r[2 * i] += __uint_as_float(w4[i] << 16);
                    r[2 * i + 1] += __uint_as_float(w4[i] & 0xffff0000u);
                  }
                }
                if (p.relu == 1) {
#pragma unroll
                  for (int i = 0; i < CPU; ++i) r[i] = fmaxf(r[i], 0.0f);
                }
                uint32_t w4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 h = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
                  w4[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                o = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
              sts128(q, o);
            }
            if (prof) c3 = clock64();
            fence_proxy_async_smem();
            // no residual: before publishing "unit u written", the leader makes sure the store that last used the NEXT
            // slot (unit u + 1 - SB) has left shared memory, so this one barrier also frees that slot for unit u + 1
            if (!has_res && leader) bulk_wait_read<SB - 2>();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (prof) c4 = clock64();
            if (prof) { pe[0] += c1 - c0; pe[1] += c2 - c1; pe[2] += c3 - c2; pe[3] += c4 - c3; }
            if (leader) {
              int gc = ch0 + sc * COLS, mi = 0;
              if (p.fused_cols) {
                mi = gc / p.fused_cols;
                gc -= mi * p.fused_cols;
              }
              tma_store_5d(&maps_out.m[mi], sbuf, gc, z0, y0, x0 + t, n);
              if constexpr (kSplit) tma_store_5d(&maps_out.m[mi], sbuf + LO_OFF, gc, z0, y0, x0 + t, p.n_outer + n);
              bulk_commit();
              if (has_res) {
                if (pf_wi < wi_end) {          // slot (u - 1) % SB: its store (unit u - 1) must have left smem
                  bulk_wait_read<1>();
                  issue_res();
                }
              }
            }
            if (prof) pe[4] += clock64() - c4;
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[accbuf]);
      }
      if (leader) bulk_wait<0>();
      if (prof && leader && grp == 0) {
        unsigned long long* o = p.prof + (size_t)blockIdx.x * kProfSlots;
        for (int i = 0; i < 5; ++i) o[8 + i] = pe[i];
      }
    };
    if constexpr (N == 16 && WD == 1) {
      if (p.sa_ws != nullptr) {
        // ---- fused soft-argmax head: logits = scale * acc + shift never leave the SM
        // scratch carved out of this group's (unused) staging ring: coordinate tables, per-warp maxima and sums
        static_assert(SB * C::kStageBytes >= (3 * 256 + 4 * 16 + 4 * 16 * 4) * 4, "staging ring too small for the head");
        float (*sa_g)[256] = reinterpret_cast<float (*)[256]>(gstage);                       // [3][256]
        float (*sa_max)[16] = reinterpret_cast<float (*)[16]>(gstage + 3 * 256 * 4);         // [4][16]
        float (*sa_sum)[16][4] = reinterpret_cast<float (*)[16][4]>(gstage + (3 * 256 + 4 * 16) * 4);   // [4][16][4]
        constexpr float kLog2e = 1.4426950408889634f;
        float m[15], s[15], wx[15], wy[15], wz[15];
        int cur_n = -1;
        const int lane = tid & 31, wq = warp & 3;
        auto reset = [&]() {
#pragma unroll
          for (int c = 0; c < 15; ++c) { m[c] = -INFINITY; s[c] = 0.f; wx[c] = 0.f; wy[c] = 0.f; wz[c] = 0.f; }
        };
        auto flush = [&](int n) {       // merge the group's 128 thread states of cube n and write one partial
#pragma unroll
          for (int c = 0; c < 15; ++c) {
            float mm = m[c];
            for (int off = 16; off > 0; off >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, off));
            if (lane == 0) sa_max[wq][c] = mm;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
          for (int c = 0; c < 15; ++c) {
            const float M = fmaxf(fmaxf(sa_max[0][c], sa_max[1][c]), fmaxf(sa_max[2][c], sa_max[3][c]));
            const float f = s[c] > 0.0f ? ex2_sfu((m[c] - M) * kLog2e) : 0.0f;
            float v4[4] = {s[c] * f, wx[c] * f, wy[c] * f, wz[c] * f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              for (int off = 16; off > 0; off >>= 1) v4[i] += __shfl_xor_sync(0xffffffffu, v4[i], off);
              if (lane == 0) sa_sum[wq][c][i] = v4[i];
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (row < p.sa_C) {
            const int c = row;
            const float M = fmaxf(fmaxf(sa_max[0][c], sa_max[1][c]), fmaxf(sa_max[2][c], sa_max[3][c]));
            double* o = p.sa_ws + (((int64_t)n * p.sa_slots + (int)blockIdx.x * EG + grp) * p.sa_C + c) * 5;
            o[0] = (double)M;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o[1 + i] = (double)sa_sum[0][c][i] + (double)sa_sum[1][c][i] + (double)sa_sum[2][c][i] +
                         (double)sa_sum[3][c][i];
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        };
        float hsc[16], hsh[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { hsc[c] = s_scale[0][c]; hsh[c] = s_shift[0][c]; }
        reset();
        for (int wi = wi_begin; wi < wi_end; wi += wi_step, ++it) {
          int n, x0, y0, z0, nt;
          item_coords(wi, n, x0, y0, z0, nt);
          if (n != cur_n) {
            if (cur_n >= 0) flush(cur_n);
            reset();
            cur_n = n;
            const float* cen = p.sa_centers + (int64_t)n * p.sa_center_stride;
            const float cx = cen[0], cy = cen[1], cz = cen[2];
            // g - centre with g = fl(lin + centre): the `grids` values of ProjectLayer.compute_grid
            for (int i = row; i < p.X; i += 128) sa_g[0][i] = __fsub_rn(__fadd_rn(p.sa_lin_x[i], cx), cx);
            for (int i = row; i < p.Y; i += 128) sa_g[1][i] = __fsub_rn(__fadd_rn(p.sa_lin_y[i], cy), cy);
            for (int i = row; i < p.Z; i += 128) sa_g[2][i] = __fsub_rn(__fadd_rn(p.sa_lin_z[i], cz), cz);
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          }
          const uint32_t accbuf = it % C::kAB;
          mbar_wait(&acc_full[accbuf], (it / C::kAB) & 1);
          tc_fence_after();
          const int y = y0 + ly, z = z0 + lz;
#pragma unroll 1
          for (int t = grp; t < TX; t += EG) {
            const int x = x0 + t;
            uint32_t v[16];
            tmem_ld_x16(tmem_base + lane_base + (accbuf * TX + t) * N, v);
            tmem_ld_wait();
            if (x < p.X && y < p.Y && z < p.Z) {
              const float gx = sa_g[0][x], gy = sa_g[1][y], gz = sa_g[2][z];
#pragma unroll
              for (int c = 0; c < 15; ++c) {
                if (c < p.sa_C) {
                  const float logit = __uint_as_float(v[c]) * hsc[c] + hsh[c];
                  const float zz = __fmul_rn(p.sa_beta, logit);
                  // branch-free online softmax: 15 independent chains keep the few epilogue warps busy
                  const float mn = fmaxf(m[c], zz);
                  const float sc = ex2_sfu((m[c] - mn) * kLog2e);     // 2^-inf = 0 on the first voxel
                  const float e = ex2_sfu((zz - mn) * kLog2e);
                  s[c] = fmaf(s[c], sc, e);
                  wx[c] = fmaf(wx[c], sc, e * gx);
                  wy[c] = fmaf(wy[c], sc, e * gy);
                  wz[c] = fmaf(wz[c], sc, e * gz);
                  m[c] = mn;
                }
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&acc_empty[accbuf]);
        }
        if (cur_n >= 0) flush(cur_n);
      }
    }
    if (p.tma_store && !p.sa_ws) {
      if (p.out_mode == 1) staged(float{});
      else if (p.out_mode == 2) staged(SplitPair{});
      else staged(__nv_bfloat16{});
    }
    // ---- direct form (rows whose byte pitch is not a multiple of 16, e.g. the 1-channel float32 score volume)
    // (compiled for the N = 16 kernels only: that is where a 1- or 15-channel float32 head lands)
    if constexpr (N == 16 && WD == 1)
    for (int wi = wi_begin; !p.tma_store && !p.sa_ws && wi < wi_end; wi += wi_step, ++it) {
      int n, x0, y0, z0, nt;
      item_coords(wi, n, x0, y0, z0, nt);
      const uint32_t accbuf = it % C::kAB;
      if (prof) e0 = clock64();
      mbar_wait(&acc_full[accbuf], (it / C::kAB) & 1);
      if (prof) e_wait += clock64() - e0;
      tc_fence_after();
      const int y = y0 + ly, z = z0 + lz;
      const int ch0 = nt * N;                  // first output channel of this tile
      if (p.n_tiles > 1) {
        // this item's scale/shift into the accumulator buffer's slot: every reader of the slot's previous
        // contents (two items ago) has arrived on acc_empty before this item's acc_full could complete
        if (row < N) {
          const int co = ch0 + row;
          s_scale[accbuf][row] = (p.scale != nullptr && co < p.cout) ? __ldg(p.scale + co) : 1.0f;
          s_shift[accbuf][row] = (p.shift != nullptr && co < p.cout) ? __ldg(p.shift + co) : 0.0f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      const float* sc_s = s_scale[p.n_tiles > 1 ? accbuf : 0];
      const float* sh_s = s_shift[p.n_tiles > 1 ? accbuf : 0];
#pragma unroll
      for (int t = grp; t < TX; t += EG) {
        const int x = x0 + t;
        const bool in_range = (x < p.X) && (y < p.Y) && (z < p.Z);
        const int64_t pos = (((int64_t)n * p.TD + (x * p.ostride[0] + p.ooff[0])) * p.TH + (y * p.ostride[1] + p.ooff[1])) *
                                p.TW + (z * p.ostride[2] + p.ooff[2]);
        const uint32_t taddr = tmem_base + lane_base + (accbuf * TX + t) * N;
#pragma unroll
        for (int n0 = 0; n0 < N; n0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(taddr + n0, v);      // warp-collective: every lane takes part, stores are predicated
          tmem_ld_wait();
          if (!in_range) continue;
          const int cbase = ch0 + n0;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float r = __uint_as_float(v[j]) * sc_s[n0 + j] + sh_s[n0 + j];
            if (p.relu == 2) r = fmaxf(r, 0.0f);
            f[j] = r;
          }
          if (p.out_mode == 1) {
            float* o = reinterpret_cast<float*>(p.out) + pos * p.cout_pitch + cbase;
            const float* rs = p.residual ? reinterpret_cast<const float*>(p.residual) + pos * p.cout_pitch + cbase : nullptr;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (cbase + j < p.cout_pitch) {
                float r = f[j];
                if (rs) r += rs[j];
                if (p.relu == 1) r = fmaxf(r, 0.0f);
                o[j] = (cbase + j < p.cout) ? r : 0.0f;
              }
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pos * p.cout_pitch + cbase;
            const __nv_bfloat16* rs =
                p.residual ? reinterpret_cast<const __nv_bfloat16*>(p.residual) + pos * p.cout_pitch + cbase : nullptr;
            if (cbase + 16 <= p.cout_pitch && (p.cout_pitch & 7) == 0) {   // full group of 16 channels: two 16-byte stores
              uint4 rv[2];
              if (rs) {
                rv[0] = *reinterpret_cast<const uint4*>(rs);
                rv[1] = *reinterpret_cast<const uint4*>(rs + 8);
              }
              const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(rv);
              __align__(16) __nv_bfloat16 ob[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float r = f[j];
                if (rs) r += __bfloat162float(rb[j]);
                if (p.relu == 1) r = fmaxf(r, 0.0f);
                ob[j] = __float2bfloat16_rn((cbase + j < p.cout) ? r : 0.0f);
              }
              *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(ob);
              *reinterpret_cast<uint4*>(o + 8) = *reinterpret_cast<const uint4*>(ob + 8);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (cbase + j < p.cout_pitch) {
                  float r = f[j];
                  if (rs) r += __bfloat162float(rs[j]);
                  if (p.relu == 1) r = fmaxf(r, 0.0f);
                  o[j] = __float2bfloat16_rn((cbase + j < p.cout) ? r : 0.0f);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[accbuf]);
    }
    if (prof && row == 0 && grp == 0) {
      unsigned long long* o = p.prof + (size_t)blockIdx.x * kProfSlots;
      o[4] = clock64() - e_begin; o[5] = e_wait;
    }
  }
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  else __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side
__global__ void softargmax_merge_kernel(const sp3d_softargmax_args a, int splits);   // csrc/softargmax.cu

// Debug only (profiles/conv_stalls.py): a device buffer of 148 * kProfSlots u64 receiving per-CTA wait cycles.
static unsigned long long* g_conv_prof = nullptr;
void set_conv_profile(void* dev) { g_conv_prof = static_cast<unsigned long long*>(dev); }
// CTA-pair weight sharing for the large 3^3 / 7^3 launches.  Off by default: bit-identical results and, measured on B200
// (profiles/r02_cta_pair_ab.log), the same step time -- every SM still takes in its full weight stream, so halving the L2
// reads buys nothing (the limit is the SM-side ingest).  Kept behind the switch as the measured experiment.
static int g_conv_pair = 0;
void set_conv_pair(int on) { g_conv_pair = on; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int rb) {
  return rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                   : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (rb == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

template <int KSX, int KS, int RB, int N, int TX, int G, int S, int HB, int SB, int SR, int EG, int F, int WD, int CL = 1>
static int launch_tc(const sp3d_conv_args* a, cudaStream_t st) {
  using C = TcCfg<KSX, KS, RB, N, TX, G, S, HB, SB, SR, EG, F, WD>;
  static_assert(CL == 1 || (C::kTapsPerLoad * N) % 16 == 0, "pair form: half boxes of whole 8-row swizzle atoms");
  EncodeTiledFn encode = get_encode();
  if (encode == nullptr) return SP3D_ERR_UNSUPPORTED;
  const int chunk_ch = C::kPosBytes / 2;
  if (F > 1 && (a->zfold != F || a->W % F || a->OW % F || a->cin_pitch != chunk_ch || a->cout_pitch * F != N ||
                a->fused_phases || a->stride[2] != 1 || a->ostride[2] != 1 || a->tap_off0[2] != -C::kPad))
    return SP3D_ERR_UNSUPPORTED;
  const int nc_block = (a->cin + chunk_ch - 1) / chunk_ch;
  const bool split = a->algo == SP3D_CONV_TC_BF16X3;
  // term pairs (activation term, weight term) per K block, small products first:
  //   3 -> (1,0) (0,1) (0,0);   6 -> (2,0) (1,1) (0,2) (1,0) (0,1) (0,0)
  //   2 -> x0 [w0 | w1] (2 N columns), x1 w0 (upper N columns): the 3 pairs in 2 K blocks (WD = 2 kernels)
  const uint32_t block_act = !split ? 0u : (a->split_terms == 3 ? 0x001u : (a->split_terms == 2 ? 0x10u : 0x001012u));
  const int act_planes = !split ? 1 : (a->split_terms == 6 ? 3 : 2);
  if (split && (a->head_softargmax != nullptr || (a->split_terms != 2 && a->split_terms != 3 && a->split_terms != 6) ||
                a->cin % chunk_ch || a->cin_pitch != a->cin))
    return SP3D_ERR_UNSUPPORTED;
  if ((WD == 2) != (split && a->split_terms == 2)) return SP3D_ERR_UNSUPPORTED;
  const int n_chunks = nc_block * (split ? a->split_terms : 1);
  const int n_tiles = a->fused_phases ? (8 * a->cout) / N : (F > 1 ? 1 : (a->cout + N - 1) / N);

  CUtensorMap map_in, map_w;
  TcStoreMaps maps_out, maps_res;
  memset(&maps_out, 0, sizeof(maps_out));
  memset(&maps_res, 0, sizeof(maps_res));
  const bool pair_out = a->out_dtype == SP3D_BF16X2;   // float32 results as two bf16 term planes [2][N, TD, TH, TW, pitch]
  const int esz = a->out_dtype == SP3D_F32 ? 4 : 2;     // bytes per stored element
  const int elt = pair_out ? 4 : esz;                   // staged bytes per accumulator column
  const bool fused = a->fused_phases != 0;
  const bool tma_store = ((int64_t)a->cout_pitch * esz) % 16 == 0 &&
                         (a->residual == nullptr || reinterpret_cast<uintptr_t>(a->residual) % 16 == 0);
  if (!tma_store && N != 16) return SP3D_ERR_UNSUPPORTED;   // the direct-store epilogue exists in the N = 16 kernels only
  if (fused && (!tma_store || a->cout_pitch != a->cout || (2 * a->cout * elt) % 128 || (8 * a->cout) % N)) return SP3D_ERR_UNSUPPORTED;
  if (pair_out && !tma_store) return SP3D_ERR_UNSUPPORTED;
  if (WD == 2 && (!tma_store || a->head_softargmax != nullptr)) return SP3D_ERR_UNSUPPORTED;
  if (tma_store) {
    // output / residual viewed through the launch's output stride and offset (transposed-convolution phases):
    // [N][OD][OH][OW][cout_pitch] with scaled strides; box = one x-slice of the brick x one staged channel chunk.
    // Fused k2/s2 transposed form: one view per (px, py); (pz, co) is contiguous, so dim 0 spans 2 * cout elements.
    const int64_t pe = (int64_t)a->cout_pitch * esz;
    for (int d = 0; d < 3; ++d) {
      const int full = d == 0 ? a->TD : (d == 1 ? a->TH : a->TW);
      const int og = d == 0 ? a->OD : (d == 1 ? a->OH : a->OW);
      const int off = fused ? a->ostride[d] - 1 : a->ooffset[d];
      if ((int64_t)(og - 1) * a->ostride[d] + off >= full) return SP3D_ERR_INVALID_ARG;
    }
    cuuint64_t gdim[5] = {(cuuint64_t)(fused ? 2 * a->cout : a->cout_pitch * F), (cuuint64_t)a->OW / F, (cuuint64_t)a->OH,
                          (cuuint64_t)a->OD, (cuuint64_t)a->N * (pair_out ? 2 : 1)};
    cuuint64_t gstr[4] = {(cuuint64_t)(pe * a->ostride[2] * F), (cuuint64_t)(pe * a->TW * a->ostride[1]),
                          (cuuint64_t)(pe * a->TW * a->TH * a->ostride[0]), (cuuint64_t)(pe * a->TW * a->TH * a->TD)};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    cuuint32_t box[5] = {(cuuint32_t)(C::kStageRow / elt), (cuuint32_t)kBZ, (cuuint32_t)kBY, 1, 1};
    const CUtensorMapDataType dt = esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle osw = swizzle_for(pair_out ? C::kStageRow / 2 : C::kStageRow);
    for (int mi = 0; mi < (fused ? 4 : 1); ++mi) {
      const int ox = fused ? (mi >> 1) : a->ooffset[0], oy = fused ? (mi & 1) : a->ooffset[1], oz = fused ? 0 : a->ooffset[2];
      const int64_t base_off = (((int64_t)ox * a->TH + oy) * a->TW + oz) * pe;
      if (encode(&maps_out.m[mi], dt, 5, static_cast<uint8_t*>(a->out) + base_off, gdim, gstr, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, osw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return SP3D_ERR_INVALID_ARG;
      if (a->residual != nullptr &&
          encode(&maps_res.m[mi], dt, 5, const_cast<uint8_t*>(static_cast<const uint8_t*>(a->residual)) + base_off, gdim,
                 gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, osw,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return SP3D_ERR_INVALID_ARG;
    }
  }
  {  // activations: [N][D][H][W][cin_pitch] bf16; box = {chunk, HZ, HY, HX, 1} positions, stepped by the conv stride
    // (split-operand mode: the term planes follow each other, plane t of cube n = outer index t * N + n)
    cuuint64_t gdim[5] = {(cuuint64_t)a->cin_pitch * F, (cuuint64_t)a->W / F, (cuuint64_t)a->H, (cuuint64_t)a->D,
                          (cuuint64_t)a->N * act_planes};
    cuuint64_t gstr[4] = {(cuuint64_t)a->cin_pitch * 2 * F, (cuuint64_t)a->cin_pitch * 2 * a->W,
                          (cuuint64_t)a->cin_pitch * 2 * a->W * a->H, (cuuint64_t)a->cin_pitch * 2 * a->W * a->H * a->D};
    cuuint32_t es[5] = {1, (cuuint32_t)a->stride[2], (cuuint32_t)a->stride[1], (cuuint32_t)a->stride[0], 1};
    // with an element stride s the box extent is given in traversed elements: ceil(box / s) elements are loaded
    cuuint32_t box[5] = {(cuuint32_t)(chunk_ch * F), (cuuint32_t)(C::HZ * a->stride[2]), (cuuint32_t)(C::HY * a->stride[1]),
                         (cuuint32_t)((C::kSliced ? 1 : C::HX) * a->stride[0]), 1};
    if (encode(&map_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a->in), gdim, gstr, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(RB), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SP3D_ERR_INVALID_ARG;
  }
  {  // weights: [n_tiles * n_chunks * taps * N rows][chunk channels] bf16 (K-major rows), box = {chunk, taps_per_load * N}
    // rows per channel tile: every chunk holds kTaps taps of N rows (2 N rows in the x0 block of the WD = 2 form)
    const int64_t rows_tile = (int64_t)C::kTaps * N * (WD == 2 ? 3 * nc_block : n_chunks);
    cuuint64_t gdim[2] = {(cuuint64_t)chunk_ch, (cuuint64_t)(rows_tile * n_tiles)};
    cuuint64_t gstr[1] = {(cuuint64_t)C::kPosBytes};
    cuuint32_t box[2] = {(cuuint32_t)chunk_ch, (cuuint32_t)(C::kTapsPerLoad * N / CL)};   // pair form: each CTA loads half
    cuuint32_t es[2] = {1, 1};
    if (encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->weight), gdim, gstr, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(C::kPosBytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SP3D_ERR_INVALID_ARG;
  }
  TcConvParams p{};
  p.n_outer = a->N; p.X = a->OD; p.Y = a->OH; p.Z = a->OW / F;
  p.bricks_x = (a->OD + TX - 1) / TX;
  p.bricks_y = (a->OH + kBY - 1) / kBY;
  p.bricks_z = (a->OW / F + kBZ - 1) / kBZ;
  p.n_bricks = a->N * p.bricks_x * p.bricks_y * p.bricks_z;
  p.n_tiles = n_tiles;
  p.n_chunks = n_chunks;
  p.nc_block = nc_block;
  p.w_rows_tile = C::kTaps * N * (WD == 2 ? 3 * nc_block : n_chunks);
  p.block_act = block_act;
  for (int d = 0; d < 3; ++d) {
    p.istride[d] = a->stride[d];
    p.origin[d] = a->tap_off0[d];
    if (F > 1 && d == 2) p.origin[d] = -((C::kPad + F - 1) / F);   // in rows of F positions
    p.ostride[d] = a->ostride[d];
    p.ooff[d] = a->ooffset[d];
  }
  p.cout = a->cout; p.cout_pitch = a->cout_pitch;
  p.TD = a->TD; p.TH = a->TH; p.TW = a->TW;
  p.relu = a->relu;
  p.out_mode = a->out_dtype == SP3D_F32 ? 1 : (pair_out ? 2 : 0);
  p.scale = a->scale; p.shift = a->shift; p.residual = a->residual; p.out = a->out;
  p.prof = g_conv_prof;
  const sp3d_softargmax_args* head = a->head_softargmax;
  if (head != nullptr) {
    if (N != 16 || F != 1 || n_tiles != 1 || a->fused_phases || head->C < 1 || head->C > 15 || head->n_cubes != a->N ||
        head->X != a->OD || head->Y != a->OH || head->Z != a->OW || a->OD > 256 || a->OH > 256 || a->OW > 256 ||
        head->centers == nullptr || head->out == nullptr || head->lin_x == nullptr || head->lin_y == nullptr ||
        head->lin_z == nullptr || head->check_flag)
      return SP3D_ERR_UNSUPPORTED;
  }
  p.tma_store = tma_store ? 1 : 0;
  p.fused_cols = fused ? 2 * a->cout : 0;
  p.col_mod = fused ? a->cout : (F > 1 ? a->cout_pitch : 0);
  p.has_res = a->residual != nullptr ? 1 : 0;

  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  auto kern = conv_tc_kernel<KSX, KS, RB, N, TX, G, S, HB, SB, SR, EG, F, WD, CL>;
  {  // opt in to the large dynamic shared memory once per (kernel instance, device)
    static unsigned long long done_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((done_mask >> dev) & 1ull)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
      if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
      if (dev < 64) done_mask |= 1ull << dev;
    }
  }
  const int items = p.n_bricks * p.n_tiles;
  int grid = items < n_sm ? items : n_sm;
  if (CL > 1) {
    // persistent CTA pairs: as many clusters as are co-resident (a second wave would double the run time); small launches
    // keep the one-CTA form
    if (head != nullptr || n_tiles != 1 || items < 4 * n_sm) return SP3D_ERR_UNSUPPORTED;
    static int max_clusters[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int mc = dev < 64 ? max_clusters[dev] : 0;
    if (mc == 0) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(n_sm / CL * CL); qc.blockDim = dim3(C::kThreads); qc.dynamicSmemBytes = C::kSmemBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&mc, kern, &qc);
      if (e != cudaSuccess || mc < 1) {          // no cluster support here (e.g. a partitioned GPU): single-CTA form
        cudaGetLastError();
        return SP3D_ERR_UNSUPPORTED;
      }
      if (dev < 64) max_clusters[dev] = mc;
    }
    grid = (items < CL * mc ? items : CL * mc) / CL * CL;
    if (grid < CL) return SP3D_ERR_UNSUPPORTED;
  }
  if (head != nullptr) {
    const int slots = grid * EG;
    const int64_t need = (int64_t)a->N * slots * head->C * 5 * (int64_t)sizeof(double);
    if (head->workspace == nullptr || head->workspace_bytes < need || (reinterpret_cast<uintptr_t>(head->workspace) % 8))
      return SP3D_ERR_WORKSPACE;
    cudaError_t me = cudaMemsetAsync(head->workspace, 0, (size_t)need, st);   // s = 0 marks "no partial from this slot"
    if (me != cudaSuccess) { set_last_error(me); return SP3D_ERR_LAUNCH; }
    p.blocked = 1;
    p.sa_ws = reinterpret_cast<double*>(head->workspace);
    p.sa_lin_x = head->lin_x; p.sa_lin_y = head->lin_y; p.sa_lin_z = head->lin_z;
    p.sa_centers = head->centers; p.sa_center_stride = head->center_stride;
    p.sa_C = head->C; p.sa_slots = slots; p.sa_beta = head->beta;
  }
  if (CL > 1) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(C::kThreads); lc.dynamicSmemBytes = C::kSmemBytes; lc.stream = st;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributeClusterDimension;
    la[0].val.clusterDim.x = CL; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
    lc.attrs = la; lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, kern, map_in, map_w, maps_out, maps_res, p);
    if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
  } else {
    kern<<<grid, C::kThreads, C::kSmemBytes, st>>>(map_in, map_w, maps_out, maps_res, p);
  }
  int rc = check_launch();
  if (rc != SP3D_OK || head == nullptr) return rc;
  const int total = head->n_cubes * head->C;
  softargmax_merge_kernel<<<(total + 3) / 4, 128, 0, st>>>(*head, grid * EG);
  return check_launch();
}

// Shapes taken by the tensor-core path.  `cin` is the padded channel count (a multiple of 16, or of 64 above 64),
// `cout_pitch_w` the output-channel tile N the weights were packed for.
int conv_tc(const sp3d_conv_args* a, cudaStream_t st) {
  if (a->algo != SP3D_CONV_TC_BF16 && a->algo != SP3D_CONV_TC_BF16X3) return SP3D_ERR_UNSUPPORTED;
  if (a->in_dtype != SP3D_BF16) return SP3D_ERR_UNSUPPORTED;
  if (a->out_dtype != SP3D_BF16 && a->out_dtype != SP3D_F32 && a->out_dtype != SP3D_BF16X2) return SP3D_ERR_UNSUPPORTED;
  if (a->out_dtype == SP3D_BF16X2 && (a->head_softargmax != nullptr || (a->cout_pitch % 8))) return SP3D_ERR_UNSUPPORTED;
  if (a->fused_phases && (a->ksize[0] != 1 || a->ksize[1] != 1 || a->ksize[2] != 1 || a->ostride[0] != 2 ||
                          a->ostride[1] != 2 || a->ostride[2] != 2 || a->cout_pitch_w != 128))
    return SP3D_ERR_UNSUPPORTED;
  const int ks = a->ksize[1], ksx = a->ksize[0];
  if (a->ksize[2] != ks || (ksx != ks && ksx != 1)) return SP3D_ERR_UNSUPPORTED;
  for (int d = 0; d < 3; ++d)
    if (a->tap_step[d] != 1 || a->stride[d] < 1 || a->stride[d] > 2) return SP3D_ERR_UNSUPPORTED;
  if (a->stride[0] != 1 || a->stride[1] != a->stride[2]) return SP3D_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a->in) % 16) || (reinterpret_cast<uintptr_t>(a->weight) % 16) ||
      (reinterpret_cast<uintptr_t>(a->out) % 16) || (a->cin_pitch % 8))
    return SP3D_ERR_INVALID_ARG;
  const int cin = a->cin, n = a->cout_pitch_w;
  const int zf = a->zfold > 1 ? a->zfold : 1;
  const int rb = cin >= 64 ? 128 : cin * 2 * zf;
  if (cin >= 64 && (cin % 64)) return SP3D_ERR_UNSUPPORTED;
  if (a->cin_pitch < cin) return SP3D_ERR_INVALID_ARG;
  // (kernel x, kernel yz, row bytes, N, x-slices per brick, taps per weight stage, weight stages)
  const int wd = (a->algo == SP3D_CONV_TC_BF16X3 && a->split_terms == 2) ? 2 : 1;
#define SP3D_TC_CASE_W(KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_, WD_) \
  if (ksx == KSX_ && ks == KS_ && rb == RB_ && n == N_ && zf == F_ && wd == WD_) \
    return launch_tc<KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_, WD_>(a, st);
  // pair-able: large launches run as CTA pairs sharing the weight stream (CL = 2), the rest as single CTAs
#define SP3D_TC_CASE_P(KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_, WD_) \
  if (ksx == KSX_ && ks == KS_ && rb == RB_ && n == N_ && zf == F_ && wd == WD_) { \
    if (g_conv_pair) { \
      const int rc2 = launch_tc<KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_, WD_, 2>(a, st); \
      if (rc2 != SP3D_ERR_UNSUPPORTED) return rc2; \
    } \
    return launch_tc<KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_, WD_>(a, st); \
  }
#define SP3D_TC_CASE_F(KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_) \
  SP3D_TC_CASE_W(KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, F_, 1)
#define SP3D_TC_CASE(KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_) \
  SP3D_TC_CASE_F(KSX_, KS_, RB_, N_, TX_, G_, S_, HB_, SB_, SR_, EG_, 1)
  // (kernel x, kernel yz, row bytes, N, x-slices per brick, taps per weight stage, weight stages, halo buffers,
  //  epilogue staging slots per group, staged row bytes cap, epilogue groups)
  // 3-D (V2VNet)
  SP3D_TC_CASE(7, 7, 32, 16, 2, 49, 2, 2, 2, 128, 2)
  // 7^3 stem, z-folded by 2: rows of 2 positions x 16 channels, N = 2 x 16, 8 windows per (dx, dy); one halo buffer
  SP3D_TC_CASE_F(7, 7, 64, 32, 4, 8, 3, 1, 2, 128, 1, 2)
  // 3^3 16->32 and 32->32, z-folded by 2 (N = 2 x 32, 4 windows per (dx, dy))
  SP3D_TC_CASE_F(3, 3, 64, 64, 4, 4, 3, 2, 2, 128, 1, 2)
  // 1-channel 7^3 stem with the x taps stacked into channels: 1 x 7 x 7, z-folded by 2
  SP3D_TC_CASE_F(1, 7, 64, 32, 4, 8, 3, 2, 2, 128, 2, 2)
  SP3D_TC_CASE_F(3, 3, 128, 64, 4, 4, 2, 1, 2, 128, 1, 2)
  SP3D_TC_CASE(3, 3, 32, 32, 4, 27, 2, 2, 3, 128, 2)
  SP3D_TC_CASE(3, 3, 64, 32, 4, 9, 2, 2, 2, 128, 2)
  SP3D_TC_CASE(3, 3, 64, 64, 4, 3, 3, 2, 2, 64, 2)
  SP3D_TC_CASE(3, 3, 128, 64, 2, 1, 3, 2, 2, 64, 1)
  // (one halo buffer handed over per x-slice leaves room for a 7-stage weight ring: with 2 stages beside two halo
  //  buffers the MMA warps waited 42 % of their time on weights, profiles/r02_conv_stalls_v3.log)
  SP3D_TC_CASE_P(3, 3, 128, 128, 2, 1, 7, 1, 2, 32, 1, 1, 1)
  // adjoint shapes of the training path's input gradients (3^3 32 -> 16 and 64 -> 32: dgrad of 16 -> 32 / 32 -> 64)
  SP3D_TC_CASE(3, 3, 64, 16, 4, 9, 2, 2, 2, 128, 2)
  SP3D_TC_CASE(3, 3, 128, 32, 2, 1, 3, 2, 2, 64, 1)
  // 1x1(x1) on any rank: a work item is a handful of MMAs, so the halo ring is deeper
  SP3D_TC_CASE(1, 1, 32, 32, 4, 1, 2, 6, 3, 128, 2)
  SP3D_TC_CASE(1, 1, 64, 64, 4, 1, 2, 4, 2, 128, 2)
  SP3D_TC_CASE(1, 1, 64, 16, 4, 1, 2, 4, 3, 128, 2)
  SP3D_TC_CASE(1, 1, 128, 16, 4, 1, 2, 3, 3, 128, 2)
  SP3D_TC_CASE(1, 1, 128, 32, 4, 1, 2, 2, 3, 128, 2)
  SP3D_TC_CASE(1, 1, 128, 64, 4, 1, 2, 2, 2, 128, 2)
  SP3D_TC_CASE(1, 1, 128, 128, 2, 1, 3, 3, 2, 128, 2)
  // 2-D (PoseResNet): 3x3, and the 2x2 sub-kernels of the 4x4 stride-2 transposed convolutions
  SP3D_TC_CASE(1, 3, 128, 64, 4, 1, 3, 2, 2, 64, 1)
  SP3D_TC_CASE(1, 3, 128, 128, 2, 3, 2, 2, 2, 128, 1)
  SP3D_TC_CASE(1, 2, 128, 128, 2, 2, 3, 2, 2, 128, 1)
  // 7x7/s2 stem on the 2x2 space-to-depth image (4x4 taps over 16 channels)
  SP3D_TC_CASE(1, 4, 32, 64, 4, 4, 3, 2, 2, 128, 2)
  // split_terms 2 (3 term pairs in 2 K blocks, 2 N accumulator columns): the layers whose narrow channel tile leaves the
  // tensor core waiting on the A operand -- 7^3 stems (z-folded, N = 32) and the 3^3 16 -> 32 layer at N = 32
  SP3D_TC_CASE_P(7, 7, 64, 32, 4, 8, 2, 1, 2, 128, 1, 2, 2)
  // 7^3 stem z-folded by FOUR (rows of 4 positions x 16 channels, N = 4 x 16, 10 windows per (dx, dy)): its 2 N = 128-column
  // MMAs run at the math floor.  TX = 4 x-slices share a weight stage (one accumulator set), the halo is a 5-slot slice ring
  SP3D_TC_CASE_W(7, 7, 128, 64, 4, 5, 3, 1, 2, 32, 2, 4, 2)
  SP3D_TC_CASE_W(1, 7, 64, 32, 4, 8, 2, 2, 2, 128, 2, 2, 2)
  SP3D_TC_CASE_W(3, 3, 64, 32, 4, 3, 3, 2, 2, 128, 2, 1, 2)
  // 3^3 64 -> 64 (N = 64): 2 x 64 accumulator columns per x-slice at TX = 2, one halo buffer.  (The z-folded 16 / 32 -> 32
  // layers would need TX = 4 with ONE accumulator set: measured slower, 14.0 against 12.5 ms per 80 cubes.)
  // (weight rings sized from the in-kernel wait counters, profiles/r02_conv_stalls.log)
  SP3D_TC_CASE_P(3, 3, 128, 64, 2, 1, 6, 1, 2, 64, 1, 1, 2)
  SP3D_TC_CASE_P(3, 3, 128, 64, 2, 4, 3, 1, 2, 128, 1, 2, 2)
  // (16 -> 32 keeps two halo buffers: one sliced buffer + a 6-stage ring measured 4 % slower, halo wait 2 -> 13 %)
  SP3D_TC_CASE_P(3, 3, 64, 64, 2, 4, 4, 2, 2, 128, 2, 2, 2)
#undef SP3D_TC_CASE
#undef SP3D_TC_CASE_F
#undef SP3D_TC_CASE_W
#undef SP3D_TC_CASE_P
  return SP3D_ERR_UNSUPPORTED;
}

}  // namespace sp3d
