// Weight gradient of stride-1 convolutions (3-D and 2-D, any tap window) and of the phases of stride-s transposed
// convolutions on tcgen05 (training path).
//
//   dW[tap, ci, co] = sum over positions p of  x[p + tap_off + tap, ci] * dy[p * g_stride + g_off, co]
//
// contracts over POSITIONS, so both operands have to be K-major in the position index.  Two steps:
//
//  1. wgrad_prep_kernel turns the channel-last float32 tensors into channel-FIRST bf16 term planes (x = hi + lo, as
//     the split-operand forward convolution uses them): every (cube, channel, x, y) owns a contiguous z-line.  For x,
//     one copy per z tap (the line shifted by dz, zero outside) so that a z tap never is a sub-16-byte shift; dx / dy
//     taps are whole-line offsets.  The dy pass also sums the bias gradient.
//  2. wgrad_tc_kernel is a plain tcgen05 GEMM over K blocks = pieces of z-lines.  Its A tile [128 rows][KB positions]
//     is ASSEMBLED BY TMA: rows are (tap, ci) -- 128 / ci taps per tile, each tap one box [ci rows][KB] of the matching
//     shifted copy at (x + dx, y + dy), out-of-range lines zero-filled by TMA (= the convolution's padding); layers
//     with more than 128 input channels use one tap x 128 channels per tile, more than 128 output channels go to
//     grid.z in blocks of 128.  Lines are padded to a multiple of 16 positions (zeros).  B = [dy_hi rows | dy_lo rows].
//     NARROW layers (kz * co <= 128: the 7^3 stem with 16 output channels, the 3^3 layers with 32) fold the z taps into N
//     instead: x stays un-shifted, dy gets one copy per z tap shifted the OTHER way (x[p + tz] dy[p] = x[p'] dy[p' - tz]),
//     the B tile holds kz * co rows per term and one MMA covers all z taps of a (dx, dy) pair -- 7 times fewer A tiles and
//     MMAs of 2 * 112 instead of 2 * 16 columns for the stem.  Per A tile and K step: one MMA of 2 co columns (x_hi * [dy_hi | dy_lo]) and one of
//     co columns (x_lo * dy_hi, onto the upper half) -- the three term pairs of the float32-faithful mode.  Every
//     (tap-tile) keeps its own TMEM accumulator for the whole launch; tap tiles that do not fit 512 columns go to other
//     CTAs (grid.y); the K blocks are split over grid.x; results are added into dW with float atomics at the end.
//
// Replaces the autograd weight gradient (cudnn wgrad in the reference) of nn.Conv3d / nn.ConvTranspose3d in
// lib/models/v2v_net.py:10-69,124 and of the stride-1 nn.Conv2d / the nn.ConvTranspose2d layers of
// lib/models/pose_resnet.py:58-93,161-207.
#include "sp3d_common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include <algorithm>

namespace sp3d {

using namespace tc;

// ------------------------------------------------------------------------------------------------ operand prep
// src: float32 channel-last [N][SX][SY][SZ][pitch]; a LINE is the z-run of one (n, x, y) of the position grid
// [N][X][Y][Z], read at src[n, x * st.x + of.x, y * st.y + of.y, z * st.z + of.z] (st = 1, of = 0 for the forward input;
// the output phase of a transposed convolution for its gradient).  dst: bf16 [2 planes][copies][N][Cp][X*Y][Zp] with
// copy j holding the line shifted by (j * shift_step + shift0): dst[.., z] = line[z + shift] (zero outside [0, Z) and in the
// padding Z .. Zp).  grid = (line walkers, channel chunks of kPrepChunk); bias: per-channel sums of the lines, one
// atomic per channel and CTA (dy pass only).
constexpr int kPrepChunk = 64;
struct PrepGeom {
  int N, X, Y, Z, Zp;
  int SX, SY, SZ;          // source tensor extents
  int st[3], of[3];        // source position = grid position * st + of
  int C, Cp, pitch, copies, shift0, shift_step;      // copy j is shifted by j * shift_step + shift0
};
__global__ void __launch_bounds__(256) wgrad_prep_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         const PrepGeom g, float* __restrict__ bias) {
  __shared__ float tile[128 * (kPrepChunk + 1)];      // [Z <= 128][chunk + 1]
  constexpr int cs = kPrepChunk + 1;
  const int c0 = blockIdx.y * kPrepChunk;
  const int cn = min(kPrepChunk, g.Cp - c0);          // channels of this chunk (Cp is a multiple of 16)
  const int zg = g.Zp / 8;                            // 16-byte groups of 8 positions
  const int XY = g.X * g.Y;
  const int64_t plane = (int64_t)g.copies * g.N * g.Cp * XY * g.Zp;
  float bsum = 0.0f;                                  // thread c < cn: running sum of channel c0 + c
  for (int line = blockIdx.x; line < g.N * XY; line += gridDim.x) {
    const int n = line / XY, xy = line % XY;
    const int x = xy / g.Y, y = xy % g.Y;
    const float* s = src + ((((int64_t)n * g.SX + (x * g.st[0] + g.of[0])) * g.SY + (y * g.st[1] + g.of[1])) * g.SZ + g.of[2]) *
                               g.pitch + c0;
    const int64_t zstep = (int64_t)g.st[2] * g.pitch;
    __syncthreads();                                  // the previous line's readers are done with the tile
    for (int i = threadIdx.x; i < g.Z * cn; i += blockDim.x) {
      const int z = i / cn, c = i % cn;
      tile[z * cs + c] = (c0 + c < g.C) ? __ldg(s + z * zstep + c) : 0.0f;
    }
    __syncthreads();
    if (bias != nullptr && (int)threadIdx.x < cn && c0 + (int)threadIdx.x < g.C) {
      for (int z = 0; z < g.Z; ++z) bsum += tile[z * cs + threadIdx.x];
    }
    for (int i = threadIdx.x; i < g.copies * cn * zg; i += blockDim.x) {
      const int q8 = i % zg, c = (i / zg) % cn, j = i / (zg * cn);
      __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int zo = q8 * 8 + q, z = zo + j * g.shift_step + g.shift0;
        const float v = (zo < g.Z && z >= 0 && z < g.Z) ? tile[z * cs + c] : 0.0f;
        hi[q] = __float2bfloat16_rn(v);
        lo[q] = __float2bfloat16_rn(v - __bfloat162float(hi[q]));
      }
      const int64_t o = ((((int64_t)j * g.N + n) * g.Cp + c0 + c) * XY + xy) * g.Zp + q8 * 8;
      *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(dst + plane + o) = *reinterpret_cast<const uint4*>(lo);
    }
  }
  if (bias != nullptr && (int)threadIdx.x < cn && c0 + (int)threadIdx.x < g.C) atomicAdd(bias + c0 + threadIdx.x, bsum);
}

// ------------------------------------------------------------------------------------------------ GEMM
struct WgradParams {
  int N, X, Y, Z, taps;          // Z: padded line length
  int kx, ky, kz, ox, oy;        // tap window; x / y offset of tap 0 (the z offset lives in the shifted copies)
  int Cr, cb;                    // rows of one tap in an A tile (min(Cp, 128)) / rows of one dy term in the B tile
  int cin, cout;
  int taps_per_tile, ci_tiles, n_tiles, tiles_per_group;   // ci_tiles = Cp / 128 tiles per tap where Cp > 128
  int copies;                    // z-shifted copies of x (kz, or 1 when the z taps are folded into N)
  int nfold;                     // z taps folded into N: the B tile holds nfold shifted copies of dy per term (else 1)
  int n_kblocks, kb_per_line;    // K blocks = (n, x, y, z piece)
  float* gw;                     // [taps][gw_cin][gw_pitch] float32, added into
  int gw_cin, gw_pitch;
};

constexpr int kWgStages = 4;      // A stages (hi tile + lo tile each)

template <int KB>                 // positions per K block: 64 / 32 / 16 -> 128 / 64 / 32-byte swizzled rows
__global__ void __launch_bounds__(128, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WgradParams p) {
  constexpr int RB = KB * 2;
  constexpr int kTile = 128 * RB;                       // one A tile (one term plane)
  constexpr uint32_t kLayout = RB == 128 ? kSwizzle128 : (RB == 64 ? kSwizzle64 : kSwizzle32);
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[kWgStages], a_empty[kWgStages], b_full[2], b_empty[2], done;
  __shared__ uint32_t tmem_base_s;
  __shared__ int3 s_tap[32 * 8];                        // <= 32 tiles per CTA x 8 tap slots per tile
  __shared__ int s_ci0[32];                             // first input channel of each tile
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_buf = smem;                                // [kWgStages][2 planes][kTile]
  uint8_t* b_buf = smem + kWgStages * 2 * kTile;        // [2][256 rows * RB]
  constexpr int kBStride = 256 * RB;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (tid == 0) {
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  const int group = blockIdx.y;
  const int tile0 = group * p.tiles_per_group;
  const int my_tiles = min(p.tiles_per_group, p.n_tiles - tile0);
  const int cbn = p.nfold * p.cb;                       // columns of one dy term in the B tile / accumulator
  const int acc_cols = 2 * cbn;
  const int co0 = blockIdx.z * p.cb;                    // first output channel of this CTA
  // tap table of this CTA's tiles: slot s of tile t -> (dz copy, y offset, x offset); taps beyond the kernel get an x
  // offset far outside the tensor (TMA zero-fills the box).  Cp > 128: one tap per tile, ci_tiles tiles per tap.
  for (int i = tid; i < my_tiles * 8; i += 128) {
    const int tile = tile0 + i / 8;
    const int tap = (tile / p.ci_tiles) * p.taps_per_tile + (i % 8);
    int3 tp = make_int3(0, 0, -100000);
    if ((i % 8) < p.taps_per_tile && tap < p.taps)
      tp = make_int3(tap % p.kz, (tap / p.kz) % p.ky + p.oy, tap / (p.kz * p.ky) + p.ox);
    s_tap[i] = tp;
    if ((i % 8) == 0) s_ci0[i / 8] = (tile % p.ci_tiles) * 128;
  }
  __syncthreads();
  auto kblock_coords = [&](int kb, int& n, int& x, int& y, int& z0) {
    z0 = (kb % p.kb_per_line) * KB;
    const int line = kb / p.kb_per_line;
    y = line % p.Y;
    x = (line / p.Y) % p.X;
    n = line / (p.Y * p.X);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: the whole warp.  Lane s issues the
    // two boxes (hi / lo plane) of tap slot s of every A tile -- one thread issuing all of a tile's boxes (up to 16 for
    // the 7^3 stem) was the bottleneck of the first version (about 300 cycles of address arithmetic per box).
    uint32_t sa = 0, sb = 0;
    const int slot = lane;
    for (int kb = blockIdx.x; kb < p.n_kblocks; kb += gridDim.x, ++sb) {
      int n, x, y, z0;
      kblock_coords(kb, n, x, y, z0);
      const uint32_t bb = sb & 1;
      mbar_wait(&b_empty[bb], ((sb >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&b_full[bb], (uint32_t)(2 * p.nfold * p.cb * RB));
        uint8_t* bdst = b_buf + bb * kBStride;
        // rows [dy_hi copy 0 .. nfold-1 | dy_lo copy 0 .. nfold-1]; outer index (plane * nfold + copy) * N + n
        for (int j = 0; j < p.nfold; ++j) {
          tma_load_5d(bdst + j * p.cb * RB, &map_dy, &b_full[bb], z0, y, x, co0, j * p.N + n);
          tma_load_5d(bdst + (p.nfold + j) * p.cb * RB, &map_dy, &b_full[bb], z0, y, x, co0, (p.nfold + j) * p.N + n);
        }
      }
      for (int t = 0; t < my_tiles; ++t, ++sa) {
        const uint32_t st = sa % kWgStages;
        mbar_wait(&a_empty[st], ((sa / kWgStages) & 1) ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&a_full[st], 2u * kTile);
        __syncwarp();
        if (slot < p.taps_per_tile) {
          uint8_t* adst = a_buf + st * 2 * kTile + slot * p.Cr * RB;
          const int3 tp = s_tap[t * 8 + slot];               // (dz copy, y offset, x offset); x far outside: no such tap
          const int ci0 = s_ci0[t];
          // outer index: (plane * copies + dz) * N + n
          tma_load_5d(adst, &map_x, &a_full[st], z0, y + tp.y, x + tp.z, ci0, tp.x * p.N + n);
          tma_load_5d(adst + kTile, &map_x, &a_full[st], z0, y + tp.y, x + tp.z, ci0, (p.copies + tp.x) * p.N + n);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform, elect-predicated)
    const uint32_t idesc_w = make_idesc(kFmtBF16, 128, (uint32_t)(2 * cbn));
    const uint32_t idesc_n = make_idesc(kFmtBF16, 128, (uint32_t)cbn);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(a_buf), 0, 8 * RB, kLayout);
    const uint64_t b_desc0 = make_smem_desc(smem_u32(b_buf), 0, 8 * RB, kLayout);
    uint32_t sa = 0, sb = 0;
    bool first = true;
    for (int kb = blockIdx.x; kb < p.n_kblocks; kb += gridDim.x, ++sb) {
      const uint32_t bb = sb & 1;
      mbar_wait(&b_full[bb], (sb >> 1) & 1);
      tc_fence_after();
      const uint64_t bd = b_desc0 + (uint64_t)((bb * kBStride) >> 4);
      for (int t = 0; t < my_tiles; ++t, ++sa) {
        const uint32_t st = sa % kWgStages;
        mbar_wait(&a_full[st], (sa / kWgStages) & 1);
        tc_fence_after();
        const uint64_t ad_hi = a_desc0 + (uint64_t)((st * 2 * kTile) >> 4);
        const uint64_t ad_lo = ad_hi + (uint64_t)(kTile >> 4);
        const uint32_t d = tmem_base + (uint32_t)(t * acc_cols);
#pragma unroll
        for (int ks = 0; ks < KB / 16; ++ks) {
          if (elect_one_sync()) {
            // x_hi * [dy_hi | dy_lo] -> 2 co columns; x_lo * dy_hi -> the upper co columns (small products together)
            mma_f16_ss(d, ad_hi + 2 * ks, bd + 2 * ks, idesc_w, (first && ks == 0) ? 0u : 1u);
            mma_f16_ss(d + (uint32_t)cbn, ad_lo + 2 * ks, bd + 2 * ks, idesc_n, 1u);
          }
        }
        if (elect_one_sync()) mma_commit(&a_empty[st]);
      }
      if (elect_one_sync()) mma_commit(&b_empty[bb]);
      first = false;
    }
    if (elect_one_sync()) mma_commit(&done);
  }
  // ------------------------------------------------------------------ epilogue: all four warps, thread <-> A row
  __syncthreads();
  const bool any = (int)blockIdx.x < p.n_kblocks;
  if (any) {
    mbar_wait(&done, 0);
    tc_fence_after();
    const int row = tid;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = tile0 + t;
      const int tap = (tile / p.ci_tiles) * p.taps_per_tile + row / p.Cr;
      const int ci = (tile % p.ci_tiles) * 128 + row % p.Cr;
      const bool ok = row / p.Cr < p.taps_per_tile && tap < p.taps && ci < p.cin;
      // folded z taps: the row's tap is the (dx, dy) pair, column block f is z tap f
      for (int f = 0; f < p.nfold; ++f) {
        float* o = p.gw + ((int64_t)(tap * p.nfold + f) * p.gw_cin + ci) * p.gw_pitch + co0;
        for (int c0 = 0; c0 < p.cb; c0 += 16) {
          uint32_t v0[16], v1[16];
          tmem_ld_x16(tmem_base + lane_base + (uint32_t)(t * acc_cols + f * p.cb + c0), v0);
          tmem_ld_x16(tmem_base + lane_base + (uint32_t)(t * acc_cols + cbn + f * p.cb + c0), v1);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float g = __uint_as_float(v0[j]) + __uint_as_float(v1[j]);
              if (co0 + c0 + j < p.cout && g != 0.0f) atomicAdd(o + c0 + j, g);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode() {
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

struct WgShape {
  int Cp, Cr, co16, cb, co_blocks, Zp, KB, taps, taps_per_tile, ci_tiles, n_tiles, tiles_per_group, groups;
  int nfold;                    // z taps folded into N (narrow layers), else 1; `taps` then counts the (dx, dy) pairs
  int64_t x_elems, dy_elems;    // bf16 elements of the two workspace regions
};

static int wg_shape(const sp3d_conv_wgrad_tc_args* a, WgShape* s) {
  if (a == nullptr || a->N < 1 || a->X < 1 || a->Y < 1 || a->Z < 1 || a->cin < 1 || a->cout < 1 || a->x_pitch < a->cin ||
      a->g_pitch < a->cout)
    return SP3D_ERR_INVALID_ARG;
  for (int d = 0; d < 3; ++d)
    if (a->ksize[d] < 1 || a->ksize[d] > 7 || a->g_stride[d] < 1 || a->g_off[d] < 0) return SP3D_ERR_INVALID_ARG;
  if (a->Z > 128) return SP3D_ERR_UNSUPPORTED;                       // one line has to fit the prep tile
  s->Cp = (a->cin + 15) / 16 * 16;
  s->co16 = (a->cout + 15) / 16 * 16;
  if (s->Cp > 128) s->Cp = (s->Cp + 127) / 128 * 128;
  if (s->Cp != 16 && s->Cp != 32 && s->Cp != 64 && (s->Cp % 128)) return SP3D_ERR_UNSUPPORTED;
  if (s->co16 > 128 && (s->co16 % 128)) return SP3D_ERR_UNSUPPORTED;
  s->Cr = s->Cp < 128 ? s->Cp : 128;
  s->cb = s->co16 < 128 ? s->co16 : 128;
  s->co_blocks = s->co16 / s->cb;
  s->Zp = (a->Z + 15) / 16 * 16;
  s->KB = s->Zp % 64 == 0 ? 64 : (s->Zp % 32 == 0 ? 32 : 16);
  // narrow layers: all z taps of a (dx, dy) pair in one MMA (kz * co columns per term; 2 * kz * co <= 256)
  s->nfold = (a->ksize[2] > 1 && a->ksize[2] * s->co16 <= 128 && s->Cp <= 128) ? a->ksize[2] : 1;
  s->taps = a->ksize[0] * a->ksize[1] * (s->nfold > 1 ? 1 : a->ksize[2]);
  s->taps_per_tile = 128 / s->Cr;
  s->ci_tiles = s->Cp <= 128 ? 1 : s->Cp / 128;
  s->n_tiles = (s->taps + s->taps_per_tile - 1) / s->taps_per_tile * s->ci_tiles;
  s->tiles_per_group = 512 / (2 * s->nfold * s->cb);
  if (s->tiles_per_group > s->n_tiles) s->tiles_per_group = s->n_tiles;
  s->groups = (s->n_tiles + s->tiles_per_group - 1) / s->tiles_per_group;
  if (s->groups > 65535 || s->co_blocks > 65535) return SP3D_ERR_UNSUPPORTED;
  const int64_t vox = (int64_t)a->N * a->X * a->Y * s->Zp;
  s->x_elems = 2 * (int64_t)(s->nfold > 1 ? 1 : a->ksize[2]) * s->Cp * vox;
  s->dy_elems = 2 * (int64_t)s->nfold * s->co16 * vox;
  if ((int64_t)a->N * a->X * a->Y > 2147483647LL || 2 * (int64_t)a->ksize[2] * a->N > 2147483647LL) return SP3D_ERR_UNSUPPORTED;
  // the gradient positions p * g_stride + g_off must lie inside grad_out
  const int ge[3] = {a->GX, a->GY, a->GZ}, pe[3] = {a->X, a->Y, a->Z};
  for (int d = 0; d < 3; ++d)
    if ((int64_t)(pe[d] - 1) * a->g_stride[d] + a->g_off[d] >= ge[d]) return SP3D_ERR_INVALID_ARG;
  return SP3D_OK;
}

template <int KB>
static int wg_launch(const sp3d_conv_wgrad_tc_args* a, const WgShape& s, const CUtensorMap& mx, const CUtensorMap& mdy,
                     cudaStream_t st) {
  WgradParams p{};
  p.N = a->N; p.X = a->X; p.Y = a->Y; p.Z = s.Zp; p.taps = s.taps;
  p.kx = a->ksize[0]; p.ky = a->ksize[1]; p.kz = s.nfold > 1 ? 1 : a->ksize[2]; p.ox = a->tap_off[0]; p.oy = a->tap_off[1];
  p.nfold = s.nfold;
  p.Cr = s.Cr; p.cb = s.cb; p.cin = a->cin; p.cout = a->cout;
  p.taps_per_tile = s.taps_per_tile; p.ci_tiles = s.ci_tiles; p.n_tiles = s.n_tiles; p.tiles_per_group = s.tiles_per_group;
  p.copies = s.nfold > 1 ? 1 : a->ksize[2];
  p.kb_per_line = s.Zp / KB;
  p.n_kblocks = a->N * a->X * a->Y * p.kb_per_line;
  p.gw = a->grad_weight; p.gw_cin = a->gw_cin; p.gw_pitch = a->gw_pitch;
  constexpr int kSmem = kWgStages * 2 * 128 * KB * 2 + 2 * 256 * KB * 2 + 1024;
  auto kern = wgrad_tc_kernel<KB>;
  {  // opt in to the large dynamic shared memory once per device
    static unsigned long long done_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((done_mask >> dev) & 1ull)) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      if (e != cudaSuccess) { set_last_error(e); return SP3D_ERR_LAUNCH; }
      if (dev < 64) done_mask |= 1ull << dev;
    }
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  int gx = n_sm / (s.groups * s.co_blocks);
  if (gx < 1) gx = 1;
  if (gx > p.n_kblocks) gx = p.n_kblocks;
  kern<<<dim3(gx, s.groups, s.co_blocks), 128, kSmem, st>>>(mx, mdy, p);
  return check_launch();
}

}  // namespace sp3d

extern "C" int64_t sp3d_conv_wgrad_tc_workspace(const sp3d_conv_wgrad_tc_args* a) {
  sp3d::WgShape s;
  if (sp3d::wg_shape(a, &s) != SP3D_OK) return -1;
  return (s.x_elems + s.dy_elems) * 2 + 1024;
}

extern "C" int sp3d_conv_wgrad_tc(const sp3d_conv_wgrad_tc_args* a, void* stream) {
  using namespace sp3d;
  WgShape s;
  int rc = wg_shape(a, &s);
  if (rc != SP3D_OK) return rc;
  if (a->x == nullptr || a->grad_out == nullptr || a->grad_weight == nullptr || a->workspace == nullptr ||
      a->gw_cin < a->cin || a->gw_pitch < a->cout)
    return SP3D_ERR_INVALID_ARG;
  if (a->workspace_bytes < (s.x_elems + s.dy_elems) * 2 + 1024 || (reinterpret_cast<uintptr_t>(a->workspace) % 16))
    return SP3D_ERR_WORKSPACE;
  EncodeTiledFn encode = wg_encode();
  if (encode == nullptr) return SP3D_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a->workspace) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* xT = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* dyT = xT + s.x_elems;
  const int lines = a->N * a->X * a->Y;
  PrepGeom gx{};
  gx.N = a->N; gx.X = a->X; gx.Y = a->Y; gx.Z = a->Z; gx.Zp = s.Zp;
  gx.SX = a->X; gx.SY = a->Y; gx.SZ = a->Z;
  for (int d = 0; d < 3; ++d) { gx.st[d] = 1; gx.of[d] = 0; }
  gx.C = a->cin; gx.Cp = s.Cp; gx.pitch = a->x_pitch; gx.shift_step = 1;
  gx.copies = s.nfold > 1 ? 1 : a->ksize[2];
  gx.shift0 = s.nfold > 1 ? 0 : a->tap_off[2];
  PrepGeom gd = gx;
  gd.SX = a->GX; gd.SY = a->GY; gd.SZ = a->GZ;
  for (int d = 0; d < 3; ++d) { gd.st[d] = a->g_stride[d]; gd.of[d] = a->g_off[d]; }
  gd.C = a->cout; gd.Cp = s.co16; gd.pitch = a->g_pitch;
  // folded z taps: copy j of dy holds dy[z - (tap_off_z + j)], so that x[z] * dy_j[z] = x[p + tap_off_z + j] * dy[p]
  gd.copies = s.nfold; gd.shift_step = -1; gd.shift0 = s.nfold > 1 ? -a->tap_off[2] : 0;
  const int chunks_x = (s.Cp + kPrepChunk - 1) / kPrepChunk, chunks_d = (s.co16 + kPrepChunk - 1) / kPrepChunk;
  const int walkers_x = std::max(1, std::min(lines, 148 * 8 / chunks_x)), walkers_d = std::max(1, std::min(lines, 148 * 8 / chunks_d));
  wgrad_prep_kernel<<<dim3(walkers_x, chunks_x), 256, 0, st>>>(a->x, xT, gx, nullptr);
  rc = check_launch();
  if (rc != SP3D_OK) return rc;
  wgrad_prep_kernel<<<dim3(walkers_d, chunks_d), 256, 0, st>>>(a->grad_out, dyT, gd, a->grad_bias);
  rc = check_launch();
  if (rc != SP3D_OK) return rc;

  CUtensorMap mx, mdy;
  const CUtensorMapSwizzle sw = s.KB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (s.KB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const cuuint64_t Zp = (cuuint64_t)s.Zp;
  {  // x^T: [2 planes * kz copies * N][Cp][X][Y][Zp] bf16; box = {KB, 1, 1, Cr, 1}
    cuuint64_t gdim[5] = {Zp, (cuuint64_t)a->Y, (cuuint64_t)a->X, (cuuint64_t)s.Cp,
                          (cuuint64_t)(2 * (s.nfold > 1 ? 1 : a->ksize[2]) * a->N)};
    cuuint64_t gstr[4] = {Zp * 2, Zp * a->Y * 2, Zp * a->Y * a->X * 2, Zp * a->Y * a->X * s.Cp * 2};
    cuuint32_t box[5] = {(cuuint32_t)s.KB, 1, 1, (cuuint32_t)s.Cr, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (encode(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, xT, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SP3D_ERR_INVALID_ARG;
  }
  {  // dy^T: [2 planes * N][co16][X][Y][Zp] bf16; box = {KB, 1, 1, cb, 1}
    cuuint64_t gdim[5] = {Zp, (cuuint64_t)a->Y, (cuuint64_t)a->X, (cuuint64_t)s.co16, (cuuint64_t)(2 * s.nfold * a->N)};
    cuuint64_t gstr[4] = {Zp * 2, Zp * a->Y * 2, Zp * a->Y * a->X * 2, Zp * a->Y * a->X * s.co16 * 2};
    cuuint32_t box[5] = {(cuuint32_t)s.KB, 1, 1, (cuuint32_t)s.cb, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (encode(&mdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dyT, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return SP3D_ERR_INVALID_ARG;
  }
  if (s.KB == 64) return wg_launch<64>(a, s, mx, mdy, st);
  if (s.KB == 32) return wg_launch<32>(a, s, mx, mdy, st);
  return wg_launch<16>(a, s, mx, mdy, st);
}
