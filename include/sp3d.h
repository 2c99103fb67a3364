/*
 * libsp3d -- C ABI of the B200 (sm_100a) kernels behind the SelfPose3d / VoxelPose
 * voxelised multi-view pose path.
 *
 * The reference (CAMMA-public/SelfPose3d) has no FFI layer: its replaceable seam is the
 * Python nn.Module API.  Each entry point below replaces the ATen /
 * cuDNN call sequence of one reference function; the citation says which.
 *
 * Conventions (all entry points):
 *   - plain C, POD argument structs, raw DEVICE pointers, explicit sizes and strides;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, performs no
 *     host synchronisation, allocates and frees nothing (scratch comes from the caller);
 *   - returns 0 (SP3D_OK) or a negative sp3d_status; never throws; re-entrant (no global state);
 *   - float tensors are IEEE fp32 unless a dtype field says otherwise;
 *   - "channel-last" means [.., spatial.., C_pitch] with the channel index fastest.
 */
#ifndef SP3D_H
#define SP3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP3D_ABI_VERSION 4   /* 2: sp3d_conv_args.split_terms, sp3d_split_bf16, the backward operators;
                                3: SP3D_BF16X2 activations, sp3d_merge_bf16, sp3d_s2d_args.dst_dtype, sp3d_gauss_render_*,
                                   split_terms 2, sp3d_target_heatmaps / sp3d_target_volume, sp3d_conv_wgrad_tc;
                                4: grouped BatchNorm statistics (sp3d_bn_*_args), sp3d_debug_conv_pair, general sp3d_conv_wgrad_tc_args */
#define SP3D_MAX_VIEWS 8
#define SP3D_CAM_FLOATS 32

typedef enum {
  SP3D_OK = 0,
  SP3D_ERR_INVALID_ARG = -1,
  SP3D_ERR_UNSUPPORTED = -2,
  SP3D_ERR_LAUNCH = -3,
  SP3D_ERR_WORKSPACE = -4
} sp3d_status;

/* SP3D_BF16X2: a float32 tensor held as TWO bf16 term planes [2][...same shape...] (plane 0 = bf16(x), plane 1 =
 * bf16(x - plane 0); x ~= plane 0 + plane 1 to 2^-17 relative) -- the activation format of the float32-faithful
 * tensor-core mode: what sp3d_split_bf16 (S = 2) produces, what SP3D_CONV_TC_BF16X3 (split_terms = 3) reads, and what
 * its epilogue, sp3d_maxpool_fwd, sp3d_unproject_fwd and sp3d_space_to_depth can write directly, so that no separate
 * split pass runs between layers.  The pointer addresses plane 0; plane 1 follows the whole of plane 0. */
typedef enum { SP3D_F32 = 0, SP3D_BF16 = 1, SP3D_F16 = 2, SP3D_BF16X2 = 3 } sp3d_dtype;

/* ABI version of the loaded library and a static description of an error code. */
int sp3d_abi_version(void);
const char* sp3d_strerror(int status);
/* Last CUDA error string recorded by a failed launch on the calling thread ("" if none). */
const char* sp3d_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * Packed camera record, one per (sample, view): SP3D_CAM_FLOATS fp32 values
 *   [0..8] R row-major   [9..11] T (camera centre, world mm)   [12,13] fx fy   [14,15] cx cy
 *   [16..18] k1 k2 k3    [19,20] p1 p2    [21..26] 2x3 affine original-image -> network input
 *   [27] width = 2*center_x   [28] height = 2*center_y   [29] flip (0/1)   [30,31] reserved
 * Replaces the per-(sample, view) host work of lib/models/project_layer.py:64-75 and
 * lib/utils/cameras.py:13-24 (unfold_camera_param).
 * ------------------------------------------------------------------------------------------ */

/* Fused multi-view un-projection: per voxel back-projection into every view, bilinear sample
 * of all channels, masked mean over views, clamp to [0,1].
 * Replaces ProjectLayer.get_voxel / forward, lib/models/project_layer.py:42-106, together with
 * compute_grid (:22-40), cameras.project_pose (lib/utils/cameras.py:27-55,111-113) and
 * affine_transform_pts_cuda (lib/utils/transforms.py:119-123). */
typedef struct {
  const float* heatmaps[SP3D_MAX_VIEWS]; /* per view: [B, C, h, w] addressed through the strides below */
  int64_t hm_stride_b, hm_stride_c, hm_stride_h, hm_stride_w; /* in elements */
  const float* cams;        /* [B, V, SP3D_CAM_FLOATS] */
  const float* centers;     /* [n_cubes, center_stride]: x, y, z, flag(, score) */
  int center_stride;        /* floats between consecutive cube centres (>= 3) */
  int check_flag;           /* 1: cubes with centers[.][3] < 0 are skipped and written as zeros */
  int cubes_per_sample;     /* sample of cube q is q / cubes_per_sample ... */
  const int32_t* cube_sample; /* ... unless this optional [n_cubes] table gives it (compacted proposals) */
  const float* lin_x;       /* [X] torch.linspace(-size/2, size/2, X) values (no centre) */
  const float* lin_y;       /* [Y] */
  const float* lin_z;       /* [Z] */
  int B, V, C, h, w;
  int n_cubes, X, Y, Z;
  float img_w, img_h;       /* cfg.NETWORK.IMAGE_SIZE */
  float hm_cfg_w, hm_cfg_h; /* cfg.NETWORK.HEATMAP_SIZE: scales network-input pixels to heat-map coordinates
                               (project_layer.py:84-90); normally == (w, h).  The tensor extents (w, h) are what
                               grid_sample un-normalises and bounds-checks against (:93). */
  int view_begin, view_end; /* views summed by this call (multi-GPU view sharding); 0, V for all */
  int partial;              /* 0: write clamp(num/(den+1e-6),0,1); 1: write raw numerators and the
                               view count as channel C (all-reduce, then sp3d_unproject_finalize) */
  void* cubes;              /* output, dtype out_dtype */
  int out_dtype;            /* sp3d_dtype */
  int64_t out_stride_cube, out_stride_c, out_stride_vox; /* in elements; voxel = (ix*Y+iy)*Z+iz */
  int out_c_pad;            /* channels [C, out_c_pad) are written as zeros (channel-last padding); 0 = none */
  float* grids;             /* optional [n_cubes, X*Y*Z, 3] voxel-centre coordinates, or NULL */
  int hm_dtype;             /* sp3d_dtype of the heat-maps: SP3D_F32, or SP3D_F16 (math_mode 1 only) */
  int math_mode;            /* 0: float32 arithmetic in the reference's operation order (bit-faithful mask, parity
                               form).  1: throughput form for the bf16 volume mode -- FMA-contracted projection,
                               half2 tap blending; needs SP3D_F16 channel-last heat-maps with 16 channels per pixel
                               (sp3d_heatmaps_to_f16), bf16 channel-last cubes with pitch 16, all views, no grids */
} sp3d_unproject_args;
int sp3d_unproject_fwd(const sp3d_unproject_args* a, void* stream);

/* float32 heat-maps of V views ([B, C, h, w] through the strides, C <= 16) -> fp16 channel-last
 * [V][B][h][w][16] (channels >= C zero): the input layout of sp3d_unproject_fwd's math_mode 1. */
typedef struct {
  const float* heatmaps[SP3D_MAX_VIEWS];
  int64_t stride_b, stride_c, stride_h, stride_w; /* in elements */
  int V, B, C, h, w;
  void* out;
} sp3d_heatmaps_f16_args;
int sp3d_heatmaps_to_f16(const sp3d_heatmaps_f16_args* a, void* stream);

/* Divide / NaN / clamp step of the un-projection after partial sums were all-reduced:
 * buf is [n, (C+1), N] channel-first or [n, N, pitch] channel-last partial output. */
typedef struct {
  float* buf;
  int64_t n_cubes, C, N;
  int64_t stride_cube, stride_c, stride_vox;
} sp3d_unproject_finalize_args;
int sp3d_unproject_finalize(const sp3d_unproject_finalize_args* a, void* stream);

/* 3-D NMS (3x3x3 local maxima, non-maxima zeroed) + top-K + index -> world location + flag.
 * Replaces core.proposal.nms/max_pool/get_index (lib/core/proposal.py:18-48) and
 * ProposalLayer[Soft].forward/get_real_loc (lib/models/cuboid_proposal_net_soft.py:46-68,
 * lib/models/cuboid_proposal_net.py:42-83, no-GT branch).
 * Tie policy: highest value, then lowest flat index. */
typedef struct {
  const float* root_cubes;  /* [B, X, Y, Z] contiguous */
  int B, X, Y, Z, K;        /* K <= 32 */
  float threshold;
  double space_size[3], space_center[3];
  int loc_f64;              /* 1: evaluate get_real_loc in float64 after the float32 idx/(n-1)
                               (config values were float64 numpy arrays), 0: all float32 (YAML lists) */
  float* grid_centers;      /* [B, K, 5]: x, y, z, flag = (score > threshold) - 1, score */
  int32_t* topk_index;      /* optional [B, K] flat voxel index, or NULL */
  void* workspace;          /* optional sp3d_nms_topk3d_workspace() bytes: 16 CTAs per sample (x slabs) + a merge
                               instead of one CTA per sample; same result */
  int64_t workspace_bytes;
} sp3d_nms_topk_args;
int64_t sp3d_nms_topk3d_workspace(const sp3d_nms_topk_args* a);
int sp3d_nms_topk3d(const sp3d_nms_topk_args* a, void* stream);

/* Soft-argmax over a voxel cube: sum_v softmax(beta * x)_v * grid_v, one (x,y,z) per channel.
 * Replaces SoftArgmaxLayer.forward, lib/models/pose_regression_net.py:19-28.  The voxel
 * coordinates are rebuilt from lin_* + centre (bit-identical to ProjectLayer's `grids`). */
typedef struct {
  const void* x;            /* [n_cubes, C, N] through strides, dtype x_dtype */
  int x_dtype;
  int64_t stride_cube, stride_c, stride_vox;
  int n_cubes, C, X, Y, Z;
  const float* centers;     /* [n_cubes, center_stride] */
  int center_stride;
  int check_flag;           /* 1: cubes with centers[.][3] < 0 produce zeros */
  const float* lin_x; const float* lin_y; const float* lin_z;
  float beta;
  float* out;               /* [n_cubes, C, 3] */
  float* workspace;         /* sp3d_softargmax3d_workspace() bytes */
  int64_t workspace_bytes;
} sp3d_softargmax_args;
int64_t sp3d_softargmax3d_workspace(const sp3d_softargmax_args* a);
int sp3d_softargmax3d_fwd(const sp3d_softargmax_args* a, void* stream);

/* Convolution family on channel-last activations, evaluated as an implicit GEMM
 *   out[n, o*ostride+ooffset, co] = act( scale[co] * sum_{t,ci} in[n, o*stride + tap_off0 + t*tap_step, ci]
 *                                        * weight[t, ci, co] + shift[co] (+ residual) )
 * with zero padding outside the input.  2-D tensors use D = 1.  One call covers a regular
 * convolution; a transposed convolution with kernel = m*stride is stride^d calls (one per output
 * phase) with ostride = stride, ooffset = phase.
 * Replaces cudnn conv3d/conv_transpose3d + batch_norm + relu (+ residual add) of
 * lib/models/v2v_net.py:10-69,124 and conv2d/conv_transpose2d + batch_norm + relu of
 * lib/models/pose_resnet.py:58-93,102-124,161-207 (evaluation-mode BatchNorm folded into
 * scale/shift by the caller). */
typedef enum {
  SP3D_CONV_SIMT_F32 = 0,   /* float32 FMA, any shape */
  SP3D_CONV_TC_BF16 = 1,    /* tcgen05, bf16 operands, float32 accumulation */
  SP3D_CONV_TC_BF16X3 = 2   /* tcgen05 on float32 data split into bf16 terms (split_terms below): float32-faithful */
} sp3d_conv_algo;
typedef struct {
  const void* in;           /* [N, D, H, W, cin_pitch] */
  const void* weight;       /* [ntaps, cin, cout_pitch_w] packed, taps ordered (kd, kh, kw) */
  const float* scale;       /* [cout] or NULL (=1) */
  const float* shift;       /* [cout] or NULL (=0) */
  const void* residual;     /* same addressing as out, or NULL */
  void* out;                /* [N, TD, TH, TW, cout_pitch] */
  int N, D, H, W, cin, cin_pitch;
  int OD, OH, OW;           /* virtual output grid of this call */
  int TD, TH, TW;           /* full output tensor extent */
  int cout, cout_pitch;     /* channels computed per position, and the distance between positions;
                               padding channels [cout, cout_pitch) are written as zeros */
  int cout_pitch_w;         /* row length of the packed weight (>= cout, multiple of 4) */
  int ksize[3];             /* taps per axis */
  int stride[3];
  int tap_off0[3], tap_step[3];
  int ostride[3], ooffset[3];
  int relu;                 /* 0: none; 1: ReLU after the residual add; 2: ReLU before the residual add */
  int algo;                 /* sp3d_conv_algo */
  int in_dtype, out_dtype;  /* sp3d_dtype (SIMT path: F32 only).  SP3D_CONV_TC_BF16X3 also takes out_dtype = SP3D_BF16X2:
                               out / residual are [2][N, TD, TH, TW, cout_pitch] bf16 term planes (cout_pitch % 8 == 0) */
  int fused_phases;         /* tensor-core path only.  1: kernel-2 stride-2 transposed 3-D convolution in ONE launch:
                               ksize = (1,1,1), ostride = (2,2,2), the packed weight has 8 * cout rows ordered
                               (px, py, pz, co), and out[2x+px, 2y+py, 2z+pz, co] is written for all 8 phases
                               (cout_pitch == cout).  0: one launch per phase through ostride / ooffset. */
  int zfold;                /* tensor-core path only.  F > 1: F consecutive output positions along W share one GEMM row
                               (3-D "same" convolutions with cin_pitch == cin, cout_pitch == channel tile, W % F == 0):
                               the packed weight holds, per (kd, kh), k + F - 1 windows e of [F * cout_pitch][cin] rows
                               ordered (ro, co) with tap kw = e - ro (zero rows where that is outside the kernel);
                               cout_pitch_w = F * cout_pitch.  0 / 1: off */
  const sp3d_softargmax_args* head_softargmax;
                            /* tensor-core path only, 1x1x1 convolutions with cout <= 15 (the V2VNet head,
                               lib/models/v2v_net.py:124 + lib/models/pose_regression_net.py:19-28 in one pass):
                               when non-NULL the convolution output is NOT stored (out may be NULL); the soft-argmax
                               of scale * acc + shift over each of the N volumes is accumulated on chip and written
                               to head->out [N, cout, 3].  Uses head->centers / lin_* / beta / out and
                               head->workspace of sp3d_conv_head_workspace() bytes; x / strides are ignored,
                               n_cubes, C, X, Y, Z must equal N, cout, OD, OH, OW; check_flag must be 0. */
  int split_terms;          /* SP3D_CONV_TC_BF16X3 only: 3, 6 or 2.  The float32 operands are used as sums of bf16 terms
                               (x = x0 + x1 (+ x2), each term the bf16 rounding of what the previous ones left) and the
                               product is accumulated in float32 over the term pairs
                                 3: x1 w0 + x0 w1 + x0 w0                          (drops ~3 * 2^-18 per product)
                                 6: x2 w0 + x1 w1 + x0 w2 + x1 w0 + x0 w1 + x0 w0   (drops ~2^-24 per product)
                                 2: the same 3 pairs in TWO K blocks -- x0 [w0 | w1] as MMAs of 2 N columns (the A
                                    operand is read once for two products: narrow channel tiles are bound by that
                                    read), x1 w0 onto the upper N columns; the epilogue adds the halves.  `weight` is
                                    packed [n_tile][chunk][tap][2 N rows: w0 rows, w1 rows][chunk channels] followed by
                                    [chunk][tap][N rows of w0][chunk channels].  Available for the layer shapes listed
                                    in csrc/conv_tc.cu (SP3D_TC_CASE_W); outputs must be 16-byte pitched.
                               as extra K blocks of ONE implicit GEMM, small products first (the tensor core's float32
                               accumulator loses ~2^-24 of its magnitude per MMA, which bounds both variants: measured
                               ~1e-5 of the output range per layer, tests/test_gpu_split.py).  `in` is the bf16 tensor
                               sp3d_split_bf16 writes: S = 2 (3 pairs) or 3 (6 pairs) term planes [S][N, D, H, W, cin],
                               cin_pitch = cin; `weight` is packed [n_tile][term pair][chunk][tap][N][chunk channels]
                               with the matching weight term per pair.  The fused soft-argmax head is not available. */
} sp3d_conv_args;
int sp3d_conv_fwd(const sp3d_conv_args* a, void* stream);
int64_t sp3d_conv_head_workspace(const sp3d_conv_args* a);
/* Debug aid (profiles/conv_stalls.py): when given a device buffer of 148 * 16 uint64, every tensor-core
 * convolution launch writes per-CTA pipeline wait cycles into it (process-global; NULL switches it off). */
void sp3d_debug_conv_profile(void* dev_u64_buffer);
/* Debug aid (A/B measurements): 0 (the default) keeps every tensor-core convolution on single CTAs; 1 lets the large
 * 3^3 / 7^3 launches run as CTA pairs (clusters of two) that share one multicast weight stream.  Same results bit for bit,
 * and on B200 the same speed. */
void sp3d_debug_conv_pair(int enable);

/* Max pooling on channel-last activations (window k, stride s, padding p per axis; -inf padding).
 * Replaces F.max_pool3d(k2,s2) (lib/models/v2v_net.py:54) and nn.MaxPool2d(3,2,1)
 * (lib/models/pose_resnet.py:105). */
typedef struct {
  const void* in; void* out;
  int N, D, H, W, C, c_pitch;
  int OD, OH, OW;
  int k[3], s[3], p[3];
  int dtype;                /* SP3D_F32, SP3D_BF16, or SP3D_BF16X2 (in = [2][N,D,H,W,c_pitch], out = [2][N,OD,OH,OW,c_pitch]:
                               the maximum of plane 0 + plane 1, re-split) */
} sp3d_maxpool_args;
int sp3d_maxpool_fwd(const sp3d_maxpool_args* a, void* stream);

/* Layout change between the reference's channel-first tensors [N, C, S] and channel-last
 * [N, S, c_pitch] (S = product of spatial extents); padding channels are written as zeros. */
typedef struct {
  const void* src; void* dst;
  int64_t N, C, S;
  int64_t c_pitch;
  int to_channel_last;      /* 1: [N,C,S] -> [N,S,c_pitch]; 0: reverse */
  int src_dtype, dst_dtype;
} sp3d_layout_args;
int sp3d_layout_convert(const sp3d_layout_args* a, void* stream);

/* 2 x 2 space-to-depth of an image batch ([N, C, H, W] through the strides, float32 or bf16, H and W even) into
 * channel-last bf16 [N, H/2, W/2, dst_pitch] with channel (py * 2 + px) * C + c; channels >= 4 C are zeros.
 * Turns the stride-2 convolutions of PoseResNet (lib/models/pose_resnet.py:58-93,102-105) into stride-1 ones. */
typedef struct {
  const void* src; void* dst;
  int src_dtype;            /* SP3D_F32 or SP3D_BF16 */
  int64_t stride_n, stride_c, stride_y, stride_x; /* source strides in elements */
  int N, C, H, W;
  int dst_pitch;            /* >= 4 C, multiple of 8 */
  int dst_dtype;            /* SP3D_BF16 (0 is read as SP3D_BF16), or SP3D_BF16X2 with a float32 source: dst is
                               [2][N, H/2, W/2, dst_pitch] term planes */
} sp3d_s2d_args;
int sp3d_space_to_depth(const sp3d_s2d_args* a, void* stream);

/* float32 channel-last activations [P, src_pitch] (C channels used) -> bf16 [S][P, c_block]: plane s holds term s
 * of the bf16 expansion x = x0 + x1 (+ x2) (x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1)); channels
 * [C, c_block) of every plane are zeros.  Input layout of SP3D_CONV_TC_BF16X3. */
typedef struct {
  const float* src; void* dst;
  int64_t P;                /* positions */
  int C, src_pitch;
  int c_block;              /* channels per position in dst (multiple of 8, >= C) */
  int S;                    /* 2 or 3 planes */
} sp3d_split_args;
int sp3d_split_bf16(const sp3d_split_args* a, void* stream);
/* The reverse for S = 2: dst[p, c] = plane0[p, c] + plane1[p, c] (float32, src_pitch = pitch of dst, c_block = pitch of
 * the planes; `src` is the float32 destination, `dst` the bf16 planes -- the struct is shared with sp3d_split_bf16). */
int sp3d_merge_bf16(const sp3d_split_args* a, void* stream);

/* Stacks the x-neighbourhood of a 1-channel bf16 volume into channels:
 * dst[n, x, y, z, j] = src[n, x + j - pad, y, z, 0] (0 outside), j < taps; dst is [N, X, Y, Z, 16] bf16.
 * Lets the 1 -> 16 channel 7^3 stem of the root V2VNet (lib/models/v2v_net.py:117, input_channels = 1) run as a
 * 1 x 7 x 7 tensor-core convolution over 7 "tap channels" instead of 7^3 taps over 15 padding zeros. */
typedef struct {
  const void* src; void* dst;
  int N, X, Y, Z;
  int src_pitch;            /* channels per voxel of src (elements); channel 0 is read */
  int taps, pad;
} sp3d_stack_args;
int sp3d_stack_x_shifts(const sp3d_stack_args* a, void* stream);

/* ==========================================================================================
 * Backward operators (unproject_bwd, softargmax3d_bwd, maxpool bwd, conv wgrad, bn stats /
 * apply / bwd).  float32, channel-last unless strides say otherwise.  They replace what torch.autograd runs for the
 * reference's training forward/backward through the same path (lib/models/multi_person_posenet_ssv.py:222-501 calls
 * the modules below under autograd).  The input gradient of a convolution (dgrad) is sp3d_conv_fwd itself on the
 * flipped / transposed weight (host-side packing, selfpose3d_b200/ops.py conv_dgrad).
 * ========================================================================================== */

/* Gradient of sp3d_unproject_fwd (math_mode 0, partial 0) with respect to the heat-maps: the adjoint of
 * F.grid_sample(bilinear, zeros, align_corners=True) + masked mean + NaN->0 + clamp(0,1) of
 * lib/models/project_layer.py:93-99 (voxel coordinates carry no gradient: grid centres are detached,
 * lib/models/cuboid_proposal_net_soft.py:57).  The clamp gate (0 <= r <= 1, not NaN) is re-evaluated from the
 * forward inputs. */
typedef struct {
  sp3d_unproject_args fwd;  /* the forward call's arguments; fwd.cubes / fwd.out_dtype / fwd.grids are not used */
  const float* grad_cubes;  /* dL/dcubes, float32, addressed through fwd.out_stride_{cube,c,vox} */
  float* grad_heatmaps[SP3D_MAX_VIEWS]; /* dL/dheatmaps per view, addressed through fwd.hm_stride_*;
                                           ACCUMULATED into (atomic adds): the caller zeroes them */
} sp3d_unproject_bwd_args;
int sp3d_unproject_bwd(const sp3d_unproject_bwd_args* a, void* stream);

/* Gradient of sp3d_softargmax3d_fwd with respect to x: dx_v = beta * p_v * sum_d (g_vd - out_d) * grad_out_d,
 * p = softmax(beta * x) (autograd of lib/models/pose_regression_net.py:22-27). */
typedef struct {
  sp3d_softargmax_args fwd; /* forward arguments (x float32; fwd.out = the forward RESULT [n_cubes, C, 3]).
                               fwd.workspace (optional, >= n_cubes * 64 * C * 8 bytes): channel-last volumes (stride_c = 1,
                               stride_vox a multiple of 4, 16-byte aligned) then take the streaming two-launch form */
  const float* grad_out;    /* [n_cubes, C, 3] */
  float* grad_x;            /* float32, addressed like fwd.x; written (cubes skipped by check_flag get zeros) */
} sp3d_softargmax_bwd_args;
int sp3d_softargmax3d_bwd(const sp3d_softargmax_bwd_args* a, void* stream);

/* Gradient of sp3d_maxpool_fwd (float32): each window's gradient goes to its first maximum in (d, h, w) scan
 * order, as ATen's max_pool backward does (F.max_pool3d of lib/models/v2v_net.py:54, nn.MaxPool2d of
 * lib/models/pose_resnet.py:105). */
typedef struct {
  sp3d_maxpool_args fwd;    /* forward arguments, dtype SP3D_F32; fwd.in = the forward input, fwd.out unused */
  const float* grad_out;    /* [N, OD, OH, OW, c_pitch] */
  float* grad_in;           /* [N, D, H, W, c_pitch]; overwritten (zero-filled, then scattered into) */
} sp3d_maxpool_bwd_args;
int sp3d_maxpool_bwd(const sp3d_maxpool_bwd_args* a, void* stream);

/* Weight (and bias) gradient of one sp3d_conv_fwd launch (float32 SIMT geometry, no scale/shift/activation:
 * grad_out is the gradient of the raw convolution result):
 *   grad_weight[t, ci, co] += sum_{n, o} in[n, o*stride + tap_off0 + t*tap_step, ci] * grad_out[n, o*ostride+ooffset, co]
 *   grad_bias[co]          += sum_{n, o} grad_out[n, o*ostride+ooffset, co]
 * Replaces cudnn's convolution backward-filter for nn.Conv{2,3}d / nn.ConvTranspose{2,3}d (one call per output
 * phase for the transposed ones, exactly as the forward). */
typedef struct {
  sp3d_conv_args fwd;       /* geometry of the forward launch; fwd.in = the forward input (float32);
                               weight / scale / shift / residual / out / relu / algo are not used */
  const float* grad_out;    /* [N, TD, TH, TW, cout_pitch] */
  float* grad_weight;       /* [ntaps, cin, cout_pitch_w], packed like the SIMT forward weight; ACCUMULATED into */
  float* grad_bias;         /* optional [cout]; ACCUMULATED into */
} sp3d_conv_wgrad_args;
int sp3d_conv_wgrad(const sp3d_conv_wgrad_args* a, void* stream);

/* Training-mode BatchNorm on channel-last activations x [P, pitch] (P = all positions of the batch; pitch a multiple of
 * 4, 16-byte aligned tensors):
 *   sp3d_bn_stats: mean[c], var[c] (biased, the normalisation's variance) over P -- F.batch_norm(training=True) of
 *                  nn.BatchNorm{2,3}d (lib/models/v2v_net.py:14,27,30, lib/models/pose_resnet.py:49-...);
 *   sp3d_bn_apply: y = act(x * scale[c] + shift[c] (+ residual)), relu as in sp3d_conv_args (0 / 1 / 2);
 *   sp3d_bn_bwd:   with xhat = (x - mean) * rsqrt(var + eps) and dz = grad_y masked by (y > 0) when `y` is given
 *                  (ReLU directly after the normalisation):
 *                    dgamma = sum dz * xhat,  dbeta = sum dz,
 *                    dx = gamma * rsqrt(var + eps) * (dz - dbeta / P - xhat * dgamma / P).
 * GROUPED statistics (item_group != NULL): x is n_items equal items of P / n_items positions (the cubes of a launch);
 * item i belongs to statistic group item_group[i] in [0, n_groups), group g holds group_items[g] items.  Every group is
 * normalised with ITS OWN batch statistics -- what the reference gets by calling the pose net once per proposal slot
 * (lib/models/multi_person_posenet.py:88-99) -- in one launch set: mean / var / scale / shift are [n_groups][C],
 * workspaces n_groups * 2 * C doubles, grad_gamma / grad_beta [C] summed over the groups (shared parameters).
 * item_group == NULL (n_groups <= 1): one group of P positions. */
typedef struct {
  const float* x; int64_t P; int C, pitch;
  float* mean; float* var;  /* [n_groups][C] outputs */
  double* workspace;        /* n_groups * 2 * C doubles */
  int64_t workspace_bytes;
  int n_items, n_groups;
  const int32_t* item_group;    /* [n_items] or NULL */
  const int32_t* group_items;   /* [n_groups] items per group (with item_group) */
  float* running_mean; float* running_var;   /* optional [C] (both or neither): the module's running statistics, updated
                                                in place as n_groups successive training-mode calls do it -- per group
                                                r <- (1 - momentum) r + momentum stat, the variance unbiased n / (n - 1) */
  float momentum;
} sp3d_bn_stats_args;
int sp3d_bn_stats(const sp3d_bn_stats_args* a, void* stream);

typedef struct {
  const float* x; const float* residual; float* y;   /* [P, pitch]; residual optional */
  int64_t P; int C, pitch;
  const float* scale; const float* shift;            /* [n_groups][C] */
  int relu;
  int n_items, n_groups;
  const int32_t* item_group;    /* [n_items] or NULL */
} sp3d_bn_apply_args;
int sp3d_bn_apply(const sp3d_bn_apply_args* a, void* stream);

typedef struct {
  const float* x; const float* grad_y; const float* y;   /* [P, pitch]; y optional (ReLU mask) */
  int64_t P; int C, pitch;
  const float* mean; const float* var; const float* gamma;   /* mean, var [n_groups][C]; gamma [C], NULL = 1 */
  float eps;
  float* grad_x;            /* [P, pitch], written; padding channels get zeros */
  float* grad_gamma; float* grad_beta;   /* [C], written */
  double* workspace;        /* n_groups * 2 * C doubles */
  int64_t workspace_bytes;
  int n_items, n_groups;
  const int32_t* item_group;    /* [n_items] or NULL */
  const int32_t* group_items;   /* [n_groups] */
} sp3d_bn_bwd_args;
int sp3d_bn_bwd(const sp3d_bn_bwd_args* a, void* stream);

/* Backward of a ReLU whose output y is at hand: grad_x[i] = y[i] > 0 ? grad_y[i] : 0 over n contiguous floats
 * (the residual branches relu(a + b) of lib/models/v2v_net.py:44 and lib/models/pose_resnet.py:90). */
typedef struct {
  const float* grad_y; const float* y; float* grad_x;
  int64_t n;
} sp3d_relu_bwd_args;
int sp3d_relu_bwd(const sp3d_relu_bwd_args* a, void* stream);

/* Gaussian joint rendering of the SSL losses (lib/models/multi_person_posenet_ssv.py:410-448): per (view, sample) and
 * joint the heat-map  clip( sum_{p < n_people[b]} exp(-((x - kx/4)/3)^2/2 - ((y - ky/4)/3)^2/2), 0, 1 )  of the
 * re-projected people, and its gradient with respect to the joint pixels (the clip passes gradient where the sum is
 * <= 1).  kps [V, B, P, J, 2] network-input pixels; heatmaps [V, B, J, h, w]. */
typedef struct {
  const float* kps;
  const int32_t* n_people;  /* [B] people per sample (<= P); rows beyond it are ignored */
  int V, B, P, J, h, w;
  float inv_scale;          /* heat-map pixels per network-input pixel (reference: 1/4) */
  float sigma;              /* in heat-map pixels (reference: 3) */
  float* heatmaps;
} sp3d_gauss_render_args;
int sp3d_gauss_render_fwd(const sp3d_gauss_render_args* a, void* stream);

typedef struct {
  sp3d_gauss_render_args fwd;   /* forward arguments; fwd.heatmaps is not used */
  const float* grad_heatmaps;   /* [V, B, J, h, w] */
  float* grad_kps;              /* [V, B, P, J, 2]; written (rows beyond n_people get zeros) */
} sp3d_gauss_render_bwd_args;
int sp3d_gauss_render_bwd(const sp3d_gauss_render_bwd_args* a, void* stream);

/* Weight (and bias) gradient on the tensor cores, for convolutions whose taps step by one input position per output
 * position -- stride-1 convolutions (3-D, or 2-D with X = 1) and each output phase of a stride-s transposed convolution:
 *   grad_weight[tap, ci, co] += sum_p x[p + tap_off + tap, ci] * grad_out[p * g_stride + g_off, co]
 *   grad_bias[co]            += sum_p grad_out[p * g_stride + g_off, co]
 * p runs over the position grid [N, X, Y, Z] (= x's extent), taps are ordered (kx, ky, kz), reads outside x are zero (the
 * padding).  float32 channel-last operands are used as sums of two bf16 terms with three term pairs accumulated in
 * float32 (the arithmetic of SP3D_CONV_TC_BF16X3); the contraction over positions runs as a tcgen05 GEMM on channel-first
 * z-lines staged in `workspace` (csrc/conv_wgrad_tc.cu).  Needs Z <= 128, round_up(cin, 16) in {16, 32, 64} or cin > 64
 * (padded to multiples of 128), cout <= 128 or a multiple of 128, taps <= 7 per axis; anything else returns
 * SP3D_ERR_UNSUPPORTED (callers fall back to sp3d_conv_wgrad).  Replaces cudnn's wgrad for nn.Conv3d / nn.ConvTranspose3d
 * of lib/models/v2v_net.py:10-69,124 and the stride-1 nn.Conv2d / nn.ConvTranspose2d of lib/models/pose_resnet.py in the
 * training step (lib/core/function.py:27-217). */
typedef struct {
  const float* x;           /* [N, X, Y, Z, x_pitch] forward input */
  const float* grad_out;    /* [N, GX, GY, GZ, g_pitch] gradient of the raw convolution result */
  int N, X, Y, Z;
  int cin, x_pitch, cout, g_pitch;
  int ksize[3];             /* taps per axis */
  int tap_off[3];           /* input offset of tap 0 per axis (-padding for a "same" convolution) */
  int GX, GY, GZ;           /* extent of grad_out */
  int g_stride[3], g_off[3];/* gradient position of grid position p: p * g_stride + g_off (1 / 0 for a convolution) */
  float* grad_weight;       /* [taps, gw_cin, gw_pitch] float32, ADDED into (zero it first) */
  int gw_cin, gw_pitch;
  float* grad_bias;         /* [cout], added into; or NULL */
  void* workspace;          /* sp3d_conv_wgrad_tc_workspace(args) bytes, 16-byte aligned */
  int64_t workspace_bytes;
} sp3d_conv_wgrad_tc_args;
int64_t sp3d_conv_wgrad_tc_workspace(const sp3d_conv_wgrad_tc_args* a);
int sp3d_conv_wgrad_tc(const sp3d_conv_wgrad_tc_args* a, void* stream);

/* ==========================================================================================
 * Training targets on the device: what the reference's DataLoader workers render with
 * numpy per item, lib/dataset/JointsDataset.py:237-341.
 * ========================================================================================== */

/* generate_target_heatmap (:237-302): per (item, joint) the pixel-wise MAXIMUM over the item's people of a sigma-wide
 * Gaussian centred on the TRUNCATED heat-map pixel  mu = int(joint / feat_stride)  and cut to the (2 * 3 sigma + 1)^2
 * window around it; joints with vis == 0 and people without any visible joint are skipped; target_weight[j] = 1 when
 * any person shows joint j.  `window` is the reference's own Gaussian window  g[dy][dx] = exp(-((dx - r)^2 + (dy - r)^2)
 * / (2 sigma^2)), r = 3 sigma, as float32 [2r+1][2r+1] (evaluated once on the host with numpy, so that the device
 * result is bit-identical to the reference's). */
typedef struct {
  const double* joints;     /* [n_items, P, J, jstride] float64 as the dataset holds them: x, y (network-input pixels) */
  const double* joints_vis; /* [n_items, P, J, vstride]: visibility in [..][0] */
  const int32_t* n_people;  /* [n_items] people per item (<= P) */
  int n_items, P, J, jstride, vstride;
  int h, w;                 /* heat-map extent */
  double stride_x, stride_y; /* feat_stride = image_size / heatmap_size (float64 division, as numpy's) */
  const float* window;      /* [2r+1][2r+1] */
  int radius;               /* r = int(3 sigma) */
  float* target;            /* [n_items, J, h, w] */
  float* target_weight;     /* [n_items, J] */
} sp3d_target_heatmaps_args;
int sp3d_target_heatmaps(const sp3d_target_heatmaps_args* a, void* stream);

/* generate_3d_target (:304-341): per item the voxel-wise maximum over the people of  exp(-|g - mu|^2 / (2 sigma^2))
 * (float64, as numpy evaluates it; stored as float32) on the voxels with |g_axis - mu_axis| <= 3 sigma on every axis,
 * where g are the np.linspace voxel centres (given as three float64 vectors) and mu the person's root. */
typedef struct {
  const double* roots;      /* [n_items, P, 3] world mm */
  const int32_t* n_people;  /* [n_items] */
  int n_items, P;
  const double* grid_x;     /* [X] np.linspace(-size/2, size/2, X) + centre */
  const double* grid_y;     /* [Y] */
  const double* grid_z;     /* [Z] */
  int X, Y, Z;
  double sigma;             /* reference: 200 mm */
  float* target;            /* [n_items, X, Y, Z] */
} sp3d_target_volume_args;
int sp3d_target_volume(const sp3d_target_volume_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SP3D_H */
