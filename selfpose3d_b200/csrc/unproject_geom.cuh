// Per-(voxel, view) sampling geometry of the un-projection, shared by the forward (unproject.cu) and backward
// (backward.cu) kernels: float32 arithmetic in the reference's operation order with explicit round-to-nearest
// intrinsics (lib/models/project_layer.py:64-93, lib/utils/cameras.py:27-55, lib/utils/transforms.py:119-123).
#pragma once
#include "sp3d_common.cuh"

namespace sp3d {

struct ViewSample {
  float wx0, wx1, wy0, wy1;  // bilinear weights
  int x0, y0;                // top-left tap
  float m;                   // in-image mask (0/1)
};

// World point -> heat-map sampling position of one view.  `cam` points at SP3D_CAM_FLOATS floats.
__device__ __forceinline__ ViewSample project_view(const float* __restrict__ cam, float gx, float gy, float gz,
                                                   float img_w, float img_h, float cfg_w, float cfg_h,
                                                   float hm_w, float hm_h) {
  const float dx = __fsub_rn(gx, cam[9]);
  const float dy = __fsub_rn(gy, cam[10]);
  const float dz = __fsub_rn(gz, cam[11]);
  const float xc = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam[0]), __fmul_rn(dy, cam[1])), __fmul_rn(dz, cam[2]));
  const float yc = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam[3]), __fmul_rn(dy, cam[4])), __fmul_rn(dz, cam[5]));
  float zc = __fadd_rn(__fadd_rn(__fmul_rn(dx, cam[6]), __fmul_rn(dy, cam[7])), __fmul_rn(dz, cam[8]));
  zc = __fadd_rn(zc, 1e-5f);
  const float y0 = __fdiv_rn(xc, zc);
  const float y1 = __fdiv_rn(yc, zc);
  const float r2 = fminf(__fadd_rn(__fmul_rn(y0, y0), __fmul_rn(y1, y1)), 1e10f);
  const float r4 = __fmul_rn(r2, r2);
  const float r6 = __fmul_rn(r4, r2);
  const float radial = __fadd_rn(
      1.0f, __fadd_rn(__fadd_rn(__fmul_rn(cam[16], r2), __fmul_rn(cam[17], r4)), __fmul_rn(cam[18], r6)));
  const float tan = __fadd_rn(__fmul_rn(cam[19], y1), __fmul_rn(cam[20], y0));
  const float corr = __fadd_rn(radial, __fmul_rn(2.0f, tan));
  const float u = __fadd_rn(__fmul_rn(y0, corr), __fmul_rn(cam[20], r2));
  const float v = __fadd_rn(__fmul_rn(y1, corr), __fmul_rn(cam[19], r2));
  float px = __fadd_rn(__fmul_rn(cam[12], u), cam[14]);
  float py = __fadd_rn(__fmul_rn(cam[13], v), cam[15]);

  const float width = cam[27], height = cam[28];
  ViewSample s;
  s.m = (px >= 0.0f && py >= 0.0f && px < width && py < height) ? 1.0f : 0.0f;  // mask on un-clamped pixels
  const float hi = fmaxf(width, height);
  // torch.clamp semantics: NaN propagates; fminf/fmaxf would drop it, so keep NaN explicitly
  px = (px != px) ? px : fminf(fmaxf(px, -1.0f), hi);
  py = (py != py) ? py : fminf(fmaxf(py, -1.0f), hi);
  float qx = __fadd_rn(__fadd_rn(__fmul_rn(cam[21], px), __fmul_rn(cam[22], py)), cam[23]);
  const float qy = __fadd_rn(__fadd_rn(__fmul_rn(cam[24], px), __fmul_rn(cam[25], py)), cam[26]);
  if (cam[29] != 0.0f) qx = __fsub_rn(img_w, qx);
  const float uu = __fdiv_rn(__fmul_rn(qx, cfg_w), img_w);
  const float vv = __fdiv_rn(__fmul_rn(qy, cfg_h), img_h);
  float sx = __fsub_rn(__fmul_rn(__fdiv_rn(uu, __fsub_rn(cfg_w, 1.0f)), 2.0f), 1.0f);
  float sy = __fsub_rn(__fmul_rn(__fdiv_rn(vv, __fsub_rn(cfg_h, 1.0f)), 2.0f), 1.0f);
  sx = (sx != sx) ? sx : fminf(fmaxf(sx, -1.1f), 1.1f);
  sy = (sy != sy) ? sy : fminf(fmaxf(sy, -1.1f), 1.1f);
  // grid_sample(align_corners=True) un-normalisation
  const float fx = __fmul_rn(__fdiv_rn(__fadd_rn(sx, 1.0f), 2.0f), __fsub_rn(hm_w, 1.0f));
  const float fy = __fmul_rn(__fdiv_rn(__fadd_rn(sy, 1.0f), 2.0f), __fsub_rn(hm_h, 1.0f));
  const float x0f = floorf(fx), y0f = floorf(fy);
  s.wx1 = __fsub_rn(fx, x0f);
  s.wy1 = __fsub_rn(fy, y0f);
  s.wx0 = __fsub_rn(1.0f, s.wx1);
  s.wy0 = __fsub_rn(1.0f, s.wy1);
  // NaN coordinates (never produced by finite cameras) sample nothing
  s.x0 = (fx == fx) ? (int)x0f : -4;
  s.y0 = (fy == fy) ? (int)y0f : -4;
  return s;
}

}  // namespace sp3d
