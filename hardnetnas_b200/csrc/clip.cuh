// The sampling arithmetic of the reference's clip_patch (FDLNet-master/utils/image_utils.py:11-158), shared by the stand-alone
// kernels (clip.cu) and by the CLIP variant of the fused front kernel (front_fused.cuh: the crop is taken straight from the image
// stack inside the first conv kernel, so the 4 KiB/patch fp32 patch tensor never exists in HBM and the images may be uint8).
// Every step is an explicitly rounded fp32 operation in the reference's order, so all users produce the same bits.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace hn {

// Where the patches come from when they are cropped on the fly: the keypoint arrays of clip_patch plus the image stack.
struct ClipSrc {
  const void* images;              // [B, 1, H, W] fp32, or uint8 when img_u8
  const long long* kpts_byxc;      // [N, 4] (b, y, x, c)
  const float* kpts_scale;         // [N]
  const float* kpts_ori;           // [N, 2] (cos, sin) or nullptr
  const float* im_info;            // [B, 2]; column 0 = rescale ratio
  long long B;
  long long kp_per_image;          // N / B: the reference's view(B, -1)
  long long n0;                    // first keypoint of this launch
  int H, W;
  int img_u8;
};

// the similarity transform of one keypoint and its centre in image pixels
struct ClipKp {
  float t00, t01, t10, t11, cx, cy;
  long long b;
};

// torch.linspace(-1, 1, steps)[i] as the CPU kernel computes it (symmetric halves), image_utils.py:30-35
__device__ __forceinline__ float linspace_pm1(int i, int steps) {
  const float step = 2.0f / static_cast<float>(steps - 1);
  return i < steps / 2 ? -1.0f + step * static_cast<float>(i) : 1.0f - step * static_cast<float>(steps - i - 1);
}

__device__ __forceinline__ ClipKp clip_keypoint(const long long* __restrict__ kpts_byxc, const float* __restrict__ kpts_scale,
                                                const float* __restrict__ kpts_ori, const float* __restrict__ im_info,
                                                long long kp_per_image, long long n) {
  ClipKp k;
  // the reference divides by im_info[:, 0] after a view(B, -1): keypoint n belongs to image n / (N / B) there
  const float ratio = im_info[(n / kp_per_image) * 2];
  const float s = __fdiv_rn(__fdiv_rn(kpts_scale[n], ratio), 2.0f);          // image_utils.py:55-56
  k.t00 = s; k.t01 = 0.f; k.t10 = 0.f; k.t11 = s;                            // thetas = diag(s, s, 1)
  if (kpts_ori != nullptr) {                                                  // thetas @ R, :67-73
    const float c = kpts_ori[n * 2], sn = kpts_ori[n * 2 + 1];
    k.t00 = __fmul_rn(s, c);
    k.t01 = __fmul_rn(s, -sn);
    k.t10 = __fmul_rn(s, sn);
    k.t11 = __fmul_rn(s, c);
  }
  k.cx = __fdiv_rn(static_cast<float>(kpts_byxc[n * 4 + 2]), ratio);
  k.cy = __fdiv_rn(static_cast<float>(kpts_byxc[n * 4 + 1]), ratio);
  k.b = kpts_byxc[n * 4];
  return k;
}

// the four bilinear taps of output pixel (px, py): offsets into one H x W image and their weights
struct ClipTaps {
  int ia, ib, ic, id;
  float wa, wb, wc, wd;
};

__device__ __forceinline__ ClipTaps clip_taps(const ClipKp& k, int px, int py, int psize, int H, int W) {
  const float xt = linspace_pm1(px, psize), yt = linspace_pm1(py, psize);
  // T_g = thetas @ grid (:77-79), then centre on the keypoint (:82-93)
  float x = __fadd_rn(__fmul_rn(k.t00, xt), __fmul_rn(k.t01, yt));
  float y = __fadd_rn(__fmul_rn(k.t10, xt), __fmul_rn(k.t11, yt));
  x = __fadd_rn(x, k.cx);
  y = __fadd_rn(y, k.cy);
  // bilinear taps with the reference's clamp-then-weight order (:98-150). The reference clamps int64 pixel indices; clamping the
  // floor in the float domain gives the same taps for every finite coordinate (image sides are far below 2^24) without the
  // multi-instruction 64-bit conversions and compares.
  const float fx = floorf(x), fy = floorf(y);
  const float max_x = static_cast<float>(W - 1), max_y = static_cast<float>(H - 1);
  const float x0f = fminf(fmaxf(fx, 0.f), max_x), x1f = fminf(fmaxf(fx + 1.f, 0.f), max_x);
  const float y0f = fminf(fmaxf(fy, 0.f), max_y), y1f = fminf(fmaxf(fy + 1.f, 0.f), max_y);
  const int x0 = static_cast<int>(x0f), x1 = static_cast<int>(x1f);
  const int y0 = static_cast<int>(y0f), y1 = static_cast<int>(y1f);
  ClipTaps t;
  t.ia = y0 * W + x0; t.ib = y1 * W + x0; t.ic = y0 * W + x1; t.id = y1 * W + x1;
  t.wa = __fmul_rn(x1f - x, y1f - y); t.wb = __fmul_rn(x1f - x, y - y0f);
  t.wc = __fmul_rn(x - x0f, y1f - y); t.wd = __fmul_rn(x - x0f, y - y0f);
  return t;
}

__device__ __forceinline__ float clip_blend(const ClipTaps& t, float Ia, float Ib, float Ic, float Id) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.wa, Ia), __fmul_rn(t.wb, Ib)), __fmul_rn(t.wc, Ic)), __fmul_rn(t.wd, Id));
}

template <typename T>
__device__ __forceinline__ float clip_pixel(const T* __restrict__ img, int i) {
  return static_cast<float>(__ldg(img + i));
}

}  // namespace hn
