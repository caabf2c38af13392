// Patch extraction upstream of the descriptor: the reference's clip_patch
// (FDLNet-master/utils/image_utils.py:11-158) crops a PSIZE x PSIZE patch around every keypoint from the raw image
// with a per-keypoint similarity transform (scale / 2 / rescale-ratio, optional rotation) and bilinear interpolation.
// Generic patch size: one thread per output pixel; psize 32: clip_patch32_kernel below. Output is [N,1,PSIZE,PSIZE] fp32 = exactly what hn_forward takes.
#include <cuda_runtime.h>

#include <algorithm>

#include "clip.cuh"
#include "host_common.h"

namespace hn {

__global__ void __launch_bounds__(256) clip_patch_kernel(const float* __restrict__ images, long long B, int H, int W,
                                                         const long long* __restrict__ kpts_byxc,
                                                         const float* __restrict__ kpts_scale,
                                                         const float* __restrict__ kpts_ori, const float* __restrict__ im_info,
                                                         long long N, long long kp_per_image, int psize,
                                                         float* __restrict__ out) {
  const long long total = N * psize * psize;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % psize);
    const int py = static_cast<int>((i / psize) % psize);
    const long long n = i / (static_cast<long long>(psize) * psize);
    const long long o = n * psize * psize + py * psize + px;
    const ClipKp k = clip_keypoint(kpts_byxc, kpts_scale, kpts_ori, im_info, kp_per_image, n);
    // The reference gathers from the flattened image stack and raises on a batch index outside [0, B); a kernel cannot
    // raise, so such a keypoint yields a NaN patch instead of an out-of-bounds read.
    if (k.b < 0 || k.b >= B) {
      out[o] = __int_as_float(0x7fc00000);
      continue;
    }
    const ClipTaps t = clip_taps(k, px, py, psize, H, W);
    const float* img = images + k.b * (static_cast<long long>(H) * W);
    out[o] = clip_blend(t, clip_pixel(img, t.ia), clip_pixel(img, t.ib), clip_pixel(img, t.ic), clip_pixel(img, t.id));
  }
}

// psize 32 (what the descriptor takes): a block per patch, the keypoint's transform computed once per thread instead of once per
// pixel (its divisions and the 64-bit index arithmetic of the generic kernel made that one instruction-bound: 12 ns/patch), four
// pixels per thread, a warp-wide load covering a compact 4 x 8 pixel block.
__global__ void __launch_bounds__(256) clip_patch32_kernel(const float* __restrict__ images, long long B, int H, int W,
                                                           const long long* __restrict__ kpts_byxc,
                                                           const float* __restrict__ kpts_scale,
                                                           const float* __restrict__ kpts_ori, const float* __restrict__ im_info,
                                                           long long N, long long kp_per_image, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int py = 4 * w + (lane >> 3), pc = lane & 7;
  for (long long n = blockIdx.x; n < N; n += gridDim.x) {
    const ClipKp k = clip_keypoint(kpts_byxc, kpts_scale, kpts_ori, im_info, kp_per_image, n);
    float* o = out + n * 1024 + py * 32 + pc;
    if (k.b < 0 || k.b >= B) {   // see clip_patch_kernel
#pragma unroll
      for (int r = 0; r < 4; ++r) o[8 * r] = __int_as_float(0x7fc00000);
      continue;
    }
    const float* img = images + k.b * (static_cast<long long>(H) * W);
    ClipTaps t[4];
    float I[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) t[r] = clip_taps(k, 8 * r + pc, py, 32, H, W);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      I[r][0] = clip_pixel(img, t[r].ia); I[r][1] = clip_pixel(img, t[r].ib);
      I[r][2] = clip_pixel(img, t[r].ic); I[r][3] = clip_pixel(img, t[r].id);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) o[8 * r] = clip_blend(t[r], I[r][0], I[r][1], I[r][2], I[r][3]);
  }
}

}  // namespace hn

using namespace hn;

extern "C" int hn_clip_patches(const float* images, long long B, int H, int W, const long long* kpts_byxc,
                               const float* kpts_scale, const float* kpts_ori, const float* im_info, long long N, int psize,
                               float* out, void* stream) {
  HN_REQUIRE(images && kpts_byxc && kpts_scale && im_info && out, "hn_clip_patches: NULL argument");
  HN_REQUIRE(B >= 1 && H >= 1 && W >= 1 && psize >= 2, "hn_clip_patches: bad image / patch size");
  HN_REQUIRE(N >= 0 && N % B == 0, "hn_clip_patches: the reference's view(B, -1) needs N (%lld) to be a multiple of B (%lld)", N, B);
  if (N == 0) return HN_OK;
  int sm = 0;
  HN_TRY(device_sm_count(&sm));
  if (psize == 32) {
    const int grid32 = static_cast<int>(std::min<long long>(N, static_cast<long long>(sm) * 64));
    clip_patch32_kernel<<<grid32, 256, 0, static_cast<cudaStream_t>(stream)>>>(images, B, H, W, kpts_byxc, kpts_scale, kpts_ori, im_info,
                                                                              N, N / B, out);
    HN_CUDA(cudaGetLastError());
    count_launch();
    return HN_OK;
  }
  const long long total = N * psize * psize;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sm) * 16));
  clip_patch_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(images, B, H, W, kpts_byxc, kpts_scale, kpts_ori, im_info, N,
                                                                        N / B, psize, out);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}
