// Patch extraction upstream of the descriptor: the reference's clip_patch
// (FDLNet-master/utils/image_utils.py:11-158) crops a PSIZE x PSIZE patch around every keypoint from the raw image
// with a per-keypoint similarity transform (scale / 2 / rescale-ratio, optional rotation) and bilinear interpolation.
// One thread per output pixel; a warp covers one patch row, so the four gathers of neighbouring pixels hit
// neighbouring addresses. Output is [N,1,PSIZE,PSIZE] fp32 = exactly what hn_forward takes.
#include <cuda_runtime.h>

#include <algorithm>

#include "host_common.h"

namespace hn {

// torch.linspace(-1, 1, steps)[i] as the CPU kernel computes it (symmetric halves), image_utils.py:30-35
__device__ __forceinline__ float linspace_pm1(int i, int steps) {
  const float step = 2.0f / static_cast<float>(steps - 1);
  return i < steps / 2 ? -1.0f + step * static_cast<float>(i) : 1.0f - step * static_cast<float>(steps - i - 1);
}

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
    // the reference divides by im_info[:, 0] after a view(B, -1): keypoint n belongs to image n / (N / B) there
    const float ratio = im_info[(n / kp_per_image) * 2];
    const float s = __fdiv_rn(__fdiv_rn(kpts_scale[n], ratio), 2.0f);          // image_utils.py:55-56
    float t00 = s, t01 = 0.f, t10 = 0.f, t11 = s;                               // thetas = diag(s, s, 1)
    if (kpts_ori != nullptr) {                                                  // thetas @ R, :67-73
      const float c = kpts_ori[n * 2], sn = kpts_ori[n * 2 + 1];
      t00 = __fmul_rn(s, c);
      t01 = __fmul_rn(s, -sn);
      t10 = __fmul_rn(s, sn);
      t11 = __fmul_rn(s, c);
    }
    const float xt = linspace_pm1(px, psize), yt = linspace_pm1(py, psize);
    // T_g = thetas @ grid (:77-79), then centre on the keypoint (:82-93)
    float x = __fadd_rn(__fmul_rn(t00, xt), __fmul_rn(t01, yt));
    float y = __fadd_rn(__fmul_rn(t10, xt), __fmul_rn(t11, yt));
    x = __fadd_rn(x, __fdiv_rn(static_cast<float>(kpts_byxc[n * 4 + 2]), ratio));
    y = __fadd_rn(y, __fdiv_rn(static_cast<float>(kpts_byxc[n * 4 + 1]), ratio));
    // bilinear taps with the reference's clamp-then-weight order (:98-150)
    const long long x0u = static_cast<long long>(floorf(x)), y0u = static_cast<long long>(floorf(y));
    const long long max_x = W - 1, max_y = H - 1;
    const long long x0 = min(max(x0u, 0LL), max_x), x1 = min(max(x0u + 1, 0LL), max_x);
    const long long y0 = min(max(y0u, 0LL), max_y), y1 = min(max(y0u + 1, 0LL), max_y);
    // The reference gathers from the flattened image stack and raises on a batch index outside [0, B); a kernel cannot
    // raise, so such a keypoint yields a NaN patch instead of an out-of-bounds read.
    const long long b = kpts_byxc[n * 4];
    if (b < 0 || b >= B) {
      out[i] = __int_as_float(0x7fc00000);
      continue;
    }
    const float* img = images + b * (static_cast<long long>(H) * W);
    const float Ia = __ldg(img + y0 * W + x0), Ib = __ldg(img + y1 * W + x0);
    const float Ic = __ldg(img + y0 * W + x1), Id = __ldg(img + y1 * W + x1);
    const float x0f = static_cast<float>(x0), x1f = static_cast<float>(x1);
    const float y0f = static_cast<float>(y0), y1f = static_cast<float>(y1);
    const float wa = __fmul_rn(x1f - x, y1f - y), wb = __fmul_rn(x1f - x, y - y0f);
    const float wc = __fmul_rn(x - x0f, y1f - y), wd = __fmul_rn(x - x0f, y - y0f);
    out[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, Ia), __fmul_rn(wb, Ib)), __fmul_rn(wc, Ic)), __fmul_rn(wd, Id));
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
  const long long total = N * psize * psize;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(sm) * 16));
  clip_patch_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(images, B, H, W, kpts_byxc, kpts_scale, kpts_ori, im_info, N,
                                                                        N / B, psize, out);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}
