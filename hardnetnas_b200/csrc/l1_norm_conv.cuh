// First HardNet stage on CUDA cores: per-patch input normalisation fused with conv 1->32 (k3, p1),
// eval-mode BatchNorm and ReLU. Reference: HardNet.input_norm (hardnet/HardNet.py:306-310) and
// features[0..2] (:281-283). K = 9 is too thin for the tensor pipe to matter at this stage, the
// work is 295k FMA per patch.
//
// in : [B, 1, 32, 32] fp32 (or u8)      out: [B, 32, 32, 32] NHWC 16-bit (fp16 / bf16)
#pragma once

#include "common.cuh"
#include "tc_conv.cuh"

namespace hn {

constexpr int kL1Threads = 256;

template <typename TIn>
__global__ void __launch_bounds__(kL1Threads) l1_norm_conv_kernel(const TIn* __restrict__ in, uint16_t* __restrict__ out,
                                                                   const float* __restrict__ w /*[9][32], BN scale folded*/,
                                                                   const float* __restrict__ bias /*[32]*/, int num_patches,
                                                                   int act_bf16, int do_norm) {
  __shared__ float sp[34][36];  // normalised patch with a zero halo
  __shared__ __align__(16) float sw[9][32];
  __shared__ __align__(16) float sb[32];
  __shared__ float red[8];
  __shared__ float stat[2];

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  for (int i = t; i < 34 * 36; i += kL1Threads) (&sp[0][0])[i] = 0.f;
  for (int i = t; i < 9 * 32; i += kL1Threads) (&sw[0][0])[i] = w[i];
  if (t < 32) sb[t] = bias[t];

  const int py = t >> 3;         // pixel row handled by this thread
  const int px0 = (t & 7) * 4;   // first of 4 consecutive pixels

  for (int patch = blockIdx.x; patch < num_patches; patch += gridDim.x) {
    __syncthreads();  // previous iteration's readers of sp are done (also covers the init above)
    float x[4];
    {
      const TIn* src = in + static_cast<size_t>(patch) * 1024 + t * 4;
      if constexpr (sizeof(TIn) == 4) {
        const float4 v = *reinterpret_cast<const float4*>(src);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
      } else {
        const uchar4 v = *reinterpret_cast<const uchar4*>(src);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
      }
    }
    float mean = 0.f, inv = 1.f;
    if (do_norm) {
      // pure pairwise trees so that a constant patch gives mean == value exactly (-> all-zero input,
      // like the reference's 0 / 1e-7)
      float s = (x[0] + x[1]) + (x[2] + x[3]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      if (t == 0) stat[0] = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) * (1.f / 1024.f);
      __syncthreads();
      mean = stat[0];
      float d0 = x[0] - mean, d1 = x[1] - mean, d2 = x[2] - mean, d3 = x[3] - mean;
      float v = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp] = v;
      __syncthreads();
      if (t == 0) {
        const float var = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) * (1.f / 1023.f);
        stat[1] = 1.f / (sqrtf(var) + 1e-7f);  // torch.std is the unbiased estimator
      }
      __syncthreads();
      inv = stat[1];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sp[py + 1][px0 + 1 + j] = (x[j] - mean) * inv;
    __syncthreads();

    float win[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) win[r][c] = sp[py + r][px0 + c];

    uint16_t* dst = out + (static_cast<size_t>(patch) * 1024 + py * 32 + px0) * 32;
#pragma unroll
    for (int cg = 0; cg < 2; ++cg) {
      float acc[4][16];
#pragma unroll
      for (int pxl = 0; pxl < 4; ++pxl)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[pxl][c] = sb[cg * 16 + c];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        float wv[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 q = *reinterpret_cast<const float4*>(&sw[tap][cg * 16 + c4 * 4]);
          wv[c4 * 4 + 0] = q.x; wv[c4 * 4 + 1] = q.y; wv[c4 * 4 + 2] = q.z; wv[c4 * 4 + 3] = q.w;
        }
#pragma unroll
        for (int pxl = 0; pxl < 4; ++pxl) {
          const float a = win[ky][pxl + kx];
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[pxl][c] = fmaf(a, wv[c], acc[pxl][c]);
        }
      }
#pragma unroll
      for (int pxl = 0; pxl < 4; ++pxl) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = pack16(fmaxf(acc[pxl][2 * j], 0.f), fmaxf(acc[pxl][2 * j + 1], 0.f), act_bf16);
        uint4* d4 = reinterpret_cast<uint4*>(dst + pxl * 32 + cg * 16);
        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

}  // namespace hn
