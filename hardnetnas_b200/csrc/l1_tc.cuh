// First HardNet stage on the tensor core: per-patch input normalisation (HardNet.input_norm,
// hardnet/HardNet.py:306-310) + conv 1->32 k3 p1 + eval BatchNorm + ReLU (features[0..2], :281-283).
//
// The conv is a GEMM with M = 1024 pixels, N = 32, K = 9 taps padded to 16. The A operand cannot come from TMA
// (a tap row is 3 contiguous values), so loader warps build the im2col tiles in shared memory directly in the
// UMMA canonical no-swizzle K-major layout (8x16-byte core matrices; element (row, k) at
// (row/8)*256 + (k/8)*128 + (row%8)*16 + (k%8)*2) and publish them with fence.proxy.async + an mbarrier.
// One CTA = one patch at a time, double buffered: loaders normalise / im2col patch i+1 while the epilogue warps
// drain the eight 128x32 accumulators of patch i (TMEM: 2 x 8 x 32 = 512 columns).
//
// Warps: 0-7 epilogue (TMEM lane quarter = warp % 4, tiles (warp / 4) * 4 ..+3), 8 = TMEM owner + UMMA issuer,
// 9-16 loaders.
#pragma once

#include "common.cuh"
#include "tc_conv.cuh"

namespace hn {

constexpr int kL1TcThreads = 17 * 32;
constexpr int kL1PPitch = 40;                               // halfs per row of the haloed patch (34 used)
constexpr uint32_t kL1ABytes = 8 * 4096;                    // eight 128x16 im2col tiles
constexpr uint32_t kL1PBytes = 34 * kL1PPitch * 2;          // 2720
constexpr uint32_t kL1BufBytes = kL1ABytes;
constexpr uint32_t kL1StageBytes = 8 * 2 * 2048;            // per epilogue warp: two 32-row x 64 B output staging buffers
// > half of the SM's shared memory on purpose: one CTA per SM, because each CTA allocates all 512 TMEM columns
constexpr size_t kL1TcSmem = 120 * 1024;
static_assert(2 * kL1BufBytes + 1024 /*W*/ + kL1StageBytes + 1024 /*align*/ + 256 /*barriers*/ + 256 /*bias*/ <= kL1TcSmem, "smem budget");

__device__ __forceinline__ uint16_t to16bits(float v, int bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// Per-patch (mean, 1 / (unbiased std + 1e-7)) of HardNet.input_norm (hardnet/HardNet.py:306-310); one warp per
// patch. Pure pairwise trees, so a constant patch has mean == value exactly and normalises to exactly 0, like the
// reference's 0 / 1e-7.
template <typename TIn>
__global__ void __launch_bounds__(256) patch_stats_kernel(const TIn* __restrict__ in, float2* __restrict__ stats,
                                                          int num_patches, float norm_eps) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int patch = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; patch < num_patches; patch += warps) {
    float x[32];
    const TIn* src = in + static_cast<size_t>(patch) * 1024;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (sizeof(TIn) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
        x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
      } else {
        const uchar4 q = *reinterpret_cast<const uchar4*>(src + (i * 32 + lane) * 4);
        x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
      }
    }
    float t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = x[i];
#pragma unroll
    for (int w = 16; w > 0; w >>= 1)
#pragma unroll
      for (int i = 0; i < w; ++i) t[i] = t[2 * i] + t[2 * i + 1];
    float s = t[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 1024.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = x[i] - mean; t[i] = d * d; }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1)
#pragma unroll
      for (int i = 0; i < w; ++i) t[i] = t[2 * i] + t[2 * i + 1];
    float v = t[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) stats[patch] = make_float2(mean, 1.f / (sqrtf(v * (1.f / 1023.f)) + norm_eps));  // torch.std is unbiased
  }
}

template <typename TIn>
__global__ void __launch_bounds__(kL1TcThreads, 1)
l1_tc_kernel(const TIn* __restrict__ in, uint16_t* __restrict__ out, const float* __restrict__ w /*[9][32] folded*/,
             const float* __restrict__ bias /*[32]*/, const float2* __restrict__ stats /*null: no normalisation*/,
             int num_patches, int act_bf16) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);
  // layout: [buf0: A (32 KB) | P (3 KB)] [buf1 ...] [W 1 KB] [barriers] [bias 32 f] [red 8 f] [stat 2 f]
  const uint32_t w_addr = base + 2 * kL1BufBytes;
  const uint32_t stage_addr = w_addr + 1024;   // epilogue output staging
  const uint32_t bar_base = stage_addr + kL1StageBytes;
  auto a_full = [&](int b) { return bar_base + 8u * b; };
  auto a_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  auto t_full = [&](int b) { return bar_base + 8u * (4 + b); };
  auto t_empty = [&](int b) { return bar_base + 8u * (6 + b); };
  const uint32_t tmem_slot = bar_base + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  float* s_bias = reinterpret_cast<float*>(gbase + (bar_base + 256 - base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8) {
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(a_full(b), 8);    // one arrive per loader warp
        mbar_init(a_empty(b), 1);   // tcgen05.commit
        mbar_init(t_full(b), 1);    // tcgen05.commit
        mbar_init(t_empty(b), 8);   // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp < 8) {
    if (threadIdx.x < 32) s_bias[threadIdx.x] = bias[threadIdx.x];
  } else if (warp >= 9) {
    const int l = threadIdx.x - 9 * 32;  // 0..255
    // weights -> canonical no-swizzle [32 x 16] tile: (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2
    uint16_t* W = reinterpret_cast<uint16_t*>(gbase + (w_addr - base));
    for (int i = l; i < 32 * 16; i += 256) {
      const int n = i >> 4, k = i & 15;
      const float v = k < 9 ? w[k * 32 + n] : 0.f;
      W[((n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) >> 1] = to16bits(v, act_bf16);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp >= 9) {
    // ============================== loaders: normalise + im2col ==============================
    // Every thread owns 4 consecutive pixels and reads its own 3 x 6 window straight from global memory (the 4 KB
    // patch is L1 resident), so the loader warps never synchronise with each other.
    const int l = threadIdx.x - 9 * 32;
    const int py = l >> 3;             // pixel row handled by this thread
    const int px0 = (l & 7) * 4;       // first of 4 consecutive pixels
    int it = 0;
    for (int patch = blockIdx.x; patch < num_patches; patch += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      float mean = 0.f, inv = 1.f;
      if (stats != nullptr) {
        const float2 st = __ldg(stats + patch);
        mean = st.x;
        inv = st.y;
      }
      const TIn* src = in + static_cast<size_t>(patch) * 1024;
      uint16_t win[3][6];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int y = py + r - 1;
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool ok[6] = {false, false, false, false, false, false};
        if (y >= 0 && y < 32) {
          const TIn* row = src + y * 32 + px0;
          if constexpr (sizeof(TIn) == 4) {
            const float4 q = *reinterpret_cast<const float4*>(row);
            v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
          } else {
            const uchar4 q = *reinterpret_cast<const uchar4*>(row);
            v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
          }
          ok[1] = ok[2] = ok[3] = ok[4] = true;
          if (px0 > 0) { v[0] = static_cast<float>(row[-1]); ok[0] = true; }
          if (px0 < 28) { v[5] = static_cast<float>(row[4]); ok[5] = true; }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) win[r][c] = ok[c] ? to16bits((v[c] - mean) * inv, act_bf16) : static_cast<uint16_t>(0);
      }
      // the MMAs that read this buffer two patches ago must have retired before A is overwritten
      mbar_wait(a_empty(b), ph ^ 1u);
      uint8_t* A = gbase + b * kL1BufBytes;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pix = py * 32 + px0 + j;
        const int tile = pix >> 7, r = pix & 127;
        uint8_t* dst = A + tile * 4096 + (r >> 3) * 256 + (r & 7) * 16;
        uint4 k0, k1;
        k0.x = win[0][j] | (static_cast<uint32_t>(win[0][j + 1]) << 16);
        k0.y = win[0][j + 2] | (static_cast<uint32_t>(win[1][j]) << 16);
        k0.z = win[1][j + 1] | (static_cast<uint32_t>(win[1][j + 2]) << 16);
        k0.w = win[2][j] | (static_cast<uint32_t>(win[2][j + 1]) << 16);
        k1.x = win[2][j + 2];
        k1.y = 0u; k1.z = 0u; k1.w = 0u;
        *reinterpret_cast<uint4*>(dst) = k0;
        *reinterpret_cast<uint4*>(dst + 128) = k1;
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(b));
    }
  } else if (warp == 8) {
    // ============================== UMMA issuer ==============================
    const uint32_t idesc = make_idesc_f16(kTileM, 32, act_bf16);
    const uint64_t b_desc = make_noswizzle_desc(w_addr, 128, 256);
    int it = 0;
    for (int patch = blockIdx.x; patch < num_patches; patch += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(a_full(b), ph);
      mbar_wait(t_empty(b), ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint64_t a_desc = make_noswizzle_desc(base + b * kL1BufBytes + t * 4096, 128, 256);
          umma_f16(tmem_base + b * 256 + t * 32, a_desc, b_desc, idesc, 0u);
        }
        umma_commit(a_empty(b));
        umma_commit(t_full(b));
      }
      __syncwarp();
    }
  } else {
    // ============================== epilogue ==============================
    // Each warp owns 32 consecutive pixels of a tile = 2 KB of contiguous NHWC output: rows are staged in shared
    // memory and leave through one bulk (async-proxy) store per warp and tile, fully coalesced and off the LSU.
    const int q = warp & 3;
    const int half = warp >> 2;
    const uint32_t my_stage = stage_addr + warp * 4096;
    uint8_t* my_stage_ptr = gbase + (my_stage - base);
    int it = 0, sbuf = 0;
    for (int patch = blockIdx.x; patch < num_patches; patch += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(t_full(b), ph);
      tc_fence_after();
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const int t = half * 4 + tt;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 256 + t * 32, r);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v0 = fmaxf(__uint_as_float(r[2 * j]) + s_bias[2 * j], 0.f);
          const float v1 = fmaxf(__uint_as_float(r[2 * j + 1]) + s_bias[2 * j + 1], 0.f);
          o[j] = pack16(v0, v1, act_bf16);
        }
        // the bulk store that last read this staging buffer (two tiles ago) must be done reading it
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        uint4* srow = reinterpret_cast<uint4*>(my_stage_ptr + sbuf * 2048 + lane * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j) srow[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          bulk_store(out + (static_cast<size_t>(patch) * 1024 + t * 128 + q * 32) * 32, my_stage + sbuf * 2048, 2048);
          bulk_commit();
        }
        sbuf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty(b));
    }
    if (lane == 0) bulk_wait_all<0>();   // outstanding stores must complete before the CTA (and its smem) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace hn
