// Persistent, warp-specialised tcgen05 kernel for the HardNet conv stack.
//
// One kernel template covers
//   * the 3x3 convs (features[3..17], reference hardnet/HardNet.py:284-298) as implicit GEMMs:
//     M = output pixels (128 per tile), N = C_out, K = 9 taps x C_in. The A operand of every tap is
//     fetched by a 4-D TMA box (C, x, y, patch) whose start coordinate is shifted by the tap offset;
//     out-of-range coordinates are zero-filled by TMA, which is exactly the conv's zero padding, and
//     because x/y are per-patch tensor dimensions nothing ever bleeds between patches.
//     Stride-2 layers use four "parity" views of the input (even/odd rows x even/odd columns), so the
//     stride never has to be expressed to TMA.
//   * the 8x8 head conv (features[19..20] + L2Norm, HardNet.py:300-301,314-315 and Utils.py:19-22) as
//     a plain [B, 8192] x [8192, 128] GEMM with bias + L2-normalise in the epilogue.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread UMMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/ReLU/pack or L2-normalise -> global).
// Two accumulator buffers in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once

#include "common.cuh"

namespace hn {

constexpr int kTileM = 128;
constexpr int kTcThreads = 192;

enum : int { LOAD_CONV3X3 = 0, LOAD_GEMM = 1 };
enum : int { EPI_BIAS_RELU_PACK16 = 0, EPI_BIAS_L2NORM = 1 };
enum : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2, DT_U8 = 3 };

struct TcParams {
  CUtensorMap tmA[4];      // conv stride 1 / gemm: [0]; conv stride 2: parity views [ypar * 2 + xpar]
  CUtensorMap tmB;         // weights, [C_out, K] K-major
  const float* bias;       // [N] folded BatchNorm shift
  void* out;               // conv: 16-bit [rows, N]; head: f32/f16/bf16 [rows, N]
  long long total_rows;    // valid output rows (pixels or patches)
  int num_tiles;
  int num_k_stages;        // conv: 9 * cin_chunks, gemm: K / (KCB / 2)
  int stride;              // conv only: 1 or 2
  int cin_chunks;          // conv only: C_in / (KCB / 2)
  int tiles_per_patch;     // conv only: (H_out * W_out) / 128, or 0 when a tile holds several patches
  int rows_per_tile;       // conv only: output image rows per tile when tiles_per_patch >= 1
  int patches_per_tile;    // conv only: patches per tile when tiles_per_patch == 0
  int act_bf16;            // 16-bit activation flavour: 0 = fp16, 1 = bf16
  int out_dtype;           // head only: DT_F32 / DT_F16 / DT_BF16
  float l2_eps;            // head only: 1e-10 (Utils.py:18)
};

template <int N>
struct TmemCols {
  static constexpr uint32_t value = (2 * N <= 32) ? 32 : (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;
};

template <int N, int KCB, int STAGES>
constexpr size_t tc_smem_bytes() {
  return size_t(STAGES) * (size_t(kTileM) * KCB + size_t(N) * KCB) + 1024 /*align slack*/ + 256 /*barriers*/ + N * 4 /*bias*/;
}

__device__ __forceinline__ uint32_t pack16(float lo, float hi, int bf16) {
  if (bf16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  // saturate instead of producing inf for out-of-range activations
  lo = fminf(lo, 65504.f);
  hi = fminf(hi, 65504.f);
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int N, int KCB, int STAGES, int LOAD, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) tc_kernel(const __grid_constant__ TcParams p) {
  static_assert(KCB == 64 || KCB == 128, "stage rows are one 64B or 128B swizzle span");
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA M=128 needs N % 16 == 0");
  constexpr uint32_t A_BYTES = kTileM * KCB;
  constexpr uint32_t B_BYTES = N * KCB;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must stay 1024B aligned");
  constexpr uint32_t TMEM_COLS = TmemCols<N>::value;
  constexpr int KC_ELEMS = KCB / 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_addr));
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - raw_addr));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    if (LOAD == LOAD_CONV3X3 && p.stride == 2) {
      tma_prefetch_desc(&p.tmA[1]);
      tma_prefetch_desc(&p.tmA[2]);
      tma_prefetch_desc(&p.tmA[3]);
    }
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < N; i += 128) s_bias[i] = p.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int patch0 = 0, y0 = 0;
        if (LOAD == LOAD_CONV3X3) {
          if (p.tiles_per_patch >= 1) {
            patch0 = tile / p.tiles_per_patch;
            y0 = (tile - patch0 * p.tiles_per_patch) * p.rows_per_tile;
          } else {
            patch0 = tile * p.patches_per_tile;
          }
        }
        for (int ks = 0; ks < p.num_k_stages; ++ks) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = base + stage * STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          if (LOAD == LOAD_CONV3X3) {
            const int tap = ks / p.cin_chunks;
            const int cc = ks - tap * p.cin_chunks;
            const int ky = tap / 3, kx = tap - ky * 3;
            if (p.stride == 1) {
              tma_load_4d(a_dst, &p.tmA[0], full_bar(stage), cc * KC_ELEMS, kx - 1, y0 + ky - 1, patch0);
            } else {
              // input x = 2*ox + kx - 1: kx=0 -> odd column ox-1, kx=1 -> even column ox, kx=2 -> odd column ox
              const int xpar = (kx != 1), ypar = (ky != 1);
              const int xs = (kx == 0) ? -1 : 0;
              const int ys = y0 + ((ky == 0) ? -1 : 0);
              tma_load_4d(a_dst, &p.tmA[ypar * 2 + xpar], full_bar(stage), cc * KC_ELEMS, xs, ys, patch0);
            }
          } else {
            tma_load_2d(a_dst, &p.tmA[0], full_bar(stage), ks * KC_ELEMS, tile * kTileM);
          }
          tma_load_2d(b_dst, &p.tmB, full_bar(stage), ks * KC_ELEMS, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== UMMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kTileM, N, p.act_bf16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N;
        for (int ks = 0; ks < p.num_k_stages; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = base + stage * STAGE_BYTES;
          const uint64_t a_desc = make_kmajor_desc(a_addr, KCB);
          const uint64_t b_desc = make_kmajor_desc(a_addr + A_BYTES, KCB);
#pragma unroll
          for (int k = 0; k < KCB / 32; ++k) {
            // advance 16 K-elements = 32 bytes inside the swizzle span: +2 in the (addr >> 4) field
            umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (ks | k) != 0);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int row_in_tile = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * N;
      const long long row = static_cast<long long>(tile) * kTileM + row_in_tile;
      const bool valid = row < p.total_rows;
      if (EPI == EPI_BIAS_RELU_PACK16) {
        uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + row * N);
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c0, r);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v0 = fmaxf(__uint_as_float(r[2 * j]) + s_bias[c0 + 2 * j], 0.f);
            const float v1 = fmaxf(__uint_as_float(r[2 * j + 1]) + s_bias[c0 + 2 * j + 1], 0.f);
            o[j] = pack16(v0, v1, p.act_bf16);
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[c0 / 8 + j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
        }
      } else {
        // bias, then x / sqrt(sum(x*x) + eps) over the N columns of the row (Utils.py:19-22)
        float ss = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]) + s_bias[c0 + j];
            ss = fmaf(v, v, ss);
          }
        }
        const float inv = 1.0f / sqrtf(ss + p.l2_eps);
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (__uint_as_float(r[j]) + s_bias[c0 + j]) * inv;
          if (valid) {
            if (p.out_dtype == DT_F32) {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + row * N + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + row * N + c0);
              const int bf = p.out_dtype == DT_BF16;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(pack16(v[8 * j], v[8 * j + 1], bf), pack16(v[8 * j + 2], v[8 * j + 3], bf),
                                    pack16(v[8 * j + 4], v[8 * j + 5], bf), pack16(v[8 * j + 6], v[8 * j + 7], bf));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace hn
