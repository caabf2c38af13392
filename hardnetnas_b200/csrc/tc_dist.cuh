// Fused descriptor-distance kernels: a tcgen05 GEMM (D = A * B^T over 128-d descriptors) whose epilogue
// reduces each 128 x 128 accumulator tile in registers, so the Na x Nb distance matrix never exists in
// HBM. Replaces distance_matrix_vector + the masking / min / sort sequences of the reference
// (hardnet/Losses.py:5-13,95-108; FDLNet-master/utils/math_utils.py:8-19, eval_utils.py:113-114,168-175).
//
// A work item is (block of MB*128 A-rows, segment of B). The A block stays resident in shared memory
// (K <= 384), B tiles of 128 rows stream through a TMA ring, accumulators are double buffered in TMEM.
//
//   EPI_EXACT : operands are 3-way split fp16 (K = 384: hi*hi + hi*lo + lo*hi ~ fp32 dot). Every element
//               goes through the reference arithmetic (norms, eps, sqrt, diagonal / duplicate masks) and
//               feeds a per-row (value, index) minimum. Used for the loss and for hn_dist_min.
//   EPI_SHORTLIST : single fp16 pass (K = 128). Per row, the TOPC best 8-column chunks by dot product
//               are kept (max3 trees + a rarely taken sorted insert); a tiny fp32 kernel then re-ranks
//               those candidates exactly. Used for bulk matching.
#pragma once

#include "../../include/hardnet_b200.h"
#include "common.cuh"

namespace hn {

constexpr int kDistTile = 128;
constexpr int kDistStages = 4;
constexpr int kTopC = 4;     // chunks kept per row and segment
constexpr int kChunk = 8;    // columns per chunk

enum : int { EPI_EXACT = 0, EPI_SHORTLIST = 1, EPI_EXACT_NEI = 2 };   // EXACT_NEI = EXACT + the keypoint-neighbour masks

struct DistSide {
  CUtensorMap tmA;             // rows of this launch direction, [Na, K] 16-bit
  CUtensorMap tmB;             // columns, [Nb, K]
  const float* norm_a;         // |a_i|^2 (hardnet form)
  const float* norm_b;
  unsigned long long* row_pack;  // EXACT: packed (float bits << 32 | col) running minimum per row
  float* pos;                  // EXACT: diagonal distance before masking (may be null)
  int* cand;                   // SHORTLIST: [Na, segments, kTopC] chunk ids
  float* cand_val;             // SHORTLIST: the chunks' approximate (fp16-operand, scaled) dot-product maxima
  float* block_max;            // SHORTLIST, optional: [ceil(Nb / 8)][ceil(Na / 32)] maxima of the approximate dot products over
                               // (32-row block, 8-column chunk) cells: what the column side of mutual NN needs from this GEMM
  int n_row_blocks;            // ceil(Na / 32)
  // neighbour mask (FDLNet-master/latency/rfnet/model/rf_des.py:72-86): (x, y) keypoint coordinates of the row / column items
  // in the anchor image (xy_a*) and in the positive image (xy_p*); a pair closer than nei_c pixels in either gets +10 each
  const float2* xy_a_rows; const float2* xy_a_cols;
  const float2* xy_p_rows; const float2* xy_p_cols;
  long long Na, Nb;
};

struct DistParams {
  DistSide side[2];            // [1] = transposed problem (anchor swap / column minima)
  int k_blocks;                // K / 64
  int segments;
  int form;                    // HN_FORM_*
  int loss_mask;               // apply +1e-8, diagonal +10, (<0.008) +10
  int nei_mask;                // diagonal +10 and the neighbour masks (no +1e-8, no duplicate mask): HardNetNeiMask.loss
  float nei_c;                 // neighbour radius C in pixels
  float dot_scale;             // undoes the power-of-two operand scaling
};

template <int MB>
constexpr size_t dist_smem_bytes(int k_blocks) {
  return size_t(MB) * k_blocks * (kDistTile * 128) + size_t(kDistStages) * (kDistTile * 128) + 1024 + 256 + 2 * kDistTile * 4 +
         2 * kDistTile * 16 /* column keypoints of the neighbour mask */;
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// c[0..7]: this lane's (= this row's) maxima over eight consecutive 8-column chunks. Returns, in every lane, the maximum over
// the warp's 32 rows of chunk (lane >> 2): a halving butterfly - each exchange keeps half of the chunks and sends the other
// half, so the 8 x 32 values cost 4 + 2 + 1 + 1 + 1 = 9 shuffles instead of 40.
__device__ __forceinline__ float warp_chunk_max8(const float (&c)[8], int lane) {
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
  float a[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float keep = b4 ? c[t + 4] : c[t], send = b4 ? c[t] : c[t + 4];
    a[t] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
  float b[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float keep = b3 ? a[t + 2] : a[t], send = b3 ? a[t] : a[t + 2];
    b[t] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
  const float keep = b2 ? b[1] : b[0], send = b2 ? b[0] : b[1];
  float v = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}

// Epilogue warps per accumulator lane quarter. The exact epilogues (loss kernels) cost ~80 instructions per column and a loss-sized
// problem (N = 1024) gives every CTA exactly ONE 128 x 128 tile: with one warp per quarter a thread walked 128 columns alone
// (~10 k dependent instructions = 13 us of a 28 us kernel). Four warps per quarter take 32 columns each and merge through the
// packed atomicMin the segments already use.
template <int EPI>
constexpr int dist_col_split() { return (EPI == EPI_EXACT || EPI == EPI_EXACT_NEI) ? 4 : 1; }

template <int MB, int EPI>
__global__ void __launch_bounds__(64 + 128 * MB * dist_col_split<EPI>(), 1) dist_kernel(const __grid_constant__ DistParams p) {
  constexpr bool EXACT = EPI == EPI_EXACT || EPI == EPI_EXACT_NEI;
  constexpr int CS = dist_col_split<EPI>();
  static_assert(CS == 1 || MB == 1, "the column split is for the single-M-block loss kernels");
  constexpr bool NEI = EPI == EPI_EXACT_NEI;   // compile-time: the plain loss kernel carries none of the mask code
  constexpr uint32_t BLK_BYTES = kDistTile * 128;  // 128 rows x 64 fp16
  constexpr uint32_t TMEM_COLS = 2 * MB * kDistTile;
  const DistSide& sd = p.side[blockIdx.y];

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = base + MB * p.k_blocks * BLK_BYTES;
  const uint32_t bar_base = b_base + kDistStages * BLK_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kDistStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kDistStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kDistStages + 2 + a); };
  const uint32_t afull_bar = bar_base + 8u * (2 * kDistStages + 4);
  const uint32_t aempty_bar = bar_base + 8u * (2 * kDistStages + 5);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kDistStages + 6);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_addr));
  float* s_nb = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - raw_addr));  // [2][128]
  float4* s_xy = reinterpret_cast<float4*>(smem_raw + (bar_base + 256u + 2 * kDistTile * 4 - raw_addr));  // [2][128] (ax, ay, px, py)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = static_cast<int>((sd.Na + MB * kDistTile - 1) / (MB * kDistTile));
  const int n_tiles = static_cast<int>((sd.Nb + kDistTile - 1) / kDistTile);
  const int num_items = m_blocks * p.segments;
  auto seg_range = [&](int seg, int& t0, int& t1) {
    t0 = static_cast<int>(static_cast<long long>(n_tiles) * seg / p.segments);
    t1 = static_cast<int>(static_cast<long long>(n_tiles) * (seg + 1) / p.segments);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&sd.tmA);
    tma_prefetch_desc(&sd.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kDistStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4 * MB * CS);
      }
      mbar_init(afull_bar, 1);
      mbar_init(aempty_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer (warp-uniform) ==============================
    {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int mblk = item / p.segments, seg = item - mblk * p.segments;
        int t0, t1;
        seg_range(seg, t0, t1);
        mbar_wait(aempty_bar, a_phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(afull_bar, MB * p.k_blocks * BLK_BYTES);
          for (int mb = 0; mb < MB; ++mb)
            for (int kb = 0; kb < p.k_blocks; ++kb)
              tma_load_2d(a_base + (mb * p.k_blocks + kb) * BLK_BYTES, &sd.tmA, afull_bar, kb * 64,
                          (mblk * MB + mb) * kDistTile);
        }
        __syncwarp();
        a_phase ^= 1u;
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (elect_one()) {
              mbar_arrive_expect_tx(full_bar(stage), BLK_BYTES);
              tma_load_2d(b_base + stage * BLK_BYTES, &sd.tmB, full_bar(stage), kb * 64, t * kDistTile);
            }
            __syncwarp();
            if (++stage == kDistStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== UMMA issuer (warp-uniform) ==============================
    // Every lane runs the control flow, elect.sync picks the issuing lane; descriptors advance by adding to their lo word.
    {
      const uint32_t idesc = make_idesc_f16(kDistTile, kDistTile, 0);
      constexpr uint32_t HI = kmajor_desc_hi(128);
      const uint32_t a_lo0 = kmajor_desc_lo(a_base);
      const uint32_t b_lo0 = kmajor_desc_lo(b_base);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int mblk = item / p.segments, seg = item - mblk * p.segments;
        int t0, t1;
        seg_range(seg, t0, t1);
        mbar_wait(afull_bar, a_phase);
        a_phase ^= 1u;
        tc_fence_after();
        for (int t = t0; t < t1; ++t, ++it) {
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(stage) * (BLK_BYTES >> 4);
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) {
                const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(mb * p.k_blocks + kb) * (BLK_BYTES >> 4);
                const uint32_t d_tmem = tmem_base + (acc * MB + mb) * kDistTile;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_w(d_tmem, a_lo + 2u * k, HI, b_lo + 2u * k, HI, idesc, (kb | k) != 0);
              }
              umma_commit(empty_bar(stage));
              if (kb == p.k_blocks - 1) umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++stage == kDistStages) { stage = 0; phase ^= 1u; }
          }
        }
        if (elect_one()) umma_commit(aempty_bar);  // the resident A block may be overwritten once these MMAs retired
        __syncwarp();
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int q = warp & 3;
    const int mb = CS == 1 ? (warp - 2) >> 2 : 0;
    const int cs = CS == 1 ? 0 : (warp - 2) >> 2;   // which 32 of the tile's 128 columns (exact epilogues)
    const int ep_tid = threadIdx.x - 64;  // 0 .. 128*MB*CS-1
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int mblk = item / p.segments, seg = item - mblk * p.segments;
      int t0, t1;
      seg_range(seg, t0, t1);
      const long long row = static_cast<long long>(mblk * MB + mb) * kDistTile + q * 32 + lane;
      const bool row_ok = row < sd.Na;

      // per-row running state
      float best = __int_as_float(0x7f800000);
      int best_col = 0x7fffffff;
      float na = 0.f;
      float tb[kTopC];
      int tc[kTopC];
      float2 r_axy = make_float2(0.f, 0.f), r_pxy = make_float2(0.f, 0.f);
      if (EXACT) {
        if (p.form == HN_FORM_HARDNET && row_ok) na = sd.norm_a[row];
        if (NEI && row_ok) { r_axy = sd.xy_a_rows[row]; r_pxy = sd.xy_p_rows[row]; }
      } else {
#pragma unroll
        for (int i = 0; i < kTopC; ++i) { tb[i] = -__int_as_float(0x7f800000); tc[i] = 0; }
      }

      for (int t = t0; t < t1; ++t, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const long long col0 = static_cast<long long>(t) * kDistTile;
        if (EXACT && p.form == HN_FORM_HARDNET) {
          // column norms of this tile -> smem (double buffered with the accumulator)
          if (ep_tid < kDistTile) {
            const long long c = col0 + ep_tid;
            s_nb[acc * kDistTile + ep_tid] = c < sd.Nb ? sd.norm_b[c] : 0.f;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(128 * MB * CS) : "memory");
        }
        if (NEI) {
          if (ep_tid < kDistTile) {
            const long long c = col0 + ep_tid;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < sd.Nb) { const float2 a = sd.xy_a_cols[c], b = sd.xy_p_cols[c]; v = make_float4(a.x, a.y, b.x, b.y); }
            s_xy[acc * kDistTile + ep_tid] = v;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(128 * MB * CS) : "memory");
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (acc * MB + mb) * kDistTile;
        const bool ragged = col0 + kDistTile > sd.Nb;
        if (EPI == EPI_SHORTLIST) {
          // all four TMEM loads are issued before the single wait: one exposed round trip per tile instead of four
          uint32_t r[4][32];
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) tmem_ld32(t_row + gq * 32, r[gq]);
          tmem_ld_wait();
          float cm[16];   // this row's maxima of the tile's sixteen chunks
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int c0 = gq * 32;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[gq][j]);
            if (ragged) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + c0 + j >= sd.Nb) v[j] = -__int_as_float(0x7f800000);
            }
#pragma unroll
            for (int s = 0; s < 32 / kChunk; ++s) {
              const float* w = v + s * kChunk;
              const float m = fmaxf(max3(w[0], w[1], w[2]), max3(max3(w[3], w[4], w[5]), w[6], w[7]));
              cm[gq * 4 + s] = m;
              if (m > tb[kTopC - 1]) {
                const int id = static_cast<int>((col0 + c0) / kChunk) + s;
                // sorted insert (descending); strict '>' keeps the earlier chunk on ties
                const bool g0 = m > tb[0], g1 = m > tb[1], g2 = m > tb[2];
                tb[3] = g2 ? tb[2] : m;               tc[3] = g2 ? tc[2] : id;
                tb[2] = g2 ? (g1 ? tb[1] : m) : tb[2]; tc[2] = g2 ? (g1 ? tc[1] : id) : tc[2];
                tb[1] = g1 ? (g0 ? tb[0] : m) : tb[1]; tc[1] = g1 ? (g0 ? tc[0] : id) : tc[1];
                tb[0] = g0 ? m : tb[0];               tc[0] = g0 ? id : tc[0];
              }
            }
          }
          if (sd.block_max != nullptr) {   // warp-uniform: column side of mutual NN (hn_match_mutual)
            const long long rb = row >> 5;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const float half8[8] = {cm[hf * 8], cm[hf * 8 + 1], cm[hf * 8 + 2], cm[hf * 8 + 3],
                                      cm[hf * 8 + 4], cm[hf * 8 + 5], cm[hf * 8 + 6], cm[hf * 8 + 7]};
              const float bm = warp_chunk_max8(half8, lane);
              const long long cb = col0 / kChunk + hf * 8 + (lane >> 2);
              if ((lane & 3) == 0 && cb * kChunk < sd.Nb && rb < sd.n_row_blocks) sd.block_max[cb * sd.n_row_blocks + rb] = bm;
            }
          }
        } else {
        // NOT unrolled: the exact epilogue is ~80 instructions per column and a loss-sized problem (N = 1024) runs it ONCE per CTA,
        // so four unrolled copies were 10 k instructions of straight-line code fetched cold - the instruction fetch (stall reason
        // no_inst in ncu) was most of the kernel's 28 us; one copy of the 32-column body stays in the instruction cache.
#pragma unroll 1
        for (int c0 = CS == 1 ? 0 : cs * 32; c0 < (CS == 1 ? kDistTile : cs * 32 + 32); c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + c0, r);
          tmem_ld_wait();
          if (EXACT) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const long long col = col0 + c0 + j;
              const float dot = __uint_as_float(r[j]) * p.dot_scale;
              float d;
              if (p.form == HN_FORM_HARDNET) {
                // sqrt((|a|^2 + |p|^2) - 2ab + 1e-6), hardnet/Losses.py:12-13
                const float s = (na + s_nb[acc * kDistTile + c0 + j]) - 2.0f * dot;
                d = sqrtf(s + 1e-6f);
              } else {
                // sqrt(clamp(2 - 2ab, 1e-8, 4)), FDLNet-master/utils/math_utils.py:15-18
                d = sqrtf(fminf(fmaxf(2.0f - 2.0f * dot, 1e-8f), 4.0f));
              }
              if (p.loss_mask) {
                d += 1e-8f;                                   // Losses.py:95
                if (col == row) {
                  if (sd.pos && row_ok) sd.pos[row] = d;      // Losses.py:99 (before masking)
                  d += 10.0f;                                 // Losses.py:100
                }
                if (d < 0.008f) d += 10.0f;                   // Losses.py:101-103
              }
              if (NEI) {
                if (col == row) {
                  if (sd.pos && row_ok) sd.pos[row] = d;      // rf_des.py:70 (before masking)
                  d += 10.0f;                                 // rf_des.py:71
                }
                // pairwise_distances(kp, kp).lt(C), math_utils.py:22-40: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 1e-8)) < C
                const float4 cxy = s_xy[acc * kDistTile + c0 + j];
                const float da = (r_axy.x * r_axy.x + r_axy.y * r_axy.y) + (cxy.x * cxy.x + cxy.y * cxy.y) - 2.0f * (r_axy.x * cxy.x + r_axy.y * cxy.y);
                const float dp = (r_pxy.x * r_pxy.x + r_pxy.y * r_pxy.y) + (cxy.z * cxy.z + cxy.w * cxy.w) - 2.0f * (r_pxy.x * cxy.z + r_pxy.y * cxy.w);
                if (sqrtf(fmaxf(da, 1e-8f)) < p.nei_c) d += 10.0f;   // rf_des.py:74-79
                if (sqrtf(fmaxf(dp, 1e-8f)) < p.nei_c) d += 10.0f;   // rf_des.py:80-85
              }
              if (col < sd.Nb && d < best) { best = d; best_col = static_cast<int>(col); }
            }
          } else {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (ragged) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + c0 + j >= sd.Nb) v[j] = -__int_as_float(0x7f800000);
            }
#pragma unroll
            for (int s = 0; s < 32 / kChunk; ++s) {
              const float* w = v + s * kChunk;
              const float m = fmaxf(max3(w[0], w[1], w[2]), max3(max3(w[3], w[4], w[5]), w[6], w[7]));
              if (m > tb[kTopC - 1]) {
                const int id = static_cast<int>((col0 + c0) / kChunk) + s;
                // sorted insert (descending); strict '>' keeps the earlier chunk on ties
                const bool g0 = m > tb[0], g1 = m > tb[1], g2 = m > tb[2];
                tb[3] = g2 ? tb[2] : m;               tc[3] = g2 ? tc[2] : id;
                tb[2] = g2 ? (g1 ? tb[1] : m) : tb[2]; tc[2] = g2 ? (g1 ? tc[1] : id) : tc[2];
                tb[1] = g1 ? (g0 ? tb[0] : m) : tb[1]; tc[1] = g1 ? (g0 ? tc[0] : id) : tc[1];
                tb[0] = g0 ? m : tb[0];               tc[0] = g0 ? id : tc[0];
              }
            }
          }
        }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }

      if (row_ok) {
        if (EXACT) {
          if (sd.row_pack && best_col != 0x7fffffff) {
            const unsigned long long pk =
                (static_cast<unsigned long long>(__float_as_uint(best)) << 32) | static_cast<unsigned int>(best_col);
            atomicMin(sd.row_pack + row, pk);  // distances are positive: bit pattern is order preserving
          }
        } else {
          int4* dst = reinterpret_cast<int4*>(sd.cand + (row * p.segments + seg) * kTopC);
          // unused slots (segment with < kTopC chunks) are marked -1
          *dst = make_int4(tb[0] > -3.0e38f ? tc[0] : -1, tb[1] > -3.0e38f ? tc[1] : -1, tb[2] > -3.0e38f ? tc[2] : -1,
                           tb[3] > -3.0e38f ? tc[3] : -1);
          *reinterpret_cast<float4*>(sd.cand_val + (row * p.segments + seg) * kTopC) = make_float4(tb[0], tb[1], tb[2], tb[3]);
        }
      }
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
