// Patch-resident execution of a run of NAS ops ("segment"): hardnetNAS/fbnet_building_blocks/fbnet_builder.py:455-570
// (IRFBlock: pw -> dw -> pwl [+ residual] [+ SE]) and :202-228 (Identity: max-pool / 1x1 conv) for a group of G patches
// whose activations never leave the SM between the segment's first load and its last store.
//
// One CTA = one group of G patches at a time (persistent over groups), ops executed one after the other by the whole CTA:
//   PW       1x1 conv as tcgen05 GEMM: A = the source activation IN PLACE (shared memory, channel-planar
//            [plane = c / 8][patch][pixel][c % 8] = the UMMA no-swizzle K-major layout: core matrix = 8 pixels x 16 B,
//            SBO = 128 B, LBO = plane pitch), B = the op's weight image (resident in shared memory for the whole launch),
//            D in tensor memory; epilogue = TMEM -> bias (+ residual from shared memory) (+ ReLU) -> 16-bit -> planar
//            destination buffer (a warp's 32 lanes = 32 consecutive pixels = 512 contiguous bytes per store).
//   DW       depthwise k x k (3 | 5, stride 1 | 2) on the CUDA cores, a thread = 8 channels x a vertical strip of 4 | 2 | 1
//            output rows (the strip shrinks with the map so that every thread of the CTA has work). fp16 activations:
//            packed HFMA2 arithmetic with fp16 weights and accumulators (twice the FP32 rate, no unpacking, 40 live
//            registers: two 512-thread CTAs per SM); the 25-term fp16 accumulation costs ~1e-4 of descriptor error against
//            the 1e-3 gate (measured 1.5e-4 .. 2.2e-4 max-abs on wang2/3/4). bf16 activations: fp32 FFMA2 path, the
//            arithmetic order of dw_conv_smem_kernel (nas.cu).
//   MAXPOOL  3x3 stride 2 pad 1 with packed 16-bit maxima;  SE  squeeze-excite in place.
// The three activation "slots" of the compiled program (hn_nas_op.src/dst/res) map to three shared-memory buffers, so the
// liveness analysis of the Python compiler carries over unchanged. HBM traffic of a segment = its first op's input + its
// last op's output (NHWC, the interchange layout of the per-op kernels), e.g. wang2 blocks 1-2: 16 KB in, 8 KB out per
// patch instead of 11 round trips.
// Several CTAs per SM (MINB) overlap one CTA's latency-bound GEMM phases with another's CUDA-core phases.
#pragma once

#include "common.cuh"
#include "tc_conv.cuh"

namespace hn {

constexpr int kSegMaxOps = 24;
constexpr int kSegThreads = 512;
enum : int { SEG_PW = 1, SEG_DW = 2, SEG_MAXPOOL = 3, SEG_SE = 4 };   // = OP_* of nas.cu

struct SegOp {
  int kind;
  int cin, cout;
  int kernel, stride;
  int hin, hout;
  int relu;
  int src_off, dst_off, res_off;      // byte offsets of the activation buffers from the aligned shared-memory base (res: -1 = none)
  int w_off, b_off, w2_off, b2_off;   // byte offsets from the aligned base (inside the weight blob)
  int mid;                            // SE hidden width
  int nt;                             // PW: N of one MMA (cout / nt chunks per accumulator tile)
  // PW, precomputed for the group size of the launch: the issuing thread only adds
  int tiles, ksteps, chunks;          // accumulator tiles of 128 pixel rows, K / 16, cout / nt
  uint32_t idesc, b_hi;               // instruction descriptor, hi word of the weight descriptor
  uint32_t pitch;                     // plane pitch in bytes of this op's (equal-sized) source / destination tensors
  int sh;                             // DW: output rows per thread (4 | 2 | 1)
};

struct SegParams {
  const uint16_t* in;     // [n][hin * hin * cin] NHWC 16-bit (input of ops[0])
  uint16_t* out;          // [n][hout * hout * cout] NHWC 16-bit (output of ops[n_ops - 1])
  const uint4* blob;      // weight image in global memory (copied once per CTA)
  int blob_off;           // where it goes (bytes from the aligned base)
  int blob_bytes;         // multiple of 16
  int n;                  // patches
  int G;                  // patches per group
  int n_ops;
  int tmem_cols;          // power of two >= 32
  int scratch_off;        // SE scratch: [C] mean, [mid] hidden, [C] gate (floats)
  int bar_off;            // mbarrier (8 B) + tensor-memory slot (4 B)
  int op_base, seg_id;    // index of ops[0] in the program / of this segment in the plan (diagnostics)
  SegOp ops[kSegMaxOps];
};

#ifdef HN_SEG_TRACE
// Diagnostic build only (-DHN_SEG_TRACE): cycles thread 0 of CTA 0 spends in every phase, summed over the launch.
//   [op index] whole op   [64 + op index] PW: up to "accumulators complete"   [32 + seg] load  [40 + seg] store  [48 + seg] total
//   [56 + seg] groups of CTA 0
__device__ unsigned long long hn_seg_trace[256];
#define HN_SEG_T(var) const long long var = clock64()
#define HN_SEG_ACC(slot, t0) do { if (blockIdx.x == 0 && threadIdx.x == 0) hn_seg_trace[slot] += static_cast<unsigned long long>(clock64() - (t0)); } while (0)
#else
#define HN_SEG_T(var) do { } while (0)
#define HN_SEG_ACC(slot, t0) do { } while (0)
#endif

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <bool BF16>
__device__ __forceinline__ float2 seg_unpack(uint32_t v) {
  if constexpr (BF16) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
  else return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// depthwise k x k, pad k / 2; planar source / destination buffers of `np` patches
template <int K, int S, int SH, bool BF16>
__device__ __forceinline__ void seg_dw_phase(const uint8_t* __restrict__ sm, const SegOp& o, int G, int np) {
  constexpr int PAD = K >> 1;
  constexpr int NR = (SH - 1) * S + K;
  const int C = o.cin, hin = o.hin, hout = o.hout;
  const int planes = C >> 3;
  const int strips = hout / SH;
  const uint32_t pitch_in = static_cast<uint32_t>(G) * hin * hin * 16;
  const uint32_t pitch_out = static_cast<uint32_t>(G) * hout * hout * 16;
  const float* s_w = reinterpret_cast<const float*>(sm + o.w_off);    // [K * K][C]
  const float* s_b = reinterpret_cast<const float*>(sm + o.b_off);    // [C]
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const int items = np * planes * strips * hout;
  for (int e = threadIdx.x; e < items; e += kSegThreads) {
    const int ox = e % hout;
    int t = e / hout;
    const int ys = t % strips;
    t /= strips;
    const int plane = t % planes;
    const int pl = t / planes;
    const int c8 = plane * 8;
    const uint8_t* map = sm + o.src_off + plane * pitch_in + static_cast<uint32_t>(pl) * hin * hin * 16;
    float2 acc[SH][4];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(s_b + c8);
      const float4 b1 = *reinterpret_cast<const float4*>(s_b + c8 + 4);
#pragma unroll
      for (int j = 0; j < SH; ++j) {
        acc[j][0] = make_float2(b0.x, b0.y); acc[j][1] = make_float2(b0.z, b0.w);
        acc[j][2] = make_float2(b1.x, b1.y); acc[j][3] = make_float2(b1.z, b1.w);
      }
    }
    const int iy0 = ys * SH * S - PAD;
#pragma unroll 1
    for (int kx = 0; kx < K; ++kx) {
      const int ix = ox * S + kx - PAD;
      const bool x_ok = ix >= 0 && ix < hin;
      float2 wk[K][4];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const float* wp = s_w + (ky * K + kx) * C + c8;
        const float4 w0 = *reinterpret_cast<const float4*>(wp);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
        wk[ky][0] = make_float2(w0.x, w0.y); wk[ky][1] = make_float2(w0.z, w0.w);
        wk[ky][2] = make_float2(w1.x, w1.y); wk[ky][3] = make_float2(w1.z, w1.w);
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int iy = iy0 + r;
        const bool ok = x_ok && iy >= 0 && iy < hin;
        uint4 xv = zero;
        if (ok) xv = *reinterpret_cast<const uint4*>(map + (iy * hin + ix) * 16);
        const float2 x[4] = {seg_unpack<BF16>(xv.x), seg_unpack<BF16>(xv.y), seg_unpack<BF16>(xv.z), seg_unpack<BF16>(xv.w)};
#pragma unroll
        for (int j = 0; j < SH; ++j) {
          const int ky = r - j * S;            // compile-time after unrolling
          if (ky >= 0 && ky < K) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = __ffma2_rn(x[q], wk[ky][q], acc[j][q]);
          }
        }
      }
    }
    uint8_t* optr = const_cast<uint8_t*>(sm) + o.dst_off + plane * pitch_out +
                    (static_cast<uint32_t>(pl) * hout * hout + ys * SH * hout + ox) * 16;
#pragma unroll
    for (int j = 0; j < SH; ++j) {
      if (o.relu) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = make_float2(fmaxf(acc[j][q].x, 0.f), fmaxf(acc[j][q].y, 0.f));
      }
      *reinterpret_cast<uint4*>(optr + j * hout * 16) =
          make_uint4(pack16(acc[j][0].x, acc[j][0].y, BF16), pack16(acc[j][1].x, acc[j][1].y, BF16),
                     pack16(acc[j][2].x, acc[j][2].y, BF16), pack16(acc[j][3].x, acc[j][3].y, BF16));
    }
  }
}

// fp16 activations: the same traversal in packed half2 arithmetic. Weights [K * K][C] and bias [C] are fp16 in the blob.
template <int K, int S, int SH>
__device__ __forceinline__ void seg_dw_phase_h2(const uint8_t* __restrict__ sm, const SegOp& o, int G, int np) {
  constexpr int PAD = K >> 1;
  constexpr int NR = (SH - 1) * S + K;
  const int C = o.cin, hin = o.hin, hout = o.hout;
  const int planes = C >> 3;
  const int strips = hout / SH;
  const uint32_t pitch_in = static_cast<uint32_t>(G) * hin * hin * 16;
  const uint32_t pitch_out = static_cast<uint32_t>(G) * hout * hout * 16;
  const uint8_t* s_w = sm + o.w_off;    // [K * K][C] fp16
  const uint8_t* s_b = sm + o.b_off;    // [C] fp16
  const int items = np * planes * strips * hout;
  const __half2 hzero = __float2half2_rn(0.f);
  for (int e = threadIdx.x; e < items; e += kSegThreads) {
    const int ox = e % hout;
    int t = e / hout;
    const int ys = t % strips;
    t /= strips;
    const int plane = t % planes;
    const int pl = t / planes;
    const uint8_t* map = sm + o.src_off + plane * pitch_in + static_cast<uint32_t>(pl) * hin * hin * 16;
    __half2 acc[SH][4];
    {
      const uint4 b = *reinterpret_cast<const uint4*>(s_b + plane * 16);
#pragma unroll
      for (int j = 0; j < SH; ++j) {
        acc[j][0] = *reinterpret_cast<const __half2*>(&b.x); acc[j][1] = *reinterpret_cast<const __half2*>(&b.y);
        acc[j][2] = *reinterpret_cast<const __half2*>(&b.z); acc[j][3] = *reinterpret_cast<const __half2*>(&b.w);
      }
    }
    const int iy0 = ys * SH * S - PAD;
#pragma unroll 1
    for (int kx = 0; kx < K; ++kx) {
      const int ix = ox * S + kx - PAD;
      const bool x_ok = ix >= 0 && ix < hin;
      uint4 wk[K];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) wk[ky] = *reinterpret_cast<const uint4*>(s_w + ((ky * K + kx) * C + plane * 8) * 2);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int iy = iy0 + r;
        const bool ok = x_ok && iy >= 0 && iy < hin;
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (ok) xv = *reinterpret_cast<const uint4*>(map + (iy * hin + ix) * 16);
        const __half2 x[4] = {*reinterpret_cast<const __half2*>(&xv.x), *reinterpret_cast<const __half2*>(&xv.y),
                              *reinterpret_cast<const __half2*>(&xv.z), *reinterpret_cast<const __half2*>(&xv.w)};
#pragma unroll
        for (int j = 0; j < SH; ++j) {
          const int ky = r - j * S;            // compile-time after unrolling
          if (ky >= 0 && ky < K) {
            const __half2 w[4] = {*reinterpret_cast<const __half2*>(&wk[ky].x), *reinterpret_cast<const __half2*>(&wk[ky].y),
                                  *reinterpret_cast<const __half2*>(&wk[ky].z), *reinterpret_cast<const __half2*>(&wk[ky].w)};
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = __hfma2(x[q], w[q], acc[j][q]);
          }
        }
      }
    }
    uint8_t* optr = const_cast<uint8_t*>(sm) + o.dst_off + plane * pitch_out +
                    (static_cast<uint32_t>(pl) * hout * hout + ys * SH * hout + ox) * 16;
#pragma unroll
    for (int j = 0; j < SH; ++j) {
      if (o.relu) {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = __hmax2(acc[j][q], hzero);
      }
      *reinterpret_cast<uint4*>(optr + j * hout * 16) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&acc[j][0]), *reinterpret_cast<const uint32_t*>(&acc[j][1]),
                     *reinterpret_cast<const uint32_t*>(&acc[j][2]), *reinterpret_cast<const uint32_t*>(&acc[j][3]));
    }
  }
}

template <int K, int S, bool BF16>
__device__ __forceinline__ void seg_dw_dispatch(const uint8_t* __restrict__ sm, const SegOp& o, int G, int np) {
  // One strip height everywhere (every map of the search space has an even side) and the kx loop kept rolled: the kernel
  // switches between op bodies every few thousand cycles, so the code of a body has to stay instruction-cache sized
  // (with every (K, S, strip) variant fully unrolled the kernel was 128 KB of SASS and ran at the i-cache miss rate).
  if constexpr (BF16) seg_dw_phase<K, S, 2, true>(sm, o, G, np);
  else seg_dw_phase_h2<K, S, 2>(sm, o, G, np);
}

// MaxPool2d(3, stride 2, padding 1): packed 16-bit maxima (exact; NaNs propagate like torch), padding = -inf
template <bool BF16>
__device__ __forceinline__ void seg_maxpool_phase(const uint8_t* __restrict__ sm, const SegOp& o, int G, int np) {
  constexpr int SH = 4, NR = (SH - 1) * 2 + 3;
  const int C = o.cin, hin = o.hin, hout = o.hout;
  const int planes = C >> 3;
  const int strips = hout / SH;
  const uint32_t pitch_in = static_cast<uint32_t>(G) * hin * hin * 16;
  const uint32_t pitch_out = static_cast<uint32_t>(G) * hout * hout * 16;
  const uint32_t ninf = BF16 ? 0xFF80FF80u : 0xFC00FC00u;
  auto vmax = [](uint32_t a, uint32_t b) -> uint32_t {
    if constexpr (BF16) {
      const __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
      return *reinterpret_cast<const uint32_t*>(&r);
    } else {
      const __half2 r = __hmax2_nan(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
      return *reinterpret_cast<const uint32_t*>(&r);
    }
  };
  const int items = np * planes * strips * hout;
  for (int e = threadIdx.x; e < items; e += kSegThreads) {
    const int ox = e % hout;
    int t = e / hout;
    const int ys = t % strips;
    t /= strips;
    const int plane = t % planes;
    const int pl = t / planes;
    const uint8_t* map = sm + o.src_off + plane * pitch_in + static_cast<uint32_t>(pl) * hin * hin * 16;
    uint4 acc[SH];
#pragma unroll
    for (int j = 0; j < SH; ++j) acc[j] = make_uint4(ninf, ninf, ninf, ninf);
    const int iy0 = ys * SH * 2 - 1;
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      const bool x_ok = ix >= 0 && ix < hin;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int iy = iy0 + r;
        const bool ok = x_ok && iy >= 0 && iy < hin;
        uint4 xv = make_uint4(ninf, ninf, ninf, ninf);
        if (ok) xv = *reinterpret_cast<const uint4*>(map + (iy * hin + ix) * 16);
#pragma unroll
        for (int j = 0; j < SH; ++j) {
          const int ky = r - j * 2;
          if (ky >= 0 && ky < 3)
            acc[j] = make_uint4(vmax(acc[j].x, xv.x), vmax(acc[j].y, xv.y), vmax(acc[j].z, xv.z), vmax(acc[j].w, xv.w));
        }
      }
    }
    uint8_t* optr = const_cast<uint8_t*>(sm) + o.dst_off + plane * pitch_out +
                    (static_cast<uint32_t>(pl) * hout * hout + ys * SH * hout + ox) * 16;
#pragma unroll
    for (int j = 0; j < SH; ++j) *reinterpret_cast<uint4*>(optr + j * hout * 16) = acc[j];
  }
}

// Squeeze-and-excite in place (fbnet_builder.py:407-421): x * sigmoid(fc2(relu(fc1(mean_hw(x))))); same summation order
// as se_kernel (nas.cu).
template <bool BF16>
__device__ __forceinline__ void seg_se_phase(uint8_t* __restrict__ sm, const SegOp& o, int G, int np, int scratch_off) {
  const int C = o.cin, pix = o.hin * o.hin, mid = o.mid;
  const uint32_t pitch = static_cast<uint32_t>(G) * pix * 16;
  float* s_mean = reinterpret_cast<float*>(sm + scratch_off);
  float* s_hid = s_mean + C;
  float* s_gate = s_hid + mid;
  const float* w1 = reinterpret_cast<const float*>(sm + o.w_off);
  const float* b1 = reinterpret_cast<const float*>(sm + o.b_off);
  const float* w2 = reinterpret_cast<const float*>(sm + o.w2_off);
  const float* b2 = reinterpret_cast<const float*>(sm + o.b2_off);
  for (int pl = 0; pl < np; ++pl) {
    uint8_t* xp = sm + o.src_off + static_cast<uint32_t>(pl) * pix * 16;
    for (int c = threadIdx.x; c < C; c += kSegThreads) {
      const uint8_t* col = xp + (c >> 3) * pitch + (c & 7) * 2;
      float s = 0.f;
      for (int i = 0; i < pix; ++i) {
        const uint16_t raw = *reinterpret_cast<const uint16_t*>(col + i * 16);
        s += BF16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&raw)) : __half2float(*reinterpret_cast<const __half*>(&raw));
      }
      s_mean[c] = s / static_cast<float>(pix);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < mid; j += kSegThreads) {
      float s = b1[j];
      for (int c = 0; c < C; ++c) s = fmaf(w1[j * C + c], s_mean[c], s);
      s_hid[j] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kSegThreads) {
      float s = b2[c];
      for (int j = 0; j < mid; ++j) s = fmaf(w2[c * mid + j], s_hid[j], s);
      s_gate[c] = 1.0f / (1.0f + expf(-s));
    }
    __syncthreads();
    const int planes = C >> 3;
    for (int i = threadIdx.x; i < pix * planes; i += kSegThreads) {
      const int plane = i / pix, px = i - plane * pix;
      uint4* ptr = reinterpret_cast<uint4*>(xp + plane * pitch + px * 16);
      uint4 v = *ptr;
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = seg_unpack<BF16>(w[t]);
        w[t] = pack16(f.x * s_gate[plane * 8 + 2 * t], f.y * s_gate[plane * 8 + 2 * t + 1], BF16);
      }
      *ptr = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
  }
}

template <int MINB, bool BF16>
__global__ void __launch_bounds__(kSegThreads, MINB) nas_seg_kernel(const __grid_constant__ SegParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const uint32_t bar = base + p.bar_off;
  const uint32_t tmem_slot = bar + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.bar_off + 8);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = p.G;

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  {
    uint4* dst = reinterpret_cast<uint4*>(sm + p.blob_off);
    for (int i = threadIdx.x; i < p.blob_bytes / 16; i += kSegThreads) dst[i] = __ldg(p.blob + i);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  uint32_t mma_phase = 0;

  const SegOp& first = p.ops[0];
  const SegOp& last = p.ops[p.n_ops - 1];
  const int in_pix = first.hin * first.hin, in_planes = first.cin >> 3;
  const int out_pix = last.hout * last.hout, out_planes = last.cout >> 3;
  const size_t in_patch_bytes = static_cast<size_t>(in_pix) * first.cin * 2;
  const size_t out_patch_bytes = static_cast<size_t>(out_pix) * last.cout * 2;
  const int num_groups = (p.n + G - 1) / G;

  HN_SEG_T(t_total);
  for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
    const int np = min(G, p.n - grp * G);
    HN_SEG_T(t_load);
    // ---- NHWC global -> channel-planar shared memory in 16-byte chunks. Chunk i = (pixel block i / (8 planes'),
    // plane, pixel % 8): 8 consecutive lanes write 8 consecutive pixels of one plane (128 contiguous bytes, conflict
    // free) while a warp as a whole still reads whole 128-byte lines of the NHWC tensor. ----
    {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.in) + static_cast<size_t>(grp) * G * in_patch_bytes;
      const uint32_t pitch = static_cast<uint32_t>(G) * in_pix * 16;
      const uint32_t dst0 = base + first.src_off;
      const int chunks = np * in_pix * in_planes;
      for (int i = threadIdx.x; i < chunks; i += kSegThreads) {
        const int blk = i / (8 * in_planes), rem = i - blk * 8 * in_planes;
        const int plane = rem >> 3, pixel = blk * 8 + (rem & 7);
        cp_async_16(dst0 + plane * pitch + pixel * 16, src + (static_cast<size_t>(pixel) * in_planes + plane) * 16);
      }
      cp_async_wait_all();
    }
    fence_proxy_async_smem();
    __syncthreads();
    HN_SEG_ACC(32 + p.seg_id, t_load);

    for (int oi = 0; oi < p.n_ops; ++oi) {
      const SegOp& o = p.ops[oi];
      HN_SEG_T(t_op);
      if (o.kind == SEG_PW) {
        const uint32_t a_pitch = o.pitch;
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_hi = noswizzle_desc_hi(128);
            const uint32_t a_step = (2u * a_pitch) >> 4;                       // one K step = two 8-channel planes
            const uint32_t b_chunk = (static_cast<uint32_t>(o.nt) * o.cin * 2) >> 4;
            uint32_t a_lo_t = noswizzle_desc_lo(base + o.src_off, a_pitch);
            const uint32_t b_lo0 = noswizzle_desc_lo(base + o.w_off, 128);
            uint32_t d_t = tmem_base;
            for (int t = 0; t < o.tiles; ++t, a_lo_t += 2048u >> 4, d_t += o.cout) {
              uint32_t b_lo_c = b_lo0, d = d_t;
              for (int nc = 0; nc < o.chunks; ++nc, b_lo_c += b_chunk, d += o.nt) {
                uint32_t a_lo = a_lo_t, b_lo = b_lo_c;
                for (int k = 0; k < o.ksteps; ++k, a_lo += a_step, b_lo += 16u)
                  umma_f16_w(d, a_lo, a_hi, b_lo, o.b_hi, o.idesc, k != 0);
              }
            }
            umma_commit(bar);
          }
          __syncwarp();
          HN_SEG_ACC(128 + p.op_base + oi, t_op);
        }
        mbar_wait(bar, mma_phase);
        mma_phase ^= 1u;
        tc_fence_after();
        HN_SEG_ACC(64 + p.op_base + oi, t_op);
        // ---- epilogue: (tile, 16-column group) tasks dealt round robin to the four groups of four warps ----
        const int q = warp & 3;
        const int n_cg = o.cout >> 4;
        const int valid_rows = np * o.hin * o.hin;
        const uint8_t* s_bias = sm + o.b_off;
        int t = 0, cg = warp >> 2;
        while (cg >= n_cg) { cg -= n_cg; ++t; }
        while (t < o.tiles) {
          if (t * kTileM + q * 32 < valid_rows) {          // warp-uniform
            const int prow = t * kTileM + q * 32 + lane;
            uint32_t r[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * o.cout + cg * 16, r);
            tmem_ld_wait();
            if (prow < valid_rows) {
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const float4 b0 = *reinterpret_cast<const float4*>(s_bias + (cg * 16 + h2 * 8) * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(s_bias + (cg * 16 + h2 * 8 + 4) * 4);
                float v[8] = {__uint_as_float(r[8 * h2]) + b0.x,     __uint_as_float(r[8 * h2 + 1]) + b0.y,
                              __uint_as_float(r[8 * h2 + 2]) + b0.z, __uint_as_float(r[8 * h2 + 3]) + b0.w,
                              __uint_as_float(r[8 * h2 + 4]) + b1.x, __uint_as_float(r[8 * h2 + 5]) + b1.y,
                              __uint_as_float(r[8 * h2 + 6]) + b1.z, __uint_as_float(r[8 * h2 + 7]) + b1.w};
                const uint32_t poff = static_cast<uint32_t>(cg * 2 + h2) * a_pitch + prow * 16;
                if (o.res_off >= 0) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(sm + o.res_off + poff);
                  const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                  for (int t2 = 0; t2 < 4; ++t2) {
                    const float2 f = seg_unpack<BF16>(w[t2]);
                    v[2 * t2] += f.x;
                    v[2 * t2 + 1] += f.y;
                  }
                }
                uint4 ov;
                if (o.relu)
                  ov = make_uint4(pack16_relu(v[0], v[1], BF16), pack16_relu(v[2], v[3], BF16), pack16_relu(v[4], v[5], BF16),
                                  pack16_relu(v[6], v[7], BF16));
                else
                  ov = make_uint4(pack16(v[0], v[1], BF16), pack16(v[2], v[3], BF16), pack16(v[4], v[5], BF16), pack16(v[6], v[7], BF16));
                *reinterpret_cast<uint4*>(sm + o.dst_off + poff) = ov;
              }
            }
          }
          cg += kSegThreads / 128;
          while (cg >= n_cg) { cg -= n_cg; ++t; }
        }
        HN_SEG_ACC(160 + p.op_base + oi, t_op);
      } else if (o.kind == SEG_DW) {
        if (o.kernel == 3 && o.stride == 1) seg_dw_dispatch<3, 1, BF16>(sm, o, G, np);
        else if (o.kernel == 3) seg_dw_dispatch<3, 2, BF16>(sm, o, G, np);
        else if (o.stride == 1) seg_dw_dispatch<5, 1, BF16>(sm, o, G, np);
        else seg_dw_dispatch<5, 2, BF16>(sm, o, G, np);
      } else if (o.kind == SEG_MAXPOOL) {
        seg_maxpool_phase<BF16>(sm, o, G, np);
      } else {
        seg_se_phase<BF16>(sm, o, G, np, p.scratch_off);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      HN_SEG_ACC(p.op_base + oi, t_op);
    }

    // ---- channel-planar shared memory -> NHWC global (same chunk order as the load) ----
    HN_SEG_T(t_store);
    {
      uint8_t* dst = reinterpret_cast<uint8_t*>(p.out) + static_cast<size_t>(grp) * G * out_patch_bytes;
      const uint32_t pitch = static_cast<uint32_t>(G) * out_pix * 16;
      const uint8_t* s0 = sm + last.dst_off;
      const int chunks = np * out_pix * out_planes;
      for (int i = threadIdx.x; i < chunks; i += kSegThreads) {
        const int blk = i / (8 * out_planes), rem = i - blk * 8 * out_planes;
        const int plane = rem >> 3, pixel = blk * 8 + (rem & 7);
        *reinterpret_cast<uint4*>(dst + (static_cast<size_t>(pixel) * out_planes + plane) * 16) =
            *reinterpret_cast<const uint4*>(s0 + plane * pitch + pixel * 16);
      }
    }
    __syncthreads();   // the buffers are free for the next group's load
    HN_SEG_ACC(40 + p.seg_id, t_store);
#ifdef HN_SEG_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) hn_seg_trace[56 + p.seg_id] += 1;
#endif
  }
  HN_SEG_ACC(48 + p.seg_id, t_total);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace hn
