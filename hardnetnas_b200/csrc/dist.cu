// C ABI for the distance / hardest-in-batch / matching path: operand packing, the fused tcgen05
// distance kernels (tc_dist.cuh), fp32 re-ranking of shortlisted candidates and the small finalisers.
#include <algorithm>
#include <vector>

#include "host_common.h"
#include "tc_dist.cuh"
#include "tc_dist_pair.cuh"

namespace hn {

constexpr float kOperandScale = 256.0f;  // descriptors are unit vectors: x256 keeps the fp16 hi/lo split clear of subnormals
constexpr float kDotScale = 1.0f / (kOperandScale * kOperandScale);

static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// fp32 [n,128] -> 16-bit K-major operand rows. split == 0: [hi] (K=128). split == 1: K=384,
// order 0 = [hi, hi, lo] (row side), order 1 = [hi, lo, hi] (column side), so that the GEMM accumulates
// hi*hi + hi*lo + lo*hi. Also emits |x|^2 (torch.sum(x*x, dim=1), hardnet/Losses.py:8-9).
__global__ void pack_desc_kernel(const float* __restrict__ x, long long n, uint16_t* __restrict__ out, int split, int order,
                                 float* __restrict__ norm) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4 v = reinterpret_cast<const float4*>(x + row * 128)[lane];
  const float f[4] = {v.x, v.y, v.z, v.w};
  float ss = (f[0] * f[0] + f[1] * f[1]) + (f[2] * f[2] + f[3] * f[3]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (norm && lane == 0) norm[row] = ss;
  uint16_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float s = f[j] * kOperandScale;
    const __half h = __float2half_rn(s);
    const __half l = __float2half_rn(s - __half2float(h));
    hi[j] = __half_as_ushort(h);
    lo[j] = __half_as_ushort(l);
  }
  const int K = split ? 384 : 128;
  uint16_t* dst = out + row * K + lane * 4;
  const uint2 H = make_uint2(hi[0] | (uint32_t(hi[1]) << 16), hi[2] | (uint32_t(hi[3]) << 16));
  const uint2 L = make_uint2(lo[0] | (uint32_t(lo[1]) << 16), lo[2] | (uint32_t(lo[3]) << 16));
  *reinterpret_cast<uint2*>(dst) = H;
  if (split) {
    *reinterpret_cast<uint2*>(dst + 128) = order == 0 ? H : L;
    *reinterpret_cast<uint2*>(dst + 256) = order == 0 ? L : H;
  }
}

// Push all-gather over NVSwitch multicast: a rank packs its shard and writes BOTH forms (fp16 operand rows for the GEMM, fp32
// rows for the exact re-rank) to multicast addresses (multimem.st): the switch replicates every 16-byte store into the
// symmetric buffer of every GPU of the group, so after one cross-GPU barrier each GPU holds the whole gallery without a
// gather kernel, a copy engine or a second pass over the data. Half a warp per row: a lane owns 8 consecutive floats.
__global__ void pack_desc_multicast_kernel(const float* __restrict__ x, long long n, uint16_t* __restrict__ mc16,
                                           float* __restrict__ mc32) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long row = gid >> 4;
  const int l = static_cast<int>(gid & 15);
  if (row >= n) return;
  const float4 a = reinterpret_cast<const float4*>(x + row * 128)[2 * l];
  const float4 b = reinterpret_cast<const float4*>(x + row * 128)[2 * l + 1];
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 v = __floats2half2_rn(f[2 * j] * kOperandScale, f[2 * j + 1] * kOperandScale);
    h[j] = *reinterpret_cast<const uint32_t*>(&v);
  }
  if (mc16 != nullptr)
    asm volatile("multimem.st.weak.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc16 + row * 128 + l * 8), "r"(h[0]), "r"(h[1]),
                 "r"(h[2]), "r"(h[3])
                 : "memory");
  if (mc32 != nullptr) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc32 + row * 128 + l * 8), "f"(a.x), "f"(a.y),
                 "f"(a.z), "f"(a.w)
                 : "memory");
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc32 + row * 128 + l * 8 + 4), "f"(b.x), "f"(b.y),
                 "f"(b.z), "f"(b.w)
                 : "memory");
  }
  // no per-thread fence: the stores are complete when the kernel retires, and the cross-GPU barrier that follows on the same
  // stream publishes them (release / acquire at system scope); a __threadfence_system() here cost 0.1 ms per 24 MiB
}

__global__ void unpack_min_kernel(const unsigned long long* __restrict__ pack, long long n, float* __restrict__ val,
                                  int* __restrict__ arg) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long pk = pack[i];
  if (val) val[i] = __uint_as_float(static_cast<unsigned int>(pk >> 32));
  if (arg) arg[i] = static_cast<int>(pk & 0xffffffffu);
}

// mean(clamp(margin + pos - min_neg, 0)) with min_neg = min(row_min, col_min) — hardnet/Losses.py:105-108,
// 142-143,153. One block, fixed summation order -> deterministic.
__global__ void loss_finalize_kernel(const float* __restrict__ pos, const unsigned long long* __restrict__ row_pack,
                                     const unsigned long long* __restrict__ col_pack, long long n, float margin,
                                     float* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    float mn = __uint_as_float(static_cast<unsigned int>(row_pack[i] >> 32));
    if (col_pack) mn = fminf(mn, __uint_as_float(static_cast<unsigned int>(col_pack[i] >> 32)));
    acc += fmaxf(margin + pos[i] - mn, 0.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[0] = v / static_cast<float>(n);
  }
}

// Exact fp32 re-rank of the shortlisted chunks: one warp per query row, one candidate gallery row per lane.
// Distances in the FDLNet form; ties resolve to the lower gallery index like torch.min / a stable sort.
// Known limit (documented precondition of hn_match): the GEMM keeps the kTopC = 4 best chunks per row and list (a list = one
// gallery segment, or one 64-column half of it in the CTA-pair kernel). Should five or more chunks of ONE list reach the
// threshold, the fifth is not re-ranked; with 15-30 lists per row this needs five near-tied best candidates inside one list
// (|d - d2| within the 2^-9 margin). Scanning the whole segment whenever a list's fourth chunk passes the threshold was
// measured: it fires on ~5 % of the rows of the BASELINE config-4 set and makes the call 5x slower (0.88 -> 4.2 ms), so the
// exact remedy has to track the largest DROPPED chunk maximum in the GEMM epilogue instead (not done).
__global__ void __launch_bounds__(128) rerank_kernel(const float* __restrict__ q, const float* __restrict__ g,
                                                     const int* __restrict__ cand, const float* __restrict__ cand_val,
                                                     float margin, long long nq, long long ng, int slots,
                                                     long long g_offset, float* __restrict__ d1, float* __restrict__ d2,
                                                     int* __restrict__ i1, int* __restrict__ i2) {
  __shared__ __align__(16) float sq[4][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + w;
  if (row >= nq) return;
  reinterpret_cast<float4*>(sq[w])[lane] = reinterpret_cast<const float4*>(q + row * 128)[lane];
  __syncwarp();
  // A chunk can hold one of the two nearest columns only if its approximate maximum reaches the second largest
  // chunk maximum minus twice the error bound of the 16-bit-operand dot product (`margin`, same scaled units).
  float thr;
  {
    float m1 = -__int_as_float(0x7f800000), m2 = m1;
    for (int c = lane; c < slots; c += 32) {
      const float v = cand[row * slots + c] >= 0 ? cand_val[row * slots + c] : -__int_as_float(0x7f800000);
      if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) { m2 = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
      const float lo = fminf(m1, o1);
      m1 = fmaxf(m1, o1);
      m2 = fmaxf(fmaxf(m2, o2), lo);
    }
    thr = m2 - margin;   // -inf when fewer than two chunks exist: everything is re-ranked
  }
  const float inf = __int_as_float(0x7f800000);
  float b1 = inf, b2 = inf;
  int j1 = 0x7fffffff, j2 = 0x7fffffff;
  // The warp walks the surviving chunks together: a gallery row is ONE coalesced 512 B read (lane = float4 index) and
  // the 128-term dot is finished with a butterfly, so every lane ends up with the same (d, index) stream. Per-lane
  // row reads (one lane per candidate) cost 32 L1 wavefronts per row instead of 4.
  const float4 qv = reinterpret_cast<const float4*>(sq[w])[lane];
  auto eval_chunk = [&](long long chunk) {
    const long long col0 = chunk * kChunk;
    float part[kChunk];
#pragma unroll
    for (int r = 0; r < kChunk; ++r) {
      const long long col = col0 + r;
      part[r] = 0.f;
      if (col < ng) {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + col * 128) + lane);
        part[r] = fmaf(qv.x, gv.x, fmaf(qv.y, gv.y, fmaf(qv.z, gv.z, qv.w * gv.w)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < kChunk; ++r) part[r] += __shfl_xor_sync(0xffffffffu, part[r], o);
    }
#pragma unroll
    for (int r = 0; r < kChunk; ++r) {
      const long long col = col0 + r;
      if (col >= ng) continue;
      const float d = sqrtf(fminf(fmaxf(2.0f - 2.0f * part[r], 1e-8f), 4.0f));
      const int ci = static_cast<int>(col);
      if (d < b1 || (d == b1 && ci < j1)) {
        b2 = b1; j2 = j1; b1 = d; j1 = ci;
      } else if (d < b2 || (d == b2 && ci < j2)) {
        b2 = d; j2 = ci;
      }
    }
  };
  for (int c = 0; c < slots; ++c) {
    const int chunk = cand[row * slots + c];
    if (chunk < 0 || cand_val[row * slots + c] < thr) continue;   // warp-uniform
    eval_chunk(chunk);
  }
  // every lane holds the same top-2 now
  const float m = b1, s = b2;
  const int mj = j1, sj = j2;
  if (lane == 0) {
    if (d1) d1[row] = m;
    if (i1) i1[row] = mj == 0x7fffffff ? -1 : static_cast<int>(mj + g_offset);
    if (d2) d2[row] = s;
    if (i2) i2[row] = sj == 0x7fffffff ? -1 : static_cast<int>(sj + g_offset);
  }
}

// Optional per-stage CUDA-event timing of hn_match (bench.py: the roofline of the matching GEMM is computed from the kernel's
// own in-run time): stage 0 = operand packing, 1 = GEMM + shortlist, 2 = exact re-rank.
struct MatchProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev[3];
  size_t used[3] = {0, 0, 0};
};
static MatchProfile g_match_profile;
struct MatchTimer {
  int stage;
  cudaStream_t s;
  bool on;
  MatchTimer(int stage_, cudaStream_t s_) : stage(stage_), s(s_), on(g_match_profile.on) { if (on) record(); }
  ~MatchTimer() { if (on) record(); }
  void record() {
    auto& v = g_match_profile.ev[stage];
    size_t& u = g_match_profile.used[stage];
    if (u == v.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
      v.push_back(e);
    }
    cudaEventRecord(v[u++], s);
  }
};


// ------------------------------------------------------------------------------------------------------------
// mutual nearest neighbours from ONE GEMM (column side: claims + verification against the block maxima)
// ------------------------------------------------------------------------------------------------------------
// claim[j] = min over the queries i whose nearest gallery row is j of (distance bits << 32 | global query row): only the best
// claimant of a column can be its mutual partner. `claim` must be initialised to an "unclaimed" value: any word whose upper
// half is >= 0x7F000000 (no real distance has such bits): bytes of 0x7F, or INT64_MAX - positive as a signed integer too, so
// that an all_reduce(MIN) over int64 orders claims like the unsigned atomicMin does.
__global__ void mutual_claim_kernel(const int* __restrict__ i1, const float* __restrict__ d1, const float* __restrict__ d2,
                                    long long nq, long long q_offset, unsigned long long* __restrict__ claim, long long ng,
                                    float* __restrict__ rb_min_d2) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // a warp is one 32-row block: its smallest second-nearest distance bounds every d(k, j) with j not the nearest column of k
  float m = i < nq ? d2[i] : __int_as_float(0x7f800000);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && i < nq) rb_min_d2[i >> 5] = m;
  if (i >= nq) return;
  const long long j = i1[i];
  if (j < 0 || j >= ng) return;
  atomicMin(claim + j, (static_cast<unsigned long long>(__float_as_uint(d1[i])) << 32) | static_cast<unsigned int>(i + q_offset));
}

// Is the claimant of column j really the nearest query of g_j (lowest row wins ties, like torch.min)? A query row k can beat
// it only if BOTH hold: (1) dot(q_k, g_j) >= dot(claimant), and then the GEMM's (32-row block, 8-column chunk) maximum of k's
// cell is at least that dot product minus the error bound of the fp16-operand product; (2) k's own second-nearest distance
// d2[k] <= d(claimant, j): j is not k's nearest column (else k is a claimant itself and lost the atomicMin), so d(k, j) is at
// least k's second-smallest distance. (2) is what makes confident matches free: a claim at distance 0.4 is never challenged
// by rows whose second-nearest gallery row is 1.1 away, whatever their cell maximum says. One warp per chunk: scan the chunk's
// block maxima and the row blocks' minimum d2 (both contiguous), evaluate the exact fp32 distances of the cells passing both
// tests, stop as soon as every claimed column is decided.
// A cell's 32 query rows are staged ONCE in shared memory with coalesced loads (row pitch 132 floats: the per-row reads below
// are bank-conflict free; per-lane global reads at a 512-byte stride cost 32 L1 wavefronts per instruction and made this
// kernel slower than a whole second GEMM) and serve every column of the chunk that needs the cell, two columns per sweep.
// The dot product is summed in the order of rerank_kernel (32 float4 partials, then the xor-butterfly tree), so a distance
// computed here is bit-identical to the one a second (gallery x queries) matching pass would produce.
constexpr int kVerifyWarps = 1;   // one warp per CTA: a chunk with a weakly claimed column evaluates ~20 cells one after the other (~70 us), so CTAs must be small and many (10 per SM by shared memory) for the hardware to balance them
constexpr int kVerifyPitch = 132;   // floats per staged query row
constexpr size_t kVerifySmem = static_cast<size_t>(kVerifyWarps) * (32 * kVerifyPitch * 4 + kChunk * 32 * 16);

__global__ void __launch_bounds__(kVerifyWarps * 32) mutual_verify_kernel(const float* __restrict__ q, long long nq, long long q_offset,
                                                                         const float* __restrict__ g, long long ng,
                                                                         const unsigned long long* __restrict__ claim,
                                                                         const float* __restrict__ block_max,
                                                                         const float* __restrict__ rb_min_d2, int n_row_blocks,
                                                                         float margin_scaled, float inv_dot_scale,
                                                                         unsigned char* __restrict__ beaten) {
  extern __shared__ __align__(16) uint8_t vsm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_q = reinterpret_cast<float*>(vsm) + static_cast<size_t>(w) * 32 * kVerifyPitch;                       // [32][132]
  float4* s_g = reinterpret_cast<float4*>(vsm + static_cast<size_t>(kVerifyWarps) * 32 * kVerifyPitch * 4) + w * kChunk * 32;   // [8][32]
  const long long cb = static_cast<long long>(blockIdx.x) * kVerifyWarps + w;
  if (cb * kChunk >= ng) return;
  unsigned long long cl = 0x7FFFFFFFFFFFFFFFull;   // unclaimed
  if (lane < kChunk && cb * kChunk + lane < ng) cl = claim[cb * kChunk + lane];
  float dcl[kChunk], thr[kChunk];
  unsigned int icl[kChunk];
  unsigned undecided = 0, lost = 0;
  const float ninf = -__int_as_float(0x7f800000);
#pragma unroll
  for (int c = 0; c < kChunk; ++c) {
    const unsigned long long cc = __shfl_sync(0xffffffffu, cl, c);
    dcl[c] = __uint_as_float(static_cast<unsigned int>(cc >> 32));
    icl[c] = static_cast<unsigned int>(cc & 0xffffffffu);
    thr[c] = __int_as_float(0x7f800000);
    if ((cc >> 32) < 0x7F000000ull) {
      undecided |= 1u << c;
      thr[c] = (1.0f - 0.5f * dcl[c] * dcl[c]) * inv_dot_scale - margin_scaled;
      s_g[c * 32 + lane] = __ldg(reinterpret_cast<const float4*>(g + (cb * kChunk + c) * 128) + lane);
    }
  }
  if (!undecided) return;
  __syncwarp();
  const float* bm_row = block_max + cb * n_row_blocks;

  // distance of this lane's staged row to column c (arithmetic of rerank_kernel)
  auto row_distance = [&](int c) {
    float p[32];
    const float* qr = s_q + lane * kVerifyPitch;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const float4 qv = *reinterpret_cast<const float4*>(qr + 4 * l);
      const float4 gv = s_g[c * 32 + l];
      p[l] = fmaf(qv.x, gv.x, fmaf(qv.y, gv.y, fmaf(qv.z, gv.z, qv.w * gv.w)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int l = 0; l < o; ++l) p[l] += p[l + o];
    }
    return sqrtf(fminf(fmaxf(2.0f - 2.0f * p[0], 1e-8f), 4.0f));
  };

  auto evaluate = [&](int rb, float vb, float min_d2) {
    unsigned need = 0;
#pragma unroll
    for (int c = 0; c < kChunk; ++c)
      if (((undecided >> c) & 1u) && thr[c] <= vb && min_d2 <= dcl[c]) need |= 1u << c;
    if (!need) return;
    // stage the cell's 32 rows: one coalesced 512-byte row per iteration
    const long long k0 = static_cast<long long>(rb) * 32;
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 8) {   // eight row loads in flight
      float4 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + r0 + r < nq) v[r] = __ldg(reinterpret_cast<const float4*>(q + (k0 + r0 + r) * 128) + lane);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) *reinterpret_cast<float4*>(s_q + (r0 + r) * kVerifyPitch + 4 * lane) = v[r];
    }
    __syncwarp();
    const long long k = k0 + lane;
    const bool valid = k < nq;
    const unsigned int gk = static_cast<unsigned int>(k + q_offset);
#pragma unroll 1
    for (int c = 0; c < kChunk; ++c) {
      if (!((need >> c) & 1u)) continue;   // warp-uniform
      const float d = row_distance(c);
      const bool beats = valid && gk != icl[c] && (d < dcl[c] || (d == dcl[c] && gk < icl[c]));
      if (__any_sync(0xffffffffu, beats)) {
        undecided &= ~(1u << c);
        lost |= 1u << c;
      }
    }
    __syncwarp();   // the staging tile is reused by the next cell
  };

  for (int base = 0; base < n_row_blocks && undecided; base += 32) {
    const int rb = base + lane;
    const float v = rb < n_row_blocks ? bm_row[rb] : ninf;
    const float md2 = rb < n_row_blocks ? rb_min_d2[rb] : __int_as_float(0x7f800000);
    bool mine = false;
#pragma unroll
    for (int c = 0; c < kChunk; ++c) mine |= ((undecided >> c) & 1u) && thr[c] <= v && md2 <= dcl[c];
    unsigned cand = __ballot_sync(0xffffffffu, mine);
    while (cand && undecided) {
      const int b = __ffs(cand) - 1;
      cand &= cand - 1;
      evaluate(base + b, __shfl_sync(0xffffffffu, v, b), __shfl_sync(0xffffffffu, md2, b));
    }
  }
  if (lane < kChunk && ((lost >> lane) & 1u)) beaten[cb * kChunk + lane] = 1;
}

// mutual[i] = the best claimant of my nearest column is me, and nobody beat it
__global__ void mutual_final_kernel(const int* __restrict__ i1, long long nq, long long q_offset,
                                    const unsigned long long* __restrict__ claim, const unsigned char* __restrict__ beaten,
                                    long long ng, unsigned char* __restrict__ mutual) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const long long j = i1[i];
  bool m = false;
  if (j >= 0 && j < ng) m = static_cast<unsigned int>(claim[j] & 0xffffffffu) == static_cast<unsigned int>(i + q_offset) && !beaten[j];
  mutual[i] = m ? 1 : 0;
}

// Which matching GEMM runs: -1 = by problem size, 0 = single-CTA kernel, 1 = CTA-pair kernel. The initial value comes from
// HN_MATCH_PAIR, read ONCE when first needed; tests switch it through hn_match_force_kernel instead of the environment.
static int& match_kernel_mode() {
  static int mode = [] { const char* pe = getenv("HN_MATCH_PAIR"); return pe ? atoi(pe) : -1; }();
  return mode;
}

static int make_desc_map(CUtensorMap* tm, const uint16_t* base, long long rows, int K, int box_rows = kDistTile) {
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
  const uint64_t str[1] = {static_cast<uint64_t>(K) * 2};
  const uint32_t box[2] = {64, static_cast<uint32_t>(box_rows)};
  return make_tmap_16bit(tm, base, 2, dims, str, box, 128);
}

template <int MB, int EPI>
static int launch_dist(const DistParams& p, int grid_x, int grid_y, cudaStream_t s) {
  auto kern = dist_kernel<MB, EPI>;
  const size_t smem = dist_smem_bytes<MB>(p.k_blocks);
  // the size depends on k_blocks (MB = 1: up to 6 for the 3-way split, MB = 2: always 2): raise the limit once per device
  static DeviceOnce attr_once;
  if (attr_once.first_time())
    HN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dist_smem_bytes<MB>(MB == 1 ? 6 : 2))));
  kern<<<dim3(grid_x, grid_y), 64 + 128 * MB * dist_col_split<EPI>(), smem, s>>>(p);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

struct ExactWs {
  uint16_t *a16, *p16;
  float *na, *np, *pos;
  unsigned long long *row_pack, *col_pack;
  size_t bytes;
};

static ExactWs carve_exact(void* ws, long long Na, long long Np) {
  ExactWs w;
  char* b = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t n) { char* r = b ? b + off : nullptr; off += align256(n); return r; };
  w.a16 = reinterpret_cast<uint16_t*>(take(static_cast<size_t>(Na) * 384 * 2));
  w.p16 = reinterpret_cast<uint16_t*>(take(static_cast<size_t>(Np) * 384 * 2));
  w.na = reinterpret_cast<float*>(take(static_cast<size_t>(Na) * 4));
  w.np = reinterpret_cast<float*>(take(static_cast<size_t>(Np) * 4));
  w.pos = reinterpret_cast<float*>(take(static_cast<size_t>(std::max(Na, Np)) * 4));
  w.row_pack = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(Na) * 8));
  w.col_pack = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(Np) * 8));
  w.bytes = off;
  return w;
}

// Number of gallery segments a block of query rows is split into: work items = row blocks x segments are dealt round robin to
// `workers` persistent CTAs (or CTA pairs), so the count is chosen for the best fill of the last round (64k x 64k on 74 CTA
// pairs: 2 segments = 256 items = 3.46 rounds, i.e. 14 % of the machine idles in the last one; 15 segments = 25.9 rounds).
// Fewer segments win ties (every segment adds a shortlist per row for the re-rank to scan). A problem with fewer items than
// workers (the loss at N = 1024) simply takes the most items it can get: there the fill grows with every extra segment.
static int pick_segments(long long rows, int rows_per_block, long long cols, int workers, int max_seg) {
  const long long m_blocks = (rows + rows_per_block - 1) / rows_per_block;
  const long long n_tiles = (cols + kDistTile - 1) / kDistTile;
  const long long cap = std::max<long long>(1, std::min<long long>(max_seg, n_tiles));   // small problems: as many items as tiles allow
  int best = 1;
  double best_eff = -1.0;
  for (long long sgm = 1; sgm <= cap; ++sgm) {
    const long long items = m_blocks * sgm;
    const long long rounds = (items + workers - 1) / workers;
    const double eff = static_cast<double>(items) / static_cast<double>(rounds * workers);
    if (eff > best_eff + 0.02) { best_eff = eff; best = static_cast<int>(sgm); }
  }
  return best;
}

// Shared body of hn_dist_min / hn_loss_hardnet. Leaves packed minima and pos in the workspace.
static int run_exact(const float* a, const float* p, long long Na, long long Np, int form, int flags, ExactWs& w,
                     cudaStream_t s, const float* a_xy = nullptr, const float* p_xy = nullptr, float nei_c = 0.f) {
  int sm = 0;
  HN_TRY(device_sm_count(&sm));
  const int threads = 256;
  pack_desc_kernel<<<static_cast<unsigned>((Na * 32 + threads - 1) / threads), threads, 0, s>>>(a, Na, w.a16, 1, 0, w.na);
  pack_desc_kernel<<<static_cast<unsigned>((Np * 32 + threads - 1) / threads), threads, 0, s>>>(p, Np, w.p16, 1, 1, w.np);
  HN_CUDA(cudaGetLastError());
  count_launch(2);
  HN_CUDA(cudaMemsetAsync(w.row_pack, 0xff, static_cast<size_t>(Na) * 8, s));
  const bool swap = (flags & HN_FLAG_SWAP) != 0;
  if (swap) HN_CUDA(cudaMemsetAsync(w.col_pack, 0xff, static_cast<size_t>(Np) * 8, s));
  DistParams dp;
  memset(&dp, 0, sizeof(dp));
  HN_TRY(make_desc_map(&dp.side[0].tmA, w.a16, Na, 384));
  HN_TRY(make_desc_map(&dp.side[0].tmB, w.p16, Np, 384));
  dp.side[0].norm_a = w.na;
  dp.side[0].norm_b = w.np;
  dp.side[0].row_pack = w.row_pack;
  dp.side[0].pos = w.pos;
  dp.side[0].Na = Na;
  dp.side[0].Nb = Np;
  // transposed problem: rows = positives, columns = anchors (same packed operands, roles exchanged)
  dp.side[1].tmA = dp.side[0].tmB;
  dp.side[1].tmB = dp.side[0].tmA;
  dp.side[1].norm_a = w.np;
  dp.side[1].norm_b = w.na;
  dp.side[1].row_pack = w.col_pack;
  dp.side[1].pos = nullptr;
  dp.side[1].Na = Np;
  dp.side[1].Nb = Na;
  dp.k_blocks = 6;
  dp.form = form;
  dp.loss_mask = (flags & HN_FLAG_LOSS_MASK) ? 1 : 0;
  dp.nei_mask = (flags & HN_FLAG_NEI_MASK) ? 1 : 0;
  dp.nei_c = nei_c;
  for (int sd = 0; sd < 2; ++sd) {   // the mask is symmetric in (row, column): both directions index the same two arrays
    dp.side[sd].xy_a_rows = dp.side[sd].xy_a_cols = reinterpret_cast<const float2*>(a_xy);
    dp.side[sd].xy_p_rows = dp.side[sd].xy_p_cols = reinterpret_cast<const float2*>(p_xy);
  }
  dp.dot_scale = kDotScale;
  const long long rows = swap ? std::max(Na, Np) : Na;
  const long long cols = swap ? std::min(Na, Np) : Np;
  dp.segments = pick_segments(rows, kDistTile, cols, sm, 16);
  const long long items = ((rows + kDistTile - 1) / kDistTile) * dp.segments;
  const int grid_x = static_cast<int>(std::min<long long>(items, sm));
  if (dp.nei_mask) HN_TRY((launch_dist<1, EPI_EXACT_NEI>(dp, grid_x, swap ? 2 : 1, s)));
  else HN_TRY((launch_dist<1, EPI_EXACT>(dp, grid_x, swap ? 2 : 1, s)));
  return HN_OK;
}

}  // namespace hn

using namespace hn;

extern "C" long long hn_dist_workspace_bytes(long long Na, long long Np, int split) {
  if (Na < 0 || Np < 0) return 0;
  (void)split;
  ExactWs w = carve_exact(nullptr, Na, Np);
  const size_t shortlist = align256(static_cast<size_t>(Na) * 128 * 2) + align256(static_cast<size_t>(Np) * 128 * 2) +
                           2 * align256(static_cast<size_t>(Na) * 32 * kTopC * 4);
  return static_cast<long long>(std::max(w.bytes, shortlist) + 256);
}

extern "C" int hn_dist_min_ex(const float* a, const float* p, long long Na, long long Np, int form, int flags, const float* a_xy,
                              const float* p_xy, float nei_c, float* pos, float* row_min, int32_t* row_arg, float* col_min,
                              int32_t* col_arg, void* workspace, long long workspace_bytes, void* stream) {
  HN_REQUIRE(a && p && workspace, "hn_dist_min: NULL argument");
  HN_REQUIRE(Na >= 1 && Np >= 1, "hn_dist_min: empty input (Na=%lld, Np=%lld)", Na, Np);
  HN_REQUIRE(Na < (1LL << 31) && Np < (1LL << 31), "hn_dist_min: more than 2^31 rows");
  HN_REQUIRE(form == HN_FORM_HARDNET || form == HN_FORM_FDL, "hn_dist_min: unknown distance form %d", form);
  HN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hn_dist_min: workspace must be 256-byte aligned");
  HN_REQUIRE(workspace_bytes >= hn_dist_workspace_bytes(Na, Np, 1), "hn_dist_min: workspace too small");
  if ((col_min || col_arg) && !(flags & HN_FLAG_SWAP)) {
    set_error("hn_dist_min: column outputs need HN_FLAG_SWAP");
    return HN_ERR_INVALID;
  }
  if (flags & HN_FLAG_NEI_MASK) {
    HN_REQUIRE(a_xy && p_xy && Na == Np, "hn_dist_min: the neighbour mask needs both keypoint arrays and a square problem");
    HN_REQUIRE(!(flags & HN_FLAG_LOSS_MASK), "hn_dist_min: HN_FLAG_LOSS_MASK and HN_FLAG_NEI_MASK are different losses");
    HN_REQUIRE((reinterpret_cast<uintptr_t>(a_xy) & 7) == 0 && (reinterpret_cast<uintptr_t>(p_xy) & 7) == 0,
               "hn_dist_min: keypoint arrays must be 8-byte aligned");
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ExactWs w = carve_exact(workspace, Na, Np);
  HN_TRY(run_exact(a, p, Na, Np, form, flags, w, s, a_xy, p_xy, nei_c));
  const int threads = 256;
  if (row_min || row_arg)
    unpack_min_kernel<<<static_cast<unsigned>((Na + threads - 1) / threads), threads, 0, s>>>(w.row_pack, Na, row_min, row_arg);
  if (col_min || col_arg)
    unpack_min_kernel<<<static_cast<unsigned>((Np + threads - 1) / threads), threads, 0, s>>>(w.col_pack, Np, col_min, col_arg);
  if (pos && (flags & (HN_FLAG_LOSS_MASK | HN_FLAG_NEI_MASK)))
    HN_CUDA(cudaMemcpyAsync(pos, w.pos, static_cast<size_t>(std::min(Na, Np)) * 4, cudaMemcpyDeviceToDevice, s));
  HN_CUDA(cudaGetLastError());
  return HN_OK;
}

extern "C" int hn_dist_min(const float* a, const float* p, long long Na, long long Np, int form, int flags, float* pos,
                           float* row_min, int32_t* row_arg, float* col_min, int32_t* col_arg, void* workspace,
                           long long workspace_bytes, void* stream) {
  return hn_dist_min_ex(a, p, Na, Np, form, flags, nullptr, nullptr, 0.f, pos, row_min, row_arg, col_min, col_arg, workspace,
                        workspace_bytes, stream);
}

extern "C" int hn_loss_hardnet(const float* anchor, const float* positive, long long N, float margin, int anchor_swap,
                               float* loss_out, void* workspace, long long workspace_bytes, void* stream) {
  HN_REQUIRE(anchor && positive && loss_out && workspace, "hn_loss_hardnet: NULL argument");
  HN_REQUIRE(N >= 1 && N < (1LL << 31), "hn_loss_hardnet: N out of range (%lld)", N);
  HN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hn_loss_hardnet: workspace must be 256-byte aligned");
  HN_REQUIRE(workspace_bytes >= hn_dist_workspace_bytes(N, N, 1), "hn_loss_hardnet: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ExactWs w = carve_exact(workspace, N, N);
  HN_TRY(run_exact(anchor, positive, N, N, HN_FORM_HARDNET, HN_FLAG_LOSS_MASK | (anchor_swap ? HN_FLAG_SWAP : 0), w, s));
  loss_finalize_kernel<<<1, 1024, 0, s>>>(w.pos, w.row_pack, anchor_swap ? w.col_pack : nullptr, N, margin, loss_out);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" int hn_pack_descriptors(const float* x, long long n, void* out16, void* stream) {
  HN_REQUIRE(x && out16, "hn_pack_descriptors: NULL argument");
  HN_REQUIRE(n >= 1 && n < (1LL << 31), "hn_pack_descriptors: n out of range (%lld)", n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  pack_desc_kernel<<<static_cast<unsigned>((n * 32 + threads - 1) / threads), threads, 0, s>>>(x, n, static_cast<uint16_t*>(out16), 0, 0, nullptr);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" int hn_pack_descriptors_multicast(const float* x, long long n, void* mc16, void* mc32, void* stream) {
  HN_REQUIRE(x && (mc16 || mc32), "hn_pack_descriptors_multicast: NULL argument (either destination may be NULL, not both)");
  HN_REQUIRE(n >= 1 && n < (1LL << 31), "hn_pack_descriptors_multicast: n out of range (%lld)", n);
  HN_REQUIRE((reinterpret_cast<uintptr_t>(mc16) & 15) == 0 && (reinterpret_cast<uintptr_t>(mc32) & 15) == 0,
             "hn_pack_descriptors_multicast: multicast addresses must be 16-byte aligned");
  const int threads = 256;
  pack_desc_multicast_kernel<<<static_cast<unsigned>((n * 16 + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n, static_cast<uint16_t*>(mc16), static_cast<float*>(mc32));
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" int hn_match_ex(const float* q, const float* g, const void* q16_in, const void* g16_in, long long Nq, long long Ng,
                           long long g_offset, float* d1, float* d2, int32_t* i1, int32_t* i2, float* block_max, void* workspace,
                           long long workspace_bytes, void* g_ready_event, void* stream) {
  HN_REQUIRE(q && g && workspace, "hn_match: NULL argument");
  HN_REQUIRE(Nq >= 1 && Ng >= 1, "hn_match: empty input (Nq=%lld, Ng=%lld)", Nq, Ng);
  HN_REQUIRE(Nq < (1LL << 31) && Ng + g_offset < (1LL << 31), "hn_match: more than 2^31 rows");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hn_match: workspace must be 256-byte aligned");
  HN_REQUIRE(workspace_bytes >= hn_dist_workspace_bytes(Nq, Ng, 0), "hn_match: workspace too small");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(q16_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(g16_in) & 15) == 0,
             "hn_match: packed operands must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int sm = 0;
  HN_TRY(device_sm_count(&sm));
  char* b = static_cast<char*>(workspace);
  uint16_t* q16 = reinterpret_cast<uint16_t*>(b);
  uint16_t* g16 = reinterpret_cast<uint16_t*>(b + align256(static_cast<size_t>(Nq) * 256));
  int* cand = reinterpret_cast<int*>(b + align256(static_cast<size_t>(Nq) * 256) + align256(static_cast<size_t>(Ng) * 256));
  float* cand_val = reinterpret_cast<float*>(reinterpret_cast<char*>(cand) + align256(static_cast<size_t>(Nq) * 32 * kTopC * 4));
  const int threads = 256;
  {
    MatchTimer timer(0, s);
    if (q16_in) q16 = const_cast<uint16_t*>(static_cast<const uint16_t*>(q16_in));
    else pack_desc_kernel<<<static_cast<unsigned>((Nq * 32 + threads - 1) / threads), threads, 0, s>>>(q, Nq, q16, 0, 0, nullptr);
    if (g16_in) g16 = const_cast<uint16_t*>(static_cast<const uint16_t*>(g16_in));
    else pack_desc_kernel<<<static_cast<unsigned>((Ng * 32 + threads - 1) / threads), threads, 0, s>>>(g, Ng, g16, 0, 1, nullptr);
    HN_CUDA(cudaGetLastError());
    count_launch((q16_in ? 0 : 1) + (g16_in ? 0 : 1));
  }
  DistParams dp;
  memset(&dp, 0, sizeof(dp));
  // CTA pairs (tc_dist_pair.cuh) once the problem is large enough to fill the machine with 512-row query blocks
  const int pair_mode = match_kernel_mode();
  const bool use_pair = pair_mode >= 0 ? pair_mode != 0 : (Nq >= 8192 && Ng >= 1024);
  HN_TRY(make_desc_map(&dp.side[0].tmA, q16, Nq, 128));
  HN_TRY(make_desc_map(&dp.side[0].tmB, g16, Ng, 128, use_pair ? 64 : kDistTile));
  dp.side[0].cand = cand;
  dp.side[0].cand_val = cand_val;
  dp.side[0].block_max = block_max;
  dp.side[0].n_row_blocks = static_cast<int>((Nq + 31) / 32);
  dp.side[0].Na = Nq;
  dp.side[0].Nb = Ng;
  dp.k_blocks = 2;
  dp.form = HN_FORM_FDL;
  dp.dot_scale = kDotScale;
  {
    MatchTimer timer(1, s);
    if (use_pair) {
      const int pairs = std::max(sm / 2, 1);
      dp.segments = pick_segments(Nq, 4 * kDistTile, Ng, pairs, 16);
      const long long items = ((Nq + 4 * kDistTile - 1) / (4 * kDistTile)) * dp.segments;
      static DeviceOnce attr_once;
      if (attr_once.first_time()) {
        HN_CUDA(cudaFuncSetAttribute(match_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMpSmem)));
      }
      match_pair_kernel<<<2 * static_cast<int>(std::min<long long>(items, pairs)), kMpThreads, kMpSmem, s>>>(dp);
      HN_CUDA(cudaGetLastError());
      count_launch();
    } else {
      dp.segments = pick_segments(Nq, 2 * kDistTile, Ng, sm, 16);
      const long long items = ((Nq + 2 * kDistTile - 1) / (2 * kDistTile)) * dp.segments;
      HN_TRY((launch_dist<2, EPI_SHORTLIST>(dp, static_cast<int>(std::min<long long>(items, sm)), 1, s)));
    }
  }
  // The re-rank reads the fp32 gallery rows of the shortlisted columns; a caller that gathered the packed gallery first
  // (half the bytes) passes the event that marks the fp32 rows' arrival, so that transfer hides behind the GEMM.
  if (g_ready_event) HN_CUDA(cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(g_ready_event), 0));
  // |fp16-operand dot - exact dot| <= 2^-10 for rows of norm <= 1 (L2-normalised descriptors, the only input this path
  // is defined for: FDLNet-master/utils/math_utils.py:15-18 clamps 2 - 2ab to [1e-8, 4]); keep 4x that as the margin.
  const float margin = 4.0f * (1.0f / 1024.0f) / kDotScale;
  {
    MatchTimer timer(2, s);
    // the CTA-pair kernel keeps two top-kTopC lists per row and segment (one per 64-column half of its tiles)
    rerank_kernel<<<static_cast<unsigned>((Nq + 3) / 4), 128, 0, s>>>(q, g, cand, cand_val, margin, Nq, Ng,
                                                                     dp.segments * kTopC * (use_pair ? 2 : 1), g_offset, d1, d2, i1, i2);
    HN_CUDA(cudaGetLastError());
    count_launch();
  }
  return HN_OK;
}

extern "C" int hn_match(const float* q, const float* g, long long Nq, long long Ng, long long g_offset, float* d1,
                        float* d2, int32_t* i1, int32_t* i2, void* workspace, long long workspace_bytes, void* stream) {
  return hn_match_ex(q, g, nullptr, nullptr, Nq, Ng, g_offset, d1, d2, i1, i2, nullptr, workspace, workspace_bytes, nullptr, stream);
}

extern "C" long long hn_block_max_elems(long long Nq, long long Ng) {
  if (Nq < 0 || Ng < 0) return 0;
  return ((Ng + kChunk - 1) / kChunk) * ((Nq + 31) / 32);
}

extern "C" int hn_mutual_claims(const int32_t* i1, const float* d1, const float* d2, long long Nq, long long q_offset,
                                unsigned long long* claim, long long Ng, float* rb_min_d2, void* stream) {
  HN_REQUIRE(i1 && d1 && d2 && claim && rb_min_d2, "hn_mutual_claims: NULL argument");
  HN_REQUIRE(Nq >= 1 && Ng >= 1 && Nq + q_offset < (1LL << 32), "hn_mutual_claims: size out of range");
  mutual_claim_kernel<<<static_cast<unsigned>((Nq + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(i1, d1, d2, Nq, q_offset, claim, Ng, rb_min_d2);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" int hn_mutual_verify(const float* q, long long Nq, long long q_offset, const float* g, long long Ng,
                                const unsigned long long* claim, const float* block_max, const float* rb_min_d2,
                                unsigned char* beaten, void* stream) {
  HN_REQUIRE(q && g && claim && block_max && rb_min_d2 && beaten, "hn_mutual_verify: NULL argument");
  HN_REQUIRE(Nq >= 1 && Ng >= 1 && Nq + q_offset < (1LL << 32), "hn_mutual_verify: size out of range");
  const long long chunks = (Ng + kChunk - 1) / kChunk;
  // twice the 2^-10 error bound of the fp16-operand dot product of unit vectors, in the GEMM's scaled units
  const float margin = 2.0f * (1.0f / 1024.0f) / kDotScale;
  static DeviceOnce attr_once;
  if (attr_once.first_time())
    HN_CUDA(cudaFuncSetAttribute(mutual_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kVerifySmem)));
  mutual_verify_kernel<<<static_cast<unsigned>((chunks + kVerifyWarps - 1) / kVerifyWarps), kVerifyWarps * 32, kVerifySmem, static_cast<cudaStream_t>(stream)>>>(
      q, Nq, q_offset, g, Ng, claim, block_max, rb_min_d2, static_cast<int>((Nq + 31) / 32), margin, 1.0f / kDotScale, beaten);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" long long hn_mutual_workspace_bytes(long long Nq, long long Ng) {
  if (Nq < 0 || Ng < 0) return 0;
  return hn_dist_workspace_bytes(Nq, Ng, 0) + static_cast<long long>(align256(static_cast<size_t>(hn_block_max_elems(Nq, Ng)) * 4) +
                                                                         align256(static_cast<size_t>(Ng) * 8) + align256(static_cast<size_t>(Ng)) +
                                                                         align256(static_cast<size_t>((Nq + 31) / 32) * 4) + 256);
}

extern "C" int hn_match_mutual(const float* q, const float* g, const void* q16, const void* g16, long long Nq, long long Ng,
                               float* d1, float* d2, int32_t* i1, int32_t* i2, unsigned char* mutual, void* workspace,
                               long long workspace_bytes, void* stream) {
  HN_REQUIRE(q && g && d1 && d2 && i1 && mutual && workspace, "hn_match_mutual: NULL argument (d1, d2, i1 and mutual are required outputs)");
  HN_REQUIRE(workspace_bytes >= hn_mutual_workspace_bytes(Nq, Ng), "hn_match_mutual: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long base = hn_dist_workspace_bytes(Nq, Ng, 0);
  char* b = static_cast<char*>(workspace) + align256(static_cast<size_t>(base));
  float* bm = reinterpret_cast<float*>(b);
  b += align256(static_cast<size_t>(hn_block_max_elems(Nq, Ng)) * 4);
  unsigned long long* claim = reinterpret_cast<unsigned long long*>(b);
  b += align256(static_cast<size_t>(Ng) * 8);
  unsigned char* beaten = reinterpret_cast<unsigned char*>(b);
  b += align256(static_cast<size_t>(Ng));
  float* rbmin = reinterpret_cast<float*>(b);
  HN_TRY(hn_match_ex(q, g, q16, g16, Nq, Ng, 0, d1, d2, i1, i2, bm, workspace, base, nullptr, stream));
  HN_CUDA(cudaMemsetAsync(claim, 0x7f, static_cast<size_t>(Ng) * 8, s));   // "unclaimed"
  HN_CUDA(cudaMemsetAsync(beaten, 0, static_cast<size_t>(Ng), s));
  HN_TRY(hn_mutual_claims(i1, d1, d2, Nq, 0, claim, Ng, rbmin, stream));
  HN_TRY(hn_mutual_verify(q, Nq, 0, g, Ng, claim, bm, rbmin, beaten, stream));
  mutual_final_kernel<<<static_cast<unsigned>((Nq + 255) / 256), 256, 0, s>>>(i1, Nq, 0, claim, beaten, Ng, mutual);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" int hn_match_force_kernel(int mode) {
  HN_REQUIRE(mode >= -1 && mode <= 1, "hn_match_force_kernel: mode must be -1 (by size), 0 (single CTA) or 1 (CTA pair)");
  match_kernel_mode() = mode;
  return HN_OK;
}

extern "C" int hn_match_profile_enable(int on) {
  g_match_profile.on = on != 0;
  for (size_t& u : g_match_profile.used) u = 0;
  return HN_OK;
}

extern "C" int hn_match_profile_read(double ms_out[3], long long launches_out[3]) {
  HN_REQUIRE(ms_out && launches_out, "hn_match_profile_read: NULL argument");
  for (int st = 0; st < 3; ++st) {
    double total = 0.0;
    const size_t pairs = g_match_profile.used[st] / 2;
    for (size_t i = 0; i < pairs; ++i) {
      HN_CUDA(cudaEventSynchronize(g_match_profile.ev[st][2 * i + 1]));
      float ms = 0.f;
      HN_CUDA(cudaEventElapsedTime(&ms, g_match_profile.ev[st][2 * i], g_match_profile.ev[st][2 * i + 1]));
      total += ms;
    }
    ms_out[st] = total;
    launches_out[st] = static_cast<long long>(pairs);
    g_match_profile.used[st] = 0;
  }
  return HN_OK;
}
