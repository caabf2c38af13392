// Evaluation metrics of the reference's test loop on the device, behind the C ABI:
//   hn_pair_distances  torch.sqrt(torch.sum((out_a - out_p) ** 2, 1))               hardnet/HardNet.py:458
//   hn_fpr95           ErrorRateAt95Recall(labels, scores)                           hardnet/EvalMetrics.py:6-19
// The reference sorts all N distances on the host to find one threshold. Here the threshold is SELECTED: the k-th smallest
// positive distance (k = ceil(0.95 * #positives), the first index where the running count of positives reaches the recall
// point) by a 4 x 8-bit radix select over the float bit patterns, then the negatives below it are counted. One CTA, no
// workspace, no sort; the counts (FP, TN) come back as integers so the caller forms FP / (FP + TN) exactly as the reference.
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_common.h"

namespace hn {

__global__ void __launch_bounds__(256) pair_distance_kernel(const float* __restrict__ a, const float* __restrict__ p, long long n,
                                                            float* __restrict__ out) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float4 x = reinterpret_cast<const float4*>(a + row * 128)[lane];
  const float4 y = reinterpret_cast<const float4*>(p + row * 128)[lane];
  const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
  float s = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = sqrtf(s);
}

// distances = 1.0 / (scores + 1e-8) in fp32 like the reference's numpy expression (EvalMetrics.py:7); positive floats order
// like their bit patterns
__device__ __forceinline__ uint32_t fpr_key(float score) { return __float_as_uint(__fdiv_rn(1.0f, score + 1e-8f)); }

constexpr int kFprThreads = 1024;

__global__ void __launch_bounds__(kFprThreads, 1) fpr95_kernel(const float* __restrict__ scores, const unsigned char* __restrict__ labels,
                                                               long long n, long long* __restrict__ out /*FP, TN, P, threshold_index*/) {
  __shared__ unsigned long long s_hist[256];
  __shared__ unsigned long long s_cnt[4];
  __shared__ unsigned int s_warp[32];
  __shared__ long long s_found;
  const int tid = threadIdx.x;
  // ---- number of positives / negatives ----
  if (tid < 4) s_cnt[tid] = 0;
  __syncthreads();
  {
    unsigned long long pos = 0;
    for (long long i = tid; i < n; i += kFprThreads) pos += labels[i] != 0;
    atomicAdd(&s_cnt[0], pos);
  }
  __syncthreads();
  const long long P = static_cast<long long>(s_cnt[0]);
  const long long N_neg = n - P;
  // first index where cumsum(labels) >= 0.95 * sum(labels): the k-th positive in sorted order (k = 0: index 0)
  const long long k = static_cast<long long>(ceil(0.95 * static_cast<double>(P)));
  if (k <= 0) {
    if (tid == 0) { out[0] = 0; out[1] = N_neg; out[2] = P; out[3] = 0; }
    return;
  }
  // ---- radix select of the k-th smallest positive key, most significant byte first ----
  uint32_t prefix = 0, mask = 0;
  long long remaining = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) s_hist[tid] = 0;
    __syncthreads();
    for (long long i = tid; i < n; i += kFprThreads) {
      if (labels[i] == 0) continue;
      const uint32_t key = fpr_key(scores[i]);
      if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xffu], 1ull);
    }
    __syncthreads();
    // every thread walks the 256 bins (identical result, no extra synchronisation)
    unsigned long long cum = 0;
    int bin = 0;
    for (; bin < 256; ++bin) {
      if (cum + s_hist[bin] >= static_cast<unsigned long long>(remaining)) break;
      cum += s_hist[bin];
    }
    remaining -= static_cast<long long>(cum);
    prefix |= static_cast<uint32_t>(bin) << shift;
    mask |= 0xffu << shift;
    __syncthreads();
  }
  const uint32_t t = prefix;   // key of the k-th positive; `remaining` = its rank among the positives with exactly this key
  // ---- the `remaining`-th positive with key == t in index order (a stable sort keeps equal keys in input order) ----
  if (tid == 0) s_found = -1;
  __syncthreads();
  long long seen = 0, idx_star = n;
  for (long long base = 0; base < n && idx_star == n; base += kFprThreads) {
    const long long i = base + tid;
    const bool hit = i < n && labels[i] != 0 && fpr_key(scores[i]) == t;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int w = 0; w < kFprThreads / 32; ++w) {
      if (w < (tid >> 5)) before += s_warp[w];
      total += s_warp[w];
    }
    const unsigned rank_in_chunk = before + __popc(bal & ((1u << (tid & 31)) - 1u));
    if (hit && seen + rank_in_chunk + 1 == remaining) s_found = i;
    __syncthreads();
    if (s_found >= 0) idx_star = s_found;
    seen += total;
    __syncthreads();
  }
  // ---- negatives sorted before that element = false positives; also its position in the sorted order ----
  unsigned long long fp = 0, before_all = 0;
  for (long long i = tid; i < n; i += kFprThreads) {
    const uint32_t key = fpr_key(scores[i]);
    const bool earlier = key < t || (key == t && i < idx_star);
    before_all += earlier;
    fp += earlier && labels[i] == 0;
  }
  atomicAdd(&s_cnt[1], fp);
  atomicAdd(&s_cnt[2], before_all);
  __syncthreads();
  if (tid == 0) {
    out[0] = static_cast<long long>(s_cnt[1]);
    out[1] = N_neg - static_cast<long long>(s_cnt[1]);
    out[2] = P;
    out[3] = static_cast<long long>(s_cnt[2]);
  }
}

}  // namespace hn

using namespace hn;

extern "C" int hn_pair_distances(const float* a, const float* p, long long n, float* out, void* stream) {
  HN_REQUIRE(n >= 0, "hn_pair_distances: negative count");
  if (n == 0) return HN_OK;
  HN_REQUIRE(a && p && out, "hn_pair_distances: NULL argument");
  const int threads = 256;
  pair_distance_kernel<<<static_cast<unsigned>((n * 32 + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(a, p, n, out);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}

extern "C" int hn_fpr95(const float* scores, const unsigned char* labels, long long n, long long* out4, void* stream) {
  HN_REQUIRE(scores && labels && out4, "hn_fpr95: NULL argument");
  HN_REQUIRE(n >= 1, "hn_fpr95: empty input");
  fpr95_kernel<<<1, kFprThreads, 0, static_cast<cudaStream_t>(stream)>>>(scores, labels, n, out4);
  HN_CUDA(cudaGetLastError());
  count_launch();
  return HN_OK;
}
