// The fused front kernel (input_norm + conv1 + conv2, front_fused.cuh) and the fused conv3 + conv4 kernel (tc_conv34.cuh) as
// the two ROLES of one persistent launch: the first `nf` CTAs produce conv2 outputs, the other CTAs (whole pairs) consume them.
//
// Why: the 64 KB/patch conv2 output is the largest tensor of the stack (half of its HBM traffic: written by one kernel, read by
// the next, 1.2 GB per 18 944-patch pass, i.e. far more than the 126 MB L2). When producer and consumer run at the same time a
// patch is read a few microseconds after it was written, out of L2. Every CTA of the launch is resident (grid = #SMs, one CTA
// per SM), so the consumer may spin on a flag: each of the eight conv2 epilogue warps of the producer stamps its flag of the
// patch after a __threadfence (release, gpu scope); the consumer's TMA producer warp polls the eight flags (acquire) before the
// patch's first TMA load, clears them, and hands "ready" to the second producer warp through a shared-memory mbarrier.
// The producer never waits for the consumer (every patch has its own place in global memory), so the flags cannot deadlock.
//
// MEASURED (round 2, 151 552 patches): bit-identical to the two launches, but SLOWER - 65 ns/patch for the three layers against
// 26.7 + 26.4 ns as two launches. The release fence costs each conv2 epilogue warp ~0.9 us per patch (it waits for the warp's
// outstanding stores to reach L2) and those warps have no slack; with the fence removed (diagnostic build, not correct) the launch
// reaches 52.4 ns, i.e. the whole upside of reading the conv2 output out of L2 is ~1.4 % of the forward (the kernels are not
// HBM-bound; only the power budget benefits). Kept as an opt-in (HN_COSCHED=1) and as the record of that experiment.
#pragma once

#include "front_fused.cuh"
#include "tc_conv34.cuh"

namespace hn {

static_assert(kFfThreads == kC34SThreads, "both roles run with the same CTA size");
constexpr size_t kFrontC34Smem = kFfSmem > C34SCfg::SMEM ? kFfSmem : C34SCfg::SMEM;

struct FrontC34Params {
  const void* in;        // raw patches (fp32 or uint8)
  uint16_t* conv2_out;   // parity-planar conv2 output of the pass (what c34.tmA[] point at)
  const float* w1;
  const float* bias1;
  const uint4* w2img;
  FfBias bias2;
  float norm_eps;
  int n_patches;
  int act_bf16;
  int nf;                // CTAs of the front role (even: roles do not share a CTA pair)
  int* ready;            // [chunk][8] flags, all zero between launches
  Conv34Params c34;
};

template <typename TIn>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFfThreads, 1) front_c34_kernel(const __grid_constant__ FrontC34Params P) {
  if (static_cast<int>(blockIdx.x) < P.nf) {
    front_fused_body<TIn, false, 0>(static_cast<const TIn*>(P.in), P.conv2_out, P.w1, P.bias1, P.w2img, P.bias2, 1, P.n_patches, P.act_bf16,
                                    P.norm_eps, P.c34.tmB3 /*unused by this variant*/, nullptr, nullptr, 0, 0, static_cast<int>(blockIdx.x), P.nf,
                                    P.ready);
  } else {
    conv34_stack_body<0, 2>(P.c34, static_cast<int>(blockIdx.x) - P.nf, static_cast<int>(gridDim.x) - P.nf, P.ready);
  }
}

}  // namespace hn
