// Micro-benchmark (diagnostic, not part of the library): what the ISSUING warp pays per tile iteration for the pieces of
// a UMMA issue loop - tcgen05.mma (N = 32, K = 16), tcgen05.commit, tcgen05.fence::after_thread_sync, elect + syncwarp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe.bin tools/issue_probe.cu -I hardnetnas_b200/csrc
#include <cstdio>

#include "common.cuh"
#include "tc_conv.cuh"

using namespace hn;

template <int PATTERN>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 65536, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar, 1);
      mbar_init(bar2, 1 << 20);   // sink for the per-iteration commits (never completes)
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(128, 32, 0);
    const uint32_t a_lo = noswizzle_desc_lo(base, 2048), b_lo = noswizzle_desc_lo(base + 32768, 128);
    constexpr uint32_t A_HI = noswizzle_desc_hi(128), B_HI = noswizzle_desc_hi(512);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (PATTERN == 2) tc_fence_after();
      if (elect_one()) {
        umma_f16_w(tmem + (i & 3) * 32, a_lo, A_HI, b_lo, B_HI, idesc, 0u);
        umma_f16_w(tmem + (i & 3) * 32, a_lo + 16, A_HI, b_lo + 16, B_HI, idesc, 1u);
        if (PATTERN == 1 || PATTERN == 4) umma_commit(bar2);
        if (PATTERN == 4) umma_commit(bar2);
      }
      __syncwarp();
      if (PATTERN == 3) {            // a (satisfied) mbarrier wait per iteration, like the real loop
        mbar_wait(bar, 1);
        tc_fence_after();
      }
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int P>
static void run(const char* name, long long* d) {
  const size_t smem = 65536 + 1024 + 64;
  cudaFuncSetAttribute(probe<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 1024;
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    probe<P><<<148, 128, smem>>>(iters, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return; }
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s issue %.1f clk/iter, complete %.1f clk/iter\n", name, double(h[0]) / iters, double(h[1]) / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<0>("2 MMA (N=32)", d);
  run<1>("2 MMA + commit", d);
  run<4>("2 MMA + 2 commits", d);
  run<2>("fence::after_thread_sync + 2 MMA", d);
  run<3>("2 MMA + satisfied mbarrier wait + fence", d);
  return 0;
}
