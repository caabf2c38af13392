// Micro-benchmark (diagnostic, not part of the library): cost of back-to-back tcgen05.mma.cta_group::2 kind::f16 instructions
// (M = 256 over a CTA pair, K = 16, SS mode) as a function of N, next to the single-CTA instruction (M = 128) of umma_probe.cu.
// Question: does a pair halve the B-operand shared-memory reads per SM (each CTA holds N / 2 rows of B)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_pair_probe.bin tools/umma_pair_probe.cu -I hardnetnas_b200/csrc
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_conv_pair.cuh"

using namespace hn;

// A: no-swizzle K-major (plane pitch 2048, as the conv kernels), B: 128B-swizzled K-major rows (as the resident weights)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair(int n, int adv, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 65536, bar = base + 131072, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_pair(slot, 512);
    tmem_relinquish_pair();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_f16(256, n, 0);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint32_t a_lo[8], b_lo[8];
#pragma unroll
      for (int off = 0; off < 8; ++off) {
        const uint32_t o = adv ? off : 0;
        a_lo[off] = noswizzle_desc_lo(a_base + o * 4096, 2048);
        b_lo[off] = kmajor_desc_lo(b_base + (o >> 2) * 16384) + 2u * (o & 3);
      }
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_f16_pair_w(tmem + (j & 1) * (n <= 256 ? n : 0), a_lo[j], noswizzle_desc_hi(128), b_lo[j], kmajor_desc_hi(128), idesc, 1u);
      }
      umma_commit_pair(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const size_t smem = 131072 + 1024 + 64;
  cudaFuncSetAttribute(probe_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2048;
  const int ns[] = {32, 64, 96, 128, 192, 256};
  printf("cycles per MMA (cta_group::2, M=256, K=16), %d back-to-back instructions, 74 pairs busy\n", iters);
  for (int adv = 0; adv < 2; ++adv) {
    printf("pair adv=%d :", adv);
    for (int n : ns) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        probe_pair<<<148, 128, smem>>>(n, adv, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("   n=%d %.1f", n, double(h) / iters);
    }
    printf("\n");
  }
  return 0;
}
