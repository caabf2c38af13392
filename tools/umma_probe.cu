// Micro-benchmark (diagnostic, not part of the library): cost of back-to-back tcgen05.mma kind::f16 instructions
// (M = 128, K = 16, SS mode) as a function of N, operand layout and accumulator reuse. One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/umma_probe tools/umma_probe.cu -I hardnetnas_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace hn;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= 1ull << 46;
  return d;
}

// mode 0: no-swizzle A (plane pitch 2048) + no-swizzle B; mode 1: 128B-swizzled A and B
// accs: number of accumulators cycled through (1 = every MMA accumulates into the same TMEM columns)
// adv: 1 = A descriptor advances every instruction (like a real k-loop), 0 = same operands every time
__global__ void __launch_bounds__(128, 1) probe(int n, int mode, int accs, int adv, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;                 // 64 KB of A operand space
  const uint32_t b_base = base + 65536;         // 64 KB of B operand space
  const uint32_t bar = base + 131072;
  const uint32_t slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) {
      mbar_init(bar, 1);
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
    const uint32_t idesc = make_idesc_f16(128, n, 0);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int off = 0; off < 8; ++off) {
        const uint32_t o = adv ? off : 0;
        if (mode == 0) {
          ad[off] = desc_noswz(a_base + o * 4096, 2048, 128);
          bd[off] = desc_noswz(b_base + o * 8192, 128, 256);
        } else {
          ad[off] = make_kmajor_desc(a_base + (o >> 2) * 16384, 128) + 2u * (o & 3);
          bd[off] = make_kmajor_desc(b_base + (o >> 2) * 32768, 128) + 2u * (o & 3);
        }
      }
      const uint32_t acc_step = (accs == 2) ? n : 0;
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) umma_f16(tmem + (j & 1) * acc_step, ad[j], bd[j], idesc, 1u);
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const size_t smem = 131072 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2048;
  const int ns[] = {32, 64, 96, 128, 192, 256};
  printf("cycles per MMA (M=128, K=16), %d back-to-back instructions, all 148 SMs busy\n", iters);
  for (int mode = 0; mode < 2; ++mode)
    for (int adv = 0; adv < 2; ++adv)
      for (int accs = 1; accs <= 2; ++accs) {
        printf("%s adv=%d accs=%d :", mode ? "swizzle128" : "noswizzle ", adv, accs);
        for (int n : ns) {
          if (accs * n > 512) { printf("   n=%d -", n); continue; }
          long long h = 0;
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<148, 128, smem>>>(n, mode, accs, adv, iters, d);
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
