// Probe (diagnostic): semantics of tcgen05.shift on sm_100a. Fills TMEM [128 lanes x 16 columns] with lane * 100 + column,
// issues one tcgen05.shift.cta_group::1.down on a column range, and prints which rows / columns moved.
#include <cstdio>
#include "common.cuh"
using namespace hn;

__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

__global__ void __launch_bounds__(128, 1) probe(int shift_col, int nshift, unsigned* out) {
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&slot), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t row = tm + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 32; ++c) tmem_st1(row + c, threadIdx.x * 100 + c);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nshift; ++i)
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tm + shift_col) : "memory");
    umma_commit(smem_u32(&bar));
  }
  __syncthreads();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c = 0; c < 32; ++c) {
    const uint32_t v = tmem_ld1(row + c);
    tmem_ld_wait();
    out[threadIdx.x * 32 + c] = v;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 32); }
}

int main() {
  unsigned* d; cudaMalloc(&d, 128 * 32 * 4);
  static unsigned h[128 * 32];
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int col = cfg == 2 ? 8 : 0, n = cfg == 1 ? 2 : 1;
    probe<<<1, 128>>>(col, n, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== shift at column %d, %d time(s): value = srcLane*100 + col\n", col, n);
    for (int r : {0, 1, 2, 3, 30, 31, 32, 33, 63, 64, 65, 126, 127}) {
      printf("lane %3d:", r);
      for (int c = 0; c < 20; ++c) printf(" %5u", h[r * 32 + c]);
      printf("\n");
    }
  }
  return 0;
}
