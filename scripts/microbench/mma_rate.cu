// tcgen05.mma issue-rate microbenchmark: cycles per kind::f16 MMA (K = 16) for SS / TS operand sources and N in {64, 128, 256},
// one CTA per SM, operands resident in shared memory (SWIZZLE_128B K-major tiles; contents irrelevant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I alg_b200/csrc -o /tmp/mma_rate scripts/microbench/mma_rate.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace alg::tc;
namespace alg { void set_error(const std::string&) {} std::atomic<int64_t> g_launches{0}; }

template <int N, int TS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t a_lo = smem_desc_lo_sw128(smem_u32(smem)), b_lo = smem_desc_lo_sw128(smem_u32(smem + 32768));
    const uint32_t idesc = make_idesc_bf16(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t off = ((ks >> 2) * 16384 + (ks & 3) * 32) >> 4;
        if (TS) mma_ts_lo(tm, tm + 256 + ks * 8, b_lo + off, idesc, 1);
        else mma_ss_lo(tm, a_lo + off, b_lo + off, idesc, 1);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int TS>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 2000, smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS><<<148, 128, smem>>>(10, d);
  k<N, TS><<<148, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (iters * 8.0);
  printf("%-28s %7.1f cycles / MMA   floor %3d   (%s)\n", name, per, 128 * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<64, 0>("SS 128x64x16");
  run<128, 0>("SS 128x128x16");
  run<256, 0>("SS 128x256x16");
  run<64, 1>("TS 128x64x16 (A in TMEM)");
  run<128, 1>("TS 128x128x16 (A in TMEM)");
  run<256, 1>("TS 128x256x16 (A in TMEM)");
  return 0;
}
