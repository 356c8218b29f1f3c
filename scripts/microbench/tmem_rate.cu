// tcgen05.ld / tcgen05.st throughput microbenchmark: bytes per clock per SM moved between TMEM and registers by 4 or 8 warps
// (one or two warps per 32-lane quadrant), alone and while another warp keeps the tensor pipe busy with 128x128x16 TS MMAs
// (A and D in TMEM, B in shared memory).  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I alg_b200/csrc -o /tmp/tmem_rate scripts/microbench/tmem_rate.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace alg::tc;
namespace alg { void set_error(const std::string&) {} std::atomic<int64_t> g_launches{0}; }

// MODE 0: ld x32 (two per iteration = 64 columns), 1: st x32 x 2, 2: ld 64 columns + st 32 columns (the softmax's mix)
template <int MODE, int WARPS, int MMA>
__global__ void __launch_bounds__(WARPS * 32 + 32, 1) k(int iters, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == WARPS) {  // optional MMA stream: D = columns [256, 384), A = [384, 448)
    if (MMA && elect_one()) {
      const uint32_t b_lo = smem_desc_lo_sw128(smem_u32(smem));
      const uint32_t idesc = make_idesc_bf16(128, 128);
      long long n = 0;
      const long long t0 = clock64();
      while (!stop) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_ts_lo(tm + 256, tm + 384 + (ks & 3) * 8, b_lo + ((ks * 32) >> 4), idesc, 1);
        n += 8;
        if ((n & 63) == 0) {  // bound the queue depth: wait for the batch
          tc_commit(&bar);
          mbar_wait(&bar, (uint32_t)((n >> 6) - 1) & 1);
        }
      }
      const long long t1 = clock64();
      tc_commit(&bar);  // drain what is still queued before TMEM goes away
      mbar_wait(&bar, (uint32_t)(n >> 6) & 1);
      if (blockIdx.x == 0) { out[2] = n; out[3] = t1 - t0; }
    }
  } else {
    const uint32_t t = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;  // second warp of a quadrant: columns 64-127
    uint32_t r[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0 || MODE == 2) {
        tmem_ld32(t, r);
        tmem_ld32(t + 32, r + 32);
        tmem_ld_wait();
      }
      if (MODE == 1) {
        tmem_st32(t, r);
        tmem_st32(t + 32, r + 32);
        tmem_st_wait();
      }
      if (MODE == 2) {
        tmem_st32(t, r + 16);
        tmem_st_wait();
      }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= r[i];
    if (acc == 0xdeadbeefu) sink[0] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; }
    __syncwarp();
    if (warp == 0 && (threadIdx.x & 31) == 0) stop = 1;  // warp 0 finishing ends the MMA stream (all warps run the same loop)
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE, int WARPS, int MMA>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  uint32_t* sink; cudaMalloc(&sink, 4);
  const int iters = 4000, smem = 32768 + 1024;
  cudaFuncSetAttribute(k<MODE, WARPS, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE, WARPS, MMA><<<148, WARPS * 32 + 32, smem>>>(10, d, sink);
  k<MODE, WARPS, MMA><<<148, WARPS * 32 + 32, smem>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[4] = {0, 0, 0, 0}; cudaMemcpy(c, d, 32, cudaMemcpyDeviceToHost);
  const double bytes_per_it = (MODE == 2 ? 96.0 : 64.0) * 4 * 32 * WARPS;  // columns x 4 B x lanes x warps
  printf("%-44s %8.1f B/clk/SM  (%6.1f clk per iteration)", name, bytes_per_it * iters / (double)c[0], (double)c[0] / iters);
  if (MMA) printf("   MMA stream: %.1f clk per 128x128x16 (floor 64)", (double)c[3] / (double)c[2]);
  printf("  (%s)\n", cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}
int main() {
  run<0, 4, 0>("ld 64 cols, 4 warps");
  run<0, 8, 0>("ld 64 cols, 8 warps");
  run<1, 4, 0>("st 64 cols, 4 warps");
  run<1, 8, 0>("st 64 cols, 8 warps");
  run<2, 4, 0>("ld 64 + st 32 cols, 4 warps");
  run<2, 8, 0>("ld 64 + st 32 cols, 8 warps");
  run<0, 4, 1>("ld 64 cols, 4 warps + MMA");
  run<0, 8, 1>("ld 64 cols, 8 warps + MMA");
  run<2, 8, 1>("ld 64 + st 32 cols, 8 warps + MMA");
  return 0;
}
