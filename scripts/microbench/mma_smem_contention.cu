// Does LSU shared-memory traffic slow the tensor core's operand fetch?  One CTA per SM: warp 1 issues SS 128x128x16 MMAs (8 KB of
// operands per 64-cycle MMA = 128 B/clk, the fetch limit), warps 2..9 meanwhile stream st.shared.v4 / ld.shared.v4 over a separate
// 32 KB region at a controllable duty.  Prints cycles per MMA with 0 / 4 / 8 traffic warps, and the traffic achieved.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I alg_b200/csrc -o /tmp/mma_smem scripts/microbench/mma_smem_contention.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace alg::tc;
namespace alg { void set_error(const std::string&) {} std::atomic<int64_t> g_launches{0}; }

template <int TS>
__global__ void __launch_bounds__(320, 1) k(int iters, int traffic_warps, int store, long long* out, unsigned long long* bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); done = 0; }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a_lo = smem_desc_lo_sw128(smem_u32(smem)), b_lo = smem_desc_lo_sw128(smem_u32(smem + 32768));
      const uint32_t idesc = make_idesc_bf16(128, 128);
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
      done = 1;
    }
  } else if (warp >= 2 && warp < 2 + traffic_warps) {
    uint4* region = reinterpret_cast<uint4*>(smem + 65536) + (warp - 2) * 256;  // 4 KB per warp
    uint4 v = make_uint4(lane, warp, 0, 0);
    unsigned long long n = 0;
    while (!done) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (store) region[i * 32 + lane] = v;
        else { uint4 r = region[i * 32 + lane]; v.x += r.x; }
      }
      n += 8 * 512;
    }
    if (lane == 0 && blockIdx.x == 0) atomicAdd(bytes, n + (v.x == 0xdeadbeef));
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int TS>
void run(const char* name, int tw, int store) {
  long long* d; unsigned long long* b;
  cudaMalloc(&d, 8); cudaMalloc(&b, 8); cudaMemset(b, 0, 8);
  const int iters = 2000, smem = 65536 + 32768 + 1024;
  cudaFuncSetAttribute(k<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<TS><<<148, 320, smem>>>(iters, tw, store, d, b);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; unsigned long long by = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&by, b, 8, cudaMemcpyDeviceToHost);
  printf("%-18s traffic warps %d (%s): %6.1f cycles / MMA (floor 64), LSU traffic %6.1f B/clk/SM  (%s)\n", name, tw, store ? "st.shared" : "ld.shared",
         (double)c / (iters * 8.0), (double)by / (double)c, cudaGetErrorString(e));
}
int main() {
  for (int ts = 0; ts < 2; ++ts)
    for (int store = 1; store >= 0; --store)
      for (int tw : {0, 2, 4, 8}) {
        if (ts) run<1>("TS 128x128x16", tw, store);
        else run<0>("SS 128x128x16", tw, store);
      }
  return 0;
}
