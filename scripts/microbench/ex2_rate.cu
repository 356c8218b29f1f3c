// MUFU ex2 throughput microbenchmark: results per clock per SM for ex2.approx.ftz.f32, ex2.approx.ftz.f16x2 and
// ex2.approx.ftz.bf16x2 (two results per lane-instruction when packed), 8 warps per SM sub-partition, ILP 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ex2_rate scripts/microbench/ex2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, float seed, long long* cyc, float* sink) {
  float f[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] = seed * (float)(threadIdx.x + i) * 1e-3f - 1.0f;
    h[i] = 0xBC00BC00u + (uint32_t)(threadIdx.x + i);
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(h[i]);
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  long long* d; float* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  const int iters = 4000;
  k<MODE><<<148, 1024>>>(10, 1.f, d, s);
  k<MODE><<<148, 1024>>>(iters, 1.f, d, s);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double results = (double)iters * 8 * 1024 * per_instr;
  printf("%-28s %6.2f results / clk / SM   (%lld cycles, %s)\n", name, results / (double)c, c, cudaGetErrorString(e));
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  return 0;
}
