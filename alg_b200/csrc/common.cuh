// Shared host/device helpers for libalg_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/alg_b200.h"

namespace alg {

void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define ALG_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      alg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                     std::to_string(__LINE__) + ")");                                          \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define ALG_REQUIRE(cond, msg)                        \
  do {                                                \
    if (!(cond)) {                                    \
      alg::set_error(std::string("alg_b200: ") + msg); \
      return 2;                                       \
    }                                                 \
  } while (0)

#define ALG_LAUNCH_OK()                          \
  do {                                           \
    alg::count_launch();                         \
    ALG_CUDA_OK(cudaGetLastError());             \
  } while (0)

inline size_t dtype_size(int dt) { return dt == ALG_F32 ? 4 : 2; }

// ---- device-side dtype helpers ---------------------------------------------------------------
template <int DT> struct Elem;
template <> struct Elem<ALG_F32> {
  using type = float;
  __device__ static float load(const void* p, size_t i) { return reinterpret_cast<const float*>(p)[i]; }
  __device__ static void store(void* p, size_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
  __device__ static float round(float v) { return v; }
  __device__ static void store4(void* p, size_t i, const float* v) {  // i % 4 == 0, p 16-byte aligned
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Elem<ALG_BF16> {
  using type = __nv_bfloat16;
  __device__ static float load(const void* p, size_t i) {
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  }
  __device__ static void store(void* p, size_t i, float v) {
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  }
  __device__ static float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  __device__ static void store4(void* p, size_t i, const float* v) {
    uint2 u;
    *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = u;
  }
};
template <> struct Elem<ALG_F16> {
  using type = __half;
  __device__ static float load(const void* p, size_t i) { return __half2float(reinterpret_cast<const __half*>(p)[i]); }
  __device__ static void store(void* p, size_t i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }
  __device__ static float round(float v) { return __half2float(__float2half_rn(v)); }
  __device__ static void store4(void* p, size_t i, const float* v) {
    uint2 u;
    *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[0], v[1]);
    *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p) + i) = u;
  }
};

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// host-side rounding emulation (for dtype-faithful filter taps)
inline float host_round_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return f;
  uint32_t lsb = (u >> 16) & 1u;
  u = (u + 0x7fffu + lsb) & 0xffff0000u;
  memcpy(&f, &u, 4);
  return f;
}
float host_round_f16(float f);  // lowpass.cu
inline float host_round(float f, int dt) {
  return dt == ALG_BF16 ? host_round_bf16(f) : (dt == ALG_F16 ? host_round_f16(f) : f);
}

}  // namespace alg
