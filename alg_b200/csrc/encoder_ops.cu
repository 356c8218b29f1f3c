// Kernels of the once-per-video conditioning encoders: UMT5 / T5 text encoder and CLIP vision / text
// transformers (reference call sites wan:185-234, cog:228-268, hy:282-452; the networks themselves live in
// transformers==4.48.1, requirements.txt).  The big linears run on the tcgen05 GEMM (gemm.cu); this file holds what is
// left around them.  None of it is on the per-step path: a few hundred tokens, once per video -- plain CUDA-core kernels,
// written for fidelity to the eager op chain (where each bf16 rounding happens), not for the tensor pipe.
//
//   t5_rms_norm_kernel       T5LayerNorm: h = bf16(x * rsqrt(mean(x^2) + eps)); y = bf16(w * h)
//   gather_rows_kernel       nn.Embedding lookup
//   small_attention_kernel   softmax(scale * q k^T + rel_bias[h, j - i] + mask) v for <= 768 keys, head_dim <= 128, bf16 or
//                            fp32 (T5 attention has an additive relative-position bias and no scaling, CLIP-ViT-H has
//                            head_dim 80 and runs in fp32, CLIP text is causal: none of it fits the flash kernel's shapes)
//   layer_norm_f32_kernel, bias_act_f32_kernel, split3_kernel   the fp32 CLIP vision path: LayerNorm and bias / GELU /
//                            residual in fp32, and the 3-term bf16 split [hi | hi | lo] x [hi | lo | hi] that lets the bf16
//                            tensor-core GEMM (fp32 accumulation) reproduce an fp32 nn.Linear to ~1e-5
#include <algorithm>

#include "common.cuh"

namespace alg {
namespace enc {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) t += red[i];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256) t5_rms_norm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                          int d, int64_t ld_x, int64_t ld_out, float eps,
                                                          const __nv_bfloat16* __restrict__ w) {
  __shared__ float red[8];
  const __nv_bfloat16* xr = x + (int64_t)blockIdx.x * ld_x;
  __nv_bfloat16* orow = out + (int64_t)blockIdx.x * ld_out;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += 256) {
    const float v = __bfloat162float(xr[i]);
    ss += v * v;
  }
  const float r = rsqrtf(block_sum<256>(ss, red) / (float)d + eps);
  for (int i = threadIdx.x; i < d; i += 256) {
    const float h = bf16_round(__bfloat162float(xr[i]) * r);  // hidden_states.to(weight.dtype)
    orow[i] = __float2bfloat16_rn(__bfloat162float(w[i]) * h);
  }
}

__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ table, const int64_t* __restrict__ ids,
                                   __nv_bfloat16* __restrict__ out, int d, int64_t vocab) {
  int64_t id = ids[blockIdx.x];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(table + id * d);
  uint4* dst = reinterpret_cast<uint4*>(out + (int64_t)blockIdx.x * d);
  for (int i = threadIdx.x; i < d / 8; i += blockDim.x) dst[i] = src[i];
}

// ---- small attention ------------------------------------------------------------------------------------------------
constexpr int kWarps = 4, kRowsPerWarp = 4, kRowsPerCta = kWarps * kRowsPerWarp * 2;  // two row groups per warp

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct SmallAttn {
  const void *q, *k, *v;
  void* out;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;  // element strides: batch, token row (head h starts at h * D)
  int L_q, L_kv, D, Lp;                                    // Lp = keys padded to 32
  float scale;
  const float* rel_bias;  // [H, 2 * L_kv - 1]: added to the score of (query i, key j) at index j - i + L_kv - 1; or NULL
  const int32_t* kv_valid;  // [B] keys >= kv_valid[b] are masked (right padding); or NULL
  int causal;
  int kv_group;             // query heads per key / value head (grouped-query attention); 1 = one K/V head per query head
  const uint8_t* key_mask;  // [B, L_kv] 0 = masked key (a padding mask that is not a prefix); or NULL
};

template <typename T, int KPL>
__global__ void __launch_bounds__(kWarps * 32) small_attention_kernel(const SmallAttn p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  T* sKt = reinterpret_cast<T*>(smem_raw);               // [D][Lp]
  T* sV = sKt + (size_t)p.D * p.Lp;                      // [Lp][D]
  float* sQ = reinterpret_cast<float*>(sV + (size_t)p.Lp * p.D);  // [kWarps][kRowsPerWarp][D]
  float* sP = sQ + kWarps * kRowsPerWarp * p.D;          // [kWarps][kRowsPerWarp][Lp]
  const int h = blockIdx.y, b = blockIdx.z, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = p.D, Lp = p.Lp;
  const T* kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.k_bs + (int64_t)h * D;
  const T* vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.v_bs + (int64_t)h * D;
  for (int idx = threadIdx.x; idx < Lp * D; idx += blockDim.x) {
    const int j = idx / D, d = idx - j * D;
    const bool ok = j < p.L_kv;
    sKt[(size_t)d * Lp + j] = ok ? kb[(int64_t)j * p.k_rs + d] : from_f<T>(0.f);
    sV[idx] = ok ? vb[(int64_t)j * p.v_rs + d] : from_f<T>(0.f);
  }
  __syncthreads();
  const int n_valid = p.kv_valid ? min(p.kv_valid[b], p.L_kv) : p.L_kv;
  float* myQ = sQ + warp * kRowsPerWarp * D;
  float* myP = sP + warp * kRowsPerWarp * Lp;
  const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_bs + (int64_t)h * D;
  T* ob = reinterpret_cast<T*>(p.out) + (int64_t)b * p.o_bs + (int64_t)h * D;
  for (int grp = 0; grp < 2; ++grp) {
    const int row0 = blockIdx.x * kRowsPerCta + (grp * kWarps + warp) * kRowsPerWarp;
    if (row0 >= p.L_q) continue;  // warp-uniform
    for (int idx = lane; idx < kRowsPerWarp * D; idx += 32) {
      const int r = idx / D, d = idx - r * D;
      myQ[idx] = row0 + r < p.L_q ? to_f<T>(qb[(int64_t)(row0 + r) * p.q_rs + d]) : 0.f;
    }
    __syncwarp();
    float s[kRowsPerWarp][KPL];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
      for (int c = 0; c < KPL; ++c) s[r][c] = 0.f;
    for (int d = 0; d < D; ++d) {
      float kv[KPL];
#pragma unroll
      for (int c = 0; c < KPL; ++c) kv[c] = (c * 32 < Lp) ? to_f<T>(sKt[(size_t)d * Lp + c * 32 + lane]) : 0.f;
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const float qv = myQ[r * D + d];
#pragma unroll
        for (int c = 0; c < KPL; ++c) s[r][c] = fmaf(qv, kv[c], s[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int i = row0 + r;
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < KPL; ++c) {
        const int j = c * 32 + lane;
        float x = s[r][c];
        if (sizeof(T) == 2) x = bf16_round(x);  // the q k^T matmul returns a bf16 tensor
        x *= p.scale;
        if (p.rel_bias && j < p.L_kv && i < p.L_q) {
          x += p.rel_bias[(int64_t)h * (2 * p.L_kv - 1) + (j - i + p.L_kv - 1)];
          if (sizeof(T) == 2) x = bf16_round(x);  // scores += position_bias, in the model dtype
        }
        if (j >= n_valid || (p.causal && j > i)) x = -INFINITY;
        s[r][c] = x;
        m = fmaxf(m, x);
      }
      m = warp_max(m);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < KPL; ++c) {
        const float e = (s[r][c] == -INFINITY) ? 0.f : __expf(s[r][c] - m);
        s[r][c] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
      for (int c = 0; c < KPL; ++c)
        if (c * 32 < Lp) myP[r * Lp + c * 32 + lane] = to_f<T>(from_f<T>(s[r][c] * inv));  // softmax(fp32).type_as(scores)
    }
    __syncwarp();
    // O = P V: lanes own head_dim columns d = lane + 32 c
    float o[kRowsPerWarp][4];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[r][c] = 0.f;
    for (int j = 0; j < n_valid; ++j) {
      float vv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) vv[c] = (c * 32 + lane < D) ? to_f<T>(sV[(size_t)j * D + c * 32 + lane]) : 0.f;
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const float pv = myP[r * Lp + j];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = fmaf(pv, vv[c], o[r][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
      if (row0 + r < p.L_q)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c * 32 + lane < D) ob[(int64_t)(row0 + r) * p.o_rs + c * 32 + lane] = from_f<T>(o[r][c]);
    __syncwarp();
  }
}

// ---- tiled attention (LLaVA's Llama: ~900 tokens, causal + padding mask, 32 query heads on 8 K/V heads, fp32) --------
// Same contract as small_attention_kernel without the 768-key limit and without rel_bias: keys stream through shared memory
// 64 at a time with a running (max, sum) per query row.  A warp owns 8 query rows; a lane owns keys lane, lane + 32 of the
// tile for the scores and head_dim columns lane + 32 c for the output.
constexpr int kTK = 64, kTRows = 8;
template <typename T>
__global__ void __launch_bounds__(kWarps * 32) tiled_attention_kernel(const SmallAttn p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int D = p.D;
  T* sKt = reinterpret_cast<T*>(smem_raw);                        // [D][kTK]
  T* sV = sKt + (size_t)D * kTK;                                  // [kTK][D]
  float* sQ = reinterpret_cast<float*>(sV + (size_t)kTK * D);     // [kWarps][kTRows][D]
  float* sP = sQ + kWarps * kTRows * D;                           // [kWarps][kTRows][kTK]
  const int h = blockIdx.y, b = blockIdx.z, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hk = h / p.kv_group;
  const T* kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * p.k_bs + (int64_t)hk * D;
  const T* vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * p.v_bs + (int64_t)hk * D;
  const T* qb = reinterpret_cast<const T*>(p.q) + (int64_t)b * p.q_bs + (int64_t)h * D;
  T* ob = reinterpret_cast<T*>(p.out) + (int64_t)b * p.o_bs + (int64_t)h * D;
  const uint8_t* km = p.key_mask ? p.key_mask + (int64_t)b * p.L_kv : nullptr;
  const int n_valid = p.kv_valid ? min(p.kv_valid[b], p.L_kv) : p.L_kv;
  const int cta_row0 = blockIdx.x * (kWarps * kTRows), row0 = cta_row0 + warp * kTRows;
  float* myQ = sQ + warp * kTRows * D;
  float* myP = sP + warp * kTRows * kTK;
  for (int idx = lane; idx < kTRows * D; idx += 32) {
    const int r = idx / D, d = idx - r * D;
    myQ[idx] = row0 + r < p.L_q ? to_f<T>(qb[(int64_t)(row0 + r) * p.q_rs + d]) : 0.f;
  }
  float m_run[kTRows], l_run[kTRows], o[kTRows][4];
#pragma unroll
  for (int r = 0; r < kTRows; ++r) {
    m_run[r] = -INFINITY;
    l_run[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[r][c] = 0.f;
  }
  // keys past the last query row of the CTA are never visible under the causal mask
  const int key_end = p.causal ? min(n_valid, min(cta_row0 + kWarps * kTRows, p.L_q)) : n_valid;
  for (int k0 = 0; k0 < key_end; k0 += kTK) {
    __syncthreads();  // the previous tile is fully consumed
    for (int idx = threadIdx.x; idx < kTK * D; idx += blockDim.x) {
      const int j = idx / D, d = idx - j * D;
      const bool ok = k0 + j < p.L_kv;
      sKt[(size_t)d * kTK + j] = ok ? kb[(int64_t)(k0 + j) * p.k_rs + d] : from_f<T>(0.f);
      sV[idx] = ok ? vb[(int64_t)(k0 + j) * p.v_rs + d] : from_f<T>(0.f);
    }
    __syncthreads();
    float s[kTRows][2];
#pragma unroll
    for (int r = 0; r < kTRows; ++r) s[r][0] = s[r][1] = 0.f;
    for (int d = 0; d < D; ++d) {
      const float k_a = to_f<T>(sKt[(size_t)d * kTK + lane]), k_b = to_f<T>(sKt[(size_t)d * kTK + 32 + lane]);
#pragma unroll
      for (int r = 0; r < kTRows; ++r) {
        const float qv = myQ[r * D + d];
        s[r][0] = fmaf(qv, k_a, s[r][0]);
        s[r][1] = fmaf(qv, k_b, s[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < kTRows; ++r) {
      const int i = row0 + r;
      float mt = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = k0 + c * 32 + lane;
        float x = s[r][c];
        if (sizeof(T) == 2) x = bf16_round(x);
        x *= p.scale;
        if (j >= n_valid || (p.causal && j > i) || (km && j < p.L_kv && !km[j])) x = -INFINITY;
        s[r][c] = x;
        mt = fmaxf(mt, x);
      }
      mt = warp_max(mt);
      const float m_new = fmaxf(m_run[r], mt);
      const float f = (m_run[r] == -INFINITY) ? 0.f : __expf(m_run[r] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float e = (s[r][c] == -INFINITY) ? 0.f : __expf(s[r][c] - m_new);
        myP[r * kTK + c * 32 + lane] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      l_run[r] = l_run[r] * f + sum;
      m_run[r] = m_new;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[r][c] *= f;
    }
    __syncwarp();
    const int jn = min(kTK, key_end - k0);
    for (int j = 0; j < jn; ++j) {
      float vv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) vv[c] = (c * 32 + lane < D) ? to_f<T>(sV[(size_t)j * D + c * 32 + lane]) : 0.f;
#pragma unroll
      for (int r = 0; r < kTRows; ++r) {
        const float pv = myP[r * kTK + j];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[r][c] = fmaf(pv, vv[c], o[r][c]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < kTRows; ++r) {
    if (row0 + r >= p.L_q) continue;
    const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c * 32 + lane < D) ob[(int64_t)(row0 + r) * p.o_rs + c * 32 + lane] = from_f<T>(o[r][c] * inv);
  }
}

// ---- fp32 pieces of the Llama decoder stack (LLaVA text encoder, hy:333-337) -----------------------------------------
// LlamaRMSNorm: out = weight * (x * rsqrt(mean(x^2) + eps))
__global__ void __launch_bounds__(256) rms_norm_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int d, float eps,
                                                           const float* __restrict__ w) {
  __shared__ float red[8];
  const float* xr = x + (int64_t)blockIdx.x * d;
  float* orow = out + (int64_t)blockIdx.x * d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += 256) ss += xr[i] * xr[i];
  const float r = rsqrtf(block_sum<256>(ss, red) / (float)d + eps);
  for (int i = threadIdx.x; i < d; i += 256) orow[i] = w[i] * (xr[i] * r);
}
// apply_rotary_pos_emb, rotate-half convention, in place on `heads` heads of every row: x = x * cos + rotate_half(x) * sin
// with cos / sin [rows, D] (the two halves of a table row are equal); products and sum rounded separately like the eager ops
__global__ void rope_half_f32_kernel(float* __restrict__ x, int64_t ld, const float* __restrict__ cs, const float* __restrict__ sn,
                                     int heads, int D, int64_t n) {
  const int half = D / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % half);
    const int64_t t = i / half;
    const int h = (int)(t % heads);
    const int64_t row = t / heads;
    float* xp = x + row * ld + (int64_t)h * D;
    const float a = xp[c], b2 = xp[c + half];
    const float* cr = cs + row * D;
    const float* sr = sn + row * D;
    xp[c] = __fadd_rn(__fmul_rn(a, cr[c]), __fmul_rn(-b2, sr[c]));
    xp[c + half] = __fadd_rn(__fmul_rn(b2, cr[c + half]), __fmul_rn(a, sr[c + half]));
  }
}
// LlamaMLP gate: out[r, c] = silu(gu[r, c]) * gu[r, f + c] for the fused [gate | up] projection gu [rows, 2 f]
__global__ void swiglu_f32_kernel(const float* __restrict__ gu, float* __restrict__ out, int f, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / f;
    const int c = (int)(i - r * f);
    const float g = gu[r * 2 * f + c], u = gu[r * 2 * f + f + c];
    out[i] = (g / (1.0f + expf(-g))) * u;
  }
}

// ---- fp32 path of the CLIP vision tower ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layer_norm_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int d,
                                                             float eps, const float* __restrict__ w,
                                                             const float* __restrict__ b) {
  __shared__ float red[8];
  const float* xr = x + (int64_t)blockIdx.x * d;
  float* orow = out + (int64_t)blockIdx.x * d;
  float s = 0.f;
  for (int i = threadIdx.x; i < d; i += 256) s += xr[i];
  const float mean = block_sum<256>(s, red) / (float)d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += 256) {
    const float c = xr[i] - mean;
    ss += c * c;
  }
  const float r = rsqrtf(block_sum<256>(ss, red) / (float)d + eps);
  for (int i = threadIdx.x; i < d; i += 256) orow[i] = (xr[i] - mean) * r * w[i] + b[i];
}

// x[r, c] = act(x[r, c] + bias[c]) (+ residual[r, c]);  act: 0 none, 1 GELU (erf), 2 quick-GELU
__global__ void bias_act_f32_kernel(float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ res,
                                    int64_t n, int cols, int act) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i] + (bias ? bias[i % cols] : 0.f);
    if (act == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    else if (act == 2) v = v / (1.0f + expf(-1.702f * v));
    if (res) v += res[i];
    x[i] = v;
  }
}

// out[r, 0:K] = hi(x), out[r, K:2K] = order ? lo(x) : hi(x), out[r, 2K:3K] = order ? hi(x) : lo(x)   (bf16)
// activations use order 0 -> [hi | hi | lo], weights order 1 -> [hi | lo | hi]: the K-concatenated product is
// hi*hi + hi*lo + lo*hi, an fp32 product up to the dropped lo*lo term (2^-16 relative).
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int K, int order) {
  const int64_t n = rows * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K;
    const int c = (int)(i - r * K);
    const float v = x[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * 3 * (int64_t)K;
    o[c] = hi;
    o[K + c] = order ? lo : hi;
    o[2 * K + c] = order ? hi : lo;
  }
}

__global__ void mul_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) * __bfloat162float(b[i]));
}

// CLIPVisionEmbeddings: the stride-P patch convolution as unfold + GEMM.  x [B, C, H, W] fp32 -> rows [B * gh * gw, ld]
// fp32, row = the (c, i, j)-flattened patch (nn.Conv2d weight.view(out, -1) order), zero-padded from C*P*P to ld
__global__ void patchify_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W, int P, int ld,
                                    int64_t n) {
  const int gh = H / P, gw = W / P;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % ld);
    const int64_t row = i / ld;
    float v = 0.f;
    if (col < C * P * P) {
      const int c = col / (P * P), ij = col - c * P * P, ii = ij / P, jj = ij - ii * P;
      const int64_t b = row / (gh * gw);
      const int pi = (int)(row - b * gh * gw), py = pi / gw, px = pi - py * gw;
      v = x[((b * C + c) * H + py * P + ii) * (int64_t)W + px * P + jj];
    }
    out[i] = v;
  }
}
// out[b, 0, :] = class_embedding + pos[0]; out[b, 1 + n, :] = patches[b * np + n, :] + pos[1 + n]
__global__ void clip_embed_f32_kernel(const float* __restrict__ patches, const float* __restrict__ cls,
                                      const float* __restrict__ pos, float* __restrict__ out, int np, int d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const int64_t tok_all = i / d;
    const int tok = (int)(tok_all % (np + 1));
    const int64_t b = tok_all / (np + 1);
    out[i] = (tok == 0 ? cls[c] : patches[(b * np + tok - 1) * (int64_t)d + c]) + pos[(int64_t)tok * d + c];
  }
}

static int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 8); }

template <typename T>
static int launch_small_attention(const SmallAttn& p, int B, int H, cudaStream_t st) {
  const size_t smem = 2 * (size_t)p.D * p.Lp * sizeof(T) + (size_t)kWarps * kRowsPerWarp * (p.D + p.Lp) * sizeof(float);
  ALG_REQUIRE(smem <= 227 * 1024, "small_attention: keys x head_dim do not fit in shared memory (this kernel serves the "
                                  "conditioning encoders: <= 768 keys)");
  const int kpl = p.Lp / 32;
  dim3 grid((unsigned)((p.L_q + kRowsPerCta - 1) / kRowsPerCta), (unsigned)H, (unsigned)B);
#define ALG_SA_CASE(K)                                                                                             \
  if (kpl <= K) {                                                                                                  \
    ALG_CUDA_OK(cudaFuncSetAttribute(small_attention_kernel<T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    small_attention_kernel<T, K><<<grid, kWarps * 32, smem, st>>>(p);                                              \
    ALG_LAUNCH_OK();                                                                                               \
    return 0;                                                                                                      \
  }
  ALG_SA_CASE(4)
  ALG_SA_CASE(9)
  ALG_SA_CASE(16)
  ALG_SA_CASE(24)
#undef ALG_SA_CASE
  ALG_REQUIRE(false, "small_attention: more than 768 keys");
}

template <typename T>
static int launch_tiled_attention(const SmallAttn& p, int B, int H, cudaStream_t st) {
  const size_t smem = 2 * (size_t)p.D * kTK * sizeof(T) + (size_t)kWarps * kTRows * (p.D + kTK) * sizeof(float);
  dim3 grid((unsigned)((p.L_q + kWarps * kTRows - 1) / (kWarps * kTRows)), (unsigned)H, (unsigned)B);
  ALG_CUDA_OK(cudaFuncSetAttribute(tiled_attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tiled_attention_kernel<T><<<grid, kWarps * 32, smem, st>>>(p);
  ALG_LAUNCH_OK();
  return 0;
}

}  // namespace enc
}  // namespace alg

using namespace alg;

extern "C" int alg_t5_rms_norm_bf16(const void* x, int64_t ld_x, void* out, int64_t ld_out, int64_t rows, int d, float eps,
                                    const void* weight, void* stream) {
  ALG_REQUIRE(x && out && weight && rows >= 0 && d > 0 && ld_x >= d && ld_out >= d, "t5_rms_norm: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::t5_rms_norm_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), d, ld_x, ld_out, eps,
      reinterpret_cast<const __nv_bfloat16*>(weight));
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_gather_rows_bf16(const void* table, int64_t vocab, const int64_t* ids, void* out, int64_t rows, int d,
                                    void* stream) {
  ALG_REQUIRE(table && ids && out && rows >= 0 && d > 0 && d % 8 == 0 && vocab > 0, "gather_rows: bad arguments (d % 8 == 0)");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "gather_rows: 16-byte alignment");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::gather_rows_kernel<<<(unsigned)rows, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(table), ids, reinterpret_cast<__nv_bfloat16*>(out), d, vocab);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_small_attention(const alg_small_attention_t* a, void* stream) {
  ALG_REQUIRE(a && a->q && a->k && a->v && a->out, "small_attention: null pointer");
  ALG_REQUIRE(a->dtype == ALG_BF16 || a->dtype == ALG_F32, "small_attention: dtype must be bf16 or f32");
  ALG_REQUIRE(a->batch > 0 && a->heads > 0 && a->n_q > 0 && a->n_kv > 0 && a->head_dim > 0 && a->head_dim <= 128,
              "small_attention: empty problem or head_dim > 128");
  ALG_REQUIRE(a->batch <= 65535 && a->heads <= 65535, "small_attention: batch / heads exceed the grid limits");
  if (int rc = alg_check_device()) return rc;
  enc::SmallAttn p;
  p.q = a->q; p.k = a->k; p.v = a->v; p.out = a->out;
  p.q_bs = a->q_bs; p.q_rs = a->q_rs; p.k_bs = a->k_bs; p.k_rs = a->k_rs;
  p.v_bs = a->v_bs; p.v_rs = a->v_rs; p.o_bs = a->o_bs; p.o_rs = a->o_rs;
  p.L_q = (int)a->n_q; p.L_kv = (int)a->n_kv; p.D = a->head_dim; p.Lp = ((int)a->n_kv + 31) / 32 * 32;
  p.scale = a->scale; p.rel_bias = a->rel_bias; p.kv_valid = a->kv_valid; p.causal = a->causal;
  p.kv_group = a->kv_group > 1 ? a->kv_group : 1;
  p.key_mask = a->key_mask;
  ALG_REQUIRE(a->heads % p.kv_group == 0, "small_attention: heads must be a multiple of kv_group");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t elem = a->dtype == ALG_BF16 ? 2 : 4;
  const size_t small_smem = 2 * (size_t)p.D * p.Lp * elem + (size_t)enc::kWarps * enc::kRowsPerWarp * (p.D + p.Lp) * sizeof(float);
  // streamed-key variant: grouped K/V heads, general key mask, long prompts, or all keys of a head do not fit in shared memory
  // (CLIP-ViT-L/14-336 inside LLaVA: 577 tokens x 64 in fp32)
  if (p.kv_group > 1 || p.key_mask || a->n_kv > 768 || (small_smem > 227 * 1024 && !a->rel_bias)) {
    ALG_REQUIRE(!a->rel_bias, "small_attention: rel_bias needs n_kv <= 768, kv_group 1 and no key_mask");
    if (a->dtype == ALG_BF16) return enc::launch_tiled_attention<__nv_bfloat16>(p, a->batch, a->heads, st);
    return enc::launch_tiled_attention<float>(p, a->batch, a->heads, st);
  }
  if (a->dtype == ALG_BF16) return enc::launch_small_attention<__nv_bfloat16>(p, a->batch, a->heads, st);
  return enc::launch_small_attention<float>(p, a->batch, a->heads, st);
}

extern "C" int alg_layer_norm_f32(const float* x, float* out, int64_t rows, int d, float eps, const float* weight,
                                  const float* bias, void* stream) {
  ALG_REQUIRE(x && out && weight && bias && rows >= 0 && d > 0, "layer_norm_f32: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::layer_norm_f32_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, d, eps, weight, bias);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_bias_act_f32(float* x, const float* bias, const float* residual, int64_t rows, int cols, int act,
                                void* stream) {
  ALG_REQUIRE(x && rows >= 0 && cols > 0 && act >= 0 && act <= 2, "bias_act_f32: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::bias_act_f32_kernel<<<enc::grid_for(rows * cols), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, bias, residual, rows * cols, cols, act);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_split3_bf16(const float* x, void* out, int64_t rows, int K, int weight_order, void* stream) {
  ALG_REQUIRE(x && out && rows >= 0 && K > 0, "split3: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::split3_kernel<<<enc::grid_for(rows * K), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out), rows, K, weight_order);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_mul_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  ALG_REQUIRE(a && b && out && n >= 0, "mul_bf16: null pointer");
  if (int rc = alg_check_device()) return rc;
  if (n == 0) return 0;
  enc::mul_bf16_kernel<<<enc::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
      reinterpret_cast<__nv_bfloat16*>(out), n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_patchify_f32(const float* x, float* out, int batch, int C, int H, int W, int P, int ld, void* stream) {
  ALG_REQUIRE(x && out && batch > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0 && ld >= C * P * P, "patchify: bad arguments");
  if (int rc = alg_check_device()) return rc;
  const int64_t n = (int64_t)batch * (H / P) * (W / P) * ld;
  enc::patchify_f32_kernel<<<enc::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, C, H, W, P, ld, n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_clip_embed_f32(const float* patches, const float* class_embedding, const float* position_embedding, float* out,
                                  int batch, int num_patches, int d, void* stream) {
  ALG_REQUIRE(patches && class_embedding && position_embedding && out && batch > 0 && num_patches > 0 && d > 0, "clip_embed: bad arguments");
  if (int rc = alg_check_device()) return rc;
  const int64_t n = (int64_t)batch * (num_patches + 1) * d;
  enc::clip_embed_f32_kernel<<<enc::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      patches, class_embedding, position_embedding, out, num_patches, d, n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_rms_norm_f32(const float* x, float* out, int64_t rows, int d, float eps, const float* weight, void* stream) {
  ALG_REQUIRE(x && out && weight && rows >= 0 && d > 0, "rms_norm_f32: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::rms_norm_f32_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, d, eps, weight);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_rope_half_f32(float* x, int64_t ld, const float* cos_table, const float* sin_table, int64_t rows, int heads,
                                 int head_dim, void* stream) {
  ALG_REQUIRE(x && cos_table && sin_table && rows >= 0 && heads > 0 && head_dim > 0 && head_dim % 2 == 0 &&
                  ld >= (int64_t)heads * head_dim, "rope_half_f32: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  const int64_t n = rows * heads * (head_dim / 2);
  enc::rope_half_f32_kernel<<<enc::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ld, cos_table, sin_table,
                                                                                                 heads, head_dim, n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_swiglu_f32(const float* gate_up, float* out, int64_t rows, int f, void* stream) {
  ALG_REQUIRE(gate_up && out && rows >= 0 && f > 0, "swiglu_f32: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  enc::swiglu_f32_kernel<<<enc::grid_for(rows * f), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gate_up, out, f, rows * f);
  ALG_LAUNCH_OK();
  return 0;
}
