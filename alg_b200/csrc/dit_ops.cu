// HBM-bound DiT building blocks exported through the C ABI (include/alg_b200.h, "HBM-bound DiT building blocks"):
// LayerNorm (+affine, +AdaLN modulate, fp32 or eager-bf16 rounding chain), per-head q/k norm + real-valued RoPE,
// strided patch gather / unpatchify, timestep embedding and the few tiny vector ops of the conditioning path.
//
// They serve the CogVideoX and HunyuanVideo forwards (reference call sites cog:1082-1090, hy:1243-1252), whose module
// sequencing lives host-side in alg_b200/cogvideox.py / alg_b200/hunyuan.py, and the Wan engine's LayerNorm.  Every
// kernel is one read + one write of its activation with 16-byte accesses; modulation / affine vectors are read as
// vectors too (scalar fp32 loads of scale/shift made the first Wan LayerNorm LSU-bound at 1.7 TB/s).
#include <algorithm>
#include <cstdlib>

#include "dit_kernels.cuh"

namespace alg {
namespace ops {

constexpr int kRowThreads = 256;
constexpr int kMaxChunks = 4;  // 8 elements per chunk per thread -> d <= 8192

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x;
    f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  return u;
}
// 8 consecutive elements of a [d] vector stored as fp32 or bf16, chunk index ci (read-only path)
__device__ __forceinline__ void load_vec8(const void* base, int dtype, int ci, float* f) {
  if (dtype == ALG_BF16) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(base) + ci), f);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(base) + 2 * ci);
    const float4 b = __ldg(reinterpret_cast<const float4*>(base) + 2 * ci + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kRowThreads / 32; ++i) t += red[i];
  return t;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm
// ------------------------------------------------------------------------------------------------
// One row per block, held in registers.  TH threads x CH chunks of 8 elements cover the row; CH is a template parameter so
// the register file is sized for the actual width (d = 5120: 84 registers at 128 x 5, 6 blocks per SM, against 64 registers
// x 256 threads = 4 blocks for the first version): 3.47 -> 3.94 TB/s.  A cp.async row-streaming variant (persistent blocks,
// 4-stage shared-memory ring, 120 KB in flight per SM) was slower (3.2 TB/s): the per-row block reductions, not the bytes in
// flight, pace this kernel.  Also measured and dropped in round 2 (profiles/hbm_r02.log notes): a warp-per-row variant with the
// row packed in registers (no barrier at all; 0.56-0.65 of the HBM peak against 0.72: 130-230 registers per thread) and
// single-barrier statistics by Chan's pairwise (count, mean, M2) combination (0.63: ~75 more instructions per thread-row than the
// two plain reductions).  The kernel is bound by its ~650 instructions per thread-row, not by barriers or bytes in flight.
template <int TH>
__device__ __forceinline__ float block_sum_t(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < TH / 32; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ float2 bf16x2_round(float2 x) { return __bfloat1622float2(__float22bfloat162_rn(x)); }

// FAST: the launch covers the common shape exactly -- d == TH * CH * 8 (no range guards), fp32 affine / modulation vectors (no
// dtype branches per chunk), one modulation set for all rows (no per-row batch / split arithmetic).
template <bool CHAIN, int TH, int CH, bool FAST = false>
__global__ void __launch_bounds__(TH) layer_norm_kernel(const alg_layer_norm_t p) {
  constexpr int kRowThreads = TH, kMaxChunks = CH;
  __shared__ float red[TH / 32];
  const int64_t row = blockIdx.x;
  const int d = p.d;
  const uint4* xr = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + row * d);
  uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * d);
  const int chunks = FAST ? TH * CH : d >> 3;
  // Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2: IEEE rn per lane, same results as the scalar ops): the kernel is
  // instruction-issue bound (ncu ln_r29: 21 instructions per element, issue slots 57 % busy, DRAM 60 %), not DRAM-bound.
  float2 v[kMaxChunks][4];
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ci = threadIdx.x + c * kRowThreads;
    if (FAST || ci < chunks) {
      const uint4 u = xr[ci];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[c][e] = __bfloat1622float2(h[e]);
        sum2 = __fadd2_rn(sum2, v[c][e]);
      }
    }
  }
  // modulation vectors of this row (resolved while the loads are in flight)
  const void *scale = nullptr, *shift = nullptr;
  if constexpr (FAST) {
    scale = p.scale;
    shift = p.shift;
  } else if (p.scale) {
    const uint32_t rpb = (uint32_t)min(p.rows_per_batch, (int64_t)0x7fffffff);  // rows <= 2^31 - 1 (grid size)
    const uint32_t b = (uint32_t)row / rpb, r_in = (uint32_t)row - b * rpb;
    const bool alt = (int64_t)r_in < p.split_row;
    const size_t esz = p.mod_dtype == ALG_BF16 ? 2 : 4;
    scale = reinterpret_cast<const char*>(alt ? p.scale_alt : p.scale) + b * p.mod_batch_stride * esz;
    shift = reinterpret_cast<const char*>(alt ? p.shift_alt : p.shift) + b * p.mod_batch_stride * esz;
  }
  const float mean = block_sum_t<TH>(sum2.x + sum2.y, red) / (float)d;
  const float2 nmean2 = make_float2(-mean, -mean);
  float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ci = threadIdx.x + c * kRowThreads;
    if (FAST || ci < chunks) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[c][e] = __fadd2_rn(v[c][e], nmean2);  // x - mean, kept for the normalisation below
        sq2 = __ffma2_rn(v[c][e], v[c][e], sq2);
      }
    }
  }
  const float rstd = rsqrtf(block_sum_t<TH>(sq2.x + sq2.y, red) / (float)d + p.eps);
  const float2 rstd2 = make_float2(rstd, rstd), one2 = make_float2(1.f, 1.f);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ci = threadIdx.x + c * kRowThreads;
    if (FAST || ci < chunks) {
      float2 o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = __fmul2_rn(v[c][e], rstd2);
      if (p.weight) {
        float w[8], b[8];
        load_vec8(p.weight, FAST ? (int)ALG_F32 : p.affine_dtype, ci, w);
        load_vec8(p.bias, FAST ? (int)ALG_F32 : p.affine_dtype, ci, b);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[e] = __fadd2_rn(__fmul2_rn(o[e], make_float2(w[2 * e], w[2 * e + 1])), make_float2(b[2 * e], b[2 * e + 1]));
      }
      if (scale) {
        float sc[8], sh[8];
        load_vec8(scale, FAST ? (int)ALG_F32 : p.mod_dtype, ci, sc);
        load_vec8(shift, FAST ? (int)ALG_F32 : p.mod_dtype, ci, sh);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 s1 = __fadd2_rn(one2, make_float2(sc[2 * e], sc[2 * e + 1]));
          const float2 sh2 = make_float2(sh[2 * e], sh[2 * e + 1]);
          if (CHAIN) {  // bf16 tensors all the way: norm(x) -> (1 + scale) -> product -> + shift, each op rounds
            const float2 prod = __fmul2_rn(bf16x2_round(o[e]), bf16x2_round(s1));
            o[e] = __fadd2_rn(bf16x2_round(prod), sh2);
          } else {
            o[e] = __fadd2_rn(__fmul2_rn(o[e], s1), sh2);
          }
        }
      }
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __float22bfloat162_rn(o[e]);
      orow[ci] = u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-head norm + RoPE: one block per row; a thread owns 16-byte chunks (8 elements = 4 rotary pairs), HD / 8 consecutive
// threads share a head and reduce with shuffles.  The chunk column inside the head is the same for every chunk of a thread
// (TH % (HD / 8) == 0), so its norm weights and cos / sin values are loaded once per row.  The first version (one warp
// per (row, head), 4-8 bytes per lane) reached 1.8 TB/s; 16-byte accesses and whole rows in flight fix that.
// ------------------------------------------------------------------------------------------------
template <int HD, int TH, int CH>
__global__ void __launch_bounds__(TH) head_norm_rope_kernel(const alg_head_norm_rope_t p) {
  constexpr int LPH = HD / 8;  // threads per head
  const int64_t row = blockIdx.x;
  const int chunks = p.heads * LPH;
  uint4* xr = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.x) + row * p.ld);
  const int sub = threadIdx.x % LPH;  // 8-element column block inside the head
  // packed fp32x2 arithmetic throughout (IEEE rn per lane: same rounding chain as the scalar ops, half the issue slots)
  float2 v[CH][4];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int ci = threadIdx.x + c * TH;
    if (ci < chunks) {
      const uint4 u = xr[ci];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[c][e] = __bfloat1622float2(h[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[c][e] = make_float2(0.f, 0.f);
    }
  }
  float w[8], bs[8], cs[8], sn[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    w[e] = 1.f;
    bs[e] = 0.f;
  }
  if (p.norm_kind != ALG_NORM_NONE) {
    load_vec8(p.weight, ALG_BF16, sub, w);
    if (p.norm_kind == ALG_NORM_LAYER && p.bias) load_vec8(p.bias, ALG_BF16, sub, bs);
  }
  const uint32_t rpb = (uint32_t)min(p.rows_per_batch, (int64_t)0x7fffffff);  // rows <= 2^31 - 1 (grid size)
  const int64_t rr = (int64_t)((uint32_t)row % rpb) - p.rope_row0;
  const bool rope = p.cos && rr >= 0 && rr < p.rope_rows;
  if (rope) {
    load_vec8(p.cos + rr * HD, ALG_F32, sub, cs);
    load_vec8(p.sin + rr * HD, ALG_F32, sub, sn);
  }
  float2 w2[4], bs2[4], cs2[4], sn2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    w2[e] = make_float2(w[2 * e], w[2 * e + 1]);
    bs2[e] = make_float2(bs[2 * e], bs[2 * e + 1]);
    cs2[e] = make_float2(cs[2 * e], cs[2 * e + 1]);
    sn2[e] = make_float2(sn[2 * e], sn[2 * e + 1]);
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int ci = threadIdx.x + c * TH;
    const bool live = ci < chunks;  // uniform within a head's LPH threads (chunks % LPH == 0); dead chunks hold zeros
    float2 y[4];
    if (p.norm_kind == ALG_NORM_RMS) {
      float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int e = 0; e < 4; ++e) sq2 = __ffma2_rn(v[c][e], v[c][e], sq2);
      float sq = sq2.x + sq2.y;
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / (float)HD + p.eps);
      const float2 rstd2 = make_float2(rstd, rstd);
#pragma unroll
      for (int e = 0; e < 4; ++e) y[e] = bf16x2_round(__fmul2_rn(bf16x2_round(__fmul2_rn(v[c][e], rstd2)), w2[e]));
    } else if (p.norm_kind == ALG_NORM_LAYER) {
      float2 sm2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int e = 0; e < 4; ++e) sm2 = __fadd2_rn(sm2, v[c][e]);
      float sm = sm2.x + sm2.y;
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
      const float mean = sm / (float)HD;
      const float2 nmean2 = make_float2(-mean, -mean);
      float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[c][e] = __fadd2_rn(v[c][e], nmean2);  // x - mean, reused below
        sq2 = __ffma2_rn(v[c][e], v[c][e], sq2);
      }
      float sq = live ? sq2.x + sq2.y : 0.f;
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq / (float)HD + p.eps);
      const float2 rstd2 = make_float2(rstd, rstd);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        y[e] = bf16x2_round(__fadd2_rn(__fmul2_rn(__fmul2_rn(v[c][e], rstd2), w2[e]), bs2[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) y[e] = v[c][e];
    }
    if (rope) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // x_rot = (-x_imag, x_real); out = x * cos + x_rot * sin, fp32 op by op
        const float2 rot = make_float2(-y[q].y, y[q].x);
        y[q] = __fadd2_rn(__fmul2_rn(y[q], cs2[q]), __fmul2_rn(rot, sn2[q]));
      }
    }
    if (live) {
      uint4 u;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __float22bfloat162_rn(y[e]);
      xr[ci] = u;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patch gather / unpatchify
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSrc = 6;  // n_pass (<= 3) x n_src (<= 2)
struct GatherArgs {
  alg_patch_src_t src[kMaxSrc];
  int n_pass, n_src, T, H, W;
  int ch_total;
  int64_t lda;
};

__device__ __forceinline__ float load_any(const void* p, int dtype, int64_t i) {
  if (dtype == ALG_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == ALG_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// one thread per (pass, token, channel): reads the 2 x 2 patch, writes 4 consecutive bf16 of the im2col row
__global__ void patch_gather_kernel(const GatherArgs a, __nv_bfloat16* __restrict__ A) {
  const int ph = a.H / 2, pw = a.W / 2;
  const int64_t N = (int64_t)a.T * ph * pw;
  const int64_t total = (int64_t)a.n_pass * N * a.ch_total;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % a.ch_total);
    const int64_t tok = idx / a.ch_total;
    const int pss = (int)(tok / N);
    const int64_t n = tok - (int64_t)pss * N;
    const int t = (int)(n / (ph * pw));
    const int rem = (int)(n - (int64_t)t * ph * pw);
    const int y = rem / pw, x = rem - y * pw;
    const int c_all = c;
    int s = 0;
    while (s + 1 < a.n_src && c >= a.src[pss * a.n_src + s].channels) {
      c -= a.src[pss * a.n_src + s].channels;
      ++s;
    }
    const alg_patch_src_t& d = a.src[pss * a.n_src + s];
    const void* base = d.ptr;
    int64_t off = (int64_t)c * d.sc + (int64_t)t * d.st;
    if (t == 0 && d.ptr_t0) {
      base = d.ptr_t0;
      off = (int64_t)c * d.sc_t0;
    }
    off += (int64_t)(2 * y) * d.sy + 2 * x;
    const float v00 = load_any(base, d.dtype, off), v01 = load_any(base, d.dtype, off + 1);
    const float v10 = load_any(base, d.dtype, off + d.sy), v11 = load_any(base, d.dtype, off + d.sy + 1);
    uint2 u;
    *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v00, v01);
    *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v10, v11);
    *reinterpret_cast<uint2*>(A + tok * a.lda + c_all * 4) = u;
  }
}

__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ proj, int64_t ld, __nv_bfloat16* __restrict__ out,
                                  int n_pass, int C, int T, int H, int W, int64_t s_pass, int64_t sc, int64_t st,
                                  int64_t sy, int channel_major) {
  const int ph = H / 2, pw = W / 2;
  const int64_t N = (int64_t)T * ph * pw;
  const int64_t total = (int64_t)n_pass * C * T * H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H); r /= H;
    const int t = (int)(r % T); r /= T;
    const int c = (int)(r % C);
    const int pss = (int)(r / C);
    const int64_t n = ((int64_t)t * ph + (y >> 1)) * pw + (x >> 1);
    const int ij = ((y & 1) * 2) + (x & 1);
    const int k = channel_major ? c * 4 + ij : ij * C + c;
    out[(int64_t)pss * s_pass + (int64_t)c * sc + (int64_t)t * st + (int64_t)y * sy + x] =
        proj[((int64_t)pss * N + n) * ld + k];
  }
}

// ------------------------------------------------------------------------------------------------
// conditioning-path vector ops
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(float timestep, int dim, void* out, int dtype) {
  const int half = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += gridDim.x * blockDim.x) {
    // exponent = -log(10000) * arange(half) / half ; emb = t * exp(exponent) ; [cos | sin] (flip_sin_to_cos)
    const float exponent = (-9.210340371976184f * (float)i) / (float)half;
    const float e = timestep * expf(exponent);
    const float cs = cosf(e), sn = sinf(e);
    if (dtype == ALG_F32) {
      reinterpret_cast<float*>(out)[i] = cs;
      reinterpret_cast<float*>(out)[half + i] = sn;
    } else {
      reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(cs);
      reinterpret_cast<__nv_bfloat16*>(out)[half + i] = __float2bfloat16_rn(sn);
    }
  }
}

__device__ __forceinline__ float gelu_tanh_exact(float x) {
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  return 0.5f * x * (1.0f + tanhf(kBeta * (x + kKappa * x * x * x)));
}

__global__ void elementwise_kernel(int op, const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                   __nv_bfloat16* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = __bfloat162float(a[i]);
    float y = x;
    if (op == ALG_EW_ADD) y = __fadd_rn(x, __bfloat162float(b[i]));
    else if (op == ALG_EW_SILU) y = x / (1.0f + expf(-x));
    else if (op == ALG_EW_GELU_TANH) y = gelu_tanh_exact(x);
    out[i] = __float2bfloat16_rn(y);
  }
}

// one block per 8-column chunk group: thread t accumulates rows t, t + 256, ... of 8 columns; fp32
__global__ void __launch_bounds__(256) mean_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int d,
                                                        int64_t ld, __nv_bfloat16* __restrict__ out) {
  __shared__ float red[256][9];
  const int ci = blockIdx.x;  // chunk of 8 columns
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + r * ld + ci * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[threadIdx.x][e] += red[threadIdx.x + s][e];
    __syncthreads();
  }
  if (threadIdx.x < 8) out[ci * 8 + threadIdx.x] = __float2bfloat16_rn(red[0][threadIdx.x] / (float)rows);
}

}  // namespace ops
}  // namespace alg

using namespace alg;

extern "C" int alg_layer_norm(const alg_layer_norm_t* p, void* stream) {
  ALG_REQUIRE(p && p->x && p->out, "layer_norm: null pointer");
  ALG_REQUIRE(p->d % 8 == 0 && p->d > 0 && p->d <= ops::kRowThreads * 8 * ops::kMaxChunks,
              "layer_norm: d must be a multiple of 8 and <= 8192");
  ALG_REQUIRE(p->rows >= 0 && p->rows <= 0x7fffffff, "layer_norm: bad row count");
  ALG_REQUIRE((p->weight == nullptr) == (p->bias == nullptr), "layer_norm: weight and bias come together");
  ALG_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "layer_norm: scale and shift come together");
  if (p->weight) ALG_REQUIRE(p->affine_dtype == ALG_F32 || p->affine_dtype == ALG_BF16, "layer_norm: affine dtype");
  if (p->scale) {
    ALG_REQUIRE(p->mod_dtype == ALG_F32 || p->mod_dtype == ALG_BF16, "layer_norm: modulation dtype");
    ALG_REQUIRE(p->rows_per_batch > 0, "layer_norm: rows_per_batch must be positive when modulating");
    ALG_REQUIRE(p->split_row == 0 || (p->scale_alt && p->shift_alt), "layer_norm: split_row needs scale_alt / shift_alt");
    ALG_REQUIRE(p->mod_batch_stride % 8 == 0, "layer_norm: modulation vectors must stay 16-byte aligned");
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  ALG_REQUIRE(al16(p->x) && al16(p->out) && al16(p->weight) && al16(p->bias) && al16(p->scale) && al16(p->shift) &&
                  al16(p->scale_alt) && al16(p->shift_alt),
              "layer_norm: pointers must be 16-byte aligned");
  if (int rc = alg_check_device()) return rc;
  if (p->rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static int th_knob = -1;
  if (th_knob < 0) {
    const char* e = getenv("ALG_LN_THREADS");
    th_knob = e ? atoi(e) : 128;
  }
  const bool f32_vectors = (!p->weight || p->affine_dtype == ALG_F32) && (!p->scale || p->mod_dtype == ALG_F32);
  const bool one_mod_set = !p->scale || (p->rows_per_batch >= p->rows && p->split_row == 0);
  if (th_knob != 256 && p->d == 128 * 5 * 8 && f32_vectors && one_mod_set && !p->chain_bf16) {  // Wan's d = 5120, fp32 chain
    ops::layer_norm_kernel<false, 128, 5, true><<<(unsigned)p->rows, 128, 0, st>>>(*p);
    ALG_LAUNCH_OK();
    return 0;
  }
#define ALG_LN(TH, CH)                                                                         \
  do {                                                                                         \
    if (p->chain_bf16) ops::layer_norm_kernel<true, TH, CH><<<(unsigned)p->rows, TH, 0, st>>>(*p);  \
    else ops::layer_norm_kernel<false, TH, CH><<<(unsigned)p->rows, TH, 0, st>>>(*p);           \
  } while (0)
  const int chunks = p->d / 8;
  if (th_knob == 256) {
    if (chunks <= 256 * 2) ALG_LN(256, 2);
    else if (chunks <= 256 * 3) ALG_LN(256, 3);
    else ALG_LN(256, 4);
  } else {
    if (chunks <= 128 * 3) ALG_LN(128, 3);
    else if (chunks <= 128 * 5) ALG_LN(128, 5);
    else ALG_LN(128, 8);
  }
#undef ALG_LN
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_head_norm_rope(const alg_head_norm_rope_t* p, void* stream) {
  ALG_REQUIRE(p && p->x, "head_norm_rope: null pointer");
  ALG_REQUIRE(p->head_dim == 64 || p->head_dim == 128, "head_norm_rope: head_dim must be 64 or 128");
  ALG_REQUIRE(p->heads > 0 && p->rows >= 0 && p->ld >= (int64_t)p->heads * p->head_dim && p->ld % 8 == 0,
              "head_norm_rope: bad shape (row stride must keep rows 16-byte aligned)");
  ALG_REQUIRE((int64_t)p->heads * p->head_dim <= 8192 && p->rows <= 0x7fffffff, "head_norm_rope: heads * head_dim <= 8192");
  ALG_REQUIRE(p->norm_kind >= ALG_NORM_NONE && p->norm_kind <= ALG_NORM_LAYER, "head_norm_rope: unknown norm kind");
  ALG_REQUIRE(p->norm_kind == ALG_NORM_NONE || p->weight, "head_norm_rope: the norm needs a weight");
  ALG_REQUIRE((p->cos == nullptr) == (p->sin == nullptr), "head_norm_rope: cos and sin come together");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->weight) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->cos) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->sin) & 15) == 0,
              "head_norm_rope: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  if (p->rows == 0) return 0;
  alg_head_norm_rope_t q = *p;
  if (q.rows_per_batch <= 0) q.rows_per_batch = q.rows;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int chunks = q.heads * q.head_dim / 8;
#define ALG_HNR(HD, CH) ops::head_norm_rope_kernel<HD, 128, CH><<<(unsigned)q.rows, 128, 0, st>>>(q)
  if (q.head_dim == 64) {
    if (chunks <= 128 * 3) ALG_HNR(64, 3);
    else if (chunks <= 128 * 5) ALG_HNR(64, 5);
    else ALG_HNR(64, 8);
  } else {
    if (chunks <= 128 * 3) ALG_HNR(128, 3);
    else if (chunks <= 128 * 5) ALG_HNR(128, 5);
    else ALG_HNR(128, 8);
  }
#undef ALG_HNR
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_patch_gather(const alg_patch_src_t* srcs, int n_pass, int n_src, int T, int H, int W, void* A,
                                int64_t lda, void* stream) {
  ALG_REQUIRE(srcs && A, "patch_gather: null pointer");
  ALG_REQUIRE(n_pass >= 1 && n_src >= 1 && n_pass * n_src <= ops::kMaxSrc, "patch_gather: at most 3 passes x 2 sources");
  ALG_REQUIRE(T >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "patch_gather: grid must be even in H and W");
  ops::GatherArgs a{};
  a.n_pass = n_pass; a.n_src = n_src; a.T = T; a.H = H; a.W = W; a.lda = lda;
  for (int p = 0; p < n_pass; ++p) {
    int ch = 0;
    for (int s = 0; s < n_src; ++s) {
      const alg_patch_src_t& d = srcs[p * n_src + s];
      ALG_REQUIRE(d.ptr && d.channels > 0, "patch_gather: empty source");
      ALG_REQUIRE(d.dtype == ALG_F32 || d.dtype == ALG_BF16 || d.dtype == ALG_F16, "patch_gather: unsupported dtype");
      a.src[p * n_src + s] = d;
      ch += d.channels;
    }
    ALG_REQUIRE(p == 0 || ch == a.ch_total, "patch_gather: passes disagree on the channel count");
    a.ch_total = ch;
  }
  ALG_REQUIRE(lda >= (int64_t)a.ch_total * 4 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 7) == 0,
              "patch_gather: lda too small or A misaligned");
  if (int rc = alg_check_device()) return rc;
  const int64_t total = (int64_t)n_pass * T * (H / 2) * (W / 2) * a.ch_total;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  ops::patch_gather_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, reinterpret_cast<__nv_bfloat16*>(A));
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_unpatchify(const void* proj, int64_t ld, void* out, int n_pass, int C, int T, int H, int W,
                              int64_t s_pass, int64_t sc, int64_t st_, int64_t sy, int channel_major, void* stream) {
  ALG_REQUIRE(proj && out, "unpatchify: null pointer");
  ALG_REQUIRE(n_pass >= 1 && C >= 1 && T >= 1 && H % 2 == 0 && W % 2 == 0 && ld >= 4 * (int64_t)C, "unpatchify: bad shape");
  if (int rc = alg_check_device()) return rc;
  const int64_t total = (int64_t)n_pass * C * T * H * W;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  ops::unpatchify_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(proj), ld, reinterpret_cast<__nv_bfloat16*>(out), n_pass, C, T, H, W, s_pass,
      sc, st_, sy, channel_major);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_timestep_embedding(float timestep, int dim, void* out, int dtype, void* stream) {
  ALG_REQUIRE(out && dim > 0 && dim % 2 == 0, "timestep_embedding: bad arguments");
  ALG_REQUIRE(dtype == ALG_F32 || dtype == ALG_BF16, "timestep_embedding: dtype must be f32 or bf16");
  if (int rc = alg_check_device()) return rc;
  ops::timestep_embedding_kernel<<<(dim / 2 + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(timestep, dim,
                                                                                                            out, dtype);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_elementwise_bf16(int op, const void* a, const void* b, void* out, int64_t n, void* stream) {
  ALG_REQUIRE(a && out && n >= 0, "elementwise: null pointer");
  ALG_REQUIRE(op >= ALG_EW_ADD && op <= ALG_EW_GELU_TANH, "elementwise: unknown op");
  ALG_REQUIRE(op != ALG_EW_ADD || b, "elementwise: ADD needs b");
  if (int rc = alg_check_device()) return rc;
  if (n == 0) return 0;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  ops::elementwise_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      op, reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
      reinterpret_cast<__nv_bfloat16*>(out), n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_mean_rows_bf16(const void* x, int64_t rows, int d, int64_t ld, void* out, void* stream) {
  ALG_REQUIRE(x && out && rows > 0 && d > 0 && d % 8 == 0 && ld % 8 == 0 && ld >= d, "mean_rows: bad arguments");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "mean_rows: x must be 16-byte aligned");
  if (int rc = alg_check_device()) return rc;
  ops::mean_rows_kernel<<<d / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), rows, d, ld, reinterpret_cast<__nv_bfloat16*>(out));
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_copy_rows_bf16(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows, int d,
                                  void* stream) {
  ALG_REQUIRE(src && dst, "copy_rows: null pointer");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
              "copy_rows: pointers must be 16-byte aligned");
  if (int rc = alg_check_device()) return rc;
  return dit::copy_rows(reinterpret_cast<const __nv_bfloat16*>(src), src_ld, reinterpret_cast<__nv_bfloat16*>(dst), dst_ld,
                        rows, d, reinterpret_cast<cudaStream_t>(stream));
}
