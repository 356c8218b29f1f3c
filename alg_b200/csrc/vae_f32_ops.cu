// Building blocks of the float32 video VAEs (AutoencoderKLWan: wan:429-434 condition encode, wan:526 per-step encode in
// pixel-space ALG, wan:959 decode; run.py:51-55 loads it in float32).  Activations are channels-last [T*H*W, C] fp32.  Every
// convolution is a patch gather that writes the bf16 3-term split [hi | hi | lo] of the fp32 patch directly (the gather and the
// split are one pass), followed by ONE alg_gemm_bf16 against the weight's [hi | lo | hi] split with fp32 accumulation and fp32
// output: hi*hi + hi*lo + lo*hi = the fp32 product up to 2^-16 (closer to fp32 than the TF32 convolutions PyTorch runs by default).
//
//   alg_im2col_split3_f32   causal (ZERO front padding, WanCausalConv3d) / strided / nearest-x2-upsampled patch gather + split
//   alg_rms_norm_cl_f32     WanRMS_norm over the channel axis (+ SiLU): x / max(||x||, 1e-12) * sqrt(C) * gamma
//   alg_softmax_rows_f32    softmax(scale * s) over rows (WanAttentionBlock: one head of C channels, scores via the split GEMM)
//   alg_nchw_to_cl_f32 / alg_cl_to_nchw_f32   [C, P] <-> [P, ld] layout changes at the VAE boundary (+ clamp on the way out)
// All HBM-bound streaming kernels; grids sized in multiples of the SM count.
#include <algorithm>

#include "common.cuh"

namespace alg {
namespace vae32 {

__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
  hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
}

constexpr int kTile = 32;  // output pixels per block

// Vector path (C % 4 == 0): a thread owns one 4-channel group of the K axis (tap decoded once) and walks the tile's pixels;
// consecutive threads write consecutive 8-byte chunks of a row in each of the three K sections.
__global__ void __launch_bounds__(256) im2col_split3_vec_kernel(const alg_im2col_f32_t p, int c4, int k4, int ld4, int64_t M) {
  const float4* __restrict__ x = reinterpret_cast<const float4*>(p.x);
  uint2* __restrict__ cols = reinterpret_cast<uint2*>(p.cols);
  const int64_t m0 = (int64_t)blockIdx.x * kTile;
  const int n = (int)min((int64_t)kTile, M - m0);
  const int wo0 = (int)(m0 % p.Wo);
  const int64_t r0 = m0 / p.Wo;
  const int ho0 = (int)(r0 % p.Ho), to0 = (int)(r0 / p.Ho) + p.to0;
  const int HL = p.H * p.up, WL = p.W * p.up;  // logical (upsampled) frame
  for (int kc = threadIdx.x; kc < ld4; kc += blockDim.x) {
    uint2* dst = cols + m0 * 3 * ld4 + kc;
    if (kc >= k4) {  // zero tail of a padded row
      for (int i = 0; i < n; ++i, dst += 3 * ld4) dst[0] = dst[ld4] = dst[2 * ld4] = make_uint2(0u, 0u);
      continue;
    }
    const int tap = kc / c4, cc = kc - tap * c4;
    const int it = tap / (p.kh * p.kw), r2 = tap - it * (p.kh * p.kw);
    const int ih = r2 / p.kw, iw = r2 - ih * p.kw;
    int wo = wo0, ho = ho0, to = to0;
    for (int i = 0; i < n; ++i, dst += 3 * ld4) {
      int t = to * p.st + it - p.pad_t;
      int y = ho * p.sh + ih - p.pad_top, xx = wo * p.sw + iw - p.pad_left;
      if (p.replicate) {  // F.pad(mode="replicate") in all three dimensions (HunyuanVideoCausalConv3d)
        t = max(t, p.t_min);
        y = min(max(y, 0), HL - 1);
        xx = min(max(xx, 0), WL - 1);
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t >= p.t_min && y >= 0 && y < HL && xx >= 0 && xx < WL)
        v = __ldg(x + (((int64_t)(p.tdup == 2 ? (t + 1) >> 1 : t) * p.H + y / p.up) * p.W + xx / p.up) * c4 + cc);
      uint2 hi, lo;
      split_pair(v.x, v.y, hi.x, lo.x);
      split_pair(v.z, v.w, hi.y, lo.y);
      dst[0] = hi;
      dst[ld4] = hi;
      dst[2 * ld4] = lo;
      if (++wo == p.Wo) {
        wo = 0;
        if (++ho == p.Ho) {
          ho = 0;
          ++to;
        }
      }
    }
  }
}

// Scalar path (any C; the RGB input convolution of the encoder): one thread per (pixel, k) element.
__global__ void __launch_bounds__(256) im2col_split3_scalar_kernel(const alg_im2col_f32_t p, int K, int64_t M) {
  const float* __restrict__ x = reinterpret_cast<const float*>(p.x);
  __nv_bfloat16* __restrict__ cols = reinterpret_cast<__nv_bfloat16*>(p.cols);
  const int HL = p.H * p.up, WL = p.W * p.up;
  const int64_t n = M * p.ld;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % p.ld);
    const int64_t m = i / p.ld;
    float v = 0.f;
    if (k < K) {
      const int tap = k / p.C, c = k - tap * p.C;
      const int it = tap / (p.kh * p.kw), r2 = tap - it * (p.kh * p.kw);
      const int ih = r2 / p.kw, iw = r2 - ih * p.kw;
      const int wo = (int)(m % p.Wo);
      const int64_t r = m / p.Wo;
      const int ho = (int)(r % p.Ho), to = (int)(r / p.Ho) + p.to0;
      int t = to * p.st + it - p.pad_t;
      int y = ho * p.sh + ih - p.pad_top, xx = wo * p.sw + iw - p.pad_left;
      if (p.replicate) {
        t = max(t, p.t_min);
        y = min(max(y, 0), HL - 1);
        xx = min(max(xx, 0), WL - 1);
      }
      if (t >= p.t_min && y >= 0 && y < HL && xx >= 0 && xx < WL)
        v = x[(((int64_t)(p.tdup == 2 ? (t + 1) >> 1 : t) * p.H + y / p.up) * p.W + xx / p.up) * p.C + c];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* row = cols + m * 3 * p.ld;
    row[k] = hi;
    row[p.ld + k] = hi;
    row[2 * p.ld + k] = lo;
  }
}

// WanRMS_norm on channels-last rows: one warp per pixel row.  F.normalize(x, dim=C) * sqrt(C) * gamma (+ bias) (+ SiLU).
__global__ void __launch_bounds__(256) rms_norm_cl_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int C,
                                                          float scale, const float* __restrict__ gamma,
                                                          const float* __restrict__ bias, int silu) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* xr = x + r * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = xr[c];
      ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    float* orow = out + r * C;
    for (int c = lane; c < C; c += 32) {
      float v = __fdiv_rn(xr[c], denom) * scale * gamma[c];
      if (bias) v += bias[c];
      if (silu) v = v / (1.0f + expf(-v));
      orow[c] = v;
    }
  }
}

// in place: row <- softmax(scale * row).  One block per row (6 240 columns at the Wan 480p latent size).
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int cols_all, int64_t ld, float scale, int block) {
  __shared__ float red[8];
  float* row = x + (int64_t)blockIdx.x * ld;
  // frame-causal mask (HunyuanVideo VAE mid block): row r of frame r / block sees the columns of frames <= its own; the masked
  // tail of the row is zeroed
  const int cols = block > 0 ? min(cols_all, (int)(blockIdx.x / block + 1) * block) : cols_all;
  for (int c = cols + threadIdx.x; c < cols_all; c += 256) row[c] = 0.f;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, row[c] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float e = expf(row[c] * scale - m);
    row[c] = e;
    s += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  const float inv = 1.0f / s;
  for (int c = threadIdx.x; c < cols; c += 256) row[c] *= inv;
}

// nn.GroupNorm on channels-last fp32 rows of ONE sample (statistics over rows x channels-of-the-group), + optional SiLU.
// Pass 1: per-block partial sums per group in fp64 -> global atomics; pass 2: y = (x - mean) * rstd * weight + bias.
__global__ void __launch_bounds__(256) group_norm_f32_stats_kernel(const float* __restrict__ x, int64_t rows, int C, int cpg,
                                                                   double* __restrict__ stats) {
  extern __shared__ double sh[];  // [groups][2]
  const int groups = C / cpg;
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const int64_t n = rows * C;
  // a thread keeps one channel (blockDim * gridDim is a multiple of C by construction), so its sums belong to one group
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float s = 0.f, q = 0.f;
  double ds = 0.0, dq = 0.0;
  int cnt = 0;
  for (int64_t i = i0; i < n; i += stride) {
    const float v = x[i];
    s += v;
    q = fmaf(v, v, q);
    if (++cnt == 64) {  // bound the fp32 partial sums
      ds += (double)s;
      dq += (double)q;
      s = q = 0.f;
      cnt = 0;
    }
  }
  ds += (double)s;
  dq += (double)q;
  if (i0 < n) {
    const int g = (int)(i0 % C) / cpg;
    atomicAdd(&sh[2 * g], ds);
    atomicAdd(&sh[2 * g + 1], dq);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) atomicAdd(&stats[i], sh[i]);
}
__global__ void __launch_bounds__(256) group_norm_f32_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows,
                                                                   int C, int cpg, float eps, const float* __restrict__ w,
                                                                   const float* __restrict__ b, int silu,
                                                                   const double* __restrict__ stats) {
  const int64_t n = rows * C;
  const double cnt = (double)rows * (double)cpg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), g = c / cpg;
    const double mean = stats[2 * g] / cnt;
    double var = stats[2 * g + 1] / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = rsqrtf((float)var + eps);
    float v = (x[i] - (float)mean) * rstd * (w ? w[c] : 1.f) + (b ? b[c] : 0.f);
    if (silu) v = v / (1.0f + expf(-v));
    y[i] = v;
  }
}

// ---- implicit-convolution operand (no patch matrix) ----------------------------------------------------------------------
// The stride-1 3x3x3 convolutions read their input as a ZERO-PADDED channels-last clip in split form: S3P = bf16
// [(T + pt) * (H + 2) * (W + 2), Cs], row of pixel (t, y, x) = ((t + pt) * (H + 2) + y + 1) * (W + 2) + x + 1, columns
// [hi(C) | hi(C) | lo(C) | 0...] (Cs = 3C rounded up to 64).  alg_gemm_bf16's tap mode then walks the 27 taps as row-shifted
// reads of this one buffer.  This kernel is the producer: (optional WanRMS_norm) (+ SiLU) + split of every interior pixel, read
// from a compact [T*H*W, C] or a padded-raster [(T+pt)(H+2)(W+2), C] fp32 activation; the padding rows / columns of the output
// are never written (the buffer is zeroed once).
__global__ void __launch_bounds__(256) norm_split_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int T, int H,
                                                             int W, int C, int pt, int in_padded, int Cs, float scale,
                                                             const float* __restrict__ gamma, int silu) {
  // one warp per pixel; a lane owns channel PAIRS (2 lane + 64 k, + 1): 8-byte loads, 4-byte bf16x2 stores, the row stays in
  // registers between the reduction and the write (C <= 512)
  const int lane = threadIdx.x & 31;
  const int64_t pixels = (int64_t)T * H * W;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int kIters = 8;
  for (int64_t pix = warp; pix < pixels; pix += nwarps) {
    const int xx = (int)(pix % W);
    const int64_t r = pix / W;
    const int y = (int)(r % H), t = (int)(r / H);
    const int64_t prow = ((int64_t)(t + pt) * (H + 2) + y + 1) * (W + 2) + xx + 1;
    const float* xr = x + (in_padded ? prow : pix) * C;
    float2 v[kIters];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kIters; ++k) {
      const int c = 2 * lane + 64 * k;
      v[k] = c < C ? *reinterpret_cast<const float2*>(xr + c) : make_float2(0.f, 0.f);
      ss = fmaf(v[k].x, v[k].x, ss);
      ss = fmaf(v[k].y, v[k].y, ss);
    }
    float mul = 1.f;
    if (gamma) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      mul = fmaxf(sqrtf(ss), 1e-12f);
    }
    __nv_bfloat16* orow = out + prow * Cs;
#pragma unroll
    for (int k = 0; k < kIters; ++k) {
      const int c = 2 * lane + 64 * k;
      if (c < C) {
        float a = v[k].x, b = v[k].y;
        if (gamma) {
          const float2 g = *reinterpret_cast<const float2*>(gamma + c);
          a = __fdiv_rn(a, mul) * scale * g.x;
          b = __fdiv_rn(b, mul) * scale * g.y;
        }
        if (silu) {
          a = a / (1.0f + expf(-a));
          b = b / (1.0f + expf(-b));
        }
        uint32_t hi, lo;
        split_pair(a, b, hi, lo);
        *reinterpret_cast<uint32_t*>(orow + c) = hi;
        *reinterpret_cast<uint32_t*>(orow + C + c) = hi;
        *reinterpret_cast<uint32_t*>(orow + 2 * C + c) = lo;
      }
    }
  }
}
// F.pad(mode="replicate") of the padded split operand (HunyuanVideoCausalConv3d): every PADDING pixel of the raster
// [(T + pt), H + 2, W + 2, Cs] -- the pt front frames and the one-pixel border -- takes the row of the nearest interior pixel
// (all three coordinates clamped, so sources are always interior pixels written by alg_norm_split_pad_f32).
__global__ void __launch_bounds__(256) replicate_border_kernel(uint4* __restrict__ buf, int T, int H, int W, int pt, int cs8) {
  const int64_t pixels = (int64_t)(T + pt) * (H + 2) * (W + 2);
  const int64_t n = pixels * cs8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cs8);
    const int64_t pix = i / cs8;
    const int xx = (int)(pix % (W + 2));
    const int64_t r = pix / (W + 2);
    const int y = (int)(r % (H + 2)), t = (int)(r / (H + 2));
    const int ts = max(t, pt), ys = min(max(y, 1), H), xs = min(max(xx, 1), W);
    if (ts == t && ys == y && xs == xx) continue;  // interior
    buf[i] = buf[(((int64_t)ts * (H + 2) + ys) * (W + 2) + xs) * cs8 + c];
  }
}

// interior pixels of a compact [T*H*W, C] clip <-> the padded raster [(T+pt)(H+2)(W+2), C] (fp32); padding is not touched
__global__ void pad_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int H, int W, int C, int pt,
                                int to_padded) {
  const int64_t n = (int64_t)T * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int xx = (int)(pix % W);
    const int64_t r = pix / W;
    const int y = (int)(r % H), t = (int)(r / H);
    const int64_t j = (((int64_t)(t + pt) * (H + 2) + y + 1) * (W + 2) + xx + 1) * C + c;
    if (to_padded) dst[j] = src[i];
    else dst[i] = src[j];
  }
}

// [C, P] (channel-major: one sample of [B, C, T, H, W]) -> [P, ld] channels-last, columns >= C zeroed
__global__ void nchw_to_cl_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int64_t P, int ld) {
  const int64_t n = P * ld;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const int64_t pix = i / ld;
    out[i] = c < C ? x[(int64_t)c * P + pix] : 0.f;
  }
}
// [P, ld] channels-last -> [C, P]; clamp to [lo, hi] when lo < hi
__global__ void cl_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int64_t P, int ld, float lo, float hi) {
  const int64_t n = P * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i % P;
    const int c = (int)(i / P);
    float v = x[pix * ld + c];
    if (lo < hi) v = fminf(fmaxf(v, lo), hi);
    out[i] = v;
  }
}

// out = wa * a + wb * b (fp32): the linear cross-fade between overlapping temporal tiles of a decoded clip (blend_t)
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n, float wa, float wb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a[i], wa), __fmul_rn(b[i], wb));
}

static int grid_for(int64_t n, int per = 256) { return (int)std::min<int64_t>((n + per - 1) / per, 148 * 16); }

}  // namespace vae32
}  // namespace alg

using namespace alg;

extern "C" int alg_im2col_split3_f32(const alg_im2col_f32_t* p, void* stream) {
  ALG_REQUIRE(p && p->x && p->cols, "im2col_split3: null pointer");
  ALG_REQUIRE(p->T > 0 && p->H > 0 && p->W > 0 && p->C > 0, "im2col_split3: empty input");
  ALG_REQUIRE(p->kt > 0 && p->kh > 0 && p->kw > 0 && p->st > 0 && p->sh > 0 && p->sw > 0, "im2col_split3: bad kernel / stride");
  ALG_REQUIRE(p->To > 0 && p->Ho > 0 && p->Wo > 0 && p->to0 >= 0 && p->pad_t >= 0 && p->pad_top >= 0 && p->pad_left >= 0,
              "im2col_split3: bad output geometry");
  ALG_REQUIRE(p->up == 1 || p->up == 2, "im2col_split3: up must be 1 or 2");
  const int64_t K = (int64_t)p->kt * p->kh * p->kw * p->C;
  ALG_REQUIRE(p->ld >= K && p->ld % 8 == 0, "im2col_split3: ld must be >= kt*kh*kw*C and a multiple of 8");
  ALG_REQUIRE(p->tdup == 1 || p->tdup == 2, "im2col_split3: tdup must be 1 or 2");
  ALG_REQUIRE((int64_t)(p->to0 + p->To - 1) * p->st + p->kt - 1 - p->pad_t < (p->tdup == 2 ? 2 * p->T - 1 : p->T),
              "im2col_split3: temporal window runs past the last frame");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->cols)) & 15) == 0, "im2col_split3: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t M = (int64_t)p->To * p->Ho * p->Wo;
  if (p->C % 4 == 0) {
    const int ld4 = (int)(p->ld / 4);
    const int threads = std::min(256, (ld4 + 31) / 32 * 32);
    const int64_t grid = (M + vae32::kTile - 1) / vae32::kTile;
    ALG_REQUIRE(grid <= 0x7fffffff, "im2col_split3: too many output pixels");
    vae32::im2col_split3_vec_kernel<<<(unsigned)grid, threads, 0, st>>>(*p, p->C / 4, (int)(K / 4), ld4, M);
  } else {
    vae32::im2col_split3_scalar_kernel<<<vae32::grid_for(M * p->ld), 256, 0, st>>>(*p, (int)K, M);
  }
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_rms_norm_cl_f32(const float* x, float* out, int64_t rows, int C, const float* gamma, const float* bias, int silu,
                                   void* stream) {
  ALG_REQUIRE(x && out && gamma && rows >= 0 && C > 0, "rms_norm_cl: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  vae32::rms_norm_cl_kernel<<<vae32::grid_for(rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, rows, C, sqrtf((float)C), gamma, bias, silu);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_softmax_rows_f32(float* x, int64_t rows, int cols, int64_t ld, float scale, int causal_block, void* stream) {
  ALG_REQUIRE(x && rows >= 0 && cols > 0 && ld >= cols && causal_block >= 0, "softmax_rows: bad arguments");
  ALG_REQUIRE(rows <= 0x7fffffff, "softmax_rows: too many rows");
  if (int rc = alg_check_device()) return rc;
  if (rows == 0) return 0;
  vae32::softmax_rows_kernel<<<(unsigned)rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, cols, ld, scale, causal_block);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_nchw_to_cl_f32(const float* x, float* out, int C, int64_t pixels, int ld, void* stream) {
  ALG_REQUIRE(x && out && C > 0 && pixels > 0 && ld >= C, "nchw_to_cl: bad arguments");
  if (int rc = alg_check_device()) return rc;
  vae32::nchw_to_cl_kernel<<<vae32::grid_for(pixels * ld), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, C, pixels, ld);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_cl_to_nchw_f32(const float* x, float* out, int C, int64_t pixels, int ld, float lo, float hi, void* stream) {
  ALG_REQUIRE(x && out && C > 0 && pixels > 0 && ld >= C, "cl_to_nchw: bad arguments");
  if (int rc = alg_check_device()) return rc;
  vae32::cl_to_nchw_kernel<<<vae32::grid_for(pixels * C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, C, pixels, ld, lo, hi);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_group_norm_f32(const float* x, float* y, int64_t rows, int C, int groups, float eps, const float* weight,
                                  const float* bias, int silu, double* stats, void* stream) {
  ALG_REQUIRE(x && y && stats && rows > 0 && C > 0 && groups > 0 && C % groups == 0 && groups <= 1024, "group_norm_f32: bad arguments");
  if (int rc = alg_check_device()) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ALG_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * groups, st));
  // grid * 256 must be a multiple of C so that a thread's strided elements stay in one channel
  int64_t unit = C;
  while (unit % 256) unit += C;  // lcm-ish: smallest multiple of C that is a multiple of 256 (C divides it)
  const int blocks_unit = (int)(unit / 256);
  const int64_t want = std::min<int64_t>((rows * C + 255) / 256, 148 * 8);
  const int grid = (int)std::max<int64_t>(blocks_unit, want / blocks_unit * blocks_unit);
  vae32::group_norm_f32_stats_kernel<<<grid, 256, sizeof(double) * 2 * groups, st>>>(x, rows, C, C / groups, stats);
  ALG_LAUNCH_OK();
  vae32::group_norm_f32_apply_kernel<<<vae32::grid_for(rows * C), 256, 0, st>>>(x, y, rows, C, C / groups, eps, weight, bias, silu, stats);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_norm_split_pad_f32(const float* x, void* out, int T, int H, int W, int C, int front_pad, int in_padded, int Cs,
                                      const float* gamma, int silu, void* stream) {
  ALG_REQUIRE(x && out && T > 0 && H > 0 && W > 0 && C > 0 && front_pad >= 0 && Cs >= 3 * C && Cs % 8 == 0, "norm_split_pad: bad arguments");
  ALG_REQUIRE(C % 2 == 0 && C <= 512, "norm_split_pad: C must be even and <= 512");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma)) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
              "norm_split_pad: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  const int64_t pixels = (int64_t)T * H * W;
  vae32::norm_split_pad_kernel<<<vae32::grid_for(pixels, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(out), T, H, W, C, front_pad, in_padded, Cs, sqrtf((float)C), gamma, silu);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_pad_copy_f32(const float* src, float* dst, int T, int H, int W, int C, int front_pad, int to_padded, void* stream) {
  ALG_REQUIRE(src && dst && T > 0 && H > 0 && W > 0 && C > 0 && front_pad >= 0, "pad_copy: bad arguments");
  if (int rc = alg_check_device()) return rc;
  vae32::pad_copy_kernel<<<vae32::grid_for((int64_t)T * H * W * C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, dst, T, H, W, C, front_pad, to_padded);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_axpby_f32(const float* a, const float* b, float* out, int64_t n, float wa, float wb, void* stream) {
  ALG_REQUIRE(a && b && out && n >= 0, "axpby: bad arguments");
  if (int rc = alg_check_device()) return rc;
  if (n == 0) return 0;
  vae32::axpby_kernel<<<vae32::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, out, n, wa, wb);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_replicate_border_bf16(void* buf, int T, int H, int W, int front_pad, int Cs, void* stream) {
  ALG_REQUIRE(buf && T > 0 && H > 0 && W > 0 && front_pad >= 0 && Cs > 0 && Cs % 8 == 0, "replicate_border: bad arguments");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "replicate_border: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  const int64_t n = (int64_t)(T + front_pad) * (H + 2) * (W + 2) * (Cs / 8);
  vae32::replicate_border_kernel<<<vae32::grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint4*>(buf), T, H, W, front_pad, Cs / 8);
  ALG_LAUNCH_OK();
  return 0;
}
