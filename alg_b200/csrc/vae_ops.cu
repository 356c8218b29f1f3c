// Building blocks of the CogVideoX VAE encoder (AutoencoderKLCogVideoX.encode), which the CogVideoX ALG pipeline calls
// on the per-step path when the low-pass filter runs in pixel space (cog:257 -- encode of the filtered image -- and
// cog:166 for the conditioning image): channels-last activations [T*H*W, C] bf16,
//   alg_im2col_bf16      causal 3-D / strided 2-D patch gather, so every convolution is ONE alg_gemm_bf16 (fp32 TMEM
//                        accumulation over all kt*kh*kw*C products, bias + optional residual in the GEMM epilogue)
//   alg_group_norm_bf16  GroupNorm (+ SiLU) with fp32 statistics, two passes over the activation
// Both are HBM-bound streaming kernels: 16-byte accesses, grids sized in multiples of the SM count.
#include "common.cuh"

namespace alg {
namespace vae {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
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

// ------------------------------------------------------------------------------------------------
// im2col: cols[(to, ho, wo)][(it, ih, iw, c)] = x[clamp_t(to * st + it - pad_t)][ho * sh + ih - pad_top][wo * sw + iw - pad_left][c]
// (zero outside the H x W frame; frames before t = 0 replicate frame 0 like CogVideoXCausalConv3d without a cache).
// One thread owns one 16-byte chunk column of the K axis (8 channels of one tap: decoded once) and walks the pixels of the
// block's tile, so the per-chunk work is a bounds test and two adds; consecutive threads write consecutive chunks of a
// row (coalesced stores) and read contiguous runs of C * 2 bytes.  (First version: one flat index per chunk with six
// integer divisions, 2.1 TB/s; the gather is 62 % of the encoder, profiles/r01_vae.md.)
// ------------------------------------------------------------------------------------------------
constexpr int kIm2colTile = 64;  // output pixels per block

__global__ void __launch_bounds__(512) im2col_kernel(const alg_im2col_t p, int c8, int k8, int ld8, int64_t M) {
  const uint4* __restrict__ x = reinterpret_cast<const uint4*>(p.x);
  uint4* __restrict__ cols = reinterpret_cast<uint4*>(p.cols);
  const int64_t m0 = (int64_t)blockIdx.x * kIm2colTile;
  const int n = (int)min((int64_t)kIm2colTile, M - m0);
  const int wo0 = (int)(m0 % p.Wo);
  const int64_t r0 = m0 / p.Wo;
  const int ho0 = (int)(r0 % p.Ho), to0 = (int)(r0 / p.Ho);
  for (int kc = threadIdx.x; kc < ld8; kc += blockDim.x) {
    uint4* dst = cols + m0 * ld8 + kc;
    if (kc >= k8) {  // zero tail of a padded row
      for (int i = 0; i < n; ++i, dst += ld8) *dst = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const int tap = kc / c8, cc = kc - tap * c8;
    const int it = tap / (p.kh * p.kw), r2 = tap - it * (p.kh * p.kw);
    const int ih = r2 / p.kw, iw = r2 - ih * p.kw;
    int wo = wo0, ho = ho0, to = to0;
    for (int i = 0; i < n; ++i, dst += ld8) {
      int t = to * p.st + it - p.pad_t;
      t = t < 0 ? 0 : t;
      const int y = ho * p.sh + ih - p.pad_top, xx = wo * p.sw + iw - p.pad_left;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (y >= 0 && y < p.H && xx >= 0 && xx < p.W) v = __ldg(x + (((int64_t)t * p.H + y) * p.W + xx) * c8 + cc);
      *dst = v;
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

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: thread = one 16-byte chunk column (8 channels) x a strided set of rows; the chunk's 8 channels
// belong to one group (C / groups >= 8) or to 8 / cpg consecutive groups (cpg = 4, 2, 1).  fp32 partial sums per thread,
// fp64 from the block reduction on (shared-memory, then global atomics), so E[x^2] - mean^2 is safe.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_norm_stats_kernel(const uint4* __restrict__ x, int64_t rows, int c8, int cpg,
                                                               int groups, double* __restrict__ stats) {
  extern __shared__ double sh[];  // [groups][2]
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sh[i] = 0.0;
  __syncthreads();
  const int rows_per_iter = blockDim.x / c8;  // launch guarantees c8 <= blockDim.x and blockDim.x % c8 == 0
  const int cc = threadIdx.x % c8, rr = threadIdx.x / c8;
  const int sub = cpg >= 8 ? 1 : 8 / cpg;  // groups touched by one chunk
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
#pragma unroll 4
  for (int64_t r = (int64_t)blockIdx.x * rows_per_iter + rr; r < rows; r += (int64_t)gridDim.x * rows_per_iter) {
    float f[8];
    unpack8(__ldg(x + r * c8 + cc), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += f[e];
      q[e] = fmaf(f[e], f[e], q[e]);
    }
  }
  const int g0 = cpg >= 8 ? (cc * 8) / cpg : cc * sub;
  const int per = 8 / sub;  // channels of the chunk per group
  for (int j = 0; j < sub; ++j) {
    double ds = 0.0, dq = 0.0;
    for (int e = 0; e < per; ++e) {
      ds += (double)s[j * per + e];
      dq += (double)q[j * per + e];
    }
    atomicAdd(&sh[(g0 + j) * 2], ds);
    atomicAdd(&sh[(g0 + j) * 2 + 1], dq);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) atomicAdd(&stats[i], sh[i]);
}

// y = a * x + b with a = rstd_g * gamma_c, b = beta_c - a * mean_g (the form ATen's CUDA GroupNorm forward uses), fp32,
// rounded to bf16; optional SiLU on the rounded value (x / (1 + exp(-x)) in fp32, rounded again).
__global__ void __launch_bounds__(256) group_norm_apply_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                               const __nv_bfloat16* __restrict__ gamma,
                                                               const __nv_bfloat16* __restrict__ beta, int64_t rows,
                                                               int c8, int cpg, float eps, int silu,
                                                               const double* __restrict__ stats) {
  const int rows_per_iter = blockDim.x / c8;
  const int cc = threadIdx.x % c8, rr = threadIdx.x / c8;
  const double n = (double)rows * (double)cpg;
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cc * 8 + e, g = c / cpg;
    const double mean = stats[2 * g] / n;
    double var = stats[2 * g + 1] / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = rsqrtf((float)var + eps);
    const float ga = gamma ? __bfloat162float(gamma[c]) : 1.f, be = beta ? __bfloat162float(beta[c]) : 0.f;
    a[e] = rstd * ga;
    b[e] = be - a[e] * (float)mean;
  }
#pragma unroll 4
  for (int64_t r = (int64_t)blockIdx.x * rows_per_iter + rr; r < rows; r += (int64_t)gridDim.x * rows_per_iter) {
    float f[8];
    unpack8(__ldg(x + r * c8 + cc), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = __bfloat162float(__float2bfloat16_rn(fmaf(a[e], f[e], b[e])));
      if (silu) v = v / (1.f + expf(-v));
      f[e] = v;
    }
    y[r * c8 + cc] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder pieces (AutoencoderKLCogVideoX.decode, cog:428-433): channels-last [T*H*W, C] bf16 like the encoder.
//   upsample_nearest_kernel      F.interpolate(mode="nearest") on (T, H, W): src = floor(dst * in / out)
//   spatial_norm_apply_kernel    CogVideoXSpatialNorm3D after the GroupNorm: out = bf16(bf16(norm_f * y) + b) (+ SiLU), with
//                                (y | b) = [conv_y(zq) | conv_b(zq)] computed at LATENT resolution ([zt*zh*zw, 2C]) and gathered
//                                at the nearest latent pixel -- a 1x1x1 convolution commutes with nearest resizing, so the
//                                resized zq and the two full-resolution convolution outputs are never materialised
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int c8,
                                                               int Ti, int Hi, int Wi, int To, int Ho, int Wo, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    int64_t r = i / c8;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho), to = (int)(r / Ho);
    const int ti = (int)((int64_t)to * Ti / To), yi = (int)((int64_t)yo * Hi / Ho), xi = (int)((int64_t)xo * Wi / Wo);
    out[i] = __ldg(x + (((int64_t)ti * Hi + yi) * Wi + xi) * c8 + c);
  }
}

__global__ void __launch_bounds__(256) spatial_norm_apply_kernel(const uint4* __restrict__ fn, const uint4* __restrict__ yb,
                                                                 uint4* __restrict__ out, int c8, int T, int H, int W, int zt,
                                                                 int zh, int zw, int silu, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    int64_t r = i / c8;
    const int xo = (int)(r % W);
    r /= W;
    const int yo = (int)(r % H), to = (int)(r / H);
    const int64_t zr = ((int64_t)((int64_t)to * zt / T) * zh + (int64_t)yo * zh / H) * zw + (int64_t)xo * zw / W;
    float f[8], y[8], b[8];
    unpack8(__ldg(fn + i), f);
    unpack8(__ldg(yb + zr * 2 * c8 + c), y);
    unpack8(__ldg(yb + zr * 2 * c8 + c8 + c), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = bf16_round(bf16_round(f[e] * y[e]) + b[e]);
      if (silu) v = v / (1.0f + __expf(-v));
      f[e] = v;
    }
    out[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------
// Spatially zero-padded raster for the implicit (patch-matrix-free) convolution, see alg_gemm_t.a_tap_kblocks:
// frames [F, H + 2, W + 2, C] with the image in [1, H] x [1, W].  to_padded: compact [n*H*W, C] -> interior of padded frames
// (borders are not written: zero the buffer once); else padded rows of `ld` elements -> compact [n*H*W, C] (+ bf16 residual:
// dst = bf16(float(src) + float(residual)), the rounding chain of the GEMM's residual epilogue).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_frames_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         const __nv_bfloat16* __restrict__ res, int n, int H, int W, int C, int ld,
                                                         int to_padded) {
  const int64_t total = (int64_t)n * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int xx = (int)(pix % W);
    const int64_t r = pix / W;
    const int y = (int)(r % H), t = (int)(r / H);
    const int64_t prow = ((int64_t)t * (H + 2) + y + 1) * (W + 2) + xx + 1;
    if (to_padded) {
      dst[prow * ld + c] = src[i];
    } else {
      float v = __bfloat162float(src[prow * ld + c]);
      if (res) v += __bfloat162float(res[i]);
      dst[i] = __float2bfloat16_rn(v);
    }
  }
}
__global__ void __launch_bounds__(256) pad_frames_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                             const uint4* __restrict__ res, int n, int H, int W, int c8, int ld8,
                                                             int to_padded) {
  const int64_t total = (int64_t)n * H * W * c8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8);
    const int64_t pix = i / c8;
    const int xx = (int)(pix % W);
    const int64_t r = pix / W;
    const int y = (int)(r % H), t = (int)(r / H);
    const int64_t prow = ((int64_t)t * (H + 2) + y + 1) * (W + 2) + xx + 1;
    if (to_padded) {
      dst[prow * ld8 + c] = __ldg(src + i);
    } else {
      uint4 v = __ldg(src + prow * ld8 + c);
      if (res) {
        float a[8], b[8];
        unpack8(v, a);
        unpack8(__ldg(res + i), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += b[e];
        v = pack8(a);
      }
      dst[i] = v;
    }
  }
}

}  // namespace vae
}  // namespace alg

extern "C" int alg_im2col_bf16(const alg_im2col_t* p, void* stream) {
  using namespace alg;
  ALG_REQUIRE(p && p->x && p->cols, "im2col: null pointer");
  ALG_REQUIRE(p->T > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->C % 8 == 0, "im2col: C must be a positive multiple of 8");
  ALG_REQUIRE(p->kt > 0 && p->kh > 0 && p->kw > 0 && p->st > 0 && p->sh > 0 && p->sw > 0, "im2col: bad kernel / stride");
  ALG_REQUIRE(p->To > 0 && p->Ho > 0 && p->Wo > 0 && p->pad_t >= 0 && p->pad_top >= 0 && p->pad_left >= 0, "im2col: bad output geometry");
  const int64_t K = (int64_t)p->kt * p->kh * p->kw * p->C;
  ALG_REQUIRE(p->ld >= K && p->ld % 8 == 0, "im2col: ld must be >= kt*kh*kw*C and a multiple of 8");
  ALG_REQUIRE((p->To - 1) * p->st + p->kt - 1 - p->pad_t < p->T, "im2col: temporal window runs past the last frame");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->cols)) & 15) == 0, "im2col: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  const int64_t M = (int64_t)p->To * p->Ho * p->Wo;
  const int ld8 = (int)(p->ld / 8);
  const int threads = std::min(512, (ld8 + 31) / 32 * 32);
  const int64_t grid = (M + vae::kIm2colTile - 1) / vae::kIm2colTile;
  ALG_REQUIRE(grid <= 0x7fffffff, "im2col: too many output pixels");
  vae::im2col_kernel<<<(unsigned)grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, p->C / 8, (int)(K / 8), ld8, M);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_group_norm_bf16(const alg_group_norm_t* p, void* stream) {
  using namespace alg;
  ALG_REQUIRE(p && p->x && p->y && p->stats, "group_norm: null pointer");
  ALG_REQUIRE(p->rows > 0 && p->C > 0 && p->groups > 0 && p->C % p->groups == 0, "group_norm: bad shape");
  const int cpg = p->C / p->groups, c8 = p->C / 8;
  ALG_REQUIRE(p->C % 8 == 0 && c8 <= 256 && 256 % c8 == 0, "group_norm: C must be 8 * a divisor of 256");
  ALG_REQUIRE(cpg % 8 == 0 || 8 % cpg == 0, "group_norm: channels per group must divide or be a multiple of 8");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->y) | reinterpret_cast<uintptr_t>(p->stats)) & 15) == 0,
              "group_norm: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ALG_CUDA_OK(cudaMemsetAsync(p->stats, 0, sizeof(double) * 2 * p->groups, st));
  const int rows_per_iter = 256 / c8;
  const int grid = (int)std::min<int64_t>((p->rows + rows_per_iter - 1) / rows_per_iter, (int64_t)148 * 8);
  vae::group_norm_stats_kernel<<<grid, 256, sizeof(double) * 2 * p->groups, st>>>(
      reinterpret_cast<const uint4*>(p->x), p->rows, c8, cpg, p->groups, p->stats);
  ALG_LAUNCH_OK();
  vae::group_norm_apply_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(p->x), reinterpret_cast<uint4*>(p->y),
                                                     reinterpret_cast<const __nv_bfloat16*>(p->weight),
                                                     reinterpret_cast<const __nv_bfloat16*>(p->bias), p->rows, c8, cpg,
                                                     p->eps, p->silu, p->stats);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_upsample_nearest_bf16(const void* x, void* out, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo, void* stream) {
  ALG_REQUIRE(x && out && C > 0 && C % 8 == 0 && Ti > 0 && Hi > 0 && Wi > 0 && To > 0 && Ho > 0 && Wo > 0, "upsample_nearest: bad arguments (C % 8 == 0)");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "upsample_nearest: 16-byte alignment");
  if (int rc = alg_check_device()) return rc;
  const int64_t n = (int64_t)To * Ho * Wo * (C / 8);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  alg::vae::upsample_nearest_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), C / 8, Ti, Hi, Wi, To, Ho, Wo, n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_spatial_norm_apply_bf16(const void* f_norm, const void* yb, void* out, int C, int T, int H, int W, int zt, int zh,
                                           int zw, int silu, void* stream) {
  ALG_REQUIRE(f_norm && yb && out && C > 0 && C % 8 == 0 && T > 0 && H > 0 && W > 0 && zt > 0 && zh > 0 && zw > 0,
              "spatial_norm_apply: bad arguments (C % 8 == 0)");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(f_norm) | reinterpret_cast<uintptr_t>(yb) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
              "spatial_norm_apply: 16-byte alignment");
  if (int rc = alg_check_device()) return rc;
  const int64_t n = (int64_t)T * H * W * (C / 8);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  alg::vae::spatial_norm_apply_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(f_norm), reinterpret_cast<const uint4*>(yb), reinterpret_cast<uint4*>(out), C / 8, T, H, W, zt,
      zh, zw, silu, n);
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_pad_frames_bf16(const void* src, void* dst, const void* residual, int frames, int H, int W, int C, int64_t ld,
                                   int to_padded, void* stream) {
  ALG_REQUIRE(src && dst && frames > 0 && H > 0 && W > 0 && C > 0 && ld >= C, "pad_frames: bad arguments");
  ALG_REQUIRE(!(residual && to_padded), "pad_frames: the residual belongs to the padded -> compact direction");
  if (int rc = alg_check_device()) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = C % 8 == 0 && ld % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0;
  const int64_t n = (int64_t)frames * H * W * (vec ? C / 8 : C);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  if (vec)
    alg::vae::pad_frames_vec_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                                          reinterpret_cast<const uint4*>(residual), frames, H, W, C / 8, (int)(ld / 8), to_padded);
  else
    alg::vae::pad_frames_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst),
                                                      reinterpret_cast<const __nv_bfloat16*>(residual), frames, H, W, C, (int)ld, to_padded);
  ALG_LAUNCH_OK();
  return 0;
}
