// HBM-bound kernels of the Wan DiT forward: patch gather, RMSNorm(+RoPE) (LayerNorm lives in dit_ops.cu),
// timestep embedding, modulation tables, unpatchify.  All fp32 statistics, bf16 activations, with the
// intermediate roundings placed where diffusers' WanTransformer3DModel places them.
#include "dit_kernels.cuh"

namespace alg {
namespace dit {

// ------------------------------------------------------------------------------------------------
__global__ void patch_gather_kernel(const CondPtrs cond, int n_pass, int lat_ch, int cond_ch, int T, int H, int W, __nv_bfloat16* __restrict__ A) {
  const int ph = H / 2, pw = W / 2;
  const int64_t N = (int64_t)T * ph * pw;
  const int Kdim = (lat_ch + cond_ch) * 4;
  const int64_t total = (int64_t)n_pass * N * Kdim;
  const int64_t plane = (int64_t)H * W;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % Kdim);
    const int64_t tok = idx / Kdim;
    const int pss = (int)(tok / N);
    const int64_t n = tok - (int64_t)pss * N;
    const int c = k >> 2, i = (k >> 1) & 1, j = k & 1;
    const int t = (int)(n / (ph * pw));
    const int rem = (int)(n - (int64_t)t * ph * pw);
    const int y = rem / pw, x = rem - y * pw;
    const float* src = c < lat_ch ? cond.lat[pss] + (int64_t)c * T * plane : cond.p[pss] + (int64_t)(c - lat_ch) * T * plane;
    A[idx] = __float2bfloat16_rn(src[(int64_t)t * plane + (int64_t)(2 * y + i) * W + 2 * x + j]);
  }
}

int patch_gather(CondPtrs cond_dev, int n_pass, int lat_ch, int cond_ch, int T, int H, int W,
                 __nv_bfloat16* A, cudaStream_t st) {
  const int64_t total = (int64_t)n_pass * T * (H / 2) * (W / 2) * (lat_ch + cond_ch) * 4;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  patch_gather_kernel<<<grid, 256, 0, st>>>(cond_dev, n_pass, lat_ch, cond_ch, T, H, W, A);
  ALG_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm across heads (+ RoPE).  One row per block, the row in registers, packed fp32x2 arithmetic (IEEE rn per lane).
//
// RoPE: diffusers multiplies in complex128 (`view_as_complex(x.to(float64)) * freqs`) and casts back to bf16.  The first
// version did exactly that in fp64 and was bound by the f32 <-> f64 conversions on the XU pipe (16 lanes/clk, 67 % busy:
// 0.82 ms per launch against 0.41 ms without RoPE, profiles/r01_lowpass_norm_ncu.txt).  Now each of re*cos - im*sin and
// re*sin + im*cos is evaluated in double-float arithmetic: cos / sin are split into fp32 (hi, lo) pairs when the row's
// table entries are staged in shared memory, the two leading products carry their exact FMA residuals, their sum goes
// through TwoSum, and all the low-order terms are added before the single final fp32 rounding.  That is ~2^-44 relative
// accuracy: after the cast to bf16 the result equals the fp64 one (0 differences in 4e7 random outputs on the host model
// of this sequence, tests/test_host_logic.py; a plain fp32 FMA evaluation differs in 2e-5 of the outputs).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 bf16x2_round(float2 x) { return __bfloat1622float2(__float22bfloat162_rn(x)); }

// a * x + b * y per lane with x = xh + xl, y = yh + yl (double-float), one final rounding to fp32.  `na`, `nb` are -a, -b:
// every subtraction is an FFMA2 with the constant pair (-1, -1) or a product formed with the negated multiplicand, so no
// packed negation (two scalar instructions each) is ever materialised -- the first form of this function made the kernel
// instruction-issue bound (549 M warp instructions per launch, ncu norm_r41).
__device__ __forceinline__ float2 dd_dot2(float2 a, float2 na, float2 xh, float2 xl, float2 b, float2 nb, float2 yh,
                                          float2 yl) {
  const float2 m1 = make_float2(-1.f, -1.f);
  const float2 p1n = __fmul2_rn(na, xh), e1 = __ffma2_rn(a, xh, p1n);  // a*xh = -p1n + e1 exactly
  const float2 p2n = __fmul2_rn(nb, yh), e2 = __ffma2_rn(b, yh, p2n);
  const float2 sn = __fadd2_rn(p1n, p2n);                               // TwoSum(p1n, p2n) = sn + errn exactly
  const float2 bbn = __ffma2_rn(p1n, m1, sn);
  const float2 errn = __fadd2_rn(__ffma2_rn(__ffma2_rn(bbn, m1, sn), m1, p1n), __ffma2_rn(bbn, m1, p2n));
  const float2 low = __fadd2_rn(__ffma2_rn(errn, m1, __fadd2_rn(e1, e2)), __ffma2_rn(a, xl, __fmul2_rn(b, yl)));
  return __ffma2_rn(sn, m1, low);                                       // -(sn + errn) + e1 + e2 + a*xl + b*yl
}

template <int TH, int CH>
__device__ __forceinline__ float block_sum_t(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < TH / 32; ++i) t += red[i];
  return t;
}

// HD > 0: head_dim is the compile-time HD (the rotary-pair index of a chunk is a mask, not an integer division per chunk);
// EXACT: d == TH * CH * 8, so no chunk is ever out of range.  The instruction count matters: with RoPE the kernel is
// issue-bound (profiles/r01_lowpass_norm_ncu.txt norm_r41: 549 M warp instructions, 0.50 of the HBM peak against 0.93 without
// RoPE), and a third of its instructions were bf16 <-> fp32 repacking, runtime `% head_dim` and range guards.
// `bf16(bf16(x * rstd) * w)`: the second product is ONE packed bf16 multiply (HMUL2.BF16: exact 16-bit product, one rounding --
// the same value as the fp32 multiply + cast it replaces).
template <int TH, int CH, int HD, bool EXACT>
__global__ void __launch_bounds__(TH)
    rms_norm_rope_kernel(__nv_bfloat16* __restrict__ x, int d, int head_dim_rt, float eps,
                         const __nv_bfloat16* __restrict__ w, RopeTables rope, int use_rope) {
  __shared__ float red[TH / 32];
  // per rotary pair p: (cos_hi, cos_hi, cos_lo, cos_lo) and (sin_hi, sin_hi, sin_lo, sin_lo), the packed operands as stored,
  // at slot p + (p >> 3): consecutive threads read pairs 4 apart, and the one-in-eight padding spreads a quarter warp's
  // eight 16-byte reads over all 32 banks (a dense [pair] layout is 4-way conflicted, an interleaved one 8-way)
  __shared__ float4 cs_c[72], cs_s[72];
  const int head_dim = HD > 0 ? HD : head_dim_rt;
  const int64_t row = blockIdx.x;
  uint4* xr = reinterpret_cast<uint4*>(x + row * d);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const int chunks = EXACT ? TH * CH : d / 8;
  float2 v[CH][4];
  float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int ci = threadIdx.x + c * TH;
    if (EXACT || ci < chunks) {
      const uint4 u = xr[ci];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[c][e] = __bfloat1622float2(h[e]);
        sq2 = __ffma2_rn(v[c][e], v[c][e], sq2);
      }
    }
  }
  // the row's head_dim / 2 (cos, sin) pairs are the same for every head: stage them once per block in shared memory
  // (the fp64 table reads were 4x the activation bytes when every head re-read them), split into fp32 hi + lo
  if (use_rope) {
    const int pi = threadIdx.x;
    if (pi < head_dim / 2) {
      const int64_t N = (int64_t)rope.ppf * rope.pph * rope.ppw;
      const int n = (int)(row % N);
      const int t = n / (rope.pph * rope.ppw);
      const int rem = n - t * rope.pph * rope.ppw;
      const int y = rem / rope.ppw, xx = rem - y * rope.ppw;
      const double* cs;
      if (pi < rope.n_t) cs = rope.t + ((int64_t)t * rope.n_t + pi) * 2;
      else if (pi < rope.n_t + rope.n_h) cs = rope.h + ((int64_t)y * rope.n_h + (pi - rope.n_t)) * 2;
      else cs = rope.w + ((int64_t)xx * rope.n_w + (pi - rope.n_t - rope.n_h)) * 2;
      const double c = cs[0], s = cs[1];
      const float ch = (float)c, sh = (float)s;
      const float cl = (float)(c - (double)ch), sl = (float)(s - (double)sh);
      cs_c[pi + (pi >> 3)] = make_float4(ch, ch, cl, cl);
      cs_s[pi + (pi >> 3)] = make_float4(sh, sh, sl, sl);
    }
  }
  const float rstd = rsqrtf(block_sum_t<TH, CH>(sq2.x + sq2.y, red) / (float)d + eps);  // its __syncthreads also publishes cs_c / cs_s
  const float2 rstd2 = make_float2(rstd, rstd);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int ci = threadIdx.x + c * TH;
    if (EXACT || ci < chunks) {
      const uint4 wu = __ldg(wr + ci);
      const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wu);
      uint4 u;
      __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e)  // hidden_states.to(weight.dtype) * weight: two bf16 roundings
        ob[e] = __hmul2(__float22bfloat162_rn(__fmul2_rn(v[c][e], rstd2)), wh[e]);
      if (use_rope) {
        const int pair0 = ((ci * 8) % head_dim) >> 1;  // 4 complex pairs per chunk, never straddling a head
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int slot = pair0 + q + ((pair0 + q) >> 3);
          const float4 tc = cs_c[slot], ts = cs_s[slot];
          const float2 o = __bfloat1622float2(ob[q]);
          const float re = o.x, im = o.y;
          // (re, im) * (cos, cos) + (-im, re) * (sin, sin): lane 0 = re*cos - im*sin, lane 1 = im*cos + re*sin; rounded to
          // bf16 (.type_as) by the pack below
          ob[q] = __float22bfloat162_rn(dd_dot2(o, make_float2(-re, -im), make_float2(tc.x, tc.y), make_float2(tc.z, tc.w),
                                                make_float2(-im, re), make_float2(im, -re), make_float2(ts.x, ts.y),
                                                make_float2(ts.z, ts.w)));
        }
      }
      xr[ci] = u;
    }
  }
}

int rms_norm_rope(__nv_bfloat16* x, int64_t rows, int d, int head_dim, float eps, const __nv_bfloat16* w,
                  const RopeTables* rope, cudaStream_t st) {
  ALG_REQUIRE(d % 8 == 0 && d <= 128 * 8 * 8 && head_dim % 8 == 0 && head_dim <= 128, "rms_norm: unsupported width");
  ALG_REQUIRE(rows <= 0x7fffffff, "rms_norm: too many rows");
  if (rows == 0) return 0;
  RopeTables r{};
  if (rope) r = *rope;
  const int chunks = d / 8;
  const int ur = rope != nullptr;
  if (head_dim == 128 && chunks == 128 * 5) rms_norm_rope_kernel<128, 5, 128, true><<<(unsigned)rows, 128, 0, st>>>(x, d, head_dim, eps, w, r, ur);
  else if (chunks <= 128 * 3) rms_norm_rope_kernel<128, 3, 0, false><<<(unsigned)rows, 128, 0, st>>>(x, d, head_dim, eps, w, r, ur);
  else if (chunks <= 128 * 5) rms_norm_rope_kernel<128, 5, 0, false><<<(unsigned)rows, 128, 0, st>>>(x, d, head_dim, eps, w, r, ur);
  else rms_norm_rope_kernel<128, 8, 0, false><<<(unsigned)rows, 128, 0, st>>>(x, d, head_dim, eps, w, r, ur);
  ALG_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void timestep_sinusoid_kernel(float timestep, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    // exponent = -log(10000) * arange(half) / half ; emb = t * exp(exponent) ; [cos | sin]
    const float exponent = (-9.210340371976184f * (float)i) / (float)half;
    const float e = timestep * expf(exponent);
    out[i] = cosf(e);
    out[half + i] = sinf(e);
  }
}
int timestep_sinusoid(float timestep, int dim, float* out, cudaStream_t st) {
  timestep_sinusoid_kernel<<<1, 128, 0, st>>>(timestep, dim, out);
  ALG_LAUNCH_OK();
  return 0;
}

__global__ void gemv_f32_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ x,
                                float* __restrict__ out, int out_f, int in_f, int act) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= out_f) return;
  const float* wr = W + (int64_t)warp * in_f;
  float acc = 0.f;
  if ((in_f & 3) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(wr);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int i = lane; i < in_f / 4; i += 32) {
      const float4 a = __ldg(w4 + i), c = x4[i];
      acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    }
  } else {
    for (int i = lane; i < in_f; i += 32) acc += wr[i] * x[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    float v = acc + (b ? b[warp] : 0.f);
    if (act == 1) v = v / (1.0f + expf(-v));
    out[warp] = v;
  }
}
int gemv_f32(const float* W, const float* b, const float* x, float* out, int out_f, int in_f, int act,
             cudaStream_t st) {
  const int warps_per_block = 8;
  gemv_f32_kernel<<<(out_f + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(W, b, x, out, out_f,
                                                                                                   in_f, act);
  ALG_LAUNCH_OK();
  return 0;
}

__global__ void temb_finish_kernel(const float* __restrict__ v, __nv_bfloat16* __restrict__ temb,
                                   __nv_bfloat16* __restrict__ silu_temb, int d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
    const __nv_bfloat16 t = __float2bfloat16_rn(v[i]);
    const float f = __bfloat162float(t);
    temb[i] = t;
    silu_temb[i] = __float2bfloat16_rn(f / (1.0f + expf(-f)));
  }
}
int temb_finish(const float* v, __nv_bfloat16* temb, __nv_bfloat16* silu_temb, int d, cudaStream_t st) {
  temb_finish_kernel<<<(d + 255) / 256, 256, 0, st>>>(v, temb, silu_temb, d);
  ALG_LAUNCH_OK();
  return 0;
}

__global__ void add_table_kernel(const float* __restrict__ table, const __nv_bfloat16* __restrict__ src,
                                 float* __restrict__ mod, int rows, int d, int broadcast) {
  const int total = rows * d;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % d;
    mod[i] = __fadd_rn(table[i], __bfloat162float(src[broadcast ? c : i]));
  }
}
int add_table(const float* table, const __nv_bfloat16* src, float* mod, int rows, int d, int broadcast,
              cudaStream_t st) {
  add_table_kernel<<<(rows * d + 255) / 256, 256, 0, st>>>(table, src, mod, rows, d, broadcast);
  ALG_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ proj, __nv_bfloat16* __restrict__ out, int n_pass,
                                  int C, int T, int H, int W) {
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
    const int k = (((y & 1) * 2) + (x & 1)) * C + c;
    out[idx] = proj[((int64_t)pss * N + n) * (4 * C) + k];
  }
}
int unpatchify(const __nv_bfloat16* proj, __nv_bfloat16* out, int n_pass, int C, int T, int H, int W,
               cudaStream_t st) {
  const int64_t total = (int64_t)n_pass * C * T * H * W;
  unpatchify_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(proj, out, n_pass, C, T, H, W);
  ALG_LAUNCH_OK();
  return 0;
}

__global__ void copy_rows_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld, __nv_bfloat16* __restrict__ dst,
                                 int64_t dst_ld, int64_t rows, int d) {
  const int chunks = d / 8;
  const int64_t total = rows * chunks;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / chunks;
    const int c = (int)(idx - r * chunks);
    reinterpret_cast<uint4*>(dst + r * dst_ld)[c] = reinterpret_cast<const uint4*>(src + r * src_ld)[c];
  }
}
int copy_rows(const __nv_bfloat16* src, int64_t src_ld, __nv_bfloat16* dst, int64_t dst_ld, int64_t rows, int d,
              cudaStream_t st) {
  ALG_REQUIRE(d % 8 == 0 && src_ld % 8 == 0 && dst_ld % 8 == 0, "copy_rows: widths must be multiples of 8");
  const int64_t total = rows * (d / 8);
  if (total == 0) return 0;
  copy_rows_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, st>>>(src, src_ld, dst, dst_ld, rows, d);
  ALG_LAUNCH_OK();
  return 0;
}

}  // namespace dit
}  // namespace alg

extern "C" int alg_wan_rms_norm_rope(void* x, int64_t rows, int d, int head_dim, float eps, const void* weight,
                                     const double* rope_t, const double* rope_h, const double* rope_w, int n_t, int n_h,
                                     int n_w, int ppf, int pph, int ppw, void* stream) {
  using namespace alg;
  ALG_REQUIRE(x && weight && rows >= 0, "wan_rms_norm_rope: null pointer");
  ALG_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(weight)) & 15) == 0, "wan_rms_norm_rope: misaligned pointer");
  if (int rc = alg_check_device()) return rc;
  dit::RopeTables r{rope_t, rope_h, rope_w, n_t, n_h, n_w, ppf, pph, ppw};
  if (rope_t) {
    ALG_REQUIRE(rope_h && rope_w && n_t + n_h + n_w == head_dim / 2 && ppf > 0 && pph > 0 && ppw > 0,
                "wan_rms_norm_rope: the three axis tables must cover head_dim / 2 rotary pairs");
  }
  return dit::rms_norm_rope(reinterpret_cast<__nv_bfloat16*>(x), rows, d, head_dim, eps,
                            reinterpret_cast<const __nv_bfloat16*>(weight), rope_t ? &r : nullptr,
                            reinterpret_cast<cudaStream_t>(stream));
}

