// Fused CFG combine + scheduler.step kernels (HBM-bound, one elementwise pass).
//
//   alg_cfg_unipc_step  wan:919-927   (CFG in bf16 with three roundings + UniPCMultistepScheduler.step)
//   alg_cfg_ddim_step   cog:1091-1123 (fp32 CFG + CogVideoXDDIMScheduler.step + cast back)
//   alg_cfg_euler_step  hy:1254-1270  (true CFG + FlowMatchEulerDiscreteScheduler.step on frames 1.. + re-prepend)
//   alg_cfg_dpm_step    cog:1091-1123 (fp32 CFG + CogVideoXDPMScheduler.step: SDE DPM-Solver++ 2M, two noise draws)
//
// Every intermediate that PyTorch materialises as a tensor is rounded here at the same point
// (__fmul_rn/__fadd_rn keep ptxas from contracting mul+add into FMA), so that given identical
// noise predictions the update is bit-identical to the eager sequence it replaces.
#include "common.cuh"

namespace alg {

template <int NDT>
__device__ __forceinline__ float cfg_combine(const void* __restrict__ noise, int64_t i, int64_t E, int n_pass,
                                             float w, bool fp32) {
  if (n_pass == 1) return Elem<NDT>::load(noise, i);
  float u0 = Elem<NDT>::load(noise, i);
  float u = n_pass == 3 ? Elem<NDT>::load(noise, E + i) : u0;
  float t = Elem<NDT>::load(noise, (int64_t)(n_pass - 1) * E + i);
  if (fp32) return __fadd_rn(u0, __fmul_rn(w, __fsub_rn(t, u)));
  float d = Elem<NDT>::round(__fsub_rn(t, u));
  float g = Elem<NDT>::round(__fmul_rn(w, d));
  return Elem<NDT>::round(__fadd_rn(u0, g));
}

template <int NDT>
__global__ void __launch_bounds__(256) unipc_kernel(const void* __restrict__ noise, const float* __restrict__ x,
                                                    float* __restrict__ x_out, float* __restrict__ last_sample,
                                                    const float* __restrict__ m_prev0, float* __restrict__ m_prev1,
                                                    int64_t E, alg_unipc_step_t p) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = cfg_combine<NDT>(noise, i, E, p.n_pass, p.guidance, p.cfg_fp32 != 0);
    float xs = x[i];
    // convert_model_output: x0 = sample - sigma_t * model_output  (0-dim fp32 * noise-dtype tensor -> noise dtype)
    const float m_t = __fsub_rn(xs, Elem<NDT>::round(__fmul_rn(p.sigma_t, v)));
    float m0 = 0.f, m1 = 0.f;
    if (p.use_corrector || p.order_p == 2) m0 = m_prev0[i];
    if (p.use_corrector) {  // multistep_uni_c_bh_update
      const float xl = last_sample[i];
      const float x_t_ = __fsub_rn(__fmul_rn(p.c_ratio, xl), __fmul_rn(p.c_a, m0));
      float corr = 0.f;
      if (p.order_c == 2) {
        m1 = m_prev1[i];
        const float d1 = __fmul_rn(__fsub_rn(m1, m0), p.c_rk_inv);
        corr = __fmul_rn(p.c_rho0, d1);
      }
      const float d1t = __fsub_rn(m_t, m0);
      const float inner = __fadd_rn(corr, __fmul_rn(p.c_rho_last, d1t));
      xs = __fsub_rn(x_t_, __fmul_rn(p.c_b, inner));
    }
    last_sample[i] = xs;
    m_prev1[i] = m_t;  // becomes model_outputs[-1]; the caller swaps the two history pointers
    // multistep_uni_p_bh_update (history now: m0' = m_t, m1' = old m0)
    const float x_t_ = __fsub_rn(__fmul_rn(p.p_ratio, xs), __fmul_rn(p.p_a, m_t));
    float out = x_t_;
    if (p.order_p == 2) {
      const float d1 = __fmul_rn(__fsub_rn(m0, m_t), p.p_rk_inv);
      out = __fsub_rn(x_t_, __fmul_rn(p.p_b, __fmul_rn(p.p_rho0, d1)));
    }
    x_out[i] = out;
  }
}

template <int NDT, int SDT>
__global__ void __launch_bounds__(256) ddim_kernel(const void* __restrict__ noise, const void* __restrict__ x,
                                                   void* __restrict__ x_out, int64_t E, int n_pass, float w, float sa,
                                                   float sb, float a, float b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = cfg_combine<NDT>(noise, i, E, n_pass, w, true);  // noise_pred.float() first (cog:1091)
    const float xs = Elem<SDT>::load(x, i);
    // pred_x0 = sqrt(a_t) * sample [sample dtype] - sqrt(1 - a_t) * v [fp32]
    const float pred = __fsub_rn(Elem<SDT>::round(__fmul_rn(sa, xs)), __fmul_rn(sb, v));
    // prev = a * sample [sample dtype] + b * pred [fp32]; then .to(sample dtype) (cog:1123)
    const float prev = __fadd_rn(Elem<SDT>::round(__fmul_rn(a, xs)), __fmul_rn(b, pred));
    Elem<SDT>::store(x_out, i, prev);
  }
}

// CogVideoXDPMScheduler.step (v-prediction).  Tensors and their dtypes as eager PyTorch materialises them:
//   pred_x0  = sa * sample [S] - sb * v [f32]                                   -> f32, returned (next step's old_pred)
//   prev     = m0 * sample [S] - m1 * pred_x0 [f32] + mn * noise1 [S]           -> f32
//   second-order (old_pred given, not the last step):
//   d        = m2 * pred_x0 - m3 * old_pred                                     -> f32
//   prev     = m0 * sample [S] - m1 * d + mn * noise2 [S]                       -> f32;  then .to(S) (cog:1123)
template <int NDT, int SDT>
__global__ void __launch_bounds__(256) dpm_kernel(const void* __restrict__ noise, const void* __restrict__ x,
                                                  void* __restrict__ x_out, const float* __restrict__ old_pred,
                                                  float* __restrict__ pred_out, const void* __restrict__ rnd, int64_t E,
                                                  alg_dpm_step_t p) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = cfg_combine<NDT>(noise, i, E, p.n_pass, p.guidance, true);  // noise_pred.float() first (cog:1091)
    const float xs = Elem<SDT>::load(x, i);
    const float pred = __fsub_rn(Elem<SDT>::round(__fmul_rn(p.sqrt_alpha_t, xs)), __fmul_rn(p.sqrt_beta_t, v));
    const float mx = Elem<SDT>::round(__fmul_rn(p.m0, xs));
    const float nz = Elem<SDT>::round(__fmul_rn(p.m_noise, Elem<SDT>::load(rnd, i)));
    const float den = p.second_order ? __fsub_rn(__fmul_rn(p.m2, pred), __fmul_rn(p.m3, old_pred[i])) : pred;
    const float prev = __fadd_rn(__fsub_rn(mx, __fmul_rn(p.m1, den)), nz);
    pred_out[i] = pred;
    Elem<SDT>::store(x_out, i, prev);
  }
}

template <int NDT>
__global__ void __launch_bounds__(256) euler_kernel(const void* __restrict__ noise, const float* __restrict__ x,
                                                    float* __restrict__ x_out, const float* __restrict__ first, int C,
                                                    int T, int64_t HW, int n_pass, float w, float dt) {
  const int64_t E = (int64_t)C * T * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / (T * HW);
    const int64_t f = (i - c * T * HW) / HW;
    if (f == 0) {  // torch.cat([image_latents, latents], dim=2) (hy:1270)
      x_out[i] = first[c * HW + (i - c * T * HW)];
      continue;
    }
    const float v = cfg_combine<NDT>(noise, i, E, n_pass, w, false);
    // sample.float() + dt * v  [0-dim fp32 * noise-dtype tensor -> noise dtype];  .to(v.dtype)
    const float prev = __fadd_rn(x[i], Elem<NDT>::round(__fmul_rn(dt, v)));
    x_out[i] = Elem<NDT>::round(prev);
  }
}

inline int ew_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 8); }

}  // namespace alg

extern "C" int alg_cfg_unipc_step(const void* noise, int noise_dtype, const float* x, float* x_out,
                                  float* last_sample, const float* m_prev0, float* m_prev1, int64_t E,
                                  const alg_unipc_step_t* p, void* stream) {
  using namespace alg;
  ALG_REQUIRE(noise && x && x_out && last_sample && m_prev0 && m_prev1 && p, "unipc_step: null pointer");
  ALG_REQUIRE(p->n_pass >= 1 && p->n_pass <= 3, "unipc_step: n_pass must be 1, 2 or 3");
  ALG_REQUIRE(p->order_p == 1 || p->order_p == 2, "unipc_step: predictor order must be 1 or 2");
  ALG_REQUIRE(!p->use_corrector || p->order_c == 1 || p->order_c == 2, "unipc_step: corrector order must be 1 or 2");
  if (E == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (noise_dtype == ALG_BF16)
    unipc_kernel<ALG_BF16><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, last_sample, m_prev0, m_prev1, E, *p);
  else if (noise_dtype == ALG_F32)
    unipc_kernel<ALG_F32><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, last_sample, m_prev0, m_prev1, E, *p);
  else if (noise_dtype == ALG_F16)
    unipc_kernel<ALG_F16><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, last_sample, m_prev0, m_prev1, E, *p);
  else
    ALG_REQUIRE(false, "unipc_step: unsupported noise dtype");
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_cfg_ddim_step(const void* noise, int noise_dtype, const void* x, void* x_out, int sample_dtype,
                                 int64_t E, int n_pass, float guidance, float sqrt_alpha_t, float sqrt_beta_t, float a,
                                 float b, void* stream) {
  using namespace alg;
  ALG_REQUIRE(noise && x && x_out, "ddim_step: null pointer");
  ALG_REQUIRE(n_pass >= 1 && n_pass <= 3, "ddim_step: n_pass must be 1, 2 or 3");
  if (E == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ALG_DDIM(N, S)                                                                                              \
  ddim_kernel<N, S><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, E, n_pass, guidance, sqrt_alpha_t, sqrt_beta_t, a, b)
  if (noise_dtype == ALG_BF16 && sample_dtype == ALG_BF16) ALG_DDIM(ALG_BF16, ALG_BF16);
  else if (noise_dtype == ALG_F32 && sample_dtype == ALG_F32) ALG_DDIM(ALG_F32, ALG_F32);
  else if (noise_dtype == ALG_BF16 && sample_dtype == ALG_F32) ALG_DDIM(ALG_BF16, ALG_F32);
  else if (noise_dtype == ALG_F32 && sample_dtype == ALG_BF16) ALG_DDIM(ALG_F32, ALG_BF16);
  else if (noise_dtype == ALG_F16 && sample_dtype == ALG_F16) ALG_DDIM(ALG_F16, ALG_F16);
  else ALG_REQUIRE(false, "ddim_step: unsupported dtype combination");
#undef ALG_DDIM
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_cfg_euler_step(const void* noise, int noise_dtype, const float* x, float* x_out,
                                  const float* first_frame, int C, int T, int64_t HW, int n_pass, float guidance,
                                  float dt, void* stream) {
  using namespace alg;
  ALG_REQUIRE(noise && x && x_out && first_frame, "euler_step: null pointer");
  ALG_REQUIRE(n_pass >= 1 && n_pass <= 3 && C > 0 && T > 0 && HW > 0, "euler_step: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t E = (int64_t)C * T * HW;
  if (noise_dtype == ALG_BF16)
    euler_kernel<ALG_BF16><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, first_frame, C, T, HW, n_pass, guidance, dt);
  else if (noise_dtype == ALG_F32)
    euler_kernel<ALG_F32><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, first_frame, C, T, HW, n_pass, guidance, dt);
  else if (noise_dtype == ALG_F16)
    euler_kernel<ALG_F16><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, first_frame, C, T, HW, n_pass, guidance, dt);
  else
    ALG_REQUIRE(false, "euler_step: unsupported noise dtype");
  ALG_LAUNCH_OK();
  return 0;
}

extern "C" int alg_cfg_dpm_step(const void* noise, int noise_dtype, const void* x, void* x_out, int sample_dtype,
                                const float* old_pred, float* pred_out, const void* rnd, int64_t E,
                                const alg_dpm_step_t* p, void* stream) {
  using namespace alg;
  ALG_REQUIRE(noise && x && x_out && pred_out && rnd && p, "dpm_step: null pointer");
  ALG_REQUIRE(p->n_pass >= 1 && p->n_pass <= 3, "dpm_step: n_pass must be 1, 2 or 3");
  ALG_REQUIRE(!p->second_order || old_pred, "dpm_step: the second-order update needs old_pred");
  if (E == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ALG_DPM(N, S) dpm_kernel<N, S><<<ew_grid(E), 256, 0, st>>>(noise, x, x_out, old_pred, pred_out, rnd, E, *p)
  if (noise_dtype == ALG_BF16 && sample_dtype == ALG_BF16) ALG_DPM(ALG_BF16, ALG_BF16);
  else if (noise_dtype == ALG_F32 && sample_dtype == ALG_F32) ALG_DPM(ALG_F32, ALG_F32);
  else if (noise_dtype == ALG_BF16 && sample_dtype == ALG_F32) ALG_DPM(ALG_BF16, ALG_F32);
  else if (noise_dtype == ALG_F32 && sample_dtype == ALG_BF16) ALG_DPM(ALG_F32, ALG_BF16);
  else ALG_REQUIRE(false, "dpm_step: unsupported dtype combination");
#undef ALG_DPM
  ALG_LAUNCH_OK();
  return 0;
}
