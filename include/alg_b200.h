/*
 * alg_b200.h -- C ABI of libalg_b200.so, the sm_100a implementation of the ALG
 * (Adaptive Low-pass Guidance) per-step denoise loop.
 *
 * The reference (choi403/ALG @ 3657bfd) has no FFI of its own: its hot path is
 * Python calling PyTorch/diffusers.  Each entry point below therefore names the
 * reference call site it replaces (file:line under /root/reference); the
 * Python-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no PyTorch types cross this boundary
 *   - every pointer argument is a DEVICE pointer unless its name ends in _host
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all work is enqueued on it, nothing synchronises unless stated
 *   - return value: 0 = OK, non-zero = error; text via alg_last_error()
 *   - not thread-safe per handle; one engine handle per device / process
 */
#ifndef ALG_B200_H_
#define ALG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALG_B200_ABI_VERSION 5

typedef enum { ALG_F32 = 0, ALG_BF16 = 1, ALG_F16 = 2 } alg_dtype_t;

int alg_abi_version(void);
/* Thread-local message of the last failing call in this thread ("" if none). */
const char* alg_last_error(void);
/* 0 when the current device is sm_100 (B200); non-zero + message otherwise. */
int alg_check_device(void);

/* ------------------------------------------------------------------------- */
/* Low-pass filters: lp_utils.apply_low_pass_filter (lp_utils.py:8-60)       */
/* ------------------------------------------------------------------------- */

/* lp_utils.py:49-54 -- two F.interpolate(bilinear, antialias=True) calls
 * (H,W)->(h1,w1)->(H,W) on `planes` independent H x W images (the 5-D view of
 * lp_utils.py:31-35 flattens to planes = B*C*K).  The intermediate small image
 * is rounded to `dtype` exactly where the reference materialises it.
 * in/out: [planes, H, W] contiguous, same dtype; out may alias in. */
int alg_lowpass_down_up(const void* in, void* out, int64_t planes, int H, int W, int h1, int w1,
                        int dtype, void* stream);

/* lp_utils.py:40-47 -> torchvision gaussian_blur: reflect pad k/2, dense k x k
 * depthwise conv with the outer-product Gaussian; the 1-D and 2-D weights are
 * built and rounded in `dtype` like the reference (bf16 weights sum to 0.99902
 * at k=13, sigma=15).  ksize odd, 1 <= ksize <= 63, ksize/2 < min(H, W).
 * out must NOT alias in. */
int alg_lowpass_gaussian(const void* in, void* out, int64_t planes, int H, int W, int ksize,
                         double sigma, int dtype, void* stream);

/* The dtype-faithful 1-D Gaussian taps used above, written to host memory. */
int alg_gaussian_kernel1d(int ksize, double sigma, int dtype, float* taps_host);

/* ------------------------------------------------------------------------- */
/* Fused CFG combine + scheduler.step                                         */
/* ------------------------------------------------------------------------- */

/* Host-computed scalars of one UniPC (bh2, predict-x0, flow-sigma) step. */
typedef struct {
  int32_t n_pass;         /* 1: no CFG; 2: u + w (t - u); 3: u0 + w (t - u)                */
  int32_t cfg_fp32;       /* 0: CFG in the noise dtype (Wan: three bf16 roundings, wan:919-924); 1: fp32 */
  float guidance;         /* w                                                           */
  float sigma_t;          /* convert_model_output: x0 = x - sigma_t * v                   */
  int32_t use_corrector;  /* step_index > 0                                              */
  int32_t order_c;        /* corrector order (1 or 2)                                    */
  float c_ratio;          /* sigma_t / sigma_s0                                          */
  float c_a;              /* alpha_t * h_phi_1                                           */
  float c_b;              /* alpha_t * B_h                                               */
  float c_rk_inv;         /* 1 / r_k           (order_c == 2)                            */
  float c_rho0;           /* rhos_c[0]         (order_c == 2)                            */
  float c_rho_last;       /* rhos_c[-1]                                                  */
  int32_t order_p;        /* predictor order (1 or 2)                                    */
  float p_ratio, p_a, p_b, p_rk_inv, p_rho0;
} alg_unipc_step_t;

/* wan:919-927 -- CFG combine + UniPCMultistepScheduler.step in one pass.
 *   noise      [n_pass, E]  noise_dtype (bf16 for Wan)       read
 *   x          [E] fp32     latents                           read, x_out written (may alias)
 *   last_sample[E] fp32     previous (corrected) sample       read if use_corrector, then overwritten
 *   m_prev0    [E] fp32     model_outputs[-1] of last step    read if corrector or order 2
 *   m_prev1    [E] fp32     model_outputs[-2] of last step    read if order_c == 2; OVERWRITTEN with
 *                           this step's x0 prediction (caller swaps the two pointers afterwards) */
int alg_cfg_unipc_step(const void* noise, int noise_dtype, const float* x, float* x_out, float* last_sample,
                       const float* m_prev0, float* m_prev1, int64_t E, const alg_unipc_step_t* p,
                       void* stream);

/* cog:1091-1123 -- fp32 CFG + CogVideoXDDIMScheduler.step (v-prediction) + cast back.
 *   noise [n_pass, E] noise_dtype; x, x_out [E] sample_dtype (bf16 for Cog).
 *   sqrt_alpha_t, sqrt_beta_t, a, b are the fp64 scalars of the scheduler rounded to fp32
 *   the way ATen passes CPU scalars to CUDA kernels. */
int alg_cfg_ddim_step(const void* noise, int noise_dtype, const void* x, void* x_out, int sample_dtype,
                      int64_t E, int n_pass, float guidance, float sqrt_alpha_t, float sqrt_beta_t,
                      float a, float b, void* stream);

/* Host-computed scalars of one CogVideoXDPMScheduler step (fp64 in the scheduler, handed to the kernel as fp32 the way
 * ATen hands 0-dim CPU tensors to CUDA kernels). */
typedef struct {
  int32_t n_pass;        /* 1: no CFG; 2: u + w (t - u); 3: u0 + w (t - u), in fp32 (cog:1091-1102)      */
  int32_t second_order;  /* old_pred_original_sample given and prev_timestep >= 0                       */
  float guidance;
  float sqrt_alpha_t, sqrt_beta_t; /* pred_x0 = sqrt(a_t) x - sqrt(1 - a_t) v                             */
  float m0, m1, m2, m3;  /* get_mult(): sqrt((1-a_prev)/(1-a_t)) e^-h ; expm1(-2h) sqrt(a_prev) ; 1 + 1/(2r) ; 1/(2r) */
  float m_noise;         /* sqrt(1 - a_prev) sqrt(1 - e^-2h)                                            */
} alg_dpm_step_t;

/* cog:1113-1123 -- fp32 CFG + CogVideoXDPMScheduler.step (SDE DPM-Solver++, v-prediction) + cast back.
 *   noise [n_pass, E] noise_dtype; x, x_out, rnd [E] sample_dtype (rnd = the N(0,1) draw THIS branch of the step uses:
 *   the reference draws once for the first-order update and once more for the second-order one -- the caller makes both
 *   draws on its torch.Generator, in that order, and passes the one that survives);
 *   old_pred [E] fp32 = previous step's pred_original_sample (NULL on the first step); pred_out [E] fp32 (may alias old_pred). */
int alg_cfg_dpm_step(const void* noise, int noise_dtype, const void* x, void* x_out, int sample_dtype,
                     const float* old_pred, float* pred_out, const void* rnd, int64_t E, const alg_dpm_step_t* p,
                     void* stream);

/* hy:1254-1270 -- (true-)CFG + FlowMatchEulerDiscreteScheduler.step on frames 1.. + re-prepend
 * of the conditioning frame.  Layout [C, T, HW]; noise holds all T frames (frame 0 ignored).
 *   x_out[c, 0] = first_frame[c, 0];  x_out[c, f>0] = fp32(noise_dtype(x + noise_dtype(dt * v)))  */
int alg_cfg_euler_step(const void* noise, int noise_dtype, const float* x, float* x_out,
                       const float* first_frame, int C, int T, int64_t HW, int n_pass, float guidance,
                       float dt, void* stream);

/* ------------------------------------------------------------------------- */
/* Tensor-core building blocks (tcgen05 / TMEM / TMA); used by the DiT engine */
/* and exported so parity tests can drive them directly.                      */
/* ------------------------------------------------------------------------- */

typedef enum {
  ALG_EPI_NONE = 0,          /* D = acc (+bias)                                               */
  ALG_EPI_GELU_TANH = 1,     /* D = gelu_tanh(bf16(acc + bias))                               */
  ALG_EPI_GATE_RESIDUAL = 2, /* D = bf16(fp32(R) + fp32(bf16(acc + bias)) * gate[row/rows_per_batch, col]) */
  ALG_EPI_RESIDUAL = 3,      /* D = bf16(fp32(R) + fp32(bf16(acc + bias)))                    */
  ALG_EPI_GELU_ERF = 4,      /* D = gelu(bf16(acc + bias)), exact erf form                    */
  ALG_EPI_SILU = 5           /* D = silu(bf16(acc + bias))                                    */
} alg_epilogue_t;

typedef struct {
  const void* A;      /* [M, K] bf16, row-major, lda elements between rows                     */
  const void* B;      /* [N, K] bf16, row-major (nn.Linear.weight layout), ldb                 */
  void* D;            /* [M, N] bf16 (or fp32 when out_f32), ldd                               */
  const void* bias;   /* bf16 [N] (or [M] when bias_per_row), may be NULL                      */
  const void* R;      /* residual, bf16 [M, N] with ldd; RESIDUAL / GATE_RESIDUAL only         */
  const void* gate;   /* fp32 (or bf16, see gate_dtype) [M / rows_per_batch, gate_ld] (GATE_RESIDUAL only) */
  int64_t M, N, K;
  int64_t lda, ldb, ldd;
  int64_t rows_per_batch;
  int64_t gate_ld;
  int32_t epilogue;
  int32_t bias_per_row;
  int32_t out_f32;
  /* --- ABI 2: gating variants of ALG_EPI_GATE_RESIDUAL (CogVideoX / HunyuanVideo blocks) --- */
  int32_t gate_dtype;        /* ALG_F32 (Wan: fp32 gate, one rounding) or ALG_BF16 (gate is a bf16 tensor)        */
  int32_t gate_round;        /* 1: D = bf16(R + bf16(gate * bf16(acc + bias))) -- eager bf16 `x + gate * y`       */
  int64_t gate_split_row;    /* rows with (row % rows_per_batch) < gate_split_row use gate_alt (0 = unused)       */
  const void* gate_alt;      /* Cog: text rows use enc_gate (cog DiT block); Hy: first-frame rows use the
                                token-replace gate (hy DiT block, image_condition_type token_replace)            */
  /* --- ABI 3 --- */
  int64_t a_k_period;        /* 0, or A holds only a_k_period columns that repeat along K: A[m, k] = A[m, k % period]
                                (multiple of 64 dividing K; lda >= period).  A causal 3-D convolution on ONE frame reads
                                the same [kh, kw, C] patch for each of its kt temporal taps, so the patch matrix is
                                gathered once and only the weights differ along K                                  */
  /* --- ABI 5: implicit convolution --- */
  int32_t a_tap_kblocks;     /* 0, or K = a_n_taps groups of a_tap_kblocks * 64 columns: group g multiplies A columns
                                [0, a_tap_kblocks * 64) of the rows SHIFTED by a_tap_offsets[g] against B columns of group g.  With
                                A = a zero-padded channels-last clip [(T+pt) * (H+2) * (W+2), C] and offsets = the taps' raster
                                distances, D row m is the stride-1 convolution at padded-raster pixel m: no patch matrix exists;
                                rows outside [0, M) read as zero (TMA out-of-bounds fill)                              */
  int32_t a_n_taps;          /* <= 32 */
  const int32_t* a_tap_offsets; /* HOST array of a_n_taps row offsets */
  int64_t a_rows;            /* rows of A in tap mode when it has more than M (offsets reach past the last output row); 0 = M */
  const float* bias_f32;     /* out_f32 only (no bf16 bias, no epilogue): D = acc + bias_f32[col] + residual_f32[row, col], all in */
  const float* residual_f32; /* fp32; residual_f32 has D's leading dimension.  The float32 VAE / encoder linears.  May be NULL. */
} alg_gemm_t;

/* D = epilogue(A * B^T + bias): every nn.Linear of the DiT (SURVEY kernel K6). */
int alg_gemm_bf16(const alg_gemm_t* g, void* stream);

typedef struct {
  const void* Q;   /* bf16; element (b, n, h, d) at Q + b*q_bs + n*q_rs + h*head_dim + d          */
  const void* K;   /* bf16; (b, n, h, d) at K + b*k_bs + n*k_rs + h*head_dim + d                  */
  const void* Vt;  /* bf16, transposed values: (b, h, d, n) at Vt + b*v_bs + (h*head_dim + d)*v_rs + n */
  void* O;         /* bf16; same addressing as Q with o_bs / o_rs                                 */
  int32_t batch, heads, head_dim; /* head_dim in {64, 128}                                        */
  int64_t n_q, n_kv;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs; /* in elements                          */
  float scale;                    /* 1/sqrt(head_dim)                                             */
  int32_t accumulate;             /* 1: O += result (Wan text + image cross-attention sum)        */
} alg_attention_t;

/* Non-causal softmax(Q K^T * scale) V, fp32 softmax, bf16 in/out (SURVEY kernel K8;
 * replaces F.scaled_dot_product_attention inside the DiT, wan:910). */
int alg_attention_bf16(const alg_attention_t* a, void* stream);


/* ------------------------------------------------------------------------- */
/* HBM-bound DiT building blocks shared by the CogVideoX and HunyuanVideo     */
/* forwards (call sites cog:1082-1090, hy:1243-1252).  The host side that     */
/* sequences them mirrors diffusers' modules (alg_b200/cogvideox.py,          */
/* alg_b200/hunyuan.py); all arithmetic is in these kernels.                  */
/* ------------------------------------------------------------------------- */

/* LayerNorm over the rows of a bf16 [rows, d] matrix (+ affine) (+ AdaLN modulate).
 *   chain_bf16 = 0: out = bf16( LN(x)[*w+b] [* (1 + scale) + shift] ), everything in fp32 (Wan FP32LayerNorm path)
 *   chain_bf16 = 1: y = bf16(LN(x)[*w+b]);  out = bf16( bf16(y * bf16(1 + scale)) + shift )  -- the eager bf16 op chain of
 *                   CogVideoXLayerNormZero / AdaLayerNormZero(Single) / AdaLayerNorm(Continuous)
 * Modulation vectors are [d] per sample: sample b = row / rows_per_batch reads scale + b * mod_batch_stride; rows whose
 * index inside the sample is < split_row read scale_alt / shift_alt instead (Cog text rows; Hy first-frame tokens). */
typedef struct {
  const void* x;
  void* out;
  int64_t rows;
  int32_t d;
  float eps;
  const void* weight; /* NULL: no affine */
  const void* bias;
  int32_t affine_dtype; /* ALG_F32 or ALG_BF16 */
  int32_t mod_dtype;    /* ALG_F32 or ALG_BF16 */
  const void* scale;    /* NULL: no modulation */
  const void* shift;
  const void* scale_alt;
  const void* shift_alt;
  int64_t rows_per_batch;
  int64_t mod_batch_stride;
  int64_t split_row;
  int32_t chain_bf16;
} alg_layer_norm_t;
int alg_layer_norm(const alg_layer_norm_t* p, void* stream);

typedef enum { ALG_NORM_NONE = 0, ALG_NORM_RMS = 1, ALG_NORM_LAYER = 2 } alg_head_norm_t;

/* Per-head q/k normalisation + rotary embedding, in place on bf16 [rows, heads * head_dim] (row stride ld).
 *   ALG_NORM_RMS   : h = bf16(x * rsqrt(mean(x^2) + eps)); y = bf16(h * w)            (diffusers RMSNorm, Hy qk_norm)
 *   ALG_NORM_LAYER : y = bf16(LN(x) * w + b) over head_dim                            (Cog attention norm_q / norm_k)
 * RoPE (diffusers apply_rotary_emb, use_real, adjacent pairs): y = bf16(y * cos + rot(y) * sin) in fp32 for rows whose
 * index inside the sample lies in [rope_row0, rope_row0 + rope_rows); cos / sin fp32 [rope_rows, head_dim]. */
typedef struct {
  void* x;
  int64_t rows, ld;
  int32_t heads, head_dim;
  int32_t norm_kind;
  float eps;
  const void* weight; /* bf16 [head_dim] */
  const void* bias;   /* bf16 [head_dim], ALG_NORM_LAYER only */
  const float* cos;   /* NULL: no RoPE */
  const float* sin;
  int64_t rows_per_batch, rope_row0, rope_rows;
} alg_head_norm_rope_t;
int alg_head_norm_rope(const alg_head_norm_rope_t* p, void* stream);

/* One source of the patch gather: `channels` planes of a [.., T, H, W] grid with element strides sc (channel), st (frame),
 * sy (row); x stride is 1.  ptr_t0 (optional) replaces frame 0 (hy:1171 first-frame token replacement), strides sc_t0 / sy. */
typedef struct {
  const void* ptr;
  const void* ptr_t0;
  int32_t dtype;
  int32_t channels;
  int64_t sc, st, sy, sc_t0;
} alg_patch_src_t;

/* Model-input assembly fused with the 2x2 im2col of the patch embedding (wan:882-891, cog:1059-1075, hy:1168-1195):
 * A[(p*N + n), (c, i, j)] = bf16(src_c[t, 2y+i, 2x+j]), channels = concatenation of the n_src sources of pass p.
 * srcs_host: n_pass * n_src descriptors (host memory).  The replicated / concatenated / cast model input is never built. */
int alg_patch_gather(const alg_patch_src_t* srcs_host, int n_pass, int n_src, int T, int H, int W, void* A, int64_t lda,
                     void* stream);

/* proj [n_pass*N, ld] bf16 (N = T*(H/2)*(W/2)) -> out bf16, element (p, c, t, y, x) at out + p*s_pass + c*sc + t*st + y*sy + x.
 * channel_major = 0: proj column (i*2 + j)*C + c  (Wan);  1: column c*4 + i*2 + j  (CogVideoX, HunyuanVideo). */
int alg_unpatchify(const void* proj, int64_t ld, void* out, int n_pass, int C, int T, int H, int W, int64_t s_pass,
                   int64_t sc, int64_t st, int64_t sy, int channel_major, void* stream);

/* diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): out[dim] = [cos | sin], dtype f32/bf16 */
int alg_timestep_embedding(float timestep, int dim, void* out, int dtype, void* stream);

typedef enum { ALG_EW_ADD = 0, ALG_EW_SILU = 1, ALG_EW_COPY = 2, ALG_EW_GELU_TANH = 3 } alg_ew_op_t;
/* Tiny bf16 vector ops of the conditioning path: out = bf16(a + b) | bf16(silu(a)) | a.  b may be NULL unless ADD. */
int alg_elementwise_bf16(int op, const void* a, const void* b, void* out, int64_t n, void* stream);

/* out[d] = bf16( sum_r x[r, :] / rows ) with fp32 accumulation (HunyuanVideo token-refiner pooled text projection). */
int alg_mean_rows_bf16(const void* x, int64_t rows, int d, int64_t ld, void* out, void* stream);

/* dst[r, 0:d] = src[r, 0:d] for rows rows, bf16, 16-byte vectorised (sequence concat). */
int alg_copy_rows_bf16(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows, int d, void* stream);

/* ------------------------------------------------------------------------- */
/* Once-per-video conditioning encoders: UMT5 / T5 text encoder  */
/* and CLIP vision / text towers -- the networks `self.text_encoder(...)`,       */
/* `self.image_encoder(...)` run at wan:212 / wan:233, cog:258, hy:333-452 (in   */
/* transformers==4.48.1).  The linears are alg_gemm_bf16; these are the ops      */
/* around them.  Host sequencing: alg_b200/encoders.py.                          */
/* ------------------------------------------------------------------------- */

/* T5LayerNorm: h = bf16(x * rsqrt(mean(x^2) + eps)) (fp32 statistics); out = bf16(weight * h).  bf16 rows of length d. */
int alg_t5_rms_norm_bf16(const void* x, int64_t ld_x, void* out, int64_t ld_out, int64_t rows, int d, float eps,
                         const void* weight, void* stream);

/* nn.Embedding: out[r, :] = table[ids[r], :] (bf16 rows of d elements, d % 8 == 0; ids int64 on the device). */
int alg_gather_rows_bf16(const void* table, int64_t vocab, const int64_t* ids, void* out, int64_t rows, int d, void* stream);

/* softmax(scale * q k^T + rel_bias + mask) v for short sequences (head_dim <= 128; <= 768 keys with rel_bias, any length
 * streamed 64 keys at a time otherwise -- LLaVA's ~900-token Llama prompt, hy:333-337), bf16 or fp32, with the
 * eager op chain's roundings in bf16 mode (q k^T -> bf16, + bias -> bf16, softmax in fp32 -> bf16, P V -> bf16).
 * q / k / v / out: element (b, token, h, d) at ptr + b * bs + token * rs + h * head_dim + d.
 * rel_bias (optional) fp32 [heads, 2 * n_kv - 1]: the T5 relative-position bias of (query i, key j) sits at j - i + n_kv - 1.
 * kv_valid (optional) int32 [batch]: keys >= kv_valid[b] are masked (right-padded prompts).  causal: keys j > i masked. */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  int32_t dtype; /* ALG_BF16 or ALG_F32 */
  int32_t batch, heads, head_dim;
  int64_t n_q, n_kv;
  int64_t q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, o_bs, o_rs;
  float scale;
  const float* rel_bias;
  const int32_t* kv_valid;
  int32_t causal;
  int32_t kv_group;        /* grouped-query attention: query head h reads K / V head h / kv_group (0 or 1 = one per head) */
  const uint8_t* key_mask; /* optional [batch, n_kv]: 0 = masked key (padding that is not a suffix) */
} alg_small_attention_t;
int alg_small_attention(const alg_small_attention_t* a, void* stream);

/* fp32 path of the CLIP vision tower (run.py:48 loads image_encoder in float32): */
/* out = LN(x) * weight + bias over fp32 rows of length d */
int alg_layer_norm_f32(const float* x, float* out, int64_t rows, int d, float eps, const float* weight, const float* bias,
                       void* stream);
/* in place: x = act(x + bias[col]) (+ residual); act 0 none, 1 GELU (erf), 2 quick-GELU; bias / residual may be NULL */
int alg_bias_act_f32(float* x, const float* bias, const float* residual, int64_t rows, int cols, int act, void* stream);
/* fp32 [rows, K] -> bf16 [rows, 3K]: [hi | hi | lo] (weight_order 0, activations) or [hi | lo | hi] (1, weights); the
 * K-concatenated bf16 GEMM with fp32 accumulation then equals the fp32 product up to the dropped lo * lo term (2^-16). */
int alg_split3_bf16(const float* x, void* out, int64_t rows, int K, int weight_order, void* stream);
/* CLIPVisionEmbeddings (fp32): the stride-P patch convolution as unfold + GEMM, then class token + position embeddings.
 * x [batch, C, H, W] -> out [batch * (H/P) * (W/P), ld]: row = (c, i, j)-flattened patch, zero-padded from C*P*P to ld */
int alg_patchify_f32(const float* x, float* out, int batch, int C, int H, int W, int P, int ld, void* stream);
/* out [batch, num_patches + 1, d]: token 0 = class_embedding + pos[0]; token 1 + n = patches[b * num_patches + n] + pos[1 + n] */
int alg_clip_embed_f32(const float* patches, const float* class_embedding, const float* position_embedding, float* out,
                       int batch, int num_patches, int d, void* stream);
/* fp32 pieces of the Llama decoder inside LLaVA (HunyuanVideo text_encoder, hy:333-337; transformers modeling_llama.py): */
/* LlamaRMSNorm: out = weight * (x * rsqrt(mean(x^2) + eps)) over fp32 rows of length d */
int alg_rms_norm_f32(const float* x, float* out, int64_t rows, int d, float eps, const float* weight, void* stream);
/* apply_rotary_pos_emb (rotate-half) in place on the first `heads` heads of every row of x [rows, ld]:
 * x = x * cos + rotate_half(x) * sin, cos / sin fp32 [rows, head_dim] */
int alg_rope_half_f32(float* x, int64_t ld, const float* cos_table, const float* sin_table, int64_t rows, int heads, int head_dim,
                      void* stream);
/* LlamaMLP: out[r, c] = silu(gate_up[r, c]) * gate_up[r, f + c] for the fused [gate | up] projection [rows, 2 f] */
int alg_swiglu_f32(const float* gate_up, float* out, int64_t rows, int f, void* stream);
/* out = bf16(a * b) elementwise (T5 gated-GELU feed-forward: gelu(wi_0 x) * wi_1 x) */
int alg_mul_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------- */
/* Wan2.1 I2V DiT engine: WanTransformer3DModel.forward (call site wan:910-917) */
/* ------------------------------------------------------------------------- */

typedef struct {
  int32_t num_heads, head_dim, in_channels, out_channels;
  int32_t text_dim, freq_dim, ffn_dim, num_layers, image_dim, text_len;
  int32_t patch_t, patch_h, patch_w;
  int32_t rope_max_seq_len;
  float eps;
} alg_wan_config_t;

typedef struct alg_wan_engine alg_wan_engine_t;

int alg_wan_create(const alg_wan_config_t* cfg, alg_wan_engine_t** out);
void alg_wan_destroy(alg_wan_engine_t* e);

/* Register one parameter by its diffusers state_dict name (e.g.
 * "blocks.0.attn1.to_q.weight").  The engine keeps the POINTER: the caller owns
 * the memory and must keep it alive.  dtype must be bf16, except fp32 for the
 * modules diffusers keeps in fp32 (time_embedder, scale_shift_table, norm2). */
int alg_wan_set_weight(alg_wan_engine_t* e, const char* name, const void* ptr, int64_t numel, int dtype);
/* 0 when every parameter of the configured model has been registered. */
int alg_wan_weights_complete(alg_wan_engine_t* e);

/* Scratch bytes needed for a forward of n_pass samples on a [T, H, W] latent grid. */
int alg_wan_workspace_bytes(alg_wan_engine_t* e, int n_pass, int T, int H, int W, int n_img_tokens,
                            size_t* bytes);

/* wan:882-917 -- model-input assembly + DiT forward for the 2 or 3 CFG passes.
 *   latents[p][16, T, H, W] fp32           per pass (the loop passes the SAME pointer n_pass times: never replicated)
 *   cond[p]   [20, T, H, W] fp32           per pass: condition / lp_image_latents
 *   text[p]   [text_len, text_dim] bf16    per pass: negative / positive prompt embeds
 *   image     [n_img_tokens, image_dim] bf16 (shared)
 *   timestep  the scheduler's int64 timestep value
 *   noise_out [n_pass, 16, T, H, W] bf16
 * The fp32 -> bf16 cast and the channel concat (wan:886-891) happen inside the patch gather. */
int alg_wan_forward(alg_wan_engine_t* e, const float* const* latents, const float* const* cond, const void* const* text,
                    const void* image, int n_img_tokens, int n_pass, int T, int H, int W, int64_t timestep,
                    void* noise_out, void* workspace, size_t workspace_bytes, void* stream);

/* Parity hook for the Wan q/k norm (WanAttnProcessor: norm_q / norm_k = RMSNorm across heads, then rotary embedding in
 * complex128): in place on x [rows, d] bf16.  rope_t / rope_h / rope_w are the fp64 tables [pos][n_*][2] = (cos, sin) of
 * the frame / height / width axes (n_t + n_h + n_w = head_dim / 2 rotary pairs per head; token row -> (t, y, x) over the
 * ppf x pph x ppw patch grid); rope_t == NULL skips the rotation (cross-attention q / k).  The engine calls the same kernel. */
int alg_wan_rms_norm_rope(void* x, int64_t rows, int d, int head_dim, float eps, const void* weight,
                          const double* rope_t, const double* rope_h, const double* rope_w, int n_t, int n_h, int n_w,
                          int ppf, int pph, int ppw, void* stream);

/* Debug / parity hook: when a device buffer is set, every forward appends to it (while space lasts), in order:
 * the patch embedding [n_pass*N, d] bf16, temb [d] bf16, timestep_proj [6d] bf16, then the residual stream
 * [n_pass*N, d] bf16 after each block.  Pass NULL to switch it off. */
int alg_wan_set_debug_buffer(alg_wan_engine_t* e, void* buf, size_t bytes);

/* Step-invariant context memoisation (SURVEY 7 "hoist", wan:844-944: prompt / image conditioning is fixed for the whole video):
 * with enable != 0, alg_wan_forward remembers the text / image embedder outputs' K and V^T projections of all layers per
 * (text pointers, image pointer, n_pass, n_img) -- two layouts are kept, the three-pass and the two-pass one -- and later forwards
 * with the same key skip those projections (~0.2 % of a step; results are bit-identical).  The KEY IS THE POINTERS: call again
 * (enable or disable) whenever the CONTENTS behind them change, which also drops what was memoised.  Default: off. */
int alg_wan_context_cache(alg_wan_engine_t* e, int enable);
/* Device timing per kernel class, measured with CUDA events on the launching stream around every launch of the
 * forward (classes: 0 self-attention, 1 cross-attention, 2 GEMM, 3 HBM-bound elementwise).  Enable, run forwards,
 * then read: the read synchronises on the recorded events, returns summed milliseconds + launch counts and resets. */
int alg_wan_profile(alg_wan_engine_t* e, int enable);
int alg_wan_profile_read(alg_wan_engine_t* e, float* ms_per_class, int32_t* launches_per_class, int n_classes);

/* ------------------------------------------------------------------------- */
/* CogVideoX VAE encoder blocks (ABI 3).  cog:257 (`self.vae.encode(image_lp)` EVERY step when the filter runs in    */
/* pixel space, BASELINE configs[2]) and cog:166 (conditioning image): AutoencoderKLCogVideoX.encode on ONE frame.    */
/* Activations are channels-last [T*H*W, C] bf16; every (causal 3-D / strided 2-D) convolution is                    */
/* alg_im2col_bf16 + ONE alg_gemm_bf16 (bias / residual in its epilogue).                                            */
/* ------------------------------------------------------------------------- */

typedef struct {
  const void* x;  /* [T, H, W, C] bf16 channels-last, C % 8 == 0                                                  */
  void* cols;     /* [To*Ho*Wo, ld] bf16; column order (it, ih, iw, c) -- conv weight [Co, Ci, kt, kh, kw]
                     permuted to [Co, kt, kh, kw, Ci] is the matching GEMM B operand; columns >= kt*kh*kw*C are zero  */
  int32_t T, H, W, C;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;        /* strides                                                                           */
  int32_t pad_t;             /* causal padding: source frame = max(to*st + it - pad_t, 0), i.e. frames before t = 0
                                replicate frame 0 (CogVideoXCausalConv3d.fake_context_parallel_forward, no cache)   */
  int32_t pad_top, pad_left; /* zero padding; bottom / right padding is implied by Ho, Wo                          */
  int32_t To, Ho, Wo;
  int64_t ld;                /* >= kt*kh*kw*C, multiple of 8                                                        */
} alg_im2col_t;
int alg_im2col_bf16(const alg_im2col_t* p, void* stream);

/* nn.GroupNorm over [rows, C] channels-last (one sample; rows = T*H*W) followed by an optional SiLU:
 *   y = bf16(a*x + b), a = rstd_g * weight_c, b = bias_c - a * mean_g, statistics in fp32/fp64; silu != 0:
 *   y = bf16(y / (1 + exp(-y))).  `stats` is a caller-provided scratch of 2*groups doubles (zeroed by the call). */
typedef struct {
  const void* x;
  void* y;            /* may alias x */
  const void* weight; /* bf16 [C] or NULL */
  const void* bias;   /* bf16 [C] or NULL */
  double* stats;
  int64_t rows;
  int32_t C, groups;
  float eps;
  int32_t silu;
} alg_group_norm_t;
int alg_group_norm_bf16(const alg_group_norm_t* p, void* stream);

/* AutoencoderKLCogVideoX.decode (cog:428-433), channels-last [T*H*W, C] bf16 like the encoder ops above. */
/* F.interpolate(mode="nearest") over (T, H, W): out[to, yo, xo, :] = x[to*Ti/To, yo*Hi/Ho, xo*Wi/Wo, :]  (C % 8 == 0) */
int alg_upsample_nearest_bf16(const void* x, void* out, int C, int Ti, int Hi, int Wi, int To, int Ho, int Wo, void* stream);
/* CogVideoXSpatialNorm3D after its GroupNorm: out = bf16(bf16(f_norm * y) + b) (+ SiLU when silu != 0), where
 * yb [zt*zh*zw, 2C] = [conv_y(zq) | conv_b(zq)] at latent resolution and (y, b) of pixel (t, yy, xx) are read at the nearest
 * latent pixel (t*zt/T, yy*zh/H, xx*zw/W): conv(resize(zq)) == resize(conv(zq)) for 1x1x1 convolutions. */
int alg_spatial_norm_apply_bf16(const void* f_norm, const void* yb, void* out, int C, int T, int H, int W, int zt, int zh,
                                int zw, int silu, void* stream);

/* ------------------------------------------------------------------------- */
/* float32 video VAE (AutoencoderKLWan: wan:429-434 encode of the condition    */
/* clip, wan:526 per-step encode in pixel-space ALG, wan:959 decode; run.py:   */
/* 51-55 loads it in float32; network in diffusers@be2fb77 autoencoder_kl_wan).*/
/* Activations channels-last [T*H*W, C] fp32.  Host sequencing: vae_wan.py.    */
/* ------------------------------------------------------------------------- */

/* Patch gather + bf16 3-term split in one pass: cols[(to, ho, wo)] = [hi | hi | lo] (three sections of `ld` bf16 each) of the
 * fp32 patch (it, ih, iw, c) read at frame (to0 + to)*st + it - pad_t, row ho*sh + ih - pad_top, column wo*sw + iw - pad_left of
 * the LOGICAL frame (H*up) x (W*up), whose pixel (y, x) is x[t][y / up][x / up] (up = 2: nn.Upsample(nearest-exact, x2) fused
 * into the following convolution).  Frames < t_min (WanCausalConv3d's zero front padding: t_min = 0; WanResample "upsample3d",
 * whose temporal windows never see frame 0: t_min = 1) and pixels outside the logical frame read as zero.  One alg_gemm_bf16
 * against the weight's [hi | lo | hi] split (alg_split3_bf16, weight_order 1) then gives the fp32 convolution to 2^-16. */
typedef struct {
  const void* x;  /* fp32 [T, H, W, C] */
  void* cols;     /* bf16 [To*Ho*Wo, 3*ld] */
  int32_t T, H, W, C;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pad_t, pad_top, pad_left;
  int32_t To, Ho, Wo; /* output frames of THIS call, output rows / columns */
  int32_t to0;        /* index of the first output frame of this call (frame-chunked convolutions) */
  int32_t up;         /* 1 or 2 */
  int32_t t_min;
  int32_t replicate;  /* != 0: F.pad(mode="replicate") instead of zeros, in time (frames < t_min read frame t_min) and space
                         (HunyuanVideoCausalConv3d) */
  int32_t tdup;       /* 1, or 2: the LOGICAL clip has 2T - 1 frames, frame t reads source frame (t + 1) / 2 (frame 0 once, every
                         later frame twice: HunyuanVideoUpsampleCausal3D's temporal nearest upsample fused into its convolution) */
  int64_t ld;         /* >= kt*kh*kw*C, multiple of 8 */
} alg_im2col_f32_t;
int alg_im2col_split3_f32(const alg_im2col_f32_t* p, void* stream);
/* WanRMS_norm over the channels of each row: out = x / max(||x||_2, 1e-12) * sqrt(C) * gamma (+ bias) (+ SiLU when silu != 0) */
int alg_rms_norm_cl_f32(const float* x, float* out, int64_t rows, int C, const float* gamma, const float* bias, int silu,
                        void* stream);
/* in place: every row <- softmax(scale * row) (WanAttentionBlock scores, one head of C channels per frame).  causal_block > 0:
 * row r only sees columns < (r / causal_block + 1) * causal_block, the rest of the row is zeroed (frame-causal mask of the
 * HunyuanVideo VAE mid block, diffusers prepare_causal_attention_mask) */
int alg_softmax_rows_f32(float* x, int64_t rows, int cols, int64_t ld, float scale, int causal_block, void* stream);
/* nn.GroupNorm (+ SiLU) over channels-last fp32 [rows, C] of one sample; `stats` = scratch of 2*groups doubles (zeroed here) */
int alg_group_norm_f32(const float* x, float* y, int64_t rows, int C, int groups, float eps, const float* weight, const float* bias,
                       int silu, double* stats, void* stream);
/* [C, pixels] (one sample of [B, C, T, H, W]) -> channels-last [pixels, ld] (columns >= C zeroed), and back with an optional
 * clamp to [lo, hi] (lo < hi; AutoencoderKLWan.decode clamps to [-1, 1]) */
int alg_nchw_to_cl_f32(const float* x, float* out, int C, int64_t pixels, int ld, void* stream);
int alg_cl_to_nchw_f32(const float* x, float* out, int C, int64_t pixels, int ld, float lo, float hi, void* stream);

/* Operand of the implicit (patch-matrix-free) stride-1 convolution, see alg_gemm_t.a_tap_kblocks: every interior pixel of a
 * clip -- x fp32 [T*H*W, C] compact, or the padded raster [(T+front_pad)(H+2)(W+2), C] when in_padded -- goes through the optional
 * WanRMS_norm (gamma != NULL) and SiLU and is written as bf16 [hi | hi | lo] into row ((t+front_pad)(H+2) + y+1)(W+2) + x+1 of
 * out [(T+front_pad)(H+2)(W+2), Cs] (Cs >= 3C).  Padding rows and columns >= 3C are not written: zero the buffer once. */
int alg_norm_split_pad_f32(const float* x, void* out, int T, int H, int W, int C, int front_pad, int in_padded, int Cs,
                           const float* gamma, int silu, void* stream);
/* F.pad(mode="replicate") of a padded split operand [(T+front_pad)(H+2)(W+2), Cs] bf16 in place: every padding pixel (front frames,
 * one-pixel border) takes the row of the nearest interior pixel (HunyuanVideoCausalConv3d) */
int alg_replicate_border_bf16(void* buf, int T, int H, int W, int front_pad, int Cs, void* stream);
/* interior pixels of compact [T*H*W, C] fp32 -> padded raster (to_padded != 0) or back; padding is not touched */
int alg_pad_copy_f32(const float* src, float* dst, int T, int H, int W, int C, int front_pad, int to_padded, void* stream);

/* bf16 frames <-> their spatially zero-padded raster [frames, H+2, W+2, ld] (image in [1,H] x [1,W]) for the implicit convolution:
 * to_padded != 0: src compact [frames*H*W, C] -> interior of dst (borders untouched: zero the buffer once);
 * to_padded == 0: src padded -> dst compact [frames*H*W, C], plus an optional bf16 residual [frames*H*W, C]:
 * dst = bf16(float(src) + float(residual)) */
int alg_pad_frames_bf16(const void* src, void* dst, const void* residual, int frames, int H, int W, int C, int64_t ld, int to_padded,
                        void* stream);

/* out = wa * a + wb * b over n fp32 elements (each product rounded, then the sum): the cross-fade of overlapping temporal tiles in
 * AutoencoderKLHunyuanVideo._temporal_tiled_decode (blend_t); out may alias a or b */
int alg_axpby_f32(const float* a, const float* b, float* out, int64_t n, float wa, float wb, void* stream);

/* Number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
int64_t alg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ALG_B200_H_ */
