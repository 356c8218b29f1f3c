// Launchers of the HBM-bound DiT kernels (implemented in dit_kernels.cu).
#pragma once
#include "common.cuh"

namespace alg {
namespace dit {

// wan:882-891 + Conv3d(1,2,2) im2col: A[(p*N + n), c*4 + i*2 + j] = bf16(src_c[t, 2y+i, 2x+j]); channels 0..lat_ch-1
// come from lat[p] (the pipeline passes the same latents for every pass), the rest from cond p[p].  Never materialises the x3 replicated input.
struct CondPtrs {
  const float* lat[3];
  const float* p[3];
};
int patch_gather(CondPtrs cond, int n_pass, int lat_ch, int cond_ch, int T, int H, int W,
                 __nv_bfloat16* A, cudaStream_t st);

// diffusers RMSNorm over the full row (across heads) with bf16 weight, then (optionally) Wan RoPE in fp64 on
// adjacent pairs of every head.  In place.  rope tables: cos/sin doubles [max_pos][n_t | n_h | n_w] per axis.
struct RopeTables {
  const double* t;  // [max_pos][2 * n_t]  (cos, sin) interleaved
  const double* h;  // [max_pos][2 * n_h]
  const double* w;  // [max_pos][2 * n_w]
  int n_t, n_h, n_w;
  int ppf, pph, ppw;  // token grid
};
int rms_norm_rope(__nv_bfloat16* x, int64_t rows, int d, int head_dim, float eps, const __nv_bfloat16* w,
                  const RopeTables* rope, cudaStream_t st);

// sinusoidal timestep embedding [cos | sin] fp32 (flip_sin_to_cos, shift 0)
int timestep_sinusoid(float timestep, int dim, float* out, cudaStream_t st);
// out[o] = act(W[o, :] . x + b[o]); fp32; act: 0 none, 1 SiLU
int gemv_f32(const float* W, const float* b, const float* x, float* out, int out_f, int in_f, int act, cudaStream_t st);
// temb = bf16(v); silu_temb = bf16(silu(float(temb)))
int temb_finish(const float* v, __nv_bfloat16* temb, __nv_bfloat16* silu_temb, int d, cudaStream_t st);
// mod[j, c] = table[j, c] + float(src[broadcast ? c : j*d + c]); fp32 out
int add_table(const float* table, const __nv_bfloat16* src, float* mod, int rows, int d, int broadcast, cudaStream_t st);
// proj [n_pass*N, ph*pw*C] bf16 -> out [n_pass, C, T, H, W] bf16
int unpatchify(const __nv_bfloat16* proj, __nv_bfloat16* out, int n_pass, int C, int T, int H, int W, cudaStream_t st);
// dst rows [n_rows, d] <- src rows, bf16, 16-byte vectorised (context concat)
int copy_rows(const __nv_bfloat16* src, int64_t src_ld, __nv_bfloat16* dst, int64_t dst_ld, int64_t rows, int d,
              cudaStream_t st);

}  // namespace dit
}  // namespace alg
