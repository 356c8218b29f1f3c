// Low-pass filters of the ALG conditioning image / latent (HBM-bound kernels).
//
//   alg_lowpass_down_up   replaces lp_utils.py:49-54   (2x ATen upsample_bilinear2d_aa)
//   alg_lowpass_gaussian  replaces lp_utils.py:40-47   (torchvision gaussian_blur: reflect pad + depthwise conv2d)
//
// down_up: one CTA stages a whole H x W plane in shared memory, applies the four banded
// resampling operators (W-down, H-down, W-up, H-up) there and writes the plane back once:
// algorithmic traffic = read once + write once.  Planes too large for shared memory take a
// generic four-pass path through global scratch.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace alg {

float host_round_f16(float f) { return __half2float(__float2half_rn(f)); }

// --------------------------------------------------------------------------------------------
// ATen anti-aliased triangle-filter weights (align_corners = false), fp32 like the CUDA kernel
// --------------------------------------------------------------------------------------------
struct Band {  // banded resampling operator: out[i] = sum_{k < cnt[i]} w[i * stride + k] * in[start[i] + k]
  std::vector<int> start, cnt;
  std::vector<float> w;  // [n_out][stride], zero padded
  int stride = 1;
};

// ATen's CUDA kernel (UpSampleBilinear2d.cu: upsample_gen2d_aa_out_frame, _compute_weights_span, _compute_weights)
// computes the taps in fp32, stores them in the tensor dtype (`wt_ptr[j] = scalar_t(w)`) and normalises in place with
// `wt_ptr[j] /= total_w` -- for a 16-bit scalar_t that is scalar_t / scalar_t, i.e. the fp32 total is itself rounded to
// the tensor dtype before the divide (torch/include/ATen/native/cuda/UpSample.cuh:322-347): rounded here the same way.
static Band make_band(int in_size, int out_size, int dt) {
  Band b;
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? (float)(1.0 / (double)scale) : 1.0f;
  std::vector<std::vector<float>> rows(out_size);
  int max_band = 1;
  for (int i = 0; i < out_size; ++i) {
    const float center = scale * ((float)i + 0.5f);
    int xmin = (int)(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5f);
    if (xmax > in_size) xmax = in_size;
    int n = xmax - xmin;
    if (n < 0) n = 0;
    const float xmin_m_center = (float)xmin - center;
    float total = 0.f;
    std::vector<float> ws(n);
    for (int j = 0; j < n; ++j) {
      float x = ((float)j + xmin_m_center + 0.5f) * invscale;
      float v = 1.0f - fabsf(x);
      ws[j] = v > 0.f ? v : 0.f;
      total += ws[j];
    }
    for (int j = 0; j < n; ++j) {
      float w = host_round(ws[j], dt);
      if (total != 0.f) w = host_round(w / host_round(total, dt), dt);
      ws[j] = w;
    }
    b.start.push_back(xmin);
    b.cnt.push_back(n);
    if (n > max_band) max_band = n;
    rows[i] = ws;
  }
  b.stride = max_band | 1;  // odd stride: column-parallel reads of the table spread over the banks
  b.w.assign((size_t)out_size * b.stride, 0.f);
  for (int i = 0; i < out_size; ++i)
    for (size_t k = 0; k < rows[i].size(); ++k) b.w[(size_t)i * b.stride + k] = rows[i][k];
  return b;
}

// Device tables of one (H, W, h1, w1) geometry: 4 operators packed in one allocation.
struct DownUpTables {
  int* ints = nullptr;  // [start | cnt] x 4
  float* weights = nullptr;
  int n_ints = 0, n_weights = 0;
  // element offsets into ints / weights, and the per-operator weight row stride
  int s_off[4], c_off[4], w_off[4], stride[4], max_cnt[4];
};

static std::mutex g_tab_mu;
static std::map<std::tuple<int, int, int, int, int, int>, DownUpTables> g_tables;

static int get_tables(int H, int W, int h1, int w1, int dt, DownUpTables* out) {
  int dev = 0;
  ALG_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto key = std::make_tuple(dev, H, W, h1, w1, dt);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) {
    *out = it->second;
    return 0;
  }
  if (g_tables.size() >= 512) {  // non-interval schedules change (h1, w1) per step: bound the cache
    ALG_CUDA_OK(cudaDeviceSynchronize());
    for (auto& kv : g_tables) {
      cudaFree(kv.second.ints);
      cudaFree(kv.second.weights);
    }
    g_tables.clear();
  }
  // order: 0 = W-down (W -> w1), 1 = H-down (H -> h1), 2 = W-up (w1 -> W), 3 = H-up (h1 -> H)
  Band bands[4] = {make_band(W, w1, dt), make_band(H, h1, dt), make_band(w1, W, dt), make_band(h1, H, dt)};
  std::vector<int> ints;
  std::vector<float> ws;
  DownUpTables t;
  for (int i = 0; i < 4; ++i) {
    t.s_off[i] = (int)ints.size();
    ints.insert(ints.end(), bands[i].start.begin(), bands[i].start.end());
    t.c_off[i] = (int)ints.size();
    ints.insert(ints.end(), bands[i].cnt.begin(), bands[i].cnt.end());
    t.w_off[i] = (int)ws.size();
    t.stride[i] = bands[i].stride;
    t.max_cnt[i] = 1;
    for (int c : bands[i].cnt) t.max_cnt[i] = std::max(t.max_cnt[i], c);
    ws.insert(ws.end(), bands[i].w.begin(), bands[i].w.end());
  }
  t.n_ints = (int)ints.size();
  t.n_weights = (int)ws.size();
  ALG_CUDA_OK(cudaMalloc(&t.ints, ints.size() * sizeof(int)));
  ALG_CUDA_OK(cudaMalloc(&t.weights, ws.size() * sizeof(float)));
  ALG_CUDA_OK(cudaMemcpy(t.ints, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice));
  ALG_CUDA_OK(cudaMemcpy(t.weights, ws.data(), ws.size() * sizeof(float), cudaMemcpyHostToDevice));
  g_tables[key] = t;
  *out = t;
  return 0;
}

struct BandPtr {
  const int* start;
  const int* cnt;
  const float* w;
  int stride;
};

// One tap of interpolate_aa_single_dim (UpSample.cuh:349-365): taps and samples are widened to fp32 and summed with
// `output += t * wts` (an FMA chain under nvcc's default contraction), for every tensor dtype.
template <int DT>
__device__ __forceinline__ float tap(float acc, float w, float x, bool first) {
  return first ? w * x : fmaf(w, x, acc);
}

// Shared-memory storage of the staged plane and of the pass results: fp32 for fp32 tensors; for 16-bit tensors the TENSOR
// dtype itself -- the input is 16-bit and ATen rounds every pass result to scalar_t, so nothing is lost, the staging
// buffers halve (twice the CTAs per SM) and one 16-byte shared-memory access carries 8 columns instead of 4.
template <int DT> struct Stage { using type = typename Elem<DT>::type; };
template <> struct Stage<ALG_F32> { using type = float; };
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half_rn(v); }
// one 16-byte vector = VW columns
__device__ __forceinline__ void ldv(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void ldv(const __nv_bfloat16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(h[e]);
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ void ldv(const __half* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h[e]);
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ void stv(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
__device__ __forceinline__ void stv(__nv_bfloat16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void stv(__half* p, const float* f) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// The passes are instruction-issue bound (ncu, profiles/r01_lowpass.md), so the number of tap slots NT that a pass unrolls
// is a template parameter chosen from the operator's longest band: up-sampling bands hold 2-3 taps, down-sampling ones
// ceil(2 * scale) + 1; bands longer than 8 read the remaining taps from the shared-memory table.

// dst[r][j] = sum_k w * src[r][start_j + k]      (resample along the contiguous axis)
// A thread owns one output column j (its taps stay in registers) and every ng-th row: the CTA's threads are laid out
// as ng = blockDim / out_w row groups x out_w columns, so narrow outputs (42 columns) still occupy 252 of 256 lanes
// and the per-column set-up is paid once.
template <int DT, int NT>
__device__ __forceinline__ void pass_w(const typename Stage<DT>::type* __restrict__ src,
                                       typename Stage<DT>::type* __restrict__ dst, int rows, int in_w, int out_w,
                                       int dst_ld, BandPtr b) {
  using S = typename Stage<DT>::type;
  const int nt = blockDim.x;
  int ng = 1, rg = 0, j = threadIdx.x, jstep = nt;
  if (out_w <= nt) {
    ng = nt / out_w;
    rg = threadIdx.x / out_w;
    j = threadIdx.x - rg * out_w;
    jstep = out_w;  // one sweep
    if (rg >= ng) return;
  }
  for (; j < out_w; j += jstep) {
    const int s = b.start[j], n = b.cnt[j];
    const float* wt = b.w + j * b.stride;
    float w[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) w[k] = k < n ? wt[k] : 0.f;
    const S* p = src + rg * in_w + s;
    S* q = dst + rg * dst_ld + j;
    const int pstep = ng * in_w, qstep = ng * dst_ld;
#pragma unroll 4
    for (int r = rg; r < rows; r += ng, p += pstep, q += qstep) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < NT; ++k)
        if (k < n) acc = tap<DT>(acc, w[k], ldf(p + k), k == 0);
      if (NT == 8)
        for (int k = NT; k < n; ++k) acc = tap<DT>(acc, wt[k], ldf(p + k), false);
      stf(q, acc);  // ATen keeps the row-pass result in a scalar_t buffer (16-bit storage rounds, fp32 keeps)
    }
  }
}
// dst[i][c] = sum_k w * src[start_i + k][c]       (resample along the strided axis)
// thread (tx, ty): output rows i = ty + 8 m (taps warp-uniform, in registers), columns c = tx + 32 m'.
template <int DT, int NT, bool ROUND, bool TO_GLOBAL>
__device__ __forceinline__ void pass_h(const typename Stage<DT>::type* __restrict__ src, void* __restrict__ dst, int cols,
                                       int out_h, BandPtr b) {
  using S = typename Stage<DT>::type;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ny = blockDim.x >> 5;
  for (int i = ty; i < out_h; i += ny) {
    const int s = b.start[i], n = b.cnt[i];
    const float* wt = b.w + i * b.stride;
    float w[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) w[k] = k < n ? wt[k] : 0.f;
#pragma unroll 2
    for (int c = tx; c < cols; c += 32) {
      const S* p = src + s * cols + c;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < NT; ++k)
        if (k < n) acc = tap<DT>(acc, w[k], ldf(p + k * cols), k == 0);
      if (NT == 8)
        for (int k = NT; k < n; ++k) acc = tap<DT>(acc, wt[k], ldf(p + k * cols), false);
      if (TO_GLOBAL) {
        Elem<DT>::store(dst, (size_t)i * cols + c, acc);
      } else {
        stf(reinterpret_cast<S*>(dst) + i * cols + c, acc);  // shared-memory results are always dtype-rounded (ROUND)
      }
    }
  }
}

// The same pass, one 16-byte vector of columns per thread (VW = 4 fp32 / 8 16-bit columns): rows of `src` / `dst` are `ld`
// elements apart (a multiple of VW, 16-byte aligned), so one LDS.128 feeds VW FMAs and the result leaves as one 16-byte
// store.  Work items (row, column group) are flattened over the CTA so narrow buffers still fill every lane.
template <int DT, int NT, bool ROUND, bool TO_GLOBAL>
__device__ __forceinline__ void pass_h4(const typename Stage<DT>::type* __restrict__ src, void* __restrict__ dst, int ld,
                                        int dst_ld, int out_h, BandPtr b) {
  using S = typename Stage<DT>::type;
  constexpr int VW = 16 / (int)sizeof(S);
  const int cvn = ld / VW, total = out_h * cvn;
  const float inv = 1.0f / (float)cvn;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = (int)(((float)idx + 0.5f) * inv);  // exact: total < 2^16
    const int c = (idx - i * cvn) * VW;
    const int s = b.start[i], n = b.cnt[i];
    const float* wt = b.w + i * b.stride;
    const S* p = src + s * ld + c;
    float acc[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < NT; ++k)
      if (k < n) {
        const float w = wt[k];
        float x[VW];
        ldv(p + k * ld, x);
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = tap<DT>(acc[e], w, x[e], k == 0);
      }
    if (NT == 8)
      for (int k = NT; k < n; ++k) {
        const float w = wt[k];
        float x[VW];
        ldv(p + k * ld, x);
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = fmaf(w, x[e], acc[e]);
      }
    if (TO_GLOBAL) {
      stv(reinterpret_cast<typename Elem<DT>::type*>(dst) + (size_t)i * dst_ld + c, acc);  // global tensor dtype == S for 16-bit
    } else {
      stv(reinterpret_cast<S*>(dst) + i * dst_ld + c, acc);
    }
  }
}

// NT from the longest band of the operator (block-uniform)
#define ALG_BAND_SWITCH(maxn, CALL) \
  do {                              \
    if ((maxn) <= 2) {              \
      constexpr int NT = 2;         \
      CALL;                         \
    } else if ((maxn) <= 3) {       \
      constexpr int NT = 3;         \
      CALL;                         \
    } else if ((maxn) <= 4) {       \
      constexpr int NT = 4;         \
      CALL;                         \
    } else if ((maxn) <= 6) {       \
      constexpr int NT = 6;         \
      CALL;                         \
    } else {                        \
      constexpr int NT = 8;         \
      CALL;                         \
    }                               \
  } while (0)

struct DownUpGeom {
  int s_off[4], c_off[4], w_off[4], stride[4], max_cnt[4];
  int n_ints, n_weights;
};

template <int DT>
__global__ void __launch_bounds__(256, 5) down_up_fused_kernel(const void* __restrict__ in, void* __restrict__ out,
                                                            int64_t planes, int H, int W, int h1, int w1,
                                                            const int* __restrict__ ints,
                                                            const float* __restrict__ weights, DownUpGeom g,
                                                            int bufA_elems, int bufB_elems, int vec, int prefetch) {
  using T = typename Elem<DT>::type;
  using S = typename Stage<DT>::type;
  constexpr int VW = 16 / (int)sizeof(S);  // columns per 16-byte shared-memory vector
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // prefetch: TWO input stages A0 | A1 -- while plane i is filtered out of one, cp.async streams plane i + gridDim.x into the other,
  // so the global-load latency (the largest stall of the single-stage version: long_scoreboard 2.5 cycles per issued
  // instruction, profiles/r01_lowpass_norm_ncu.txt) is off the critical path
  S* A0 = reinterpret_cast<S*>(smem_raw);    // in [H, W]      -> small [h1, w1]
  S* B = A0 + (prefetch ? 2 : 1) * bufA_elems;  // t1 [H, w1]   -> t2 [h1, W]     (bufA_elems, bufB_elems: multiples of 8)
  float* sw = reinterpret_cast<float*>(B + bufB_elems);  // operator taps
  int* si = reinterpret_cast<int*>(sw + g.n_weights);    // operator starts / counts
  for (int i = threadIdx.x; i < g.n_weights; i += blockDim.x) sw[i] = weights[i];
  for (int i = threadIdx.x; i < g.n_ints; i += blockDim.x) si[i] = ints[i];
  BandPtr bw_down{si + g.s_off[0], si + g.c_off[0], sw + g.w_off[0], g.stride[0]};
  BandPtr bh_down{si + g.s_off[1], si + g.c_off[1], sw + g.w_off[1], g.stride[1]};
  BandPtr bw_up{si + g.s_off[2], si + g.c_off[2], sw + g.w_off[2], g.stride[2]};
  BandPtr bh_up{si + g.s_off[3], si + g.c_off[3], sw + g.w_off[3], g.stride[3]};
  const int HW = H * W;
  auto stage_async = [&](int64_t plane, S* dstA) {  // prefetch mode only: planes are 16-byte aligned and HW % VW == 0
    const char* src = reinterpret_cast<const char*>(reinterpret_cast<const T*>(in) + plane * HW);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dstA);
    for (int i = threadIdx.x; i < HW / VW; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(src + (size_t)i * 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if (prefetch && (int64_t)blockIdx.x < planes) stage_async(blockIdx.x, A0);
  for (int64_t plane = blockIdx.x; plane < planes; plane += gridDim.x) {
    S* A = A0 + (prefetch ? buf * bufA_elems : 0);
    T* dst = reinterpret_cast<T*>(out) + plane * HW;
    if (prefetch) {
      const int64_t next = plane + gridDim.x;
      if (next < planes) {
        stage_async(next, A0 + (buf ^ 1) * bufA_elems);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      buf ^= 1;
    } else {
      const T* src = reinterpret_cast<const T*>(in) + plane * HW;
      // ---- stage the plane: 16-byte vector loads when the plane start is aligned (S == T: a plain copy) -------------
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (HW % VW) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* a4 = reinterpret_cast<uint4*>(A);
        for (int i = threadIdx.x; i < HW / VW; i += blockDim.x) a4[i] = __ldg(s4 + i);
      } else {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) A[i] = src[i];
      }
    }
    __syncthreads();
    if (vec) {  // W % VW == 0 and 16-byte aligned planes: H-passes run one 16-byte vector of columns per thread
      const int w1p = (w1 + VW - 1) / VW * VW;
      ALG_BAND_SWITCH(g.max_cnt[0], (pass_w<DT, NT>(A, B, H, W, w1, w1p, bw_down)));  // t1 [H, w1p]
      __syncthreads();
      // small [h1, w1p], rounded to dtype (reference materialises it)
      ALG_BAND_SWITCH(g.max_cnt[1], (pass_h4<DT, NT, true, false>(B, A, w1p, w1p, h1, bh_down)));
      __syncthreads();
      ALG_BAND_SWITCH(g.max_cnt[2], (pass_w<DT, NT>(A, B, h1, w1p, W, W, bw_up)));  // t2 [h1, W]
      __syncthreads();
      ALG_BAND_SWITCH(g.max_cnt[3], (pass_h4<DT, NT, false, true>(B, dst, W, W, H, bh_up)));  // out [H, W]
    } else {
      ALG_BAND_SWITCH(g.max_cnt[0], (pass_w<DT, NT>(A, B, H, W, w1, w1, bw_down)));  // t1 [H, w1]
      __syncthreads();
      ALG_BAND_SWITCH(g.max_cnt[1], (pass_h<DT, NT, true, false>(B, A, w1, h1, bh_down)));
      __syncthreads();
      ALG_BAND_SWITCH(g.max_cnt[2], (pass_w<DT, NT>(A, B, h1, w1, W, W, bw_up)));  // t2 [h1, W]
      __syncthreads();
      ALG_BAND_SWITCH(g.max_cnt[3], (pass_h<DT, NT, false, true>(B, dst, W, H, bh_up)));  // out [H, W]
    }
    __syncthreads();
  }
}

// ---- generic path: one resampling pass per launch through global memory -----------------------
template <int DT_IN, int DT_OUT, bool ALONG_W, bool ROUND_BF>
__global__ void resample_pass_kernel(const void* __restrict__ in, void* __restrict__ out, int64_t planes, int in_h,
                                     int in_w, int out_h, int out_w, const int* __restrict__ start,
                                     const int* __restrict__ cnt, const float* __restrict__ w, int stride, int round_dt) {
  int64_t total = planes * out_h * out_w;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t plane = idx / ((int64_t)out_h * out_w);
    int rem = (int)(idx - plane * (int64_t)out_h * out_w);
    int i = rem / out_w, j = rem - i * out_w;
    const size_t base = (size_t)plane * in_h * in_w;
    const int o = ALONG_W ? j : i;
    const int s = start[o], n = cnt[o];
    const float* wt = w + (size_t)o * stride;
    float acc = 0.f;
    for (int k = 0; k < n; ++k) {
      const float x = ALONG_W ? Elem<DT_IN>::load(in, base + (size_t)i * in_w + s + k)
                              : Elem<DT_IN>::load(in, base + (size_t)(s + k) * in_w + j);
      if (round_dt == ALG_BF16) acc = tap<ALG_BF16>(acc, wt[k], x, k == 0);
      else if (round_dt == ALG_F16) acc = tap<ALG_F16>(acc, wt[k], x, k == 0);
      else acc = tap<ALG_F32>(acc, wt[k], x, k == 0);
    }
    if (ROUND_BF) acc = round_dt == ALG_BF16 ? Elem<ALG_BF16>::round(acc) : (round_dt == ALG_F16 ? Elem<ALG_F16>::round(acc) : acc);
    Elem<DT_OUT>::store(out, idx, acc);
  }
}

template <int DT>
static int down_up_generic(const void* in, void* out, int64_t planes, int H, int W, int h1, int w1,
                           const DownUpTables& t, cudaStream_t st) {
  float *t1 = nullptr, *sm = nullptr, *t2 = nullptr;
  ALG_CUDA_OK(cudaMallocAsync(&t1, sizeof(float) * planes * H * w1, st));
  ALG_CUDA_OK(cudaMallocAsync(&sm, sizeof(float) * planes * h1 * w1, st));
  ALG_CUDA_OK(cudaMallocAsync(&t2, sizeof(float) * planes * h1 * W, st));
  auto grid = [](int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 16); };
  resample_pass_kernel<DT, ALG_F32, true, true><<<grid(planes * H * w1), 256, 0, st>>>(
      in, t1, planes, H, W, H, w1, t.ints + t.s_off[0], t.ints + t.c_off[0], t.weights + t.w_off[0], t.stride[0], DT);
  ALG_LAUNCH_OK();
  resample_pass_kernel<ALG_F32, ALG_F32, false, true><<<grid(planes * h1 * w1), 256, 0, st>>>(
      t1, sm, planes, H, w1, h1, w1, t.ints + t.s_off[1], t.ints + t.c_off[1], t.weights + t.w_off[1], t.stride[1], DT);
  ALG_LAUNCH_OK();
  resample_pass_kernel<ALG_F32, ALG_F32, true, true><<<grid(planes * h1 * W), 256, 0, st>>>(
      sm, t2, planes, h1, w1, h1, W, t.ints + t.s_off[2], t.ints + t.c_off[2], t.weights + t.w_off[2], t.stride[2], DT);
  ALG_LAUNCH_OK();
  resample_pass_kernel<ALG_F32, DT, false, false><<<grid(planes * H * W), 256, 0, st>>>(
      t2, out, planes, h1, W, H, W, t.ints + t.s_off[3], t.ints + t.c_off[3], t.weights + t.w_off[3], t.stride[3], DT);
  ALG_LAUNCH_OK();
  ALG_CUDA_OK(cudaFreeAsync(t1, st));
  ALG_CUDA_OK(cudaFreeAsync(sm, st));
  ALG_CUDA_OK(cudaFreeAsync(t2, st));
  return 0;
}

template <int DT>
static int down_up_dispatch(const void* in, void* out, int64_t planes, int H, int W, int h1, int w1,
                            cudaStream_t st) {
  DownUpTables t;
  if (int rc = get_tables(H, W, h1, w1, DT, &t)) return rc;
  constexpr int kStageBytes = DT == ALG_F32 ? 4 : 2;  // sizeof(Stage<DT>::type)
  constexpr int VW = 16 / kStageBytes;                // columns per 16-byte shared-memory vector
  const int vec = (W % VW == 0) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int w1p = vec ? (w1 + VW - 1) / VW * VW : w1;
  const int64_t a_elems = (std::max<int64_t>((int64_t)H * W, (int64_t)h1 * w1p) + 7) & ~(int64_t)7;
  const int64_t b_elems = (std::max<int64_t>((int64_t)H * w1p, (int64_t)h1 * W) + 7) & ~(int64_t)7;
  size_t smem = (size_t)(a_elems + b_elems) * kStageBytes + (size_t)(t.n_weights + t.n_ints) * sizeof(float);
  if (smem > 200 * 1024) return down_up_generic<DT>(in, out, planes, H, W, h1, w1, t, st);
  // second input stage for the cp.async prefetch of the next plane: when planes are 16-byte aligned, there is more than one plane
  // per CTA to overlap, and two stages still leave >= 3 CTAs per SM (measured: 60 x 104 fp32, 8 192 planes 141 -> 137 us; with only
  // 2 CTAs left -- 90 x 160 bf16 -- the lost occupancy costs more than the hidden load latency gains: 133 -> 141 us)
  static int prefetch_knob = -1;
  if (prefetch_knob < 0) {
    const char* e = getenv("ALG_DOWNUP_PREFETCH");
    prefetch_knob = e ? atoi(e) : 1;
  }
  const size_t smem2 = smem + (size_t)a_elems * kStageBytes;
  const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && ((int64_t)H * W) % VW == 0;
  const int prefetch = prefetch_knob && aligned && smem2 <= 72 * 1024 /* >= 3 CTAs per SM */ && planes > (int64_t)148 * ((220 * 1024) / (smem2 + 1024));
  if (prefetch) smem = smem2;
  // once per denoise step: set every time (the attribute is per-device state; no per-process cache to go stale)
  ALG_CUDA_OK(cudaFuncSetAttribute(down_up_fused_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
  const int grid = (int)std::min<int64_t>(planes, (int64_t)148 * ctas_per_sm);
  DownUpGeom g;
  for (int i = 0; i < 4; ++i) {
    g.s_off[i] = t.s_off[i];
    g.c_off[i] = t.c_off[i];
    g.w_off[i] = t.w_off[i];
    g.stride[i] = t.stride[i];
    g.max_cnt[i] = t.max_cnt[i];
  }
  g.n_ints = t.n_ints;
  g.n_weights = t.n_weights;
  down_up_fused_kernel<DT><<<grid, 256, smem, st>>>(in, out, planes, H, W, h1, w1, t.ints, t.weights, g, (int)a_elems,
                                                    (int)b_elems, vec, prefetch);
  ALG_LAUNCH_OK();
  return 0;
}

// --------------------------------------------------------------------------------------------
// gaussian blur
// --------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 63;
struct Taps {
  float w[kMaxTaps + 1];
};

// torchvision _get_gaussian_kernel1d with every op rounded to the tensor dtype
static void gaussian_taps(int k, double sigma, int dt, float* out) {
  const float half = (float)((k - 1) * 0.5);
  std::vector<float> x(k), pdf(k);
  if (k == 1) {
    x[0] = -half;
  } else {
    const float step = (half - (-half)) / (float)(k - 1);
    for (int i = 0; i < k; ++i)  // torch.linspace: two-sided formula
      x[i] = i < k / 2 ? (-half + step * (float)i) : (half - step * (float)(k - 1 - i));
  }
  float total = 0.f;
  for (int i = 0; i < k; ++i) {
    float v = host_round(x[i], dt);
    float q = host_round(v / (float)sigma, dt);
    q = host_round(q * q, dt);
    q = host_round(-0.5f * q, dt);
    pdf[i] = host_round(expf(q), dt);
    total += pdf[i];
  }
  total = host_round(total, dt);
  for (int i = 0; i < k; ++i) out[i] = host_round(pdf[i] / total, dt);
}

// Dense k x k taps (torchvision multiplies the two 1-D kernels into a [k, k] matrix in the tensor dtype and runs a
// depthwise conv2d), so the kernel is FMA-bound: k * k FMAs per output against 2 elements of HBM traffic.  Each thread owns
// 8 consecutive outputs of one row and slides a 12-float register window along the tile row: one 16-byte shared-memory
// load feeds 32 FMAs, and the four taps of a group arrive in one broadcast 16-byte load.  (The first version, one output
// column per thread with one LDS per FMA, was shared-memory-issue bound at 0.16 TB/s -- profiles/r01_lowpass.md.)
constexpr int GT_W = 64, GT_H = 32;  // output tile per CTA: 256 threads x 8 outputs

// The 8 threads of a quarter warp read 16-byte chunks 32 bytes apart: flipping the low chunk bit of every other group of
// eight chunks makes those reads bank-conflict free.
__device__ __forceinline__ int gswz(int x) { return x ^ (((x >> 5) & 1) << 2); }

__device__ __forceinline__ void lds128(float* v, uint32_t addr) {
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}

__host__ __device__ inline int gauss_tile_w(int k) { return (GT_W + 4 * (k >> 2) + 8 + 7) & ~7; }

template <int DT>
__global__ void __launch_bounds__(256) gaussian_kernel(const void* __restrict__ in, void* __restrict__ out, int H,
                                                       int W, int k, Taps taps) {
  extern __shared__ __align__(16) float smem[];
  const int r = k / 2;
  const int tw = gauss_tile_w(k), th = GT_H + k - 1;
  const int kp = (k + 3) & ~3;
  float* tile = smem;          // [th][tw], columns swizzled by gswz
  float* w2 = smem + th * tw;  // [k][kp], each product rounded to dtype (torch.mm in the tensor dtype), zero padded
  const int64_t plane = blockIdx.z;
  const int x0 = blockIdx.x * GT_W, y0 = blockIdx.y * GT_H;
  const size_t base = (size_t)plane * H * W;
  for (int i = threadIdx.x; i < k * kp; i += blockDim.x) {
    const int dy = i / kp, dx = i - dy * kp;
    w2[i] = dx < k ? Elem<DT>::round(taps.w[dy] * taps.w[dx]) : 0.f;
  }
  // tile fill, one warp per tile row: reflect (no edge repeat): -1 -> 1, H -> H-2.  Tiles may overhang the image:
  // clamp after reflecting.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ty = warp; ty < th; ty += 8) {
    int y = y0 + ty - r;
    y = y < 0 ? -y : y;
    y = y >= H ? 2 * H - 2 - y : y;
    y = min(max(y, 0), H - 1);
    const size_t rbase = base + (size_t)y * W;
    float* trow = tile + ty * tw;
    for (int tx = lane; tx < tw; tx += 32) {
      int x = x0 + tx - r;
      x = x < 0 ? -x : x;
      x = x >= W ? 2 * W - 2 - x : x;
      x = min(max(x, 0), W - 1);
      trow[gswz(tx)] = Elem<DT>::load(in, rbase + x);
    }
  }
  __syncthreads();
  const int xt = (threadIdx.x & 7) * 8, yy = threadIdx.x >> 3;
  const int kg = k >> 2, rem = k & 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  // shared-window byte addresses, advanced incrementally (the compiler otherwise rebuilds them every tile row)
  uint32_t row_s = (uint32_t)__cvta_generic_to_shared(tile + yy * tw);
  uint32_t w_s = (uint32_t)__cvta_generic_to_shared(w2);
  const uint32_t off0 = gswz(xt) * 4, off1 = gswz(xt + 4) * 4;
  for (int dy = 0; dy < k; ++dy, row_s += tw * 4, w_s += kp * 4) {
    float a[12];
    lds128(a, row_s + off0);
    lds128(a + 4, row_s + off1);
    int xw = xt + 8;
#pragma unroll 3
    for (int g = 0; g < kg; ++g, xw += 4) {
      lds128(a + 8, row_s + gswz(xw) * 4);
      float wq[4];
      lds128(wq, w_s + g * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q)  // taps in ascending dx, like the one-output-per-thread chain
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wq[q], a[i + q], acc[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i + 4];
    }
    {
      lds128(a + 8, row_s + gswz(xw) * 4);
      float wq[4];
      lds128(wq, w_s + kg * 16);
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (q < rem)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(wq[q], a[i + q], acc[i]);
    }
  }
  const int y = y0 + yy;
  if (y < H) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int x = x0 + xt + i;
      if (x < W) Elem<DT>::store(out, base + (size_t)y * W + x, acc[i]);
    }
  }
}

// float32 tensors (Wan's pixel-space ALG filters the fp32 image, wan:500-506) take the SEPARABLE form: a row pass into shared
// memory, then a column pass -- 2 k FMAs per output instead of k * k.  torchvision multiplies the two 1-D kernels into a dense
// [k, k] matrix first; in fp32 that product carries one more rounding (6e-8) than the separable evaluation, far inside the 2e-6 the
// fp32 filters are held to, so only the 16-bit dtypes (whose taps round to 8 / 11 bits and must be reproduced product by product)
// keep the dense kernel above.
__global__ void __launch_bounds__(256) gaussian_sep_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                               int k, Taps taps) {
  extern __shared__ __align__(16) float smem[];
  const int r = k / 2;
  const int th = GT_H + k - 1, tw = (GT_W + k - 1 + 3) & ~3;
  float* tile = smem;           // [th][tw]  reflect-padded input tile
  float* tmp = smem + th * tw;  // [th][GT_W] row-filtered
  const int64_t plane = blockIdx.z;
  const int x0 = blockIdx.x * GT_W, y0 = blockIdx.y * GT_H;
  const size_t base = (size_t)plane * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int ty = warp; ty < th; ty += 8) {
    int y = y0 + ty - r;
    y = y < 0 ? -y : y;
    y = y >= H ? 2 * H - 2 - y : y;
    y = min(max(y, 0), H - 1);
    const float* rsrc = in + base + (size_t)y * W;
    float* trow = tile + ty * tw;
    for (int tx = lane; tx < tw; tx += 32) {
      int x = x0 + tx - r;
      x = x < 0 ? -x : x;
      x = x >= W ? 2 * W - 2 - x : x;
      x = min(max(x, 0), W - 1);
      trow[tx] = __ldg(rsrc + x);
    }
  }
  __syncthreads();
  const int kg = k >> 2, rem = k & 3;
  // row pass: a work item = 8 consecutive columns of one tile row, 12-float sliding register window (16-byte shared loads)
  for (int item = threadIdx.x; item < th * (GT_W / 8); item += blockDim.x) {
    const int ty = item / (GT_W / 8), xs = (item - ty * (GT_W / 8)) * 8;
    const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(tile + ty * tw + xs);
    float a[12], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    lds128(a, row_s);
    lds128(a + 4, row_s + 16);
    int off = 32;
    for (int g = 0; g < kg; ++g, off += 16) {
      lds128(a + 8, row_s + off);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float w = taps.w[4 * g + q];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, a[i + q], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i + 4];
    }
    if (rem) {
      if (xs + 8 + 4 * kg < tw) lds128(a + 8, row_s + off);  // the last window may end exactly at the tile edge
      for (int q = 0; q < rem; ++q) {
        const float w = taps.w[4 * kg + q];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, a[i + q], acc[i]);
      }
    }
    float4* d = reinterpret_cast<float4*>(tmp + ty * GT_W + xs);
    d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  // column pass: thread = 8 consecutive columns of one output row
  const int xt = (threadIdx.x & 7) * 8, yy = threadIdx.x >> 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  uint32_t col_s = (uint32_t)__cvta_generic_to_shared(tmp + yy * GT_W + xt);
  for (int dy = 0; dy < k; ++dy, col_s += GT_W * 4) {
    float v[8];
    lds128(v, col_s);
    lds128(v + 4, col_s + 16);
    const float w = taps.w[dy];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, v[i], acc[i]);
  }
  const int y = y0 + yy;
  if (y < H) {
    float* orow = out + base + (size_t)y * W + x0 + xt;
    if (x0 + xt + 8 <= W && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0)) {
      reinterpret_cast<float4*>(orow)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(orow)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (x0 + xt + i < W) orow[i] = acc[i];
    }
  }
}

template <int DT>
static int gaussian_dispatch(const void* in, void* out, int64_t planes, int H, int W, int k, double sigma,
                             cudaStream_t st) {
  Taps taps;
  memset(&taps, 0, sizeof(taps));
  gaussian_taps(k, sigma, DT, taps.w);
  if (DT == ALG_F32) {
    static int dense = -1;
    if (dense < 0) {
      const char* e = getenv("ALG_GAUSS_DENSE");
      dense = e ? atoi(e) : 0;
    }
    if (!dense) {
      const int th = GT_H + k - 1, tw = (GT_W + k - 1 + 3) & ~3;
      const size_t smem_sep = ((size_t)th * tw + (size_t)th * GT_W) * sizeof(float);
      ALG_CUDA_OK(cudaFuncSetAttribute(gaussian_sep_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
        const int64_t np = std::min<int64_t>(65535, planes - p0);
        dim3 grid((W + GT_W - 1) / GT_W, (H + GT_H - 1) / GT_H, (unsigned)np);
        const size_t off = (size_t)p0 * H * W;
        gaussian_sep_f32_kernel<<<grid, 256, smem_sep, st>>>(reinterpret_cast<const float*>(in) + off, reinterpret_cast<float*>(out) + off,
                                                             H, W, k, taps);
        ALG_LAUNCH_OK();
      }
      return 0;
    }
  }
  const size_t smem = ((size_t)gauss_tile_w(k) * (GT_H + k - 1) + (size_t)k * ((k + 3) & ~3)) * sizeof(float);
  ALG_CUDA_OK(cudaFuncSetAttribute(gaussian_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
    const int64_t np = std::min<int64_t>(65535, planes - p0);
    dim3 grid((W + GT_W - 1) / GT_W, (H + GT_H - 1) / GT_H, (unsigned)np);
    const size_t off = (size_t)p0 * H * W * dtype_size(DT);
    gaussian_kernel<DT><<<grid, 256, smem, st>>>(reinterpret_cast<const char*>(in) + off,
                                                 reinterpret_cast<char*>(out) + off, H, W, k, taps);
    ALG_LAUNCH_OK();
  }
  return 0;
}

}  // namespace alg

extern "C" int alg_lowpass_down_up(const void* in, void* out, int64_t planes, int H, int W, int h1, int w1, int dtype,
                                   void* stream) {
  using namespace alg;
  ALG_REQUIRE(in && out, "down_up: null pointer");
  ALG_REQUIRE(planes >= 0 && H > 0 && W > 0 && h1 > 0 && w1 > 0, "down_up: bad geometry");
  if (planes == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case ALG_F32: return down_up_dispatch<ALG_F32>(in, out, planes, H, W, h1, w1, st);
    case ALG_BF16: return down_up_dispatch<ALG_BF16>(in, out, planes, H, W, h1, w1, st);
    case ALG_F16: return down_up_dispatch<ALG_F16>(in, out, planes, H, W, h1, w1, st);
  }
  ALG_REQUIRE(false, "down_up: unsupported dtype");
}

extern "C" int alg_lowpass_gaussian(const void* in, void* out, int64_t planes, int H, int W, int ksize, double sigma,
                                    int dtype, void* stream) {
  using namespace alg;
  ALG_REQUIRE(in && out && in != out, "gaussian: null or aliased pointers");
  ALG_REQUIRE(planes >= 0 && H > 0 && W > 0, "gaussian: bad geometry");
  ALG_REQUIRE(ksize >= 1 && ksize <= kMaxTaps && (ksize & 1), "gaussian: kernel size must be odd and <= 63");
  ALG_REQUIRE(ksize / 2 < H && ksize / 2 < W, "gaussian: reflect padding must be smaller than the image");
  ALG_REQUIRE(sigma > 0, "gaussian: sigma must be positive");
  if (planes == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case ALG_F32: return gaussian_dispatch<ALG_F32>(in, out, planes, H, W, ksize, sigma, st);
    case ALG_BF16: return gaussian_dispatch<ALG_BF16>(in, out, planes, H, W, ksize, sigma, st);
    case ALG_F16: return gaussian_dispatch<ALG_F16>(in, out, planes, H, W, ksize, sigma, st);
  }
  ALG_REQUIRE(false, "gaussian: unsupported dtype");
}

extern "C" int alg_gaussian_kernel1d(int ksize, double sigma, int dtype, float* taps_host) {
  using namespace alg;
  ALG_REQUIRE(taps_host && ksize >= 1 && ksize <= kMaxTaps, "gaussian_kernel1d: bad arguments");
  gaussian_taps(ksize, sigma, dtype, taps_host);
  return 0;
}
