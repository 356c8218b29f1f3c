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
  int s_off[4], c_off[4], w_off[4], stride[4];
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

constexpr int kRegTaps = 8;  // taps cached in registers; longer bands read the rest from the table

// One tap of interpolate_aa_single_dim (UpSample.cuh:349-365): taps and samples are widened to fp32 and summed with
// `output += t * wts` (an FMA chain under nvcc's default contraction), for every tensor dtype.
template <int DT>
__device__ __forceinline__ float tap(float acc, float w, float x, bool first) {
  return first ? w * x : fmaf(w, x, acc);
}

// dst[r][j] = sum_k w * src[r][start_j + k]      (resample along the contiguous axis)
// thread (tx, ty): columns j = tx + 32 m with that column's taps in registers, rows r = ty + 8 m'.
template <int DT>
__device__ __forceinline__ void pass_w(const float* __restrict__ src, float* __restrict__ dst, int rows, int in_w,
                                       int out_w, BandPtr b) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ny = blockDim.x >> 5;
  for (int j = tx; j < out_w; j += 32) {
    const int s = b.start[j], n = b.cnt[j];
    const float* wt = b.w + j * b.stride;
    float w[kRegTaps];
#pragma unroll
    for (int k = 0; k < kRegTaps; ++k) w[k] = k < n ? wt[k] : 0.f;
    for (int r = ty; r < rows; r += ny) {
      const float* p = src + r * in_w + s;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kRegTaps; ++k)
        if (k < n) acc = tap<DT>(acc, w[k], p[k], k == 0);
      for (int k = kRegTaps; k < n; ++k) acc = tap<DT>(acc, wt[k], p[k], false);
      dst[r * out_w + j] = Elem<DT>::round(acc);  // ATen keeps the row-pass result in a scalar_t buffer
    }
  }
}
// dst[i][c] = sum_k w * src[start_i + k][c]       (resample along the strided axis)
// thread (tx, ty): output rows i = ty + 8 m (taps warp-uniform, in registers), columns c = tx + 32 m'.
template <int DT, bool ROUND, bool TO_GLOBAL>
__device__ __forceinline__ void pass_h(const float* __restrict__ src, void* __restrict__ dst, int cols, int out_h,
                                       BandPtr b) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, ny = blockDim.x >> 5;
  for (int i = ty; i < out_h; i += ny) {
    const int s = b.start[i], n = b.cnt[i];
    const float* wt = b.w + i * b.stride;
    float w[kRegTaps];
#pragma unroll
    for (int k = 0; k < kRegTaps; ++k) w[k] = k < n ? wt[k] : 0.f;
    for (int c = tx; c < cols; c += 32) {
      const float* p = src + s * cols + c;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kRegTaps; ++k)
        if (k < n) acc = tap<DT>(acc, w[k], p[k * cols], k == 0);
      for (int k = kRegTaps; k < n; ++k) acc = tap<DT>(acc, wt[k], p[k * cols], false);
      if (TO_GLOBAL) {
        Elem<DT>::store(dst, (size_t)i * cols + c, acc);
      } else {
        reinterpret_cast<float*>(dst)[i * cols + c] = ROUND ? Elem<DT>::round(acc) : acc;
      }
    }
  }
}

struct DownUpGeom {
  int s_off[4], c_off[4], w_off[4], stride[4];
  int n_ints, n_weights;
};

template <int DT>
__global__ void __launch_bounds__(256) down_up_fused_kernel(const void* __restrict__ in, void* __restrict__ out,
                                                            int64_t planes, int H, int W, int h1, int w1,
                                                            const int* __restrict__ ints,
                                                            const float* __restrict__ weights, DownUpGeom g,
                                                            int bufA_elems, int bufB_elems) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;                           // in [H, W]      -> small [h1, w1]
  float* B = smem + bufA_elems;              // t1 [H, w1]     -> t2 [h1, W]
  float* sw = B + bufB_elems;                // operator taps
  int* si = reinterpret_cast<int*>(sw + g.n_weights);  // operator starts / counts
  for (int i = threadIdx.x; i < g.n_weights; i += blockDim.x) sw[i] = weights[i];
  for (int i = threadIdx.x; i < g.n_ints; i += blockDim.x) si[i] = ints[i];
  BandPtr bw_down{si + g.s_off[0], si + g.c_off[0], sw + g.w_off[0], g.stride[0]};
  BandPtr bh_down{si + g.s_off[1], si + g.c_off[1], sw + g.w_off[1], g.stride[1]};
  BandPtr bw_up{si + g.s_off[2], si + g.c_off[2], sw + g.w_off[2], g.stride[2]};
  BandPtr bh_up{si + g.s_off[3], si + g.c_off[3], sw + g.w_off[3], g.stride[3]};
  using T = typename Elem<DT>::type;
  const int HW = H * W;
  for (int64_t plane = blockIdx.x; plane < planes; plane += gridDim.x) {
    const T* src = reinterpret_cast<const T*>(in) + plane * HW;
    T* dst = reinterpret_cast<T*>(out) + plane * HW;
    // ---- stage the plane: 16-byte vector loads when the plane start is aligned -------------
    constexpr int VEC = 16 / sizeof(T);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (HW % VEC) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      for (int i = threadIdx.x; i < HW / VEC; i += blockDim.x) {
        uint4 v = __ldg(s4 + i);
        const T* e = reinterpret_cast<const T*>(&v);
        if (VEC == 4) {
          *reinterpret_cast<float4*>(A + i * 4) = make_float4((float)e[0], (float)e[1], (float)e[2], (float)e[3]);
        } else {
#pragma unroll
          for (int k = 0; k < VEC; ++k) A[i * VEC + k] = (float)e[k];
        }
      }
    } else {
      for (int i = threadIdx.x; i < HW; i += blockDim.x) A[i] = (float)src[i];
    }
    __syncthreads();
    pass_w<DT>(A, B, H, W, w1, bw_down);  // t1 [H, w1]
    __syncthreads();
    pass_h<DT, true, false>(B, A, w1, h1, bh_down);  // small [h1, w1], rounded to dtype (reference materialises it)
    __syncthreads();
    pass_w<DT>(A, B, h1, w1, W, bw_up);  // t2 [h1, W]
    __syncthreads();
    pass_h<DT, false, true>(B, dst, W, H, bh_up);  // out [H, W]
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
  const int64_t a_elems = std::max<int64_t>((int64_t)H * W, (int64_t)h1 * w1);
  const int64_t b_elems = std::max<int64_t>((int64_t)H * w1, (int64_t)h1 * W);
  const size_t smem = (size_t)(a_elems + b_elems + t.n_weights + t.n_ints) * sizeof(float);
  if (smem > 200 * 1024) return down_up_generic<DT>(in, out, planes, H, W, h1, w1, t, st);
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
  }
  g.n_ints = t.n_ints;
  g.n_weights = t.n_weights;
  down_up_fused_kernel<DT><<<grid, 256, smem, st>>>(in, out, planes, H, W, h1, w1, t.ints, t.weights, g, (int)a_elems,
                                                    (int)b_elems);
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

constexpr int GT_W = 32, GT_H = 32;  // output tile per CTA

template <int DT>
__global__ void __launch_bounds__(256) gaussian_kernel(const void* __restrict__ in, void* __restrict__ out, int H,
                                                       int W, int k, Taps taps) {
  extern __shared__ __align__(16) float smem[];
  const int r = k / 2;
  const int tw = GT_W + k - 1, th = GT_H + k - 1;
  float* tile = smem;          // [th][tw]
  float* w2 = smem + th * tw;  // [k][k], each product rounded to dtype (torch.mm in the tensor dtype)
  const int64_t plane = blockIdx.z;
  const int x0 = blockIdx.x * GT_W, y0 = blockIdx.y * GT_H;
  const size_t base = (size_t)plane * H * W;
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) w2[i] = Elem<DT>::round(taps.w[i / k] * taps.w[i % k]);
  for (int i = threadIdx.x; i < th * tw; i += blockDim.x) {
    int ty = i / tw, tx = i - ty * tw;
    int y = y0 + ty - r, x = x0 + tx - r;
    // reflect (no edge repeat): -1 -> 1, H -> H-2.  Tiles may overhang the image: clamp after reflecting.
    if (y < 0) y = -y;
    if (y >= H) y = 2 * H - 2 - y;
    if (x < 0) x = -x;
    if (x >= W) x = 2 * W - 2 - x;
    y = min(max(y, 0), H - 1);
    x = min(max(x, 0), W - 1);
    tile[i] = Elem<DT>::load(in, base + (size_t)y * W + x);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int dy = 0; dy < k; ++dy) {
    for (int dx = 0; dx < k; ++dx) {
      const float w = w2[dy * k + dx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(w, tile[(ty0 + 8 * q + dy) * tw + tx + dx], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int y = y0 + ty0 + 8 * q, x = x0 + tx;
    if (y < H && x < W) Elem<DT>::store(out, base + (size_t)y * W + x, acc[q]);
  }
}

template <int DT>
static int gaussian_dispatch(const void* in, void* out, int64_t planes, int H, int W, int k, double sigma,
                             cudaStream_t st) {
  Taps taps;
  memset(&taps, 0, sizeof(taps));
  gaussian_taps(k, sigma, DT, taps.w);
  const size_t smem = ((size_t)(GT_W + k - 1) * (GT_H + k - 1) + (size_t)k * k) * sizeof(float);
  ALG_CUDA_OK(cudaFuncSetAttribute(gaussian_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
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
