// bf16 GEMM  D = epilogue(A * B^T + bias)  on tcgen05 tensor cores (SURVEY kernel K6).
//
// Replaces every nn.Linear of the DiT (cuBLASLt in the reference; call site wan:910).
//   * A [M, K] and B [N, K] (nn.Linear.weight layout) are both K-major: TMA loads 64-element
//     (128-byte, SWIZZLE_128B) K-slabs into a 4-stage shared-memory ring
//   * one elected thread issues tcgen05.mma (128 x BN x 16), accumulating in TMEM
//   * two TMEM accumulator stages: the epilogue warps drain tile i (tcgen05.ld -> bias / GELU /
//     gate*x + residual -> global) while the tensor core already works on tile i + 1
//   * persistent CTAs (one per SM), grouped tile order so a wave re-uses A and B out of L2
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp % 4 selects the TMEM lane quadrant it may read).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace alg {
namespace tc {

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box) {
  PFN_encodeTiled enc = get_encode_fn();
  ALG_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA: base address must be 16-byte aligned");
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstride[i - 1] = strides_elems[i] * 2;
      ALG_REQUIRE(gstride[i - 1] % 16 == 0, "TMA: row stride must be a multiple of 16 bytes");
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bdim,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ALG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return 0;
}

}  // namespace tc

#ifndef ALG_GEMM_CLUSTER_DEFAULT
#define ALG_GEMM_CLUSTER_DEFAULT 2
#endif

namespace gemm {
using namespace tc;

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int kGroupM = 16;

template <int BN, int CL = 1>
struct Cfg {
  static constexpr int kBytesA = BM * BK * 2;
  static constexpr int kBytesB = (BN / CL) * BK * 2;  // CL = 2: each CTA of the pair holds half of the B tile
  static constexpr int kStages = CL == 2 ? 6 : (BN == 256 ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : (BN == 96 ? 7 : 8))));
  // two accumulator stages at columns 0 and BN; the allocation is the next power of two >= 32
  static constexpr int kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * (kBytesA + kBytesB) + 1024 /*align*/ + 256 /*barriers*/;
};

struct Params {
  void* D;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* R;
  const void* gate;
  const void* gate_alt;
  int64_t M, N, K, ldd, rows_per_batch, gate_ld, gate_split_row;
  int epilogue, bias_per_row, out_f32, gate_bf16, gate_round;
  int m_tiles, n_tiles, k_blocks;
  int a_k_period;  // 0, or the A operand repeats along K with this period (multiple of BK)
  int a_tap_kb;    // 0, or implicit-convolution mode: K is n_taps groups of a_tap_kb k-blocks; group g reads A columns
                   // [0, a_tap_kb * BK) at rows shifted by a_tap_off[g] (out-of-range rows are zero-filled by TMA)
  int a_tap_off[32];
  const float* bias32;  // out_f32 only: fp32 bias [N] and fp32 residual [M, ldd] added in fp32 (the float32 VAE / encoder linears)
  const float* R32;
};

__device__ __forceinline__ void tile_coords(int t, int m_tiles, int n_tiles, int& m_blk, int& n_blk) {
  const int per_group = kGroupM * n_tiles;
  const int g = t / per_group;
  const int first_m = g * kGroupM;
  const int gsize = min(m_tiles - first_m, kGroupM);
  const int local = t - g * per_group;
  m_blk = first_m + local % gsize;
  n_blk = local / gsize;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  const float inner = kBeta * (x + kKappa * x * x * x);
  const float t = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * inner));  // tanh(inner)
  return 0.5f * x * (1.0f + t);
}

// CL = 2: CTA pairs (2-CTA clusters on one TPC) with cta_group::2 MMAs.  A pair works on a 256 x BN output tile: each CTA
// loads its own 128 rows of A and HALF of the B tile (BN / 2 rows), the leader CTA (cluster rank 0) issues ONE
// tcgen05.mma.cta_group::2 of shape 256 x BN x 16 per k-step that reads A and its B half from BOTH CTAs' shared memory
// (the B halves are broadcast to both tensor cores) and accumulates each CTA's 128 rows in that CTA's own TMEM.  Per
// k-block a CTA moves 32 KB L2 -> SM instead of 48 KB and its shared memory serves 32 KB of operand reads instead of
// 48 KB: less data movement per FLOP, which is what the power-capped clock responds to.
//   full[stage]      lives in the leader: one arrive.expect_tx for both CTAs' bytes; the peer's TMA completes on it too
//   empty[stage]     in each CTA, released by the leader's multicast commit
//   tmem_full[acc]   in each CTA (multicast commit); tmem_empty[acc] in the leader, 4 epilogue warps x 2 CTAs arrive
template <int BN, int CL>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  using C = Cfg<BN, CL>;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const int cluster_id = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int num_clusters = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;
  const int m_units = (p.m_tiles + CL - 1) / CL;  // scheduling unit = CL consecutive M tiles of one N tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kStages * C::kBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * (C::kBytesA + C::kBytesB));
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = m_units * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4 * CL);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CL > 1) {
      tmem_alloc_pair(tmem_slot, C::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, C::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything completes / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int m_blk, n_blk;
        tile_coords(t, m_units, p.n_tiles, m_blk, n_blk);
        m_blk = m_blk * CL + (int)crank;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          int ka = p.a_k_period ? (kb * BK) % p.a_k_period : kb * BK;
          int arow = m_blk * BM;
          if (p.a_tap_kb) {  // tap g of a convolution: the same activation columns, rows shifted by the tap's raster offset
            const int g = kb / p.a_tap_kb;
            ka = (kb - g * p.a_tap_kb) * BK;
            arow += p.a_tap_off[g];
          }
          if (CL > 1) {  // own A rows + own half of the B tile; the bytes of both CTAs complete on the leader's barrier
            if (crank == 0) mbar_arrive_expect_tx(&full[stage], CL * (C::kBytesA + C::kBytesB));
            tma_load_2d_pair(sA + stage * C::kBytesA, &tmA, &full[stage], ka, arow);
            tma_load_2d_pair(sB + stage * C::kBytesB, &tmB, &full[stage], kb * BK, n_blk * BN + (int)crank * (BN / CL));
          } else {
            mbar_arrive_expect_tx(&full[stage], C::kBytesA + C::kBytesB);
            tma_load_2d(sA + stage * C::kBytesA, &tmA, &full[stage], ka, arow);
            tma_load_2d(sB + stage * C::kBytesB, &tmB, &full[stage], kb * BK, n_blk * BN);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0 && elect_one()) {  // ===== MMA issuer (the leader CTA of a pair): elect.sync tells ptxas a single lane runs this region, so descriptors stay in
                        // uniform registers (a plain lane == 0 test wraps every UTCHMMA in an ELECT / R2UR loop) =====
      constexpr uint32_t idesc = make_idesc_bf16(BM * CL, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * C::kBytesA);
          const uint32_t b_addr = smem_u32(sB + stage * C::kBytesB);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (CL > 1)
              mma_ss_lo_pair(d_tmem, smem_desc_lo_sw128(a_addr + k * UMMA_K * 2), smem_desc_lo_sw128(b_addr + k * UMMA_K * 2),
                             idesc, (kb | k) != 0);
            else
              mma_ss(d_tmem, make_smem_desc_sw128(a_addr + k * UMMA_K * 2), make_smem_desc_sw128(b_addr + k * UMMA_K * 2),
                     idesc, (kb | k) != 0);
          }
          if (CL > 1) tc_commit_pair(&empty[stage]);  // the slot is free in BOTH CTAs once these MMAs have read it
          else tc_commit(&empty[stage]);
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CL > 1) tc_commit_pair(&tmem_full[acc]);  // accumulator complete -> both CTAs' epilogues
        else tc_commit(&tmem_full[acc]);
      }
    }
  } else {  // ===== epilogue warps 2..5 =====
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      int m_blk, n_blk;
      tile_coords(t, m_units, p.n_tiles, m_blk, n_blk);
      m_blk = m_blk * CL + (int)crank;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int64_t row = (int64_t)m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int64_t safe_row = row_ok ? row : 0;
      float row_bias = 0.f;
      if (p.bias && p.bias_per_row && row_ok) row_bias = __bfloat162float(p.bias[row]);
      const char* gate_row = nullptr;  // fp32 or bf16 [N] gate vector of this row's sample (alt vector below split_row)
      if (p.epilogue == ALG_EPI_GATE_RESIDUAL) {
        const int64_t b = safe_row / p.rows_per_batch, r_in = safe_row - b * p.rows_per_batch;
        gate_row = reinterpret_cast<const char*>(r_in < p.gate_split_row ? p.gate_alt : p.gate) +
                   b * p.gate_ld * (p.gate_bf16 ? 2 : 4);
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int64_t col0 = (int64_t)n_blk * BN + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        tmem_ld_wait();
        if (!row_ok) continue;
        const bool full_chunk = col0 + 32 <= p.N;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // ---- bias -------------------------------------------------------------------------
        if (p.bias) {
          if (p.bias_per_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += row_bias;
          } else if (full_chunk) {
            const uint4* b4 = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u = __ldg(b4 + g);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(h[e]);
                v[g * 8 + 2 * e] += f.x;
                v[g * 8 + 2 * e + 1] += f.y;
              }
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += __bfloat162float(p.bias[col0 + j]);
          }
        }
        // ---- activation / residual --------------------------------------------------------
        if (p.epilogue != ALG_EPI_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);  // the nn.Linear output tensor is bf16
          if (p.epilogue == ALG_EPI_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
          } else if (p.epilogue == ALG_EPI_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
          } else if (p.epilogue == ALG_EPI_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-v[j]));
          } else {  // RESIDUAL / GATE_RESIDUAL
            const __nv_bfloat16* rrow = p.R + row * p.ldd + col0;
            const bool gated = p.epilogue == ALG_EPI_GATE_RESIDUAL;
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 u = *reinterpret_cast<const uint4*>(rrow + g * 8);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
                float gg[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
                if (gated) {
                  if (p.gate_bf16) {
                    const uint4 gu = __ldg(reinterpret_cast<const uint4*>(gate_row + (col0 + g * 8) * 2));
                    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gu);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 f = __bfloat1622float2(gh[e]);
                      gg[2 * e] = f.x;
                      gg[2 * e + 1] = f.y;
                    }
                  } else {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate_row + (col0 + g * 8) * 4));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate_row + (col0 + g * 8 + 4) * 4));
                    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w;
                    gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
                  }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 f = __bfloat1622float2(h[e]);
                  const int j = g * 8 + 2 * e;
                  float y0 = gated ? __fmul_rn(v[j], gg[2 * e]) : v[j];
                  float y1 = gated ? __fmul_rn(v[j + 1], gg[2 * e + 1]) : v[j + 1];
                  if (gated && p.gate_round) {  // eager bf16 `x + gate * y`: the product is a bf16 tensor
                    y0 = bf16_round(y0);
                    y1 = bf16_round(y1);
                  }
                  v[j] = __fadd_rn(f.x, y0);
                  v[j + 1] = __fadd_rn(f.y, y1);
                }
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) {
                  const float x = __bfloat162float(rrow[j]);
                  float y = v[j];
                  if (gated) {
                    const float gv = p.gate_bf16
                                         ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(gate_row)[col0 + j])
                                         : reinterpret_cast<const float*>(gate_row)[col0 + j];
                    y = __fmul_rn(y, gv);
                    if (p.gate_round) y = bf16_round(y);
                  }
                  v[j] = __fadd_rn(x, y);
                }
            }
          }
        }
        // ---- store ------------------------------------------------------------------------
        if (p.out_f32) {
          float* drow = reinterpret_cast<float*>(p.D) + row * p.ldd + col0;
          if (p.bias32) {
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias32 + col0) + g);
                v[4 * g] += b4.x; v[4 * g + 1] += b4.y; v[4 * g + 2] += b4.z; v[4 * g + 3] += b4.w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] += p.bias32[col0 + j];
            }
          }
          if (p.R32) {
            const float* rrow32 = p.R32 + row * p.ldd + col0;
            if (full_chunk) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 r4 = *reinterpret_cast<const float4*>(rrow32 + g * 4);
                v[4 * g] += r4.x; v[4 * g + 1] += r4.y; v[4 * g + 2] += r4.z; v[4 * g + 3] += r4.w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] += rrow32[j];
            }
          }
          if (full_chunk) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(drow + g * 4) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) drow[j] = v[j];
          }
        } else {
          __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(p.D) + row * p.ldd + col0;
          if (full_chunk) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(drow + g * 8) = u;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) drow[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL > 1) mbar_arrive_remote(&tmem_empty[acc], 0);  // the leader's issuer waits for both CTAs' epilogues
        else mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA leaves while the pair's MMAs / commits / remote arrives may still touch it
  if (warp == 1) {
    tc_fence_after();
    if (CL > 1) tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int CL>
static int launch(const alg_gemm_t* g, cudaStream_t st) {
  using C = Cfg<BN, CL>;
  static std::atomic<uint64_t> attr_done{0};  // per-device bit: the attribute is device state
  int dev = 0;
  ALG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 64 || !(attr_done.load(std::memory_order_relaxed) >> dev & 1)) {
    ALG_CUDA_OK(cudaFuncSetAttribute(gemm_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    if (dev < 64) attr_done.fetch_or(uint64_t(1) << dev, std::memory_order_relaxed);
  }
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)(g->a_k_period ? g->a_k_period : g->K), (uint64_t)g->M}, strides[2] = {1, (uint64_t)g->lda};
    if (g->a_tap_kblocks) {
      dims[0] = (uint64_t)g->a_tap_kblocks * BK;
      if (g->a_rows > 0) dims[1] = (uint64_t)g->a_rows;
    }
    uint32_t box[2] = {BK, BM};
    if (int rc = make_tmap_bf16(&tmA, g->A, 2, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g->K, (uint64_t)g->N}, strides[2] = {1, (uint64_t)g->ldb};
    uint32_t box[2] = {BK, BN / CL};
    if (int rc = make_tmap_bf16(&tmB, g->B, 2, dims, strides, box)) return rc;
  }
  Params p;
  p.D = g->D;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(g->bias);
  p.R = reinterpret_cast<const __nv_bfloat16*>(g->R);
  p.gate = g->gate;
  p.M = g->M;
  p.N = g->N;
  p.K = g->K;
  p.ldd = g->ldd;
  p.rows_per_batch = g->rows_per_batch > 0 ? g->rows_per_batch : g->M;
  p.gate_ld = g->gate_ld;
  p.gate_alt = g->gate_alt;
  p.gate_split_row = g->gate_alt ? g->gate_split_row : 0;
  p.gate_bf16 = g->gate_dtype == ALG_BF16;
  p.gate_round = g->gate_round;
  p.epilogue = g->epilogue;
  p.bias_per_row = g->bias_per_row;
  p.out_f32 = g->out_f32;
  p.m_tiles = (int)((g->M + BM - 1) / BM);
  p.n_tiles = (int)((g->N + BN - 1) / BN);
  p.k_blocks = (int)((g->K + BK - 1) / BK);
  p.a_k_period = (int)g->a_k_period;
  p.bias32 = g->out_f32 ? g->bias_f32 : nullptr;
  p.R32 = g->out_f32 ? g->residual_f32 : nullptr;
  p.a_tap_kb = g->a_tap_kblocks;
  for (int i = 0; i < 32; ++i) p.a_tap_off[i] = (g->a_tap_kblocks && i < g->a_n_taps) ? g->a_tap_offsets[i] : 0;
  const int units = ((p.m_tiles + CL - 1) / CL) * p.n_tiles;
  if (CL == 1) {
    const int grid = std::min(units, num_sms());
    gemm_kernel<BN, CL><<<grid, kThreads, C::kSmemBytes, st>>>(tmA, tmB, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(std::min(units, num_sms() / CL) * CL));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ALG_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, CL>, tmA, tmB, p));
  }
  ALG_LAUNCH_OK();
  return 0;
}

}  // namespace gemm
}  // namespace alg

extern "C" int alg_gemm_bf16(const alg_gemm_t* g, void* stream) {
  using namespace alg;
  ALG_REQUIRE(g && g->A && g->B && g->D, "gemm: null pointer");
  ALG_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm: empty problem");
  ALG_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (16-byte TMA strides)");
  ALG_REQUIRE(g->a_k_period >= 0 && g->a_k_period < (int64_t(1) << 30) &&
                  (g->a_k_period == 0 || (g->a_k_period % 64 == 0 && g->K % g->a_k_period == 0)),
              "gemm: a_k_period must be a multiple of 64 that divides K");
  if (g->a_tap_kblocks) {
    ALG_REQUIRE(g->a_tap_kblocks > 0 && g->a_n_taps > 0 && g->a_n_taps <= 32 && g->a_tap_offsets && g->a_k_period == 0 &&
                    (int64_t)g->a_n_taps * g->a_tap_kblocks * 64 == g->K,
                "gemm: implicit-convolution mode needs 1..32 taps, K = taps * a_tap_kblocks * 64 and no a_k_period");
    ALG_REQUIRE(g->lda >= (int64_t)g->a_tap_kblocks * 64 && g->ldb >= g->K && g->ldd >= g->N, "gemm: leading dimension too small");
  } else
  ALG_REQUIRE(g->lda >= (g->a_k_period ? g->a_k_period : g->K) && g->ldb >= g->K && g->ldd >= g->N,
              "gemm: leading dimension too small");
  ALG_REQUIRE(g->ldd % (g->out_f32 ? 4 : 8) == 0, "gemm: ldd must keep rows 16-byte aligned");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(g->D) & 15) == 0, "gemm: D must be 16-byte aligned");
  ALG_REQUIRE(g->epilogue >= ALG_EPI_NONE && g->epilogue <= ALG_EPI_SILU, "gemm: unknown epilogue");
  if (g->epilogue == ALG_EPI_RESIDUAL || g->epilogue == ALG_EPI_GATE_RESIDUAL) {
    ALG_REQUIRE(g->R && !g->out_f32, "gemm: residual epilogues need R and bf16 output");
    ALG_REQUIRE((reinterpret_cast<uintptr_t>(g->R) & 15) == 0, "gemm: R must be 16-byte aligned");
  }
  if (g->epilogue == ALG_EPI_GATE_RESIDUAL)
  {
    ALG_REQUIRE(g->gate_dtype == ALG_F32 || g->gate_dtype == ALG_BF16, "gemm: gate dtype must be f32 or bf16");
    ALG_REQUIRE(g->gate && g->gate_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(g->gate) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(g->gate_alt) & 15) == 0,
                "gemm: gate vectors must be 16-byte aligned with gate_ld % 8 == 0");
    ALG_REQUIRE(g->gate_split_row == 0 || g->gate_alt, "gemm: gate_split_row needs gate_alt");
  }
  if (g->bias && !g->bias_per_row)
    ALG_REQUIRE((reinterpret_cast<uintptr_t>(g->bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  if (g->bias_f32 || g->residual_f32) {
    ALG_REQUIRE(g->out_f32 && !g->bias && g->epilogue == ALG_EPI_NONE, "gemm: bias_f32 / residual_f32 need out_f32, no bf16 bias and no epilogue");
    ALG_REQUIRE(((reinterpret_cast<uintptr_t>(g->bias_f32) | reinterpret_cast<uintptr_t>(g->residual_f32)) & 15) == 0,
                "gemm: bias_f32 / residual_f32 must be 16-byte aligned");
  }
  if (int rc = alg_check_device()) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static int force_bn = -1, cluster = -1;  // experiment knobs
  static int64_t pair_min_m = 128;
  if (force_bn < 0) {
    const char* e = getenv("ALG_GEMM_BN");
    force_bn = e ? atoi(e) : 0;
    e = getenv("ALG_GEMM_CLUSTER");
    cluster = e ? atoi(e) : ALG_GEMM_CLUSTER_DEFAULT;
    e = getenv("ALG_GEMM_PAIR_MIN_M");
    if (e) pair_min_m = atoll(e);
  }
  if (force_bn == 128) return gemm::launch<128, 1>(g, st);
  if (force_bn == 64) return gemm::launch<64, 1>(g, st);
  if (g->N % 256 == 0 || g->N > 512) {
    // CTA pairs (cta_group::2 MMAs on 256 x 256 tiles)
    if (cluster == 2 && g->M > pair_min_m) return gemm::launch<256, 2>(g, st);
    return gemm::launch<256, 1>(g, st);
  }
  // 64 < N < 128 (the 96-channel level of the Wan VAE) takes the 128-wide tile too: a 128 x 64 x 16 SS MMA is operand-fetch
  // bound (48 cycles for 32 cycles of math), so one 3/4-used 128-wide tile beats a 64 + a half-used 64 tile
  // tiles as wide as the problem for the 96- and 192-channel levels of the float32 VAEs (N = 96: one 96-wide tile instead of a
  // 3/4-used 128-wide one; N = 192: one 192-wide tile instead of 128 + a half-used 128), which also reads A once
  if (g->N > 64 && g->N <= 96) return gemm::launch<96, 1>(g, st);
  if ((g->N > 128 && g->N <= 192) || (g->N > 256 && g->N <= 384)) return gemm::launch<192, 1>(g, st);
  if (g->N % 128 == 0 || g->N > 64) return gemm::launch<128, 1>(g, st);
  return gemm::launch<64, 1>(g, st);
}
