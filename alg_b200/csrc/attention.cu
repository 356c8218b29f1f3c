// Non-causal flash attention forward on tcgen05 / TMEM (SURVEY kernel K8).
//
// Replaces F.scaled_dot_product_attention inside the DiT (reference call site wan:910; Wan self-attention
// N = 32 760 tokens, 40 heads x 128).  One CTA owns TWO 128-row query tiles of one (batch, head) and streams the
// K / V^T tiles of that head, 64 keys per step, through a TMA-fed shared-memory ring:
//
//   warp 8  : TMA producer      Q0, Q1 once; K_j and V^T_j rings (SWIZZLE_128B, K-major), 4 stages each
//   warps 9, 10 : MMA issuers (one per query tile)
//                               S_i(j) = Q_i K_j^T   (tcgen05.mma SS, 128 x 64 x D    -> TMEM S_i[j & 1])
//                               O_i   += P_i(j) V_j  (tcgen05.mma TS, A = P_i in TMEM -> TMEM O_i)
//   warps 0-3 / 4-7 : softmax warpgroup for tile 0 / 1 (one query row per thread):
//                               tcgen05.ld S -> online softmax in fp32 (exp2; O in TMEM is rescaled only when the
//                               running max grows by > 2^8) -> bf16 P written back over S with tcgen05.st
//
// S is double-buffered in TMEM and issued TWO steps ahead (S_i(j+2) right behind P_i(j) V_j), so a softmax warpgroup
// never waits for the tensor pipe in steady state: when it has published P_i(j), S_i(j+1) is already complete.  The
// first version of this kernel used 128-key steps with a single S buffer per tile, which made every step a serial
// chain  softmax -> PV -> S -> softmax  and capped the tensor pipe at ~56 % busy (profiles/r01_attention.md).
//
// TMEM (512 columns): S0 [0,64) [64,128)   S1 [128,192) [192,256)   O0 [256,256+D)   O1 [384,384+D).
// P_i(j) (bf16 pairs) overwrites the first 32 columns of the S buffer it was computed from.
// V is consumed TRANSPOSED ([head_dim, n_kv], produced directly by the V projection GEMM with swapped operands) so
// that both MMAs use K-major operands.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "tc_common.cuh"

namespace alg {
namespace attn {
using namespace tc;

constexpr int BQ = 128;  // query rows per tile (two tiles per CTA)
constexpr int BKV = 64;  // keys per pipeline step
constexpr int kStages = 4;  // barrier-array stride (the SHORT variant uses only the first two stages)
constexpr float kRescaleThreshold = 8.0f;  // log2 units
#ifndef ALG_ATTN_POLY_DEFAULT
#define ALG_ATTN_POLY_DEFAULT 8
#endif
#ifndef ALG_ATTN_SHORT_MAX_DEFAULT
#define ALG_ATTN_SHORT_MAX_DEFAULT 1024
#endif
#ifndef ALG_ATTN_SPLIT_DEFAULT
#define ALG_ATTN_SPLIT_DEFAULT 0
#endif
#ifndef ALG_ATTN_PAIR_DEFAULT
#define ALG_ATTN_PAIR_DEFAULT 0
#endif
#ifndef ALG_ATTN_PS_DEFAULT
#define ALG_ATTN_PS_DEFAULT 0
#endif
#ifndef ALG_ATTN_S128_DEFAULT
#define ALG_ATTN_S128_DEFAULT 0
#endif
#ifndef ALG_ATTN_MC_DEFAULT
#define ALG_ATTN_MC_DEFAULT 1  // r02: +0.7 % (Wan) / +1.0 % (Hunyuan shape) alone, neutral to +0.5 % in the step; half the L2 -> SM reads
#endif

// PAIR = 1: the CTA is one half of a CTA pair (cta_group::2 MMAs, see attention_kernel): a K stage holds this CTA's 32 of
// the step's 64 keys and a V^T stage this CTA's half of the head_dim rows.
template <int D, int PAIR = 0>
struct Cfg {
  static constexpr int kCtas = PAIR ? 2 : 1;
  static constexpr int kBytesQ = BQ * D * 2;   // one query tile
  static constexpr int kBytesK = BKV * D * 2 / kCtas;  // one K stage  [64 (32) keys][D]
  static constexpr int kBytesV = D * BKV * 2 / kCtas;  // one V^T stage [D (D / 2)][64 keys]
  static constexpr int kXchBytes = 2 * 2 * 2 * BQ * 4;  // SPLIT: [tile][step parity][half][row] partial row maxima
  static constexpr int smem_bytes(int tiles, int stages) {
    return tiles * kBytesQ + stages * (kBytesK + kBytesV) + 1024 + 512 + kXchBytes;
  }
  static constexpr int kSubQ = BQ * 128;   // bytes of one [128 rows][64 elem] swizzle sub-tile of Q
  static constexpr int kSubK = BKV * 128 / kCtas;  // bytes of one [64 (32) keys][64 elem] sub-tile of K
};

// Timeline probe (scripts/attn_trace.py builds a second library with -DALG_ATTN_TRACE): clock64() stamps of one CTA's phases,
// 32 slots per CTA.  Compiled out of the product library.
#ifdef ALG_ATTN_TRACE
__device__ long long* g_attn_trace = nullptr;
#define ATTN_TRACE(slot)                                                                                                     \
  do {                                                                                                                       \
    if (g_attn_trace)                                                                                                        \
      g_attn_trace[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + (slot)] = clock64();      \
  } while (0)
#else
#define ATTN_TRACE(slot) \
  do {                   \
  } while (0)
#endif

struct Params {
  __nv_bfloat16* O;
  int64_t o_bs, o_rs;
  int n_q, n_kv, heads;
  float scale_log2;  // scale * log2(e)
  int accumulate;
  int stagger;  // S128: cycles tile 1's issuer waits before its first S (experiment knob ALG_ATTN_STAGGER)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}
// 2^x for a pair on the FMA pipe instead of MUFU (relieves the 16/clk/SM ex2 unit, which at 128 x 128 exps per
// 128 x 128 x 128 MMA pair is exactly as busy as the tensor pipe): Cody-Waite split x = n + f, |f| <= 0.5, degree-3
// minimax 2^f (7.5e-5 relative, below the bf16 rounding of P), exponent patched in with integer adds.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);  // 1.5 * 2^23: low mantissa bits of t hold rint(x)
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(0.0551716685295105f, 0.0551716685295105f), f,
                        make_float2(0.2426111251115799f, 0.2426111251115799f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}

// ---- MMA issuer (one elected thread) -----------------------------------------------------------------------------
// The first version rebuilt both 64-bit shared-memory descriptors for every tcgen05.mma and indexed stages / barriers
// with runtime j % kStages: ~330 SASS instructions per 64-key step on ONE thread, i.e. ~1 460 cycles per step against
// 1 024 cycles of tensor work -- the issuer, not the softmax, paced the kernel (profiles/r01_attention.md, a04).  Here the
// step is unrolled over j % 4 so stage, S-buffer, barrier offsets and barrier parities are compile-time, and every
// operand descriptor is (base low word) + (compile-time offset).
struct MmaCtx {
  uint32_t tmem, tmem_o, bar, q_lo, k_lo, v_lo;  // tmem_o = first O column (behind the S buffers of all tiles)
  int n_steps;
  int stagger;
};
constexpr int kBarKFull = 1, kBarKEmpty = 1 + kStages, kBarVFull = 1 + 2 * kStages, kBarVEmpty = 1 + 3 * kStages,
              kBarSFull = 1 + 4 * kStages, kBarPFull = kBarSFull + 4, kBarODone = kBarPFull + 4, kBarOFull = kBarODone + 2,
              kBarTurn = kBarOFull + 2;  // S128: the two tiles' issuers take turns with their PV + S blocks
static_assert(kStages == 4, "the issuer's compile-time parities assume a four-step unroll");
// MmaCtx::tiles = query tiles of the CTA: O accumulators start behind the S buffers of all tiles

// NOTE on the S shape: with 64-key (N = 64) S MMAs both operands stream from shared memory at 6 KB per MMA; the tensor
// core fetches 128 B/clk, so a 128x64x16 SS MMA takes 48 cycles instead of its 32-cycle floor (measured:
// scripts/microbench/mma_rate.cu; TS and N >= 128 run at the floor).  A paired 128-key S MMA runs at the floor but needs
// both S buffers of the tile at once, which serialises S -> softmax -> softmax -> PV per tile; measured slower
// (1 120 vs 1 270 TFLOP/s, profiles/r01_attention.md) because the softmax warps, not the tensor pipe, pace the kernel.
template <int PAIR>
__device__ __forceinline__ void commit(uint32_t bar_addr) {  // PAIR: the arrive lands in both CTAs of the pair
  if constexpr (PAIR) tc_commit_pair_a(bar_addr, 3);
  else tc_commit_a(bar_addr);
}
// K / V stage release.  MC (TMA-multicast cluster of two independent CTAs): a stage is refilled by BOTH CTAs' producers (each
// multicasts half of the tile into both shared memories), so the release has to reach both CTAs' `empty` barriers.
template <int PAIR, int MC>
__device__ __forceinline__ void commit_ring(uint32_t bar_addr) {
  if constexpr (MC) {  // own barrier + the peer's (shared-window addresses of a 2-CTA cluster differ in the rank bit, 1 << 24)
    tc_commit_a(bar_addr);
    tc_commit_a(bar_addr ^ 0x01000000u);
  } else {
    commit<PAIR>(bar_addr);
  }
}
template <int D, int I, int BUF, int ST, int PAIR>
__device__ __forceinline__ void issue_s(const MmaCtx& c) {  // S_I = Q_I K^T (stage ST) into S buffer BUF
  using C = Cfg<D, PAIR>;
  constexpr uint32_t idesc_s = make_idesc_bf16(BQ * C::kCtas, BKV);
  const uint32_t d = c.tmem + I * 128 + BUF * 64;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const uint32_t qoff = (I * C::kBytesQ + (ks >> 2) * C::kSubQ + (ks & 3) * 32) >> 4;
    const uint32_t koff = (ST * C::kBytesK + (ks >> 2) * C::kSubK + (ks & 3) * 32) >> 4;
    if constexpr (PAIR) mma_ss_lo_pair(d, c.q_lo + qoff, c.k_lo + koff, idesc_s, ks != 0);
    else mma_ss_lo(d, c.q_lo + qoff, c.k_lo + koff, idesc_s, ks != 0);
  }
  commit<PAIR>(c.bar + 8 * (kBarSFull + I * 2 + BUF));
}
template <int D, int I, int BUF, int ST, int PAIR>
__device__ __forceinline__ void issue_pv(const MmaCtx& c, uint32_t acc_first) {  // O_I (+)= P_I (buffer BUF) V (stage ST)
  using C = Cfg<D, PAIR>;
  constexpr uint32_t idesc_o = make_idesc_bf16(BQ * C::kCtas, D);
  const uint32_t d = c.tmem_o + I * 128, a = c.tmem + I * 128 + BUF * 64;
#pragma unroll
  for (int ks = 0; ks < BKV / 16; ++ks) {
    if constexpr (PAIR) mma_ts_lo_pair(d, a + ks * 8, c.v_lo + ((ST * C::kBytesV + ks * 32) >> 4), idesc_o, ks == 0 ? acc_first : 1u);
    else mma_ts_lo(d, a + ks * 8, c.v_lo + ((ST * C::kBytesV + ks * 32) >> 4), idesc_o, ks == 0 ? acc_first : 1u);
  }
  commit<PAIR>(c.bar + 8 * (kBarODone + I));
}
// step j = 4 m + JJ of query tile I; ph = m & 1 (parity of the K/V stage ring at this step).  Each tile has its OWN issuer
// warp: one thread issuing for both tiles still needed ~230 instructions (~1 000+ cycles) per step; two threads halve that,
// and the tiles' MMA streams are independent (disjoint TMEM), sharing only the K/V stage barriers (two arrivals each).
template <int D, int I, int JJ, int STAGES, int PAIR, int MC = 0>
__device__ __forceinline__ void mma_tile_step(const MmaCtx& c, const int j, uint32_t ph) {
  constexpr int BUF = JJ & 1, ST = JJ % STAGES, STN = (JJ + 2) % STAGES;
  constexpr uint32_t p_par = (JJ >> 1) & 1;                      // ((4 m + JJ) >> 1) & 1
  uint32_t phn = (JJ + 2 >= kStages) ? (ph ^ 1u) : ph;           // parity of (j + 2) / STAGES (four stages)
  if constexpr (STAGES == 2) {                                   // two stages: (4 m + JJ) / 2 = 2 m + JJ / 2 -> compile-time
    ph = (JJ >> 1) & 1;
    phn = ((JJ + 2) >> 1) & 1;
  }
  const bool has_next = j + 2 < c.n_steps, last = j == c.n_steps - 1;
  mbar_wait_a(c.bar + 8 * (kBarPFull + I * 2 + BUF), p_par);
  mbar_wait_a(c.bar + 8 * (kBarVFull + ST), ph);
  tc_fence_after();
  issue_pv<D, I, BUF, ST, PAIR>(c, j > 0);
  commit_ring<PAIR, MC>(c.bar + 8 * (kBarVEmpty + ST));
  if (last) commit<PAIR>(c.bar + 8 * (kBarOFull + I));
  if (has_next) {
    mbar_wait_a(c.bar + 8 * (kBarKFull + STN), phn);
    tc_fence_after();
    issue_s<D, I, BUF, STN, PAIR>(c);  // reuses the S buffer whose P was consumed by the PV just issued (in-order tensor pipe)
    commit_ring<PAIR, MC>(c.bar + 8 * (kBarKEmpty + STN));
  }
}
template <int D, int I, int STAGES, int PAIR, int MC = 0>
__device__ __forceinline__ void mma_tile_loop(const MmaCtx& c) {
  mbar_wait_a(c.bar, 0);  // q_full
  if (I == 0) ATTN_TRACE(4);
  mbar_wait_a(c.bar + 8 * (kBarKFull + 0), 0);
  tc_fence_after();
  issue_s<D, I, 0, 0, PAIR>(c);
  commit_ring<PAIR, MC>(c.bar + 8 * (kBarKEmpty + 0));
  if (I == 0) ATTN_TRACE(5);
  if (c.n_steps > 1) {
    mbar_wait_a(c.bar + 8 * (kBarKFull + 1), 0);
    tc_fence_after();
    issue_s<D, I, 1, 1, PAIR>(c);
    commit_ring<PAIR, MC>(c.bar + 8 * (kBarKEmpty + 1));
  }
  uint32_t ph = 0;
#pragma unroll 1
  for (int j0 = 0; j0 < c.n_steps; j0 += 4, ph ^= 1u) {
    mma_tile_step<D, I, 0, STAGES, PAIR, MC>(c, j0, ph);
    if (j0 + 1 < c.n_steps) mma_tile_step<D, I, 1, STAGES, PAIR, MC>(c, j0 + 1, ph);
    if (j0 + 2 < c.n_steps) mma_tile_step<D, I, 2, STAGES, PAIR, MC>(c, j0 + 2, ph);
    if (j0 + 3 < c.n_steps) mma_tile_step<D, I, 3, STAGES, PAIR, MC>(c, j0 + 3, ph);
  }
  if (I == 0) ATTN_TRACE(6);
}

// ---- S128: one 128-key S MMA per PAIR of steps ---------------------------------------------------------------------------
// The two S buffers of a tile are adjacent TMEM columns, so S(j) and S(j + 1) can be ONE 128 x 128 x 16 MMA chain over a 128-key K
// stage: it runs at the 32-cycle-per-64-keys floor instead of the 48 cycles of the operand-fetch-bound 128 x 64 x 16 shape (tensor
// work per 128 keys and tile: 1 024 instead of 1 280 cycles, i.e. the structural cap of the tensor pipe goes from 80 % to 100 %).
// The softmax side is unchanged: it still consumes 64-key buffers with per-buffer barriers, and P(j) / P(j + 1) are released and
// multiplied separately.  What changes is the issue order of a tile -- PV(j), PV(j + 1), then S(j + 2, j + 3) -- so a tile's softmax
// warpgroup idles while its own PV + S block runs.  K stages hold 128 keys (two per ring, same 64 KB), V^T stages stay 64 keys.
// MEASURED SLOWER (r02, profiles/r02_attention_s128.md): 1 100 vs 1 285 TFLOP/s at the Wan shape, tensor pipe 49 % active
// instead of 78 %, and insensitive to the softmax speed (POLY 0 / 4 / 8, one or two threads per row), to taking turns between the
// tiles and to an initial stagger: with S single-buffered per pair, every S -> softmax -> P -> PV -> S hand-over (commit, mbarrier
// wake-up, tcgen05.ld / st round trips) sits on the critical path, ~1 500 cycles per 128 keys that the 64-key double-buffered
// schedule hides by keeping S two steps ahead.  512 TMEM columns cannot hold two 128-key S buffers next to O for two tiles, so the
// 64-key schedule stays the default; this variant is kept behind ALG_ATTN_S128=1.
template <int D, int I, int KST>
__device__ __forceinline__ void issue_s128(const MmaCtx& c) {  // S_I(j, j + 1) = Q_I K^T (128-key stage KST) into both S buffers
  constexpr uint32_t idesc_s = make_idesc_bf16(BQ, 2 * BKV);
  constexpr int kBytesQ = BQ * D * 2, kBytesK = 2 * BKV * D * 2, kSub = 2 * BKV * 128;  // K slab: [128 keys][64 elem]
  const uint32_t d = c.tmem + I * 128;
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const uint32_t qoff = (I * kBytesQ + (ks >> 2) * (BQ * 128) + (ks & 3) * 32) >> 4;
    const uint32_t koff = (KST * kBytesK + (ks >> 2) * kSub + (ks & 3) * 32) >> 4;
    mma_ss_lo(d, c.q_lo + qoff, c.k_lo + koff, idesc_s, ks != 0);
  }
  tc_commit_a(c.bar + 8 * (kBarSFull + I * 2 + 0));
  tc_commit_a(c.bar + 8 * (kBarSFull + I * 2 + 1));
}
// pair jp = 2 m + PP of query tile I (steps j = 2 jp, 2 jp + 1); ph = m & 1 (parity of the V ring and, with PP, of the K ring)
template <int D, int I, int PP>
__device__ __forceinline__ void mma_tile_pair(const MmaCtx& c, const int jp, const uint32_t ph) {
  constexpr int VS0 = 2 * PP, VS1 = 2 * PP + 1, KSTN = (PP + 1) & 1;
  const uint32_t p_par = PP;                      // p_full[buf] completes once per pair: parity = jp & 1
  const uint32_t k_phn = PP ? (ph ^ 1u) : ph;     // pair jp + 1 is the ((jp + 1) >> 1)-th use of its K stage
  const int j = 2 * jp;
  mbar_wait_a(c.bar + 8 * (kBarPFull + I * 2 + 0), p_par);
  mbar_wait_a(c.bar + 8 * (kBarVFull + VS0), ph);
  tc_fence_after();
  issue_pv<D, I, 0, VS0, 0>(c, j > 0);
  tc_commit_a(c.bar + 8 * (kBarVEmpty + VS0));
  if (j == c.n_steps - 1) tc_commit_a(c.bar + 8 * (kBarOFull + I));
  if (j + 1 < c.n_steps) {
    mbar_wait_a(c.bar + 8 * (kBarPFull + I * 2 + 1), p_par);
    mbar_wait_a(c.bar + 8 * (kBarVFull + VS1), ph);
  }
  if (j + 2 < c.n_steps) mbar_wait_a(c.bar + 8 * (kBarKFull + KSTN), k_phn);
  // The PV(j + 1) + S(j + 2, j + 3) block is 768 tensor-pipe cycles during which this tile's softmax warps have nothing to do.
  // Two tiles that reach their blocks together interleave them MMA by MMA, finish together and stay in lockstep: every block
  // then takes twice as long (measured: 3 100 cycles per pair instead of ~2 200).  So the tiles take turns: tile 0 issues its
  // n-th block, hands the turn to tile 1, and waits for tile 1's n-th block to be ISSUED before its (n + 1)-th -- after the first
  // round the tiles run half a period apart and one tile's block covers the other tile's softmax.
  if (I == 0) {
    if (jp > 0) mbar_wait_a(c.bar + 8 * (kBarTurn + 0), (uint32_t)(jp - 1) & 1u);
  } else {
    mbar_wait_a(c.bar + 8 * (kBarTurn + 1), (uint32_t)jp & 1u);
  }
  tc_fence_after();
  if (j + 1 < c.n_steps) {
    issue_pv<D, I, 1, VS1, 0>(c, 1u);
    tc_commit_a(c.bar + 8 * (kBarVEmpty + VS1));
    if (j + 1 == c.n_steps - 1) tc_commit_a(c.bar + 8 * (kBarOFull + I));
  }
  if (j + 2 < c.n_steps) {
    issue_s128<D, I, KSTN>(c);  // both S buffers are free: their P was consumed by the two PVs just issued (in-order pipe)
    tc_commit_a(c.bar + 8 * (kBarKEmpty + KSTN));
  }
  mbar_arrive_a(c.bar + 8 * (kBarTurn + (I ^ 1)));
}
template <int D, int I>
__device__ __forceinline__ void mma_tile_loop128(const MmaCtx& c) {
  mbar_wait_a(c.bar, 0);  // q_full
  mbar_wait_a(c.bar + 8 * (kBarKFull + 0), 0);
  if (I == 1 && c.stagger > 0) {  // start tile 1 half a period behind tile 0 (see mma_tile_pair)
    const long long t0 = clock64();
    while (clock64() - t0 < c.stagger) {
    }
  }
  tc_fence_after();
  issue_s128<D, I, 0>(c);
  tc_commit_a(c.bar + 8 * (kBarKEmpty + 0));
  const int n_pairs = (c.n_steps + 1) >> 1;
  uint32_t ph = 0;
#pragma unroll 1
  for (int jp = 0; jp < n_pairs; jp += 2, ph ^= 1u) {
    mma_tile_pair<D, I, 0>(c, jp, ph);
    if (jp + 1 < n_pairs) mma_tile_pair<D, I, 1>(c, jp + 1, ph);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// POLY = 0: every exp2 on MUFU; POLY = n > 0: one pair in every n pairs of a row goes through ex2_poly2.
// SPLIT = 0: one thread per query row (8 softmax warps, 2 per SM sub-partition).
// SPLIT = 1: TWO threads per query row -- warps w and w + 4 of a tile own the same 32 TMEM lanes and take the key
//            columns [0, 32) / [32, 64) of every step (16 softmax warps, 4 per sub-partition, so the schedulers can hide
//            the MUFU / FMA latencies the 2-warp version exposed: 0.44 IPC, profiles/r01_attention.md).  The halves
//            exchange their partial row maxima through shared memory behind one 256-thread named barrier per step,
//            keep partial row sums, and each rescales / writes half of the O columns.
// TILES = 2, STAGES = 4: the long-sequence layout described above (one CTA per SM).
// TILES = 1, STAGES = 2: SHORT variant for a few hundred keys (Wan cross-attention: 257 image / 512 text keys; the Hunyuan
//            token refiner): one query tile, two K/V stages, ~100 KB of shared memory and 256 TMEM columns, so TWO CTAs
//            share an SM and one CTA's prologue (TMEM alloc, Q load) and epilogue overlap the other's steps -- with 5-8
//            steps per CTA those fixed costs were more than half of the long layout's time per CTA.
// PAIR = 1: the kernel runs as 2-CTA clusters (one TPC) and every MMA is a cta_group::2 instruction issued by the leader CTA
//            (cluster rank 0): M = 256 = query tile i of BOTH CTAs; the K tile of a step is split by keys and the V^T tile by
//            head_dim rows between the two CTAs' shared memories and broadcast to both tensor cores, so a CTA fetches and
//            stores HALF of every K / V tile (L2 -> SM traffic and shared-memory reads of the B operands halve; the
//            128 x 64 x 16 S MMA, which at 6 KB of operands per 32 cycles out-runs the 128 B/clk shared-memory port, drops to
//            5 KB).  K/V "full" barriers live in the leader (both CTAs' TMA bytes complete there), "P ready" collects the
//            softmax warps of both CTAs by remote arrives, everything the issuer signals is a multicast commit.
// MC = 1: the kernel runs as clusters of TWO INDEPENDENT CTAs (neighbouring query-tile pairs of one head: same K / V^T).  Each
//            CTA's producer fetches HALF of every K / V^T stage and TMA-multicasts it into both shared memories, so the L2 -> SM
//            traffic of the K / V stream (5 TB/s at the Wan shape: every CTA streams the whole head) halves; MMAs stay
//            cta_group::1, every CTA keeps its own barriers, only the stage releases are multicast commits (see commit_ring).
// FX = 1 (EXPERIMENT, ALG_ATTN_FX=1): fixed softmax reference.  P is a floating-point number, so the reference only has to keep
//            exp2 in range, not near 1: take the row maximum of the FIRST step plus 32 as the reference for the whole row, never
//            look at a maximum again (no FMNMX3 tree, no rescale path), and let P range over 2^-126 .. 2^127: rows whose later
//            scores exceed the first step's maximum by more than ~159 log2 units would overflow (not handled here: experiment).
template <int D, int POLY, int SPLIT, int TILES, int STAGES, int PAIR, int S128 = 0, int MC = 0, int FX = 0>
__global__ void __launch_bounds__(((SPLIT ? 8 : 4) * TILES + 1 + TILES) * 32, TILES == 1 ? 2 : 1)
    attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const Params p) {
  static_assert(!MC || (!PAIR && !S128 && TILES == 2), "MC is a variant of the long layout");
  using C = Cfg<D, PAIR>;
  const uint32_t crank = (PAIR || MC) ? cluster_ctarank() : 0u;
  constexpr int kSoftmaxWarps = (SPLIT ? 8 : 4) * TILES;
  constexpr int kTmaWarp = kSoftmaxWarps, kMmaWarp = kSoftmaxWarps + 1;  // MMA issuers: kMmaWarp (tile 0), kMmaWarp + 1 (tile 1)
  constexpr int kWarpsPerTile = kSoftmaxWarps / TILES;
  constexpr uint32_t kTmemCols = 256 * TILES;  // per tile: two 64-column S buffers + 128 O columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // [2][BQ x D]
  uint8_t* sK = sQ + TILES * C::kBytesQ;       // [STAGES][BKV x D]
  uint8_t* sV = sK + STAGES * C::kBytesK;      // [STAGES][D x BKV]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * C::kBytesV);
  uint64_t* q_full = bars;                     // 1
  uint64_t* k_full = bars + 1;                 // kStages
  uint64_t* k_empty = k_full + kStages;        // kStages
  uint64_t* v_full = k_empty + kStages;        // kStages
  uint64_t* v_empty = v_full + kStages;        // kStages
  uint64_t* s_full = v_empty + kStages;        // [tile][buffer] = 4
  uint64_t* p_full = s_full + 4;               // [tile][buffer] = 4 (a softmax warpgroup may run one step ahead of
                                               // the MMA warp's wait: one barrier per S buffer keeps the parity unambiguous)
  uint64_t* o_done = p_full + 4;               // 2: committed behind every PV (the rare O-rescale path waits on it)
  uint64_t* o_full = o_done + 2;               // 2: committed behind the last PV only (epilogue)
  uint64_t* turn = o_full + 2;                 // 2 (S128 only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // SPLIT only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = blockIdx.x * TILES * BQ;
  const int n_steps = (p.n_kv + BKV - 1) / BKV;
  if (threadIdx.x == 0) ATTN_TRACE(0);

  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], TILES * (MC ? 2 : 1));  // one commit per issuer warp (MC: of both CTAs)
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], TILES * (MC ? 2 : 1));
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kWarpsPerTile * C::kCtas);  // one arrival per softmax warp of the tile (of both CTAs of a pair)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&o_done[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&turn[i], 1);
    }
    fence_barrier_init();
    if constexpr (!PAIR && !S128) {
      // Q and the first two K stages go out BEFORE the CTA-wide setup barrier: the TMEM allocation and the barrier cost ~550
      // cycles that the ~1 900-cycle load latency can overlap (scripts/attn_trace.py; it matters for the SHORT variant, whose
      // CTAs live ~20 000 cycles).  Only this thread has touched the barriers so far, and nothing else uses sQ / sK yet.
      mbar_arrive_expect_tx(q_full, TILES * C::kBytesQ);
      for (int i = 0; i < TILES; ++i)
        for (int s = 0; s < D / 64; ++s)
          tma_load_3d(sQ + i * C::kBytesQ + s * C::kSubQ, &tmQ, q_full, head * D + s * 64, q0 + i * BQ, batch);
      for (int j = 0; j < 2 && j < n_steps && !MC; ++j) {  // (MC: the peer's barriers must exist first)
        mbar_arrive_expect_tx(&k_full[j], C::kBytesK);
        for (int s = 0; s < D / 64; ++s)
          tma_load_3d(sK + j * C::kBytesK + s * C::kSubK, &tmK, &k_full[j], head * D + s * 64, j * BKV, batch);
      }
    }
  }
  if (warp == kMmaWarp) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR || MC) cluster_sync_all();  // the peer's barriers exist before anything completes / arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    ATTN_TRACE(1);
#ifdef ALG_ATTN_TRACE
    uint32_t smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    if (g_attn_trace) g_attn_trace[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32 + 2] = smid;
#endif
  }

  if (warp == kTmaWarp) {
    if (elect_one()) {  // ===== TMA producer: Q0 Q1 | K0 K1 | V0 K2 | V1 K3 | ... (the order the MMA warp consumes) =====
      // PAIR: the bytes of both CTAs complete on the LEADER's barriers (one expect_tx there covers the pair); each CTA
      // loads its own query tiles, its 32 keys of the K tile and its half of the V^T rows
      auto load = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
        if constexpr (PAIR) tma_load_3d_pair(dst, m, bar, c0, c1, c2);
        else tma_load_3d(dst, m, bar, c0, c1, c2);
      };
      if constexpr (PAIR || S128) {  // (otherwise issued before the setup barrier, above)
        if (crank == 0) mbar_arrive_expect_tx(q_full, C::kCtas * TILES * C::kBytesQ);
        for (int i = 0; i < TILES; ++i)
          for (int s = 0; s < D / 64; ++s)
            load(sQ + i * C::kBytesQ + s * C::kSubQ, &tmQ, q_full, head * D + s * 64, q0 + i * BQ, batch);
      }
      auto load_mc = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], "
            "[%2], %6;" ::"r"(smem_u32(dst)),
            "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"((uint16_t)3)
            : "memory");
      };
      auto load_k = [&](int j) {
        const int st = j % STAGES;
        mbar_wait(&k_empty[st], ((j / STAGES) & 1) ^ 1);
        if constexpr (MC) {  // my 32 keys of the tile, into both CTAs; each CTA's barrier expects the whole stage
          mbar_arrive_expect_tx(&k_full[st], C::kBytesK);
          for (int s = 0; s < D / 64; ++s)
            load_mc(sK + st * C::kBytesK + s * C::kSubK + (int)crank * (BKV / 2) * 128, &tmK, &k_full[st], head * D + s * 64,
                    j * BKV + (int)crank * (BKV / 2), batch);
          return;
        }
        if (crank == 0) mbar_arrive_expect_tx(&k_full[st], C::kCtas * C::kBytesK);
        for (int s = 0; s < D / 64; ++s)
          load(sK + st * C::kBytesK + s * C::kSubK, &tmK, &k_full[st], head * D + s * 64,
               j * BKV + (int)crank * (BKV / C::kCtas), batch);
      };
      auto load_v = [&](int j) {
        const int st = j % STAGES;
        mbar_wait(&v_empty[st], ((j / STAGES) & 1) ^ 1);
        if constexpr (MC) {  // my half of the head_dim rows of the V^T tile
          mbar_arrive_expect_tx(&v_full[st], C::kBytesV);
          load_mc(sV + st * C::kBytesV + (int)crank * (D / 2) * 128, &tmV, &v_full[st], j * BKV, head * D + (int)crank * (D / 2), batch);
          return;
        }
        if (crank == 0) mbar_arrive_expect_tx(&v_full[st], C::kCtas * C::kBytesV);
        load(sV + st * C::kBytesV, &tmV, &v_full[st], j * BKV, head * D + (int)crank * (D / C::kCtas), batch);
      };
      if constexpr (S128) {  // K stages of 128 keys (two per ring); consumption order: K(0) | V(0) V(1) K(1) | V(2) V(3) K(2) | ...
        static_assert(!S128 || (TILES == 2 && STAGES == 4 && !PAIR), "S128 is a variant of the long layout");
        const int n_pairs = (n_steps + 1) >> 1;
        auto load_k128 = [&](int jp) {
          const int st = jp & 1;
          mbar_wait(&k_empty[st], ((jp >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], 2 * C::kBytesK);
          for (int s = 0; s < D / 64; ++s)
            tma_load_3d(sK + st * 2 * C::kBytesK + s * (2 * BKV * 128), &tmK, &k_full[st], head * D + s * 64, jp * 2 * BKV, batch);
        };
        load_k128(0);
        for (int jp = 0; jp < n_pairs; ++jp) {
          load_v(2 * jp);
          if (2 * jp + 1 < n_steps) load_v(2 * jp + 1);
          if (jp + 1 < n_pairs) load_k128(jp + 1);
        }
      } else {
        if constexpr (PAIR || MC) {
          load_k(0);
          if (n_steps > 1) load_k(1);
        }
        for (int j = 0; j < n_steps; ++j) {
          load_v(j);
          if (j + 2 < n_steps) load_k(j + 2);
        }
      }
    }
  } else if (warp >= kMmaWarp) {
    if ((!PAIR || crank == 0) && elect_one()) {  // ===== MMA issuers (of the leader CTA when PAIR).  elect.sync (not `lane == 0`) tells ptxas that a single lane runs this region, so
                        // the descriptors stay in uniform registers; otherwise every UTCHMMA is wrapped in an ELECT / R2UR loop =====
      MmaCtx c;
      c.tmem = tmem_base;
      c.tmem_o = tmem_base + TILES * 128;
      c.bar = smem_u32(bars);
      c.q_lo = smem_desc_lo_sw128(smem_u32(sQ));
      c.k_lo = smem_desc_lo_sw128(smem_u32(sK));
      c.v_lo = smem_desc_lo_sw128(smem_u32(sV));
      c.n_steps = n_steps;
      c.stagger = p.stagger;
      if constexpr (S128) {
        if (warp == kMmaWarp) mma_tile_loop128<D, 0>(c);
        else mma_tile_loop128<D, 1>(c);
      } else {
        if (warp == kMmaWarp) mma_tile_loop<D, 0, STAGES, PAIR, MC>(c);
        else if constexpr (TILES == 2) mma_tile_loop<D, 1, STAGES, PAIR, MC>(c);
      }
    }
  } else {  // ===== softmax warps =====
    constexpr int W = SPLIT ? BKV / 2 : BKV;      // key columns of a step owned by this thread
    constexpr int OC = SPLIT ? D / 2 : D;         // O columns this thread rescales / writes
    const int i = warp / kWarpsPerTile;           // query tile
    const int hh = SPLIT ? (warp >> 2) & 1 : 0;   // column half (SPLIT)
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + i * 128;
    const uint32_t t_o = tmem_base + lane_base + TILES * 128 + i * 128 + hh * OC;
    const int row_in_tile = quad * 32 + lane;
    const int row = q0 + i * BQ + row_in_tile;
    float m_used = -INFINITY, l = 0.f;
    const float c = p.scale_log2;
    if (p.accumulate && row < p.n_q) {
      // the epilogue reads this tile's previous O rows (written by the launch before: HBM by now): pull them into L2 while the
      // warpgroup waits for Q and S(0) anyway -- no registers held, and the epilogue's 16 loads per lane then hit L2
      const char* prow = reinterpret_cast<const char*>(p.O + (int64_t)batch * p.o_bs + (int64_t)row * p.o_rs + head * D);
#pragma unroll
      for (int b = 0; b < D * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + b));
    }
    // one 64-key step of online softmax; `ragged` (compile-time) is the last, partially filled step -- kept out of the
    // main loop body, where the compiler would otherwise if-convert the mask into always-executed selects
    // barrier addresses of this tile's two S buffers; the loop is unrolled by two so the buffer index is compile-time
    const uint32_t bar_s0 = smem_u32(&s_full[i * 2]), bar_p0 = smem_u32(&p_full[i * 2]);
    auto softmax_step = [&](const int j, auto ragged, auto buf_c) {
      constexpr int BUF = decltype(buf_c)::value;
      const uint32_t t_sj = t_s + BUF * 64;
      mbar_wait_a(bar_s0 + 8 * BUF, (j >> 1) & 1);
      tc_fence_after();
      if (j == 0 && warp == 0 && lane == 0) ATTN_TRACE(7);
      float s[W];
      tmem_ld32(t_sj + hh * W, reinterpret_cast<uint32_t*>(s));
      if constexpr (W == 64) tmem_ld32(t_sj + 32, reinterpret_cast<uint32_t*>(s) + 32);
      tmem_ld_wait();
      if constexpr (decltype(ragged)::value) {  // TMA zero-filled the tail of the tile: mask it out
        const int valid = p.n_kv - j * BKV - hh * W;
#pragma unroll
        for (int k = 0; k < W; ++k)
          if (k >= valid) s[k] = -INFINITY;
      }
      if (!FX || j == 0) {
      // row max: independent FMNMX3 chains (one dependent chain of max ops is pure latency)
      float mc[W / 16];
#pragma unroll
      for (int g = 0; g < W / 16; ++g) {
        float m = s[g * 16];
#pragma unroll
        for (int k = 1; k + 1 < 16; k += 2) m = max3(m, s[g * 16 + k], s[g * 16 + k + 1]);
        mc[g] = fmaxf(m, s[g * 16 + 15]);
      }
      float mraw;
      if constexpr (SPLIT) {
        // exchange the partial maxima of the two column halves: this also orders "both halves have read S(j)" before
        // either half overwrites the buffer with P(j)
        float* slot = xch + ((i * 2 + BUF) * 2) * BQ;
        const float mine = fmaxf(mc[0], mc[1]);
        slot[hh * BQ + row_in_tile] = mine;
        named_bar_sync(1 + i, kWarpsPerTile * 32);
        mraw = fmaxf(mine, slot[(hh ^ 1) * BQ + row_in_tile]);
      } else {
        mraw = fmaxf(max3(mc[0], mc[1], mc[2]), mc[W / 16 - 1]);
      }
      const float mx = mraw * c;
      if (j == 0) {
        m_used = FX ? mx + 32.f : mx;
      } else {
        const float m_new = fmaxf(m_used, mx);
        const bool need = (m_new - m_used) > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {  // warp-uniform (and identical in both halves: same rows, same maxima)
          // S runs two steps ahead of PV: O_i may still be receiving P_i(j-1) V_(j-1)
          mbar_wait(&o_done[i], (j - 1) & 1);
          tc_fence_after();
          const float f = ex2(m_used - m_new);
          l *= f;
          m_used = m_new;
#pragma unroll 1
          for (int ch = 0; ch < OC / 16; ++ch) {
            uint32_t o[16];
            tmem_ld16(t_o + ch * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
            tmem_st16(t_o + ch * 16, o);
          }
        }
      }
      }
      float l_step;
      for (;;) {
        const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_used, -m_used);
        float2 sum0 = make_float2(0.f, 0.f), sum1 = sum0;
#pragma unroll
        for (int ch = 0; ch < W / 32; ++ch) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float2 x = __ffma2_rn(make_float2(s[ch * 32 + 2 * k], s[ch * 32 + 2 * k + 1]), c2, nm2);
            float2 e;
            if (POLY > 0 && (k % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
              // FX: the exponent patch of the polynomial wraps past 2^128 instead of saturating; 2^128 patches to 3.4e38
              e = ex2_poly2(FX ? make_float2(fminf(x.x, 128.f), fminf(x.y, 128.f)) : x);
            } else {
              e.x = ex2(x.x);
              e.y = ex2(x.y);
            }
            if (k & 1) sum1 = __fadd2_rn(sum1, e);
            else sum0 = __fadd2_rn(sum0, e);
            __nv_bfloat162 h = __floats2bfloat162_rn(e.x, e.y);
            pk[k] = *reinterpret_cast<uint32_t*>(&h);
          }
          // P (bf16 pairs) overwrites the first 32 columns of this S buffer: 16 columns per 32 keys
          tmem_st16(t_sj + hh * 16 + ch * 16, pk);
        }
        l_step = (sum0.x + sum0.y) + (sum1.x + sum1.y);
        if (!FX || !__any_sync(0xffffffffu, l_step > 7.9228163e28f)) break;
        // Range guard of the fixed reference (cold): a step whose probabilities sum past 2^96 (or overflowed to inf) raises the
        // reference by exactly 96 log2 units -- O and l scale by the exact power of two -- and the step's P is computed again
        // from the scores still in registers, by the same code; repeated until the step fits.
        constexpr float kDown = 1.2621774e-29f;  // 2^-96
        if (j > 0) {
          mbar_wait(&o_done[i], (j - 1) & 1);  // O_i may still be receiving P_i(j-1) V_(j-1)
          tc_fence_after();
#pragma unroll 1
          for (int ch = 0; ch < OC / 16; ++ch) {
            uint32_t o[16];
            tmem_ld16(t_o + ch * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * kDown);
            tmem_st16(t_o + ch * 16, o);
          }
        }
        l *= kDown;
        m_used += 96.f;
      }
      l += l_step;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_remote_a(bar_p0 + 8 * BUF, 0);  // the leader's issuer waits for both CTAs' P
        else mbar_arrive_a(bar_p0 + 8 * BUF);
        if (j == 0 && warp == 0) ATTN_TRACE(8);
      }
    };
    const int n_full = p.n_kv / BKV;
    using B0 = std::integral_constant<int, 0>;
    using B1 = std::integral_constant<int, 1>;
    int j = 0;
#pragma unroll 1
    for (; j + 1 < n_full; j += 2) {
      softmax_step(j, std::false_type{}, B0{});
      softmax_step(j + 1, std::false_type{}, B1{});
    }
    if (j < n_full) {  // odd number of full steps: one more on buffer 0, a ragged tail (if any) lands on buffer 1
      softmax_step(j, std::false_type{}, B0{});
      if (n_full < n_steps) softmax_step(n_full, std::true_type{}, B1{});
    } else if (n_full < n_steps) {
      softmax_step(n_full, std::true_type{}, B0{});
    }
    if constexpr (SPLIT) {  // row sum = sum of the two halves' partial sums (slot of the parity the last step did not use)
      float* slot = xch + ((i * 2 + (n_steps & 1)) * 2) * BQ;
      slot[hh * BQ + row_in_tile] = l;
      named_bar_sync(1 + i, kWarpsPerTile * 32);
      l += slot[(hh ^ 1) * BQ + row_in_tile];
    }
    // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------------
    if (warp == 0 && lane == 0) ATTN_TRACE(9);
    // coalesced write-back geometry (see below): lane -> (row rl of RPI, 16-byte chunk cl) of each store instruction
    constexpr int RB = D * 2, CPR = RB / 16, RPI = 32 / CPR;  // row bytes, chunks per row, rows per warp instruction
    constexpr int NIT = 32 / RPI;                              // store instructions per warp (its 32 rows)
    const int rl = lane / CPR, cl = lane % CPR;
    const int r0 = quad * 32 + rl;                              // this lane's row in iteration 0; + RPI per iteration
    const int n_live = p.n_q - (q0 + i * BQ + r0);              // rows r0 + it * RPI with it * RPI < n_live exist
    uint4* dst0 = reinterpret_cast<uint4*>(p.O + (int64_t)batch * p.o_bs + (int64_t)(q0 + i * BQ + r0) * p.o_rs + head * D + cl * 8);
    const int64_t dstep = (int64_t)RPI * p.o_rs / 8;            // uint4 units (o_rs % 8 == 0)
    [[maybe_unused]] uint4 prev[NIT];
    if constexpr (!SPLIT) {
      if (p.accumulate) {  // the previous O rows: all loads in flight while the last PV drains and the rows are staged
#pragma unroll
        for (int it = 0; it < NIT; ++it)
          if (it * RPI < n_live) prev[it] = __ldcs(dst0 + it * dstep);
      }
    }
    mbar_wait(&o_full[i], 0);
    tc_fence_after();
    if (warp == 0 && lane == 0) ATTN_TRACE(10);
    const float inv = 1.0f / l;
    if constexpr (!SPLIT) {
      // Staged epilogue.  A thread owns one O row, so storing straight from its registers makes every 16-byte store of a warp hit
      // 32 different rows: 2 048 L1 wavefronts per tile instead of 256, measured at 4 400 cycles per CTA (11 400 with the
      // read-modify-write of `accumulate`) -- a quarter to 40 % of a SHORT CTA's life (scripts/attn_trace.py).  The tile's Q
      // region is dead by now (every S MMA completed before the commit behind the last PV), so each warp parks its 32 rows
      // there as bf16 (16-byte chunks XOR-swizzled by the row, conflict-free both ways) and writes them back out with
      // 32 / CPR whole rows per instruction.
      const uint32_t stage = smem_u32(sQ + i * C::kBytesQ);       // explicit shared-space accesses (a generic pointer costs
                                                                 // ~70 instructions per 16 bytes here and serialises the loop)
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t o[32];
        tmem_ld32(t_o + ch * 32, o);
        tmem_ld_wait();
        if (ch == 0 && warp == 0 && lane == 0) ATTN_TRACE(14);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            h[e] = __floats2bfloat162_rn(__uint_as_float(o[g * 8 + 2 * e]) * inv, __uint_as_float(o[g * 8 + 2 * e + 1]) * inv);
          const int c16 = ch * 4 + g;
          sts_v4(stage + row_in_tile * RB + ((c16 ^ (row_in_tile & (CPR - 1))) << 4), u);
        }
      }
      __syncwarp();  // a warp re-reads only the 32 rows it wrote
      if (warp == 0 && lane == 0) ATTN_TRACE(13);
      uint4 u[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int r_t = r0 + it * RPI;
        u[it] = lds_v4(stage + r_t * RB + ((cl ^ (r_t & (CPR - 1))) << 4));
      }
      if (p.accumulate) {  // hidden_states(text) + hidden_states_img: both already bf16 tensors
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev[it]);
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u[it]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = __bfloat1622float2(h[e]), b = __bfloat1622float2(ph[e]);
            h[e] = __floats2bfloat162_rn(a.x + b.x, a.y + b.y);
          }
        }
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it)
        if (it * RPI < n_live) dst0[it * dstep] = u[it];
      if (warp == 0 && lane == 0) ATTN_TRACE(11);
    } else {
    __nv_bfloat16* orow = p.O + (int64_t)batch * p.o_bs + (int64_t)row * p.o_rs + head * D + hh * OC;
#pragma unroll
    for (int ch = 0; ch < OC / 32; ++ch) {
      uint32_t o[32];
      tmem_ld32(t_o + ch * 32, o);
      tmem_ld_wait();
      if (row < p.n_q) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
          uint4 prev;
          if (p.accumulate) prev = *reinterpret_cast<const uint4*>(orow + ch * 32 + g * 8);
          const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = __uint_as_float(o[g * 8 + 2 * e]) * inv;
            float y = __uint_as_float(o[g * 8 + 2 * e + 1]) * inv;
            if (p.accumulate) {  // hidden_states(text) + hidden_states_img: both already bf16 tensors
              float2 q = __bfloat1622float2(ph[e]);
              x = bf16_round(x) + q.x;
              y = bf16_round(y) + q.y;
            }
            h[e] = __floats2bfloat162_rn(x, y);
          }
          *reinterpret_cast<uint4*>(orow + ch * 32 + g * 8) = u;
        }
      }
    }
    if (warp == 0 && lane == 0) ATTN_TRACE(11);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ATTN_TRACE(12);
  if constexpr (PAIR || MC) cluster_sync_all();  // no CTA leaves while the pair's MMAs / commits / remote arrives may touch it
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int D, int POLY, int SPLIT, int TILES, int STAGES, int PAIR = 0, int S128 = 0, int MC = 0, int FX = 0>
static int launch(const alg_attention_t* a, cudaStream_t st) {
  using C = Cfg<D, PAIR>;
  constexpr int kThreads = ((SPLIT ? 8 : 4) * TILES + 1 + TILES) * 32;
  constexpr int kSmem = C::smem_bytes(TILES, STAGES);
  static std::atomic<uint64_t> attr_done{0};  // per-device bit: the attribute is device state
  int dev = 0;
  ALG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 64 || !(attr_done.load(std::memory_order_relaxed) >> dev & 1)) {
    ALG_CUDA_OK(cudaFuncSetAttribute(attention_kernel<D, POLY, SPLIT, TILES, STAGES, PAIR, S128, MC, FX>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    if (dev < 64) attr_done.fetch_or(uint64_t(1) << dev, std::memory_order_relaxed);
  }
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t hd = (uint64_t)a->heads * D;
  {
    uint64_t dims[3] = {hd, (uint64_t)a->n_q, (uint64_t)a->batch}, strides[3] = {1, (uint64_t)a->q_rs, (uint64_t)a->q_bs};
    uint32_t box[3] = {64, BQ, 1};
    if (int rc = make_tmap_bf16(&tmQ, a->Q, 3, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[3] = {hd, (uint64_t)a->n_kv, (uint64_t)a->batch}, strides[3] = {1, (uint64_t)a->k_rs, (uint64_t)a->k_bs};
    uint32_t box[3] = {64, S128 ? 2 * BKV : BKV / (MC ? 2 : C::kCtas), 1};
    if (int rc = make_tmap_bf16(&tmK, a->K, 3, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->n_kv, hd, (uint64_t)a->batch}, strides[3] = {1, (uint64_t)a->v_rs, (uint64_t)a->v_bs};
    uint32_t box[3] = {BKV, (uint32_t)D / (MC ? 2 : C::kCtas), 1};
    if (int rc = make_tmap_bf16(&tmV, a->Vt, 3, dims, strides, box)) return rc;
  }
  Params p;
  p.O = reinterpret_cast<__nv_bfloat16*>(a->O);
  p.o_bs = a->o_bs;
  p.o_rs = a->o_rs;
  p.n_q = (int)a->n_q;
  p.n_kv = (int)a->n_kv;
  p.heads = a->heads;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  {
    static int stagger = -1;
    if (stagger < 0) {
      const char* e = getenv("ALG_ATTN_STAGGER");
      stagger = e ? atoi(e) : 0;
    }
    p.stagger = stagger;
  }
  dim3 grid((unsigned)((a->n_q + TILES * BQ - 1) / (TILES * BQ)), (unsigned)a->heads, (unsigned)a->batch);
  if constexpr (PAIR || MC) {
    grid.x = (grid.x + 1) / 2 * 2;  // whole pairs: a CTA past the last query rows loads zero-filled tiles and stores nothing
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ALG_CUDA_OK(cudaLaunchKernelEx(&cfg, attention_kernel<D, POLY, SPLIT, TILES, STAGES, PAIR, S128, MC, FX>, tmQ, tmK, tmV, p));
  } else {
    attention_kernel<D, POLY, SPLIT, TILES, STAGES, PAIR, S128, MC, FX><<<grid, kThreads, kSmem, st>>>(tmQ, tmK, tmV, p);
  }
  ALG_LAUNCH_OK();
  return 0;
}


// ====================================================================================================================
// PS variant (head_dim 128, long sequences): P travels through SHARED memory, S is a full-rate 128-key MMA, and S stays one
// step AHEAD of the softmax.
//
// Why: the default kernel's 128 x 64 x 16 S MMAs are operand-fetch bound (80 % structural cap, and it sits there); the S128
// variant above runs them at full rate but exposes the S -> softmax -> P -> PV -> S hand-over chain, because with P aliased
// onto S in TMEM an S buffer is only free once PV has CONSUMED P.  Here the softmax copies S(j) (128 keys, 128 fp32 per
// row) into registers and releases the buffer at once -- S(j+1) is issued while the softmax of step j is still running -- and P is
// written as bf16 into a small shared-memory operand (SW128 K-major rows, two 32-key slots per tile) from which PV runs as an SS
// MMA.  LSU stores do not slow the tensor core's operand fetch (scripts/microbench/mma_smem_contention.cu: 64 -> 65 cycles per
// 128 x 128 x 16 MMA under 60 B/clk of st.shared), and every MMA of the kernel is now a full-rate N = 128 shape.
//
//   TMEM: S0 [0,128)  S1 [128,256)  O0 [256,384)  O1 [384,512)          (one S buffer per tile)
//   smem: Q0 Q1 (64 KB) | K ring 2 x 128 keys (64 KB) | V^T ring 4 x 64 keys (64 KB) | P0 P1 (2 x 16 KB: [128 rows][2 slots x 32 keys])
//   per step j and tile:  issuer: wait s_free(j) -> S(j+1);  for quarter q = 0..3: wait p_ready[q] -> PV(j, q) (2 MMAs) -> commit p_free[q]
//                         softmax: wait s_full(j) -> ld S -> arrive s_free -> row max (lazy O rescale) -> for q: 32 exps -> wait
//                                  p_free[q - 2] (slot reuse) -> st.shared (swizzled) -> fence.proxy.async -> arrive p_ready[q]
// ====================================================================================================================
namespace ps {
constexpr int BK = 128, QK = 32;  // keys per step, keys per P quarter
constexpr int kBytesQ = BQ * 128 * 2, kBytesK = BK * 128 * 2, kBytesV = 128 * 64 * 2, kBytesP = BQ * 64 * 2;
constexpr int kSmem = 2 * kBytesQ + 2 * kBytesK + 4 * kBytesV + 2 * kBytesP + 512 + 2048;  // + barriers + SPLIT's exchange
constexpr int bQFull = 0, bKFull = 1, bKEmpty = 3, bVFull = 5, bVEmpty = 9, bSFull = 13, bSFree = 15, bPReady = 17, bPFree = 25,
              bOFull = 33, bCount = 35;

struct Ctx {
  uint32_t tmem, bar, q_lo, k_lo, v_lo, p_lo;
  int n_steps;
};

template <int I, int KST>
__device__ __forceinline__ void issue_s(const Ctx& c) {  // S_I = Q_I K^T over the 128-key stage KST
  constexpr uint32_t idesc = make_idesc_bf16(BQ, BK);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t qoff = (I * kBytesQ + (ks >> 2) * (BQ * 128) + (ks & 3) * 32) >> 4;
    const uint32_t koff = (KST * kBytesK + (ks >> 2) * (BK * 128) + (ks & 3) * 32) >> 4;
    mma_ss_lo(c.tmem + I * 128, c.q_lo + qoff, c.k_lo + koff, idesc, ks != 0);
  }
  tc_commit_a(c.bar + 8 * (bSFull + I));
  tc_commit_a(c.bar + 8 * (bKEmpty + KST));
}
// PV quarter Q of a step whose V^T stages are VS0 (keys 0-63) and VS0 + 1 (keys 64-127)
template <int I, int Q, int VS0>
__device__ __forceinline__ void issue_pv(const Ctx& c, uint32_t acc_first) {
  constexpr uint32_t idesc = make_idesc_bf16(BQ, 128);
  constexpr int VS = VS0 + (Q >> 1), SLOT = Q & 1;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint32_t poff = (I * kBytesP + (SLOT * 2 + ks) * 32) >> 4;
    const uint32_t voff = (VS * kBytesV + (SLOT * 2 + ks) * 32) >> 4;
    mma_ss_lo(c.tmem + 256 + I * 128, c.p_lo + poff, c.v_lo + voff, idesc, ks == 0 ? acc_first : 1u);
  }
  tc_commit_a(c.bar + 8 * (bPFree + I * 4 + Q));
  if (Q & 1) tc_commit_a(c.bar + 8 * (bVEmpty + VS));
}
// step j = 2 m + JJ of tile I; ph = m & 1
template <int I, int JJ>
__device__ __forceinline__ void issuer_step(const Ctx& c, const int j, const uint32_t ph) {
  constexpr int KSTN = (JJ + 1) & 1, VS0 = 2 * JJ;
  const uint32_t par = JJ;                          // barriers that complete once per step: parity j & 1
  const uint32_t k_phn = JJ ? (ph ^ 1u) : ph;       // step j + 1 is the ((j + 1) >> 1)-th use of its K stage
  if (j + 1 < c.n_steps) {  // S runs one step ahead: as soon as the softmax has copied S(j) out of TMEM
    mbar_wait_a(c.bar + 8 * (bSFree + I), par);
    mbar_wait_a(c.bar + 8 * (bKFull + KSTN), k_phn);
    tc_fence_after();
    issue_s<I, KSTN>(c);
  }
  const bool last = j == c.n_steps - 1;
  mbar_wait_a(c.bar + 8 * (bPReady + I * 4 + 0), par);
  mbar_wait_a(c.bar + 8 * (bVFull + VS0), ph);
  tc_fence_after();
  issue_pv<I, 0, VS0>(c, j > 0);
  mbar_wait_a(c.bar + 8 * (bPReady + I * 4 + 1), par);
  tc_fence_after();
  issue_pv<I, 1, VS0>(c, 1u);
  mbar_wait_a(c.bar + 8 * (bPReady + I * 4 + 2), par);
  mbar_wait_a(c.bar + 8 * (bVFull + VS0 + 1), ph);
  tc_fence_after();
  issue_pv<I, 2, VS0>(c, 1u);
  mbar_wait_a(c.bar + 8 * (bPReady + I * 4 + 3), par);
  tc_fence_after();
  issue_pv<I, 3, VS0>(c, 1u);
  if (last) tc_commit_a(c.bar + 8 * (bOFull + I));
}
template <int I>
__device__ __forceinline__ void issuer_loop(const Ctx& c) {
  mbar_wait_a(c.bar + 8 * bQFull, 0);
  mbar_wait_a(c.bar + 8 * (bKFull + 0), 0);
  tc_fence_after();
  issue_s<I, 0>(c);
  uint32_t ph = 0;
#pragma unroll 1
  for (int j = 0; j < c.n_steps; j += 2, ph ^= 1u) {
    issuer_step<I, 0>(c, j, ph);
    if (j + 1 < c.n_steps) issuer_step<I, 1>(c, j + 1, ph);
  }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float bf16_ceil(float x) {  // smallest bf16 >= x (x finite or -inf)
  const float r = __bfloat162float(__float2bfloat16_rd(x)), u = __bfloat162float(__float2bfloat16_ru(x));
  return x > r ? u : r;
}

// SPLIT = 1: TWO threads per query row (16 softmax warps, 4 per SM sub-partition): warps w and w + 4 of a tile own the same 32
// TMEM lanes and take the keys [0, 64) / [64, 128) of every step, i.e. the P quarters {0, 1} / {2, 3}.  The softmax of this kernel is
// bound by the instruction issue of its warps (ncu: 1 460 warp instructions per sub-partition and step at 0.40 IPC with two warps
// per sub-partition), so doubling the warps doubles the latency hiding.  The halves agree on the row maximum by exchanging their
// partial maxima ROUNDED UP to bf16 through 2 KB of shared memory (any common upper bound of the row works as the softmax shift).
template <int POLY, int SPLIT>
__global__ void __launch_bounds__((SPLIT ? 19 : 11) * 32, 1)
    attention_ps_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const Params p) {
  constexpr int kSoftmaxWarps = SPLIT ? 16 : 8, kTmaWarp = kSoftmaxWarps, kMmaWarp = kSoftmaxWarps + 1;
  constexpr int kWarpsPerTile = kSoftmaxWarps / 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();  // the SW128 tiles need 1 KB alignment; dynamic shared memory starts 1 KB-aligned
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * kBytesQ;
  uint8_t* sV = sK + 2 * kBytesK;
  uint8_t* sP = sV + 4 * kBytesV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kBytesP);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(bars) + 512);  // SPLIT: [tile][parity][half][row]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = blockIdx.x * 2 * BQ;
  const int n_steps = (p.n_kv + BK - 1) / BK;

  if (warp == kTmaWarp && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(&bars[bQFull], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[bKFull + i], 1);
      mbar_init(&bars[bKEmpty + i], 2);  // one commit per issuer
      mbar_init(&bars[bSFull + i], 1);
      mbar_init(&bars[bSFree + i], kWarpsPerTile);  // every softmax warp of the tile has copied S out of TMEM
      mbar_init(&bars[bOFull + i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars[bVFull + i], 1);
      mbar_init(&bars[bVEmpty + i], 2);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bars[bPReady + i], 4);  // a quarter is written by four warps in both layouts
      mbar_init(&bars[bPFree + i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    if (elect_one()) {  // ===== TMA producer: Q0 Q1 | K(0) K(1) | V(0) V(1) K(2) | V(2) V(3) K(3) | ... =====
      mbar_arrive_expect_tx(&bars[bQFull], 2 * kBytesQ);
      for (int i = 0; i < 2; ++i)
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sQ + i * kBytesQ + s * (BQ * 128), &tmQ, &bars[bQFull], head * 128 + s * 64, q0 + i * BQ, batch);
      auto load_k = [&](int j) {
        const int st = j & 1;
        mbar_wait(&bars[bKEmpty + st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[bKFull + st], kBytesK);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sK + st * kBytesK + s * (BK * 128), &tmK, &bars[bKFull + st], head * 128 + s * 64, j * BK, batch);
      };
      auto load_v = [&](int h) {  // h = index of the 64-key half
        const int st = h & 3;
        mbar_wait(&bars[bVEmpty + st], ((h >> 2) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[bVFull + st], kBytesV);
        tma_load_3d(sV + st * kBytesV, &tmV, &bars[bVFull + st], h * 64, head * 128, batch);
      };
      load_k(0);
      if (n_steps > 1) load_k(1);
      for (int j = 0; j < n_steps; ++j) {
        load_v(2 * j);
        load_v(2 * j + 1);
        if (j + 2 < n_steps) load_k(j + 2);
      }
    }
  } else if (warp >= kMmaWarp) {
    if (elect_one()) {  // ===== MMA issuers, one per query tile =====
      Ctx c;
      c.tmem = tmem_base;
      c.bar = smem_u32(bars);
      c.q_lo = smem_desc_lo_sw128(smem_u32(sQ));
      c.k_lo = smem_desc_lo_sw128(smem_u32(sK));
      c.v_lo = smem_desc_lo_sw128(smem_u32(sV));
      c.p_lo = smem_desc_lo_sw128(smem_u32(sP));
      c.n_steps = n_steps;
      if (warp == kMmaWarp) issuer_loop<0>(c);
      else issuer_loop<1>(c);
    }
  } else {  // ===== softmax warps =====
    constexpr int W = SPLIT ? 64 : 128;             // keys of a step owned by this thread
    constexpr int NQ = W / QK;                      // P quarters it produces
    constexpr int OC = SPLIT ? 64 : 128;            // O columns it rescales / writes
    const int i = warp / kWarpsPerTile, quad = warp & 3;
    const int hh = SPLIT ? (warp >> 2) & 1 : 0;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + i * 128 + hh * W;
    const uint32_t t_o = tmem_base + lane_base + 256 + i * 128 + hh * OC;
    const int row_in_tile = quad * 32 + lane;
    const int row = q0 + i * BQ + row_in_tile;
    const uint32_t p_row = smem_u32(sP) + i * kBytesP + row_in_tile * 128;  // this row of the tile's P operand (2 slots x 32 keys)
    const int sw = row_in_tile & 7;                                          // SWIZZLE_128B: chunk c of row r sits at c ^ (r & 7)
    const uint32_t bar = smem_u32(bars);
    float m_used = -INFINITY, l = 0.f;
    const float c = p.scale_log2;
#pragma unroll 1
    for (int j = 0; j < n_steps; ++j) {
      const uint32_t par = (uint32_t)j & 1u;
      mbar_wait_a(bar + 8 * (bSFull + i), par);
      tc_fence_after();
      float s[W];
#pragma unroll
      for (int g = 0; g < W / 32; ++g) tmem_ld32(t_s + g * 32, reinterpret_cast<uint32_t*>(s) + g * 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar + 8 * (bSFree + i));  // S(j) is in registers: the issuer may overwrite the buffer with S(j+1)
      if (j == n_steps - 1) {  // TMA zero-filled the tail of the last K tile: mask it out
        const int valid = p.n_kv - j * BK - hh * W;
#pragma unroll
        for (int k = 0; k < W; ++k)
          if (k >= valid) s[k] = -INFINITY;
      }
      if (p.stagger & 1) {  // DEBUG (ALG_ATTN_STAGGER=1): skeleton only -- no softmax arithmetic, P slots published untouched
#pragma unroll
        for (int qq = 0; qq < NQ; ++qq) {
          const int q = hh * NQ + qq;
          if (q >= 2) mbar_wait_a(bar + 8 * (bPFree + i * 4 + q - 2), par);
          else if (j > 0) mbar_wait_a(bar + 8 * (bPFree + i * 4 + q + 2), par ^ 1u);
          l += s[qq];
          __syncwarp();
          if (lane == 0) mbar_arrive_a(bar + 8 * (bPReady + i * 4 + q));
        }
        continue;
      }
      float mc[W / 16];
#pragma unroll
      for (int g = 0; g < W / 16; ++g) {
        float m = s[g * 16];
#pragma unroll
        for (int k = 1; k + 1 < 16; k += 2) m = max3(m, s[g * 16 + k], s[g * 16 + k + 1]);
        mc[g] = fmaxf(m, s[g * 16 + 15]);
      }
      float mraw = fmaxf(max3(mc[0], mc[1], mc[2]), mc[3]);
      if constexpr (!SPLIT) mraw = fmaxf(mraw, fmaxf(max3(mc[W / 16 - 4], mc[W / 16 - 3], mc[W / 16 - 2]), mc[W / 16 - 1]));
      if constexpr (SPLIT) {
        // the two halves of a row agree on a common upper bound: each partial maximum rounded UP to bf16
        __nv_bfloat16* slot = xch + ((i * 2 + (int)par) * 2) * BQ;
        const float mine = bf16_ceil(mraw);
        slot[hh * BQ + row_in_tile] = __float2bfloat16_rn(mine);  // exact: already a bf16 value
        named_bar_sync(1 + i, kWarpsPerTile * 32);
        mraw = fmaxf(mine, __bfloat162float(slot[(hh ^ 1) * BQ + row_in_tile]));
      }
      const float mx = mraw * c;
      if (j == 0) {
        m_used = mx;
      } else {
        const float m_new = fmaxf(m_used, mx);
        const bool need = (m_new - m_used) > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {  // warp-uniform, and identical in both halves (same rows, same maxima)
          // every PV of step j - 1 must have landed in O; none of step j is issued before P(j) is published
          mbar_wait_a(bar + 8 * (bPFree + i * 4 + 3), par ^ 1u);
          tc_fence_after();
          const float f = ex2(m_used - m_new);
          l *= f;
          m_used = m_new;
#pragma unroll 1
          for (int ch = 0; ch < OC / 16; ++ch) {
            uint32_t o[16];
            tmem_ld16(t_o + ch * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
            tmem_st16(t_o + ch * 16, o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
      }
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_used, -m_used);
      float2 sum0 = make_float2(0.f, 0.f), sum1 = sum0;
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) {
        const int q = hh * NQ + qq;  // quarter of the step (compile-time when !SPLIT; hh * 2 + qq otherwise)
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 x = __ffma2_rn(make_float2(s[qq * 32 + 2 * k], s[qq * 32 + 2 * k + 1]), c2, nm2);
          float2 e;
          if (POLY > 0 && (k % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
            e = ex2_poly2(x);
          } else {
            e.x = ex2(x.x);
            e.y = ex2(x.y);
          }
          if (k & 1) sum1 = __fadd2_rn(sum1, e);
          else sum0 = __fadd2_rn(sum0, e);
          __nv_bfloat162 h = __floats2bfloat162_rn(e.x, e.y);
          pk[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        // slot q & 1 was last read by PV quarter q - 2 (this step) or q + 2 (previous step)
        if (q >= 2) mbar_wait_a(bar + 8 * (bPFree + i * 4 + q - 2), par);
        else if (j > 0) mbar_wait_a(bar + 8 * (bPFree + i * 4 + q + 2), par ^ 1u);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((qq & 1) * 4 + ch) ^ sw;  // q & 1 == qq & 1 (NQ is even)
          st_shared_v4(p_row + chunk * 16, pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bar + 8 * (bPReady + i * 4 + q));
      }
      l += (sum0.x + sum0.y) + (sum1.x + sum1.y);
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait_a(bar + 8 * (bOFull + i), 0);
    tc_fence_after();
    if constexpr (SPLIT) {  // row sum = the two halves' partial sums, exchanged through the (now idle) P operand of the tile
      float* lx = reinterpret_cast<float*>(sP + i * kBytesP);
      lx[hh * BQ + row_in_tile] = l;
      named_bar_sync(1 + i, kWarpsPerTile * 32);
      l += lx[(hh ^ 1) * BQ + row_in_tile];
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = p.O + (int64_t)batch * p.o_bs + (int64_t)row * p.o_rs + head * 128 + hh * OC;
#pragma unroll
    for (int ch = 0; ch < OC / 32; ++ch) {
      uint32_t o[32];
      tmem_ld32(t_o + ch * 32, o);
      tmem_ld_wait();
      if (row < p.n_q) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
          uint4 prev;
          if (p.accumulate) prev = *reinterpret_cast<const uint4*>(orow + ch * 32 + g * 8);
          const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = __uint_as_float(o[g * 8 + 2 * e]) * inv;
            float y = __uint_as_float(o[g * 8 + 2 * e + 1]) * inv;
            if (p.accumulate) {
              float2 qv = __bfloat1622float2(ph[e]);
              x = bf16_round(x) + qv.x;
              y = bf16_round(y) + qv.y;
            }
            h[e] = __floats2bfloat162_rn(x, y);
          }
          *reinterpret_cast<uint4*>(orow + ch * 32 + g * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int POLY, int SPLIT>
static int launch(const alg_attention_t* a, cudaStream_t st) {
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  ALG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 64 || !(attr_done.load(std::memory_order_relaxed) >> dev & 1)) {
    ALG_CUDA_OK(cudaFuncSetAttribute(attention_ps_kernel<POLY, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    if (dev < 64) attr_done.fetch_or(uint64_t(1) << dev, std::memory_order_relaxed);
  }
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t hd = (uint64_t)a->heads * 128;
  {
    uint64_t dims[3] = {hd, (uint64_t)a->n_q, (uint64_t)a->batch}, strides[3] = {1, (uint64_t)a->q_rs, (uint64_t)a->q_bs};
    uint32_t box[3] = {64, BQ, 1};
    if (int rc = make_tmap_bf16(&tmQ, a->Q, 3, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[3] = {hd, (uint64_t)a->n_kv, (uint64_t)a->batch}, strides[3] = {1, (uint64_t)a->k_rs, (uint64_t)a->k_bs};
    uint32_t box[3] = {64, BK, 1};
    if (int rc = make_tmap_bf16(&tmK, a->K, 3, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->n_kv, hd, (uint64_t)a->batch}, strides[3] = {1, (uint64_t)a->v_rs, (uint64_t)a->v_bs};
    uint32_t box[3] = {64, 128, 1};
    if (int rc = make_tmap_bf16(&tmV, a->Vt, 3, dims, strides, box)) return rc;
  }
  Params p;
  p.O = reinterpret_cast<__nv_bfloat16*>(a->O);
  p.o_bs = a->o_bs;
  p.o_rs = a->o_rs;
  p.n_q = (int)a->n_q;
  p.n_kv = (int)a->n_kv;
  p.heads = a->heads;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("ALG_ATTN_STAGGER");
      dbg = e ? atoi(e) : 0;
    }
    p.stagger = dbg;
  }
  dim3 grid((unsigned)((a->n_q + 2 * BQ - 1) / (2 * BQ)), (unsigned)a->heads, (unsigned)a->batch);
  attention_ps_kernel<POLY, SPLIT><<<grid, (SPLIT ? 19 : 11) * 32, kSmem, st>>>(tmQ, tmK, tmV, p);
  ALG_LAUNCH_OK();
  return 0;
}
}  // namespace ps

}  // namespace attn
}  // namespace alg

#ifdef ALG_ATTN_TRACE
extern "C" int alg_attention_trace_buffer(void* buf) {  // device buffer of 32 x int64 per CTA of the next launches (NULL = off)
  return cudaMemcpyToSymbol(alg::attn::g_attn_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int alg_attention_bf16(const alg_attention_t* a, void* stream) {
  using namespace alg;
  ALG_REQUIRE(a && a->Q && a->K && a->Vt && a->O, "attention: null pointer");
  ALG_REQUIRE(a->head_dim == 64 || a->head_dim == 128, "attention: head_dim must be 64 or 128");
  ALG_REQUIRE(a->batch > 0 && a->heads > 0 && a->n_q > 0 && a->n_kv > 0, "attention: empty problem");
  ALG_REQUIRE(a->batch <= 65535 && a->heads <= 65535, "attention: batch/heads exceed the grid limits");
  ALG_REQUIRE(a->q_rs % 8 == 0 && a->k_rs % 8 == 0 && a->v_rs % 8 == 0 && a->o_rs % 8 == 0 && a->q_bs % 8 == 0 &&
                  a->k_bs % 8 == 0 && a->v_bs % 8 == 0 && a->o_bs % 8 == 0,
              "attention: strides must be multiples of 8 elements (16 bytes)");
  ALG_REQUIRE(a->v_rs >= a->n_kv, "attention: Vt row stride smaller than n_kv");
  ALG_REQUIRE((reinterpret_cast<uintptr_t>(a->O) & 15) == 0, "attention: O must be 16-byte aligned");
  if (int rc = alg_check_device()) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static int poly = -1, split = -1, short_max = -1, pair = 0, s128 = 0, ps = 0, mc = 0, fx = 0;  // tuning knobs; defaults from profiling
  if (poly < 0) {
    const char* e = getenv("ALG_ATTN_POLY");
    poly = e ? atoi(e) : ALG_ATTN_POLY_DEFAULT;
    e = getenv("ALG_ATTN_SPLIT");
    split = e ? atoi(e) : ALG_ATTN_SPLIT_DEFAULT;
    e = getenv("ALG_ATTN_SHORT_MAX");  // key counts up to this use the SHORT (one tile, two CTAs per SM) variant; 0 = never
    short_max = e ? atoi(e) : ALG_ATTN_SHORT_MAX_DEFAULT;
    e = getenv("ALG_ATTN_PAIR");  // CTA pairs (cta_group::2 MMAs) for the long variant
    pair = e ? atoi(e) : ALG_ATTN_PAIR_DEFAULT;
    e = getenv("ALG_ATTN_S128");  // one 128-key S MMA per pair of steps (long variant, one thread per row)
    s128 = e ? atoi(e) : ALG_ATTN_S128_DEFAULT;
    e = getenv("ALG_ATTN_PS");  // P through shared memory, 128-key S one step ahead (head_dim 128, long variant)
    ps = e ? atoi(e) : ALG_ATTN_PS_DEFAULT;
    e = getenv("ALG_ATTN_MC");  // long variant: two-CTA clusters that TMA-multicast the K / V^T stream (half the L2 reads)
    mc = e ? atoi(e) : ALG_ATTN_MC_DEFAULT;
    e = getenv("ALG_ATTN_FX");  // EXPERIMENT: fixed softmax reference (see attention_kernel); 0 in the product
    fx = e ? atoi(e) : 0;
  }
  const bool use_short = a->n_kv <= short_max;
  // the cluster variant pads an odd grid with one CTA that does a whole CTA's work for nothing: only where that is < 2 %
  const int64_t gx = (a->n_q + 2 * attn::BQ - 1) / (2 * attn::BQ);
  const bool mc_ok = mc == 2 || (mc && (gx % 2 == 0 || gx >= 64));  // ALG_ATTN_MC=2 forces it (tests of the padded grid)
  if (ps && !use_short && a->head_dim == 128) {
    if (ps == 2) {  // two threads per row
      switch (poly) {
        case 0: return attn::ps::launch<0, 1>(a, st);
        case 4: return attn::ps::launch<4, 1>(a, st);
        default: return attn::ps::launch<8, 1>(a, st);
      }
    }
    switch (poly) {
      case 0: return attn::ps::launch<0, 0>(a, st);
      case 4: return attn::ps::launch<4, 0>(a, st);
      default: return attn::ps::launch<8, 0>(a, st);
    }
  }
#define ALG_ATTN_DISPATCH(DD)                                                        \
  if (use_short) return attn::launch<DD, 8, 0, 1, 2>(a, st);                         \
  if (fx && !s128 && !pair && !split) {                                              \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 0, 2, 4, 0, 0, 0, 1>(a, st);                \
      case 2: return attn::launch<DD, 2, 0, 2, 4, 0, 0, 0, 1>(a, st);                \
      case 4: return attn::launch<DD, 4, 0, 2, 4, 0, 0, 0, 1>(a, st);                \
      default: return attn::launch<DD, 8, 0, 2, 4, 0, 0, 0, 1>(a, st);               \
    }                                                                                \
  }                                                                                  \
  if (mc_ok && !s128 && !pair && !split) {                                           \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 0, 2, 4, 0, 0, 1>(a, st);                   \
      case 4: return attn::launch<DD, 4, 0, 2, 4, 0, 0, 1>(a, st);                   \
      default: return attn::launch<DD, 8, 0, 2, 4, 0, 0, 1>(a, st);                  \
    }                                                                                \
  }                                                                                  \
  if (s128 && !pair && split) {                                                      \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 1, 2, 4, 0, 1>(a, st);                      \
      case 4: return attn::launch<DD, 4, 1, 2, 4, 0, 1>(a, st);                      \
      default: return attn::launch<DD, 8, 1, 2, 4, 0, 1>(a, st);                     \
    }                                                                                \
  }                                                                                  \
  if (s128 && !pair && !split) {                                                     \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 0, 2, 4, 0, 1>(a, st);                      \
      case 2: return attn::launch<DD, 2, 0, 2, 4, 0, 1>(a, st);                      \
      case 3: return attn::launch<DD, 3, 0, 2, 4, 0, 1>(a, st);                      \
      case 4: return attn::launch<DD, 4, 0, 2, 4, 0, 1>(a, st);                      \
      default: return attn::launch<DD, 8, 0, 2, 4, 0, 1>(a, st);                     \
    }                                                                                \
  }                                                                                  \
  if (pair && !split) {                                                              \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 0, 2, 4, 1>(a, st);                         \
      case 4: return attn::launch<DD, 4, 0, 2, 4, 1>(a, st);                         \
      default: return attn::launch<DD, 8, 0, 2, 4, 1>(a, st);                        \
    }                                                                                \
  }                                                                                  \
  if (split) {                                                                       \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 1, 2, 4>(a, st);                            \
      case 4: return attn::launch<DD, 4, 1, 2, 4>(a, st);                            \
      default: return attn::launch<DD, 8, 1, 2, 4>(a, st);                           \
    }                                                                                \
  } else {                                                                           \
    switch (poly) {                                                                  \
      case 0: return attn::launch<DD, 0, 0, 2, 4>(a, st);                            \
      case 2: return attn::launch<DD, 2, 0, 2, 4>(a, st);                            \
      case 3: return attn::launch<DD, 3, 0, 2, 4>(a, st);                            \
      case 4: return attn::launch<DD, 4, 0, 2, 4>(a, st);                            \
      default: return attn::launch<DD, 8, 0, 2, 4>(a, st);                           \
    }                                                                                \
  }
  if (a->head_dim == 128) {
    ALG_ATTN_DISPATCH(128)
  }
  ALG_ATTN_DISPATCH(64)
#undef ALG_ATTN_DISPATCH
}
