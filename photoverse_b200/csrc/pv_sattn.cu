// photoverse_b200 -- self-attention (the `attn1` layers of the same transformer blocks).
//
// The reference installs diffusers' stock AttnProcessor2_0 on attn1 (reference models/unet.py:20-24), i.e.
//   O = softmax(Q K^T / sqrt(d)) V   per (sample, head), Q/K/V = the three bias-free projections of the hidden states.
// Two kernels:
//   sattn_pack_kernel  [B,S,*] row-major Q/K/V  ->  per-(sample, head, tile) UMMA operand images (canonical no-swizzle
//                      core-matrix layout [16-byte chunk][row][8 bf16]); V gets an extra column of ones, so that the
//                      P.V contraction also produces the softmax denominator.  (The 1/sqrt(d) scale is applied to the
//                      fp32 scores inside the exponent's FMA: scaling Q would round it to bf16 a second time.)
//   self_attn_fwd_kernel  persistent, one CTA per SM, 12 warps: TMA producer / tcgen05 issuer / 2 softmax groups.  A work
//                      unit is 2 x 128 query rows of one (sample, head) against all keys in tiles of BN; S = Q K^T and
//                      O += P V on the tensor cores with fp32 accumulators in TMEM, online softmax with one thread per
//                      query row and LAZY rescaling (the O accumulator is only touched when a row maximum grows by more
//                      than 2^8), P handed to the tensor core through shared memory.
//
// The same K / V image serves both contractions: as the K-major B operand of Q K^T (rows = keys, 16-byte chunks along d)
// and as the MN-major B operand of P V (rows = the contraction index, chunks along the output columns) -- no transpose.
#include "pv_common.cuh"
#include "pv_host.h"
#include "pv_softmax.cuh"
#include "../../include/photoverse_b200.h"

namespace pv {

template <int D>
struct SaCfg {
  static constexpr int DK = (D + 15) / 16 * 16;          // contraction length of Q K^T: 48 / 80 / 160
  static constexpr int NQ = DK / 8;                      // 16-byte chunks per Q / K row
  static constexpr int DN = (D + 1 + 15) / 16 * 16;      // output columns of P V incl. the ones column: 48 / 96 / 176
  static constexpr int NV = DN / 8;
  static constexpr int BN = 64;                          // keys per tile
  // head_dim 40: TWO CTAs per SM.  The softmax is a long serial instruction stream per warp (TMEM load, maximum,
  // exponentials, packing, barrier round trips); with one CTA there are two such warps per scheduler and the issue slots
  // stay two thirds empty (device timeline: the groups' phases do not even overlap).  Two CTAs double the warps in
  // flight; each gets half the registers, shared memory and TMEM: P overwrites the first columns of its own S
  // (ALIAS), so S_t(j+1) can only be issued behind P_t V(j) -- the other three tiles of the SM cover that bubble.
  static constexpr bool ALIAS = D == 40;
  static constexpr int CTAS_PER_SM = ALIAS ? 2 : 1;
  // P as the A operand of P V: in TMEM (packed bf16, BN / 2 columns per tile) where the 512 columns allow it.  With a
  // small head_dim the shared-memory port is the bottleneck of the SS form -- 32 KB of P written and read back per
  // 128 x 128 tile for 48 output columns -- so head_dim 40 / 80 keep P in TMEM; head_dim 160 (O alone takes 352 columns)
  // hands P over through shared memory.
  static constexpr bool P_TMEM = D <= 80;
  static constexpr int Q_BYTES = NQ * 128 * 16;
  static constexpr int K_BYTES = NQ * BN * 16;
  static constexpr int V_BYTES = NV * BN * 16;
  static constexpr int O_STAGE_BYTES = 128 * D * 2;      // bf16 output rows of one 128-row tile
  static constexpr int P_BYTES = P_TMEM ? O_STAGE_BYTES : (BN / 8) * 128 * 16;   // P_TMEM: only the output staging
  static constexpr int STAGES = D == 160 ? 2 : 4;
  // No setmaxnreg: with 64 keys per tile every role fits the launch-bound register count (168, resp. 80 with two CTAs
  // per SM), and a kernel that re-allocates registers must serve the increase from its own launch allocation.
  static constexpr bool STAGE_IN_Q = O_STAGE_BYTES > P_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_P = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_KV = OFF_P + 2 * P_BYTES;
  static constexpr int OFF_BAR = OFF_KV + STAGES * (K_BYTES + V_BYTES);
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int TM_S = 0;                         // S_t at TM_S + t * BN
  static constexpr int TM_P = ALIAS ? 0 : 2 * BN;        // P_t at TM_P + t * P_STRIDE   (P_TMEM)
  static constexpr int P_STRIDE = ALIAS ? BN : BN / 2;
  static constexpr int TM_O = (P_TMEM && !ALIAS) ? 3 * BN : 2 * BN;  // O_t at TM_O + t * DN
  static constexpr int TM_COLS = ALIAS ? 256 : 512;
  static_assert(TM_O + 2 * DN <= TM_COLS, "TMEM columns");
  static_assert(SMEM_BYTES <= (ALIAS ? 113 : 227) * 1024, "shared memory per CTA");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(O_STAGE_BYTES <= (STAGE_IN_Q ? Q_BYTES : P_BYTES), "output staging");
};

constexpr int SA_THREADS = 384;

struct SaParams {
  const uint8_t* qimg;
  const uint8_t* kimg;
  const uint8_t* vimg;
  int S, H;
  int nQT;        // 128-row query tiles per (sample, head)
  int nKT;        // key tiles per (sample, head)
  int nU;         // units (pairs of query tiles) per (sample, head)
  int units;      // B * H * nU
  unsigned long long* trace;   // debug timeline (PV_TRACE builds)
  int trace_cap, trace_block;
};

// ------------------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) sattn_pack_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                         const __nv_bfloat16* __restrict__ v, long long ld, uint8_t* __restrict__ qimg,
                                                         uint8_t* __restrict__ kimg, uint8_t* __restrict__ vimg, int S, int H, int nQT,
                                                         int nKT) {
  using Cfg = SaCfg<D>;
  const int rt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const long long bh = static_cast<long long>(b) * H + h;
  const long long row_base = static_cast<long long>(b) * S;
  const int col0 = h * D;
  // Q image of this 128-row tile
  {
    uint4* dst = reinterpret_cast<uint4*>(qimg + (bh * nQT + rt) * Cfg::Q_BYTES);
    for (int i = threadIdx.x; i < Cfg::NQ * 128; i += 256) {
      const int ch = i >> 7, r = i & 127, row = rt * 128 + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (row < S && ch * 8 < D)
        val = __ldg(reinterpret_cast<const uint4*>(q + (row_base + row) * ld + col0 + ch * 8));
      dst[i] = val;
    }
  }
  constexpr int TPR = 128 / Cfg::BN;                    // key tiles per 128-row group
#pragma unroll
  for (int sub = 0; sub < TPR; ++sub) {
    const int kt = rt * TPR + sub;
    if (kt >= nKT) break;
    uint4* dk = reinterpret_cast<uint4*>(kimg + (bh * nKT + kt) * Cfg::K_BYTES);
    for (int i = threadIdx.x; i < Cfg::NQ * Cfg::BN; i += 256) {
      const int ch = i / Cfg::BN, r = i % Cfg::BN, row = kt * Cfg::BN + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (row < S && ch * 8 < D) val = __ldg(reinterpret_cast<const uint4*>(k + (row_base + row) * ld + col0 + ch * 8));
      dk[i] = val;
    }
    uint4* dv = reinterpret_cast<uint4*>(vimg + (bh * nKT + kt) * Cfg::V_BYTES);
    for (int i = threadIdx.x; i < Cfg::NV * Cfg::BN; i += 256) {
      const int ch = i / Cfg::BN, r = i % Cfg::BN, row = kt * Cfg::BN + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (row < S) {
        if (ch * 8 < D) val = __ldg(reinterpret_cast<const uint4*>(v + (row_base + row) * ld + col0 + ch * 8));
        else if (ch * 8 == D) val.x = 0x00003f80u;     // column D = 1.0: the P V contraction also sums the row of P
      }
      dv[i] = val;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------------
// Two fp32 -> packed bf16 by TRUNCATION (one byte permute on the ALU pipe).  cvt.rn.bf16x2.f32 (F2FP) shares the XU pipe
// with the exponentials, which is the pipe that bounds this kernel; the one-sided error (< 2^-8 relative) is common to
// the numerator sum(P V) and the denominator sum(P) -- both are contracted from the same packed P -- and cancels there.
__device__ __forceinline__ uint32_t pack_bf16x2_trunc(float lo, float hi) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// 2^x for a pair on the FMA pipe (the MUFU unit delivers 16 exponentials per clock and SM, which is what bounds an
// attention kernel of small head_dim): x = n + r with n = rint(x) taken from the low mantissa bits of x + 1.5*2^23,
// 2^r by a degree-3 minimax polynomial on [-0.5, 0.5] (7.5e-5 relative -- P is rounded to bf16 anyway), the exponent
// added to the bit pattern.  x <= 8 here (lazy rescaling); anything below 2^-126 is clamped to it.
__device__ __forceinline__ uint32_t exp2_pair_poly_bf16(float a, float b) {
  a = fmaxf(a, -126.f);
  b = fmaxf(b, -126.f);
  const uint64_t x2 = f2_pack(a, b);
  const uint64_t magic = f2_pack(12582912.f, 12582912.f);
  const uint64_t t2 = f2_add(x2, magic);
  const uint64_t n2 = f2_add(t2, f2_pack(-12582912.f, -12582912.f));
  const uint64_t r2 = f2_fma(n2, f2_pack(-1.f, -1.f), x2);
  uint64_t p2 = f2_fma(r2, f2_pack(0.055171459913253784f, 0.055171459913253784f), f2_pack(0.2426108568906784f, 0.2426108568906784f));
  p2 = f2_fma(p2, r2, f2_pack(0.6932609677314758f, 0.6932609677314758f));
  p2 = f2_fma(p2, r2, f2_pack(0.9999281167984009f, 0.9999281167984009f));
  float pa, pb, ta, tb;
  f2_unpack(p2, pa, pb);
  f2_unpack(t2, ta, tb);
  const float ea = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ta) << 23));
  const float eb = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(tb) << 23));
  return pack_bf16x2_trunc(ea, eb);
}

template <int D, int PF>     // PF: pairs out of every 8 whose exponentials run on the FMA pipe instead of MUFU
__global__ void __launch_bounds__(SA_THREADS, SaCfg<D>::CTAS_PER_SM) self_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmO, const SaParams p) {
  using Cfg = SaCfg<D>;
  constexpr int BN = Cfg::BN;
  constexpr float SCALE_LOG2E = D == 40 ? 0.22811485f : (D == 80 ? 0.16130156f : 0.11405742f);   // log2(e) / sqrt(D)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_free = bars + 2;        // [2]
  uint64_t* s_full = bars + 4;        // [2]
  uint64_t* s_free = bars + 6;        // [2]
  uint64_t* p_ready = bars + 8;       // [2]
  uint64_t* pv_done = bars + 10;      // [2]
  uint64_t* k_full = bars + 12;       // [STAGES]   K and V tiles travel in separate rings: a K tile is released as soon as
  uint64_t* k_empty = bars + 16;      // [STAGES]   both S = Q K^T of its step are done, a V tile only after both P V
  uint64_t* v_full = bars + 20;       // [STAGES]
  uint64_t* v_empty = bars + 24;      // [STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_free[t], 4);
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      mbar_init(&p_ready[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmO);
  }
  if (warp == 2) tmem_alloc<Cfg::TM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                             // the images are written by the pack kernel just before

  const int S = p.S, nKT = p.nKT;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    if (elect_one()) {
      uint32_t kvc = 0, it0 = 0, it1 = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const long long bh = u / p.nU;
        const int qu = u % p.nU;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int qt = 2 * qu + t;
          if (qt >= p.nQT) continue;
          uint32_t& it = t == 0 ? it0 : it1;
          if (it > 0) mbar_wait(&q_free[t], (it - 1) & 1);
          mbar_expect_tx(&q_full[t], Cfg::Q_BYTES);
          bulk_load_1d(smem + Cfg::OFF_Q + t * Cfg::Q_BYTES, p.qimg + (bh * p.nQT + qt) * Cfg::Q_BYTES, Cfg::Q_BYTES, &q_full[t]);
          ++it;
        }
        for (int j = 0; j < nKT; ++j, ++kvc) {
          const uint32_t s = kvc % Cfg::STAGES;
          uint8_t* dst = smem + Cfg::OFF_KV + s * (Cfg::K_BYTES + Cfg::V_BYTES);
          if (kvc >= Cfg::STAGES) mbar_wait(&k_empty[s], (kvc / Cfg::STAGES - 1) & 1);
          mbar_expect_tx(&k_full[s], Cfg::K_BYTES);
          bulk_load_1d(dst, p.kimg + (bh * nKT + j) * Cfg::K_BYTES, Cfg::K_BYTES, &k_full[s]);
          if (kvc >= Cfg::STAGES) mbar_wait(&v_empty[s], (kvc / Cfg::STAGES - 1) & 1);
          mbar_expect_tx(&v_full[s], Cfg::V_BYTES);
          bulk_load_1d(dst + Cfg::K_BYTES, p.vimg + (bh * nKT + j) * Cfg::V_BYTES, Cfg::V_BYTES, &v_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- tcgen05 issuer
    constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, BN);
    constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, Cfg::DN) | (1u << 16);       // B operand MN-major
    uint32_t kvc = 0;
    uint32_t g[2] = {0, 0};          // key steps completed per tile (phase counters of s_full / s_free / p_ready / pv_done)
    uint32_t it[2] = {0, 0};
    const uint32_t q_addr = smem_u32(smem + Cfg::OFF_Q), p_addr = smem_u32(smem + Cfg::OFF_P);
    const uint32_t kv_addr = smem_u32(smem + Cfg::OFF_KV);
    auto issue_qk = [&](int t, uint32_t stage) {
      const uint64_t a = umma_desc(q_addr + t * Cfg::Q_BYTES, 128 * 16, 128, UMMA_LAYOUT_NONE);
      const uint64_t b = umma_desc(kv_addr + stage * (Cfg::K_BYTES + Cfg::V_BYTES), BN * 16, 128, UMMA_LAYOUT_NONE);
#pragma unroll
      for (int ks = 0; ks < Cfg::DK / 16; ++ks)
        umma_bf16_ss(tmem + Cfg::TM_S + t * BN, a + static_cast<uint64_t>(ks * (2 * 128 * 16 >> 4)),
                     b + static_cast<uint64_t>(ks * (2 * BN * 16 >> 4)), IDESC_QK, ks > 0);
      umma_commit(&s_full[t]);
    };
    auto issue_pv = [&](int t, uint32_t stage, bool first) {
      // MN-major, no swizzle: LBO = distance between 8-row groups of the contraction index, SBO = between 8-column groups
      const uint64_t b = umma_desc(kv_addr + stage * (Cfg::K_BYTES + Cfg::V_BYTES) + Cfg::K_BYTES, 128, BN * 16, UMMA_LAYOUT_NONE);
      if constexpr (Cfg::P_TMEM) {
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks)
          umma_bf16_ts(tmem + Cfg::TM_O + t * Cfg::DN, tmem + Cfg::TM_P + t * Cfg::P_STRIDE + ks * 8,
                       b + static_cast<uint64_t>(ks * (256 >> 4)), IDESC_PV, (!first || ks > 0) ? 1u : 0u);
      } else {
        const uint64_t a = umma_desc(p_addr + t * Cfg::P_BYTES, 128 * 16, 128, UMMA_LAYOUT_NONE);
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks)
          umma_bf16_ss(tmem + Cfg::TM_O + t * Cfg::DN, a + static_cast<uint64_t>(ks * (2 * 128 * 16 >> 4)),
                       b + static_cast<uint64_t>(ks * (256 >> 4)), IDESC_PV, (!first || ks > 0) ? 1u : 0u);
      }
      umma_commit(&pv_done[t]);
    };
    A3Trace tr = a3_trace_init_raw(p.trace, p.trace_cap, 0, p.trace_block);
    auto qk = [&](int t, uint32_t stage) {
      tc_fence_after();
      if (elect_one()) issue_qk(t, stage);
      __syncwarp();
      a3_trace(tr, 10 + t, 0);
    };
    auto pv = [&](int t, uint32_t stage, bool first) {
      tc_fence_after();
      if (elect_one()) issue_pv(t, stage, first);
      __syncwarp();
      a3_trace(tr, 12 + t, 0);
    };
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int qu = u % p.nU;
      const int nt = (2 * qu + 1 < p.nQT) ? 2 : 1;
      const uint32_t b0 = g[0], b1 = g[1];              // phase counters of this unit's step 0
      auto stage_of = [&](int j) { return (kvc + j) % Cfg::STAGES; };
      auto k_wait = [&](int j) { mbar_wait(&k_full[stage_of(j)], ((kvc + j) / Cfg::STAGES) & 1); };
      auto v_wait = [&](int j) { mbar_wait(&v_full[stage_of(j)], ((kvc + j) / Cfg::STAGES) & 1); };
      auto release = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      if constexpr (Cfg::ALIAS) {
        // P_t overwrites S_t: S_t(j+1) goes behind P_t V(j) (the tensor pipe executes in issue order)
        k_wait(0);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&q_full[t], it[t] & 1);
          ++it[t];
          qk(t, stage_of(0));
        }
        release(&k_empty[stage_of(0)]);
        for (int j = 0; j < nKT; ++j) {
          v_wait(j);
          if (j + 1 < nKT) k_wait(j + 1);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&p_ready[t], ((t ? b1 : b0) + j) & 1);
            pv(t, stage_of(j), j == 0);
            if (j + 1 < nKT) qk(t, stage_of(j + 1));
          }
          release(&v_empty[stage_of(j)]);
          if (j + 1 < nKT) release(&k_empty[stage_of(j + 1)]);
        }
        g[0] += nKT;
        if (nt == 2) g[1] += nKT;
        kvc += nKT;
        continue;
      }
      k_wait(0);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(&q_full[t], it[t] & 1);
        ++it[t];
        if (g[t] > 0) mbar_wait(&s_free[t], (g[t] - 1) & 1);     // the previous unit's last S_t has been read
        qk(t, stage_of(0));
      }
      release(&k_empty[stage_of(0)]);
      // both S(j+1) as soon as S(j) has been read, then both P V (j)
      for (int j = 0; j < nKT; ++j) {
        if (j + 1 < nKT) {
          k_wait(j + 1);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&s_free[t], ((t ? b1 : b0) + j) & 1);
            qk(t, stage_of(j + 1));
          }
          release(&k_empty[stage_of(j + 1)]);
        }
        v_wait(j);
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&p_ready[t], ((t ? b1 : b0) + j) & 1);
          pv(t, stage_of(j), j == 0);
        }
        release(&v_empty[stage_of(j)]);
      }
      g[0] += nKT;
      if (nt == 2) g[1] += nKT;
      kvc += nKT;
    }
    a3_trace_done_raw(p.trace, tr, 0);
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax groups: one thread per query row
    const int t = (warp - 4) >> 2;                 // tile / group
    const int wq = warp & 3;                       // TMEM lane quarter
    const int row = wq * 32 + lane;                // row inside the 128-row tile
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t s_taddr = tmem + lane_base + Cfg::TM_S + t * BN;
    const uint32_t o_taddr = tmem + lane_base + Cfg::TM_O + t * Cfg::DN;
    const uint32_t p_taddr = tmem + lane_base + Cfg::TM_P + t * Cfg::P_STRIDE;
    const uint32_t p_row = smem_u32(smem + Cfg::OFF_P + t * Cfg::P_BYTES) + row * 16;
    uint8_t* stage_base = smem + (Cfg::STAGE_IN_Q ? Cfg::OFF_Q + t * Cfg::Q_BYTES : Cfg::OFF_P + t * Cfg::P_BYTES) + wq * (32 * D * 2);
    uint32_t g = 0;
    A3Trace tr = a3_trace_init_raw(p.trace, p.trace_cap, 1 + t, wq == 0 ? p.trace_block : -2);
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int bh = u / p.nU, qu = u % p.nU;
      const int qt = 2 * qu + t;
      if (qt >= p.nQT) continue;
      float m_ref = 0.f;
      for (int j = 0; j < nKT; ++j, ++g) {
        a3_trace(tr, 20, j);
        mbar_wait(&s_full[t], g & 1);
        a3_trace(tr, 21, j);
        tc_fence_after();
        // CH columns of S are held in registers at a time: all of them with one CTA per SM; 32 with two CTAs per SM
        // (80 registers per thread, no setmaxnreg: a kernel that re-allocates registers gets one CTA per SM) -- then the
        // row is read from TMEM twice, once for the maximum and once for the exponentials.
        constexpr int CH = Cfg::ALIAS ? 32 : BN;
        uint32_t s[CH];
        const int valid = S - j * BN;
        auto load_chunk = [&](int h) {
#pragma unroll
          for (int c = 0; c < CH; c += 32) tmem_ld32_raw(s_taddr + h * CH + c, s + c);
          tmem_ld_wait();
          if (valid < BN) {
#pragma unroll
            for (int i = 0; i < CH; ++i)
              if (h * CH + i >= valid) s[i] = 0xff800000u;    // -inf
          }
        };
        float mx = -INFINITY;
#pragma unroll
        for (int h = 0; h < BN / CH; ++h) {
          load_chunk(h);
          if (h == BN / CH - 1 && BN == CH) {
            tc_fence_before();
            __syncwarp();
            if (!Cfg::ALIAS && lane == 0) mbar_arrive(&s_free[t]);
          }
          // row maximum: four independent chains of 3-input maxima
          float mxa = fmax3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
          float mxb = fmax3(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
          float mxc = fmax3(__uint_as_float(s[6]), __uint_as_float(s[7]), __uint_as_float(s[8]));
          float mxd = fmax3(__uint_as_float(s[9]), __uint_as_float(s[10]), __uint_as_float(s[11]));
#pragma unroll
          for (int i = 12; i + 7 < CH; i += 8) {
            mxa = fmax3(mxa, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
            mxb = fmax3(mxb, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
            mxc = fmax3(mxc, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
            mxd = fmax3(mxd, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
          }
          static_assert((CH - 12) % 8 == 4, "tail of the maximum pass");
          mxa = fmax3(mxa, __uint_as_float(s[CH - 4]), __uint_as_float(s[CH - 3]));
          mxb = fmax3(mxb, __uint_as_float(s[CH - 2]), __uint_as_float(s[CH - 1]));
          mx = fmaxf(mx, fmaxf(fmax3(mxa, mxb, mxc), mxd));
        }
        a3_trace(tr, 22, j);
        if constexpr (!Cfg::STAGE_IN_Q) {
          // the last S of the unit has been produced: Q_t is no longer read by the tensor core.  (Where the staging of the
          // output re-uses Q_t it is released to the producer only after the output has left, below.)
          if (j == nKT - 1 && lane == 0) mbar_arrive(&q_free[t]);
        }
        // P_t (shared memory) and O_t are quiescent once P V (j-1) has completed -- long before this point
        a3_trace(tr, 23, j);
        if (j > 0) mbar_wait(&pv_done[t], (g - 1) & 1);
        a3_trace(tr, 24, j);
        if (j == 0) {
          m_ref = mx;
        } else {
          const bool need = (mx - m_ref) * SCALE_LOG2E > 8.f;
          if (__any_sync(0xffffffffu, need)) {
            tc_fence_after();
            const float alpha = need ? fast_exp2((m_ref - mx) * SCALE_LOG2E) : 1.f;
            if (need) m_ref = mx;
#pragma unroll
            for (int c = 0; c < Cfg::DN; c += 16) {
              uint32_t o[16];
              tmem_ld16_raw(o_taddr + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(o_taddr + c, o);
            }
            tmem_st_wait();
          }
        }
        a3_trace(tr, 25, j);
        // P = 2^(S - m_ref), stored 8 keys (16 bytes) at a time as soon as they are packed
        // P = 2^((S - m_ref) * log2(e) / sqrt(d))
        const uint64_t nm = f2_pack(-m_ref * SCALE_LOG2E, -m_ref * SCALE_LOG2E), sc2 = f2_pack(SCALE_LOG2E, SCALE_LOG2E);
#pragma unroll
        for (int h = 0; h < BN / CH; ++h) {
          if (BN != CH) load_chunk(h);          // (P overwrites columns of S that are already in registers or consumed)
#pragma unroll
          for (int c = 0; c < CH / 8; ++c) {
            uint32_t pk[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int i = 8 * c + 2 * q4;
              float a, b;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, nm), a, b);
              // PF eighths of the pairs take the FMA-pipe polynomial: ceil(PF / 2) of 4 in even chunks, floor in odd ones
              if (q4 < ((c & 1) ? PF / 2 : (PF + 1) / 2)) pk[q4] = exp2_pair_poly_bf16(a, b);
              else pk[q4] = pack_bf16x2_trunc(fast_exp2(a), fast_exp2(b));
            }
            if constexpr (Cfg::P_TMEM) tmem_st_x4(p_taddr + (h * CH / 8 + c) * 4, pk);
            else st_shared_v4_a(p_row + (h * CH / 8 + c) * 2048, pk[0], pk[1], pk[2], pk[3]);
          }
        }
        a3_trace(tr, 26, j);
        if constexpr (Cfg::P_TMEM) tmem_st_wait();
        else fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        a3_trace(tr, 27, j);
      }
      // ---- epilogue: O_t / l -> bf16 -> staging -> TMA store
      mbar_wait(&pv_done[t], (g - 1) & 1);
      tc_fence_after();
      float inv_l;
      {
        uint32_t lr[8];
        tmem_ld_x8(o_taddr + D, lr);
        tmem_ld_wait();
        inv_l = 1.f / __uint_as_float(lr[0]);
      }
      uint8_t* my_row = stage_base + lane * (D * 2);
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        uint32_t o[8];
        tmem_ld_x8(o_taddr + c, o);
        tmem_ld_wait();
        // rotate the order of the 16-byte pieces by lane so that the 32 rows (pitch D*2 bytes) spread over the banks
        st_shared_v4(my_row + c * 2, pack_bf16x2(__uint_as_float(o[0]) * inv_l, __uint_as_float(o[1]) * inv_l),
                     pack_bf16x2(__uint_as_float(o[2]) * inv_l, __uint_as_float(o[3]) * inv_l),
                     pack_bf16x2(__uint_as_float(o[4]) * inv_l, __uint_as_float(o[5]) * inv_l),
                     pack_bf16x2(__uint_as_float(o[6]) * inv_l, __uint_as_float(o[7]) * inv_l));
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int h = bh % p.H, b = bh / p.H;
        tma_store_3d(&tmO, stage_base, h * D, qt * 128 + wq * 32, b);
        bulk_commit();
        bulk_wait_read<0>();
        if constexpr (Cfg::STAGE_IN_Q) mbar_arrive(&q_free[t]);
      }
      __syncwarp();
      // the staging area overlaps the other warps' rows of P_t (resp. Q_t): nobody of the group may go on before all
      // four stores have been read
      // (constant barrier ids: the SM's 16 named barriers are shared by its resident CTAs, and a computed id makes the
      // compiler reserve all 16 for one CTA)
      if (t == 0) named_bar_sync(1, 128);
      else named_bar_sync(2, 128);
      a3_trace(tr, 28, 0);
    }
    a3_trace_done_raw(p.trace, tr, 1 + t);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<Cfg::TM_COLS>(tmem);
}

extern int g_opt_sattn_poly;
extern int g_opt_trace_block;
extern unsigned long long* g_attn3_trace;
extern int g_attn3_trace_cap;

template <int D, int PF>
static int launch_sattn(const void* q, const void* k, const void* v, long long ld, void* out, void* ws, int B, int S, int C, int H,
                        cudaStream_t stream) {
  using Cfg = SaCfg<D>;
  const int nQT = (S + 127) / 128, nKT = (S + Cfg::BN - 1) / Cfg::BN, nU = (nQT + 1) / 2;
  const long long BH = static_cast<long long>(B) * H;
  uint8_t* qimg = static_cast<uint8_t*>(ws);
  uint8_t* kimg = qimg + BH * nQT * Cfg::Q_BYTES;
  uint8_t* vimg = kimg + BH * nKT * Cfg::K_BYTES;
  sattn_pack_kernel<D><<<dim3(nQT, H, B), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                                           static_cast<const __nv_bfloat16*>(v), ld, qimg, kimg, vimg, S, H, nQT, nKT);
  PV_LAUNCHED();
  CUtensorMap tmO;
  if (make_tmap_3d(&tmO, out, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, D, 32, 1, Swz::None)) return PV_ERR_CUDA;
  SaParams p;
  p.qimg = qimg; p.kimg = kimg; p.vimg = vimg;
  p.S = S; p.H = H; p.nQT = nQT; p.nKT = nKT; p.nU = nU;
  p.trace = g_attn3_trace; p.trace_cap = g_attn3_trace_cap; p.trace_block = g_opt_trace_block;
  const long long units = BH * nU;
  PV_REQUIRE(units < (1ll << 30), "too many work units");
  p.units = static_cast<int>(units);
  auto kern = self_attn_fwd_kernel<D, PF>;
  // (cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for every kernel that allocates tensor memory; two CTAs of
  // 256 columns each do share an SM -- tools/ubench/occ_test.cu)
  PV_CUDA(set_max_smem_once(kern, Cfg::SMEM_BYTES, Cfg::CTAS_PER_SM > 1));
  const long long sms = static_cast<long long>(sm_count()) * Cfg::CTAS_PER_SM;
  const int grid = static_cast<int>(units < sms ? units : sms);
  PV_CUDA(launch_pdl(kern, dim3(grid), dim3(SA_THREADS), Cfg::SMEM_BYTES, stream, tmO, p));
  PV_LAUNCHED();
  return PV_OK;
}

long long self_attn_ws_bytes(int B, int S, int C, int H) {
  if (B <= 0 || S <= 0 || H <= 0 || C % H) return -1;
  const int d = C / H;
  const long long BH = static_cast<long long>(B) * H;
  const long long nQT = (S + 127) / 128;
  auto total = [&](auto cfg) {
    using Cfg = decltype(cfg);
    const long long nKT = (S + Cfg::BN - 1) / Cfg::BN;
    return BH * (nQT * Cfg::Q_BYTES + nKT * (Cfg::K_BYTES + Cfg::V_BYTES));
  };
  if (d == 40) return total(SaCfg<40>{});
  if (d == 80) return total(SaCfg<80>{});
  if (d == 160) return total(SaCfg<160>{});
  return -1;
}

int self_attn_fwd_bf16(const void* q, const void* k, const void* v, long long ld, void* out, void* ws, int B, int S, int C, int H,
                       cudaStream_t stream) {
  PV_REQUIRE(B > 0 && S > 0 && H > 0 && C % H == 0, "bad shape B=%d S=%d C=%d H=%d", B, S, C, H);
  const int d = C / H;
  PV_REQUIRE(d == 40 || d == 80 || d == 160, "head_dim %d unsupported (40/80/160)", d);
  PV_REQUIRE(ld >= C && ld % 8 == 0, "row stride %lld must be >= C and a multiple of 8 elements", ld);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
              reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ws)) % 16 == 0, "pointers must be 16-byte aligned");
  const int pf = g_opt_sattn_poly;
#define PV_SA_DISPATCH(DD)                                                                          \
  switch (pf) {                                                                                     \
    case 0: return launch_sattn<DD, 0>(q, k, v, ld, out, ws, B, S, C, H, stream);                   \
    case 4: return launch_sattn<DD, 4>(q, k, v, ld, out, ws, B, S, C, H, stream);                   \
    default: return launch_sattn<DD, 2>(q, k, v, ld, out, ws, B, S, C, H, stream);                  \
  }
  if (d == 40) { PV_SA_DISPATCH(40) }
  if (d == 80) { PV_SA_DISPATCH(80) }
  PV_SA_DISPATCH(160)
#undef PV_SA_DISPATCH
}

}  // namespace pv
