// photoverse_b200 -- self-attention (the `attn1` layers of the same transformer blocks).
//
// The reference installs diffusers' stock AttnProcessor2_0 on attn1 (reference models/unet.py:20-24), i.e.
//   O = softmax(Q K^T / sqrt(d)) V   per (sample, head), Q/K/V = the three bias-free projections of the hidden states.
// Two kernels:
//   sattn_pack_kernel  [B,S,*] row-major Q/K/V  ->  per-(sample, head, tile) UMMA operand images (canonical no-swizzle
//                      core-matrix layout [16-byte chunk][row][8 bf16]); Q pre-multiplied by log2(e)/sqrt(d); V gets an
//                      extra column of ones, so that the P.V contraction also produces the softmax denominator.
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
  static constexpr int BN = D > 80 ? 64 : 128;           // keys per tile
  static constexpr int Q_BYTES = NQ * 128 * 16;
  static constexpr int K_BYTES = NQ * BN * 16;
  static constexpr int V_BYTES = NV * BN * 16;
  static constexpr int P_BYTES = (BN / 8) * 128 * 16;
  static constexpr int STAGES = D == 40 ? 4 : 2;
  static constexpr int O_STAGE_BYTES = 128 * D * 2;      // bf16 output rows of one 128-row tile
  static constexpr bool STAGE_IN_Q = O_STAGE_BYTES > P_BYTES;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_P = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_KV = OFF_P + 2 * P_BYTES;
  static constexpr int OFF_BAR = OFF_KV + STAGES * (K_BYTES + V_BYTES);
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int TM_S = 0;                         // S_t at TM_S + t * BN
  static constexpr int TM_O = 2 * BN;                    // O_t at TM_O + t * DN
  static_assert(TM_O + 2 * DN <= 512, "TMEM columns");
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
};

// ------------------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 scale_bf16x8(uint4 v, float s) {
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = __uint_as_float(w[i] << 16) * s;
    const float hi = __uint_as_float(w[i] & 0xffff0000u) * s;
    w[i] = pack_bf16x2(lo, hi);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int D>
__global__ void __launch_bounds__(256) sattn_pack_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                         const __nv_bfloat16* __restrict__ v, long long ld, uint8_t* __restrict__ qimg,
                                                         uint8_t* __restrict__ kimg, uint8_t* __restrict__ vimg, int S, int H, int nQT,
                                                         int nKT, float qscale) {
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
        val = scale_bf16x8(__ldg(reinterpret_cast<const uint4*>(q + (row_base + row) * ld + col0 + ch * 8)), qscale);
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
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int D>
__global__ void __launch_bounds__(SA_THREADS, 1) self_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmO, const SaParams p) {
  using Cfg = SaCfg<D>;
  constexpr int BN = Cfg::BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_free = bars + 2;        // [2]
  uint64_t* s_full = bars + 4;        // [2]
  uint64_t* s_free = bars + 6;        // [2]
  uint64_t* p_ready = bars + 8;       // [2]
  uint64_t* pv_done = bars + 10;      // [2]
  uint64_t* kv_full = bars + 12;      // [STAGES]
  uint64_t* kv_empty = bars + 16;     // [STAGES]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

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
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmO);
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                             // the images are written by the pack kernel just before

  const int S = p.S, nKT = p.nKT;

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");        // 128*56 + 256*216 <= 64 K registers
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
          if (kvc >= Cfg::STAGES) mbar_wait(&kv_empty[s], (kvc / Cfg::STAGES - 1) & 1);
          uint8_t* dst = smem + Cfg::OFF_KV + s * (Cfg::K_BYTES + Cfg::V_BYTES);
          mbar_expect_tx(&kv_full[s], Cfg::K_BYTES + Cfg::V_BYTES);
          bulk_load_1d(dst, p.kimg + (bh * nKT + j) * Cfg::K_BYTES, Cfg::K_BYTES, &kv_full[s]);
          bulk_load_1d(dst + Cfg::K_BYTES, p.vimg + (bh * nKT + j) * Cfg::V_BYTES, Cfg::V_BYTES, &kv_full[s]);
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
      const uint64_t a = umma_desc(p_addr + t * Cfg::P_BYTES, 128 * 16, 128, UMMA_LAYOUT_NONE);
      // MN-major, no swizzle: LBO = distance between 8-row groups of the contraction index, SBO = between 8-column groups
      const uint64_t b = umma_desc(kv_addr + stage * (Cfg::K_BYTES + Cfg::V_BYTES) + Cfg::K_BYTES, 128, BN * 16, UMMA_LAYOUT_NONE);
#pragma unroll
      for (int ks = 0; ks < BN / 16; ++ks)
        umma_bf16_ss(tmem + Cfg::TM_O + t * Cfg::DN, a + static_cast<uint64_t>(ks * (2 * 128 * 16 >> 4)),
                     b + static_cast<uint64_t>(ks * (256 >> 4)), IDESC_PV, (!first || ks > 0) ? 1u : 0u);
      umma_commit(&pv_done[t]);
    };
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int qu = u % p.nU;
      const int nt = (2 * qu + 1 < p.nQT) ? 2 : 1;
      // first key tile: S_t(0) as soon as Q_t and the tile are there and the previous unit's last S_t has been read
      mbar_wait(&kv_full[kvc % Cfg::STAGES], (kvc / Cfg::STAGES) & 1);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(&q_full[t], it[t] & 1);
        ++it[t];
        if (g[t] > 0) mbar_wait(&s_free[t], (g[t] - 1) & 1);
        tc_fence_after();
        if (elect_one()) issue_qk(t, kvc % Cfg::STAGES);
        __syncwarp();
      }
      for (int j = 0; j < nKT; ++j, ++kvc) {
        const uint32_t stage = kvc % Cfg::STAGES;
        if (j + 1 < nKT) {
          const uint32_t nstage = (kvc + 1) % Cfg::STAGES;
          mbar_wait(&kv_full[nstage], ((kvc + 1) / Cfg::STAGES) & 1);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&s_free[t], g[t] & 1);
            tc_fence_after();
            if (elect_one()) issue_qk(t, nstage);
            __syncwarp();
          }
        }
        for (int t = 0; t < nt; ++t) {
          mbar_wait(&p_ready[t], g[t] & 1);
          tc_fence_after();
          if (elect_one()) issue_pv(t, stage, j == 0);
          __syncwarp();
          ++g[t];
        }
        if (elect_one()) umma_commit(&kv_empty[stage]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax groups: one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int t = (warp - 4) >> 2;                 // tile / group
    const int wq = warp & 3;                       // TMEM lane quarter
    const int row = wq * 32 + lane;                // row inside the 128-row tile
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t s_taddr = tmem + lane_base + Cfg::TM_S + t * BN;
    const uint32_t o_taddr = tmem + lane_base + Cfg::TM_O + t * Cfg::DN;
    const uint32_t p_row = smem_u32(smem + Cfg::OFF_P + t * Cfg::P_BYTES) + row * 16;
    uint8_t* stage_base = smem + (Cfg::STAGE_IN_Q ? Cfg::OFF_Q + t * Cfg::Q_BYTES : Cfg::OFF_P + t * Cfg::P_BYTES) + wq * (32 * D * 2);
    uint32_t g = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int bh = u / p.nU, qu = u % p.nU;
      const int qt = 2 * qu + t;
      if (qt >= p.nQT) continue;
      float m_ref = 0.f;
      for (int j = 0; j < nKT; ++j, ++g) {
        mbar_wait(&s_full[t], g & 1);
        tc_fence_after();
        uint32_t s[BN];
#pragma unroll
        for (int c = 0; c < BN; c += 32) tmem_ld32_raw(s_taddr + c, s + c);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        if constexpr (Cfg::STAGE_IN_Q) {
          // the last S of the unit has been produced: Q_t is no longer read by the tensor core; the staging of the output
          // re-uses it, so it is released to the producer only after the output has left (below)
        } else {
          if (j == nKT - 1 && lane == 0) mbar_arrive(&q_free[t]);
        }
        const int valid = S - j * BN;
        if (valid < BN) {
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= valid) s[i] = 0xff800000u;    // -inf
        }
        float mx = fmax3(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
#pragma unroll
        for (int i = 3; i + 1 < BN; i += 2) mx = fmax3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mx = fmaxf(mx, __uint_as_float(s[BN - 1]));
        bool waited = false;
        if (j == 0) {
          m_ref = mx;
        } else {
          const bool need = mx > m_ref + 8.f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(&pv_done[t], (g - 1) & 1);           // O_t is quiescent
            waited = true;
            tc_fence_after();
            const float alpha = need ? fast_exp2(m_ref - mx) : 1.f;
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
        // P = 2^(S - m_ref)
        const uint64_t nm = f2_pack(-m_ref, -m_ref);
        uint32_t pk[BN / 2];
#pragma unroll
        for (int i = 0; i < BN; i += 2) {
          float a, b;
          f2_unpack(f2_add(f2_pack(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), nm), a, b);
          pk[i >> 1] = pack_bf16x2(fast_exp2(a), fast_exp2(b));
        }
        if (j > 0 && !waited) mbar_wait(&pv_done[t], (g - 1) & 1);     // P V (j-1) has finished reading P_t
#pragma unroll
        for (int c = 0; c < BN / 8; ++c) st_shared_v4_a(p_row + c * 2048, pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
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
      named_bar_sync(1 + t, 128);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
    asm volatile("setmaxnreg.dec.sync.aligned.u32 168;");
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <int D>
static int launch_sattn(const void* q, const void* k, const void* v, long long ld, void* out, void* ws, int B, int S, int C, int H,
                        cudaStream_t stream) {
  using Cfg = SaCfg<D>;
  const int nQT = (S + 127) / 128, nKT = (S + Cfg::BN - 1) / Cfg::BN, nU = (nQT + 1) / 2;
  const long long BH = static_cast<long long>(B) * H;
  uint8_t* qimg = static_cast<uint8_t*>(ws);
  uint8_t* kimg = qimg + BH * nQT * Cfg::Q_BYTES;
  uint8_t* vimg = kimg + BH * nKT * Cfg::K_BYTES;
  const float qscale = 1.4426950408889634f / sqrtf(static_cast<float>(D));
  sattn_pack_kernel<D><<<dim3(nQT, H, B), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                                           static_cast<const __nv_bfloat16*>(v), ld, qimg, kimg, vimg, S, H, nQT, nKT, qscale);
  PV_LAUNCHED();
  CUtensorMap tmO;
  if (make_tmap_3d(&tmO, out, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, D, 32, 1, Swz::None)) return PV_ERR_CUDA;
  SaParams p;
  p.qimg = qimg; p.kimg = kimg; p.vimg = vimg;
  p.S = S; p.H = H; p.nQT = nQT; p.nKT = nKT; p.nU = nU;
  const long long units = BH * nU;
  PV_REQUIRE(units < (1ll << 30), "too many work units");
  p.units = static_cast<int>(units);
  auto kern = self_attn_fwd_kernel<D>;
  PV_CUDA(set_max_smem_once(kern, Cfg::SMEM_BYTES));
  const long long sms = sm_count();
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
  if (d == 40) return launch_sattn<40>(q, k, v, ld, out, ws, B, S, C, H, stream);
  if (d == 80) return launch_sattn<80>(q, k, v, ld, out, ws, B, S, C, H, stream);
  return launch_sattn<160>(q, k, v, ld, out, ws, B, S, C, H, stream);
}

}  // namespace pv
