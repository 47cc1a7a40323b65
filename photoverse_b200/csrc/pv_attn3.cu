// photoverse_b200 -- fused Q-projection + dual-branch cross-attention, variant 3: PERSISTENT, software-pipelined.
//
// Same math as pv_attn.cu / pv_attn2.cu (reference models/attention_processor.py:297, 307-322, 400-407, 411-420):
//   O[b, rows, g*160:(g+1)*160] = sum over the heads h of group g of  P_h V_h,   Q = X Wq^T on the fly,
//   P_h = [ w_text * softmax(Q_h K_text,h^T / sqrt(d)) | w_img * softmax(Q_h K_img,h^T / sqrt(d)) ]
//
// What is different: one CTA per SM walks a contiguous range of work units (unit = 128 query rows x 160 channels =
// 4 / 2 / 1 heads of d = 40 / 80 / 160) and keeps THREE pipelines running against each other:
//
//   warp 0        TMA producer      X [128x64] + Wq [160x64] K-blocks into a 4-stage ring (runs units ahead);
//                                   packed K / V^T tiles of the unit's heads, re-loaded only when (sample, head group)
//                                   changes -- consecutive units of a CTA are row tiles of the same (b, g)
//   warp 1        projection MMA    Q(unit i+1) = X Wq^T into TMEM slot (i+1)&1   ... while ...
//   warp 2        attention MMA     S = Q_h K_h^T and O = P_h V_h of unit i (A operands Q / P live in TMEM)
//   warps 4..7    softmax group 0   even heads: segment softmax straight out of TMEM, P written back in place,
//   warps 8..11   softmax group 1   odd heads : O drained -> scaled -> bf16 -> HBM; both groups convert Q to bf16
//
// TMEM (512 columns): two 160-column slots (Q accumulator of unit i, i+1; after conversion the slot holds the packed
// bf16 Q and the O accumulator) + two 96-column S/P buffers (one per softmax group):
//            slot s = [160 s, 160 s + 160)                       S/P buffers
//   d=40     Q bf16 4 x 24 = [0,96)    O [96,144)                [320,416) [416,512)
//   d=80     Q bf16 2 x 40 = [0,80)    O [80,160)                [320,416) [416,512)
//   d=160    Q bf16 80     = [0,80)    O dims 0..79 [80,160)     [320,416) ; O dims 80..159 at [416,496)
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int A3_BM = 128;
constexpr int A3_BN = 160;
constexpr int A3_BK = 64;
constexpr int A3_KEYS = PV_KEYS_PAD;
constexpr int A3_IMG_OFF = PV_IMG_KEY_OFFSET;
constexpr int A3_WARPS = 12;       // 0 TMA, 1 projection MMA, 2 attention MMA, 3 spare, 4..7 / 8..11 softmax groups
constexpr int A3_THREADS = A3_WARPS * 32;
constexpr int A3_A_BYTES = A3_BM * A3_BK * 2;
constexpr int A3_W_BYTES = A3_BN * A3_BK * 2;
constexpr int A3_STAGE_BYTES = A3_A_BYTES + A3_W_BYTES;
constexpr int A3_STAGES = 4;

template <int D>
struct Attn3Cfg {
  static constexpr int HPC = A3_BN / D;
  static constexpr int D_PAD = (D + 15) / 16 * 16;
  static constexpr int QB_COLS = D_PAD / 2;
  static constexpr int KV_TILE_BYTES = A3_KEYS * D_PAD * 2;
  static constexpr int KV_BYTES = HPC * 2 * KV_TILE_BYTES;
  static constexpr int OFF_KV = A3_STAGES * A3_STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_KV + KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr uint32_t TM_SBUF0 = 320;
  static constexpr uint32_t TM_SBUF1 = 416;
  static constexpr uint32_t TM_O_OFF = (D == 40) ? 96 : 80;     // inside the slot
  static constexpr uint32_t TM_OHI = 416;                        // d = 160 only
};

struct Attn3Params {
  const uint8_t* Kp;
  const uint8_t* Vp;
  __nv_bfloat16* O;        // [B,S,C]
  float* stats;            // optional [B,H,S,4]
  int S, C, H, Lt, Li;
  int G, MT;               // head groups per sample (C/160), row tiles per sample
  long long units;         // B * G * MT
  float w_text, w_img, scale_log2e;
};

template <int N>
__device__ __forceinline__ void pack_pairs3(const uint32_t* v, uint32_t* out) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) out[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
}

template <int D>
__global__ void __launch_bounds__(A3_THREADS, 1)
dual_attn_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                                const Attn3Params p) {
  using Cfg = Attn3Cfg<D>;
  constexpr int HPC = Cfg::HPC;
  constexpr int D_PAD = Cfg::D_PAD;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kv = smem + Cfg::OFF_KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                        // [STAGES]  TMA -> projection MMA
  uint64_t* empty = full + A3_STAGES;           // [STAGES]
  uint64_t* kv_full = empty + A3_STAGES;        // 1
  uint64_t* kv_free = kv_full + 1;              // 1
  uint64_t* q_full = kv_free + 1;               // [2] projection MMA -> softmax groups (Q accumulator complete)
  uint64_t* q_ready = q_full + 2;               // [2] softmax groups -> attention MMA (packed bf16 Q in TMEM)
  uint64_t* slot_free = q_ready + 2;            // [2] softmax group -> projection MMA (slot fully consumed)
  uint64_t* s_full = slot_free + 2;             // [2] per softmax group
  uint64_t* p_ready = s_full + 2;               // [2]
  uint64_t* o_full = p_ready + 2;               // [2]
  uint64_t* o_free = o_full + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = p.C / A3_BK;
  const int u0 = static_cast<int>((p.units * blockIdx.x) / gridDim.x);
  const int u1 = static_cast<int>((p.units * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    for (int s = 0; s < A3_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(kv_full, 1);
    mbar_init(kv_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_ready[i], 256);
      mbar_init(&slot_free[i], 128);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // register re-distribution (warpgroup granular): the producer / issuer warps need few, the row threads many
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      int prev_bg = -1;
      uint32_t kv_gen = 0;
      for (int u = u0; u < u1; ++u) {
        const int bg = u / p.MT;
        const int mt = u - bg * p.MT;
        const int b = bg / p.G;
        const int g = bg - b * p.G;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % A3_STAGES;
          const uint32_t ph = (it / A3_STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* a_dst = smem + s * A3_STAGE_BYTES;
          mbar_expect_tx(&full[s], A3_STAGE_BYTES);
          tma_load_3d(a_dst, &tmX, &full[s], kb * A3_BK, mt * A3_BM, b);
          tma_load_3d(a_dst + A3_A_BYTES, &tmWq, &full[s], kb * A3_BK, g * A3_BN, 0);
        }
        if (bg != prev_bg) {
          // K / V^T tiles of this (sample, head group): needed only once the projection above has completed
          if (kv_gen > 0) mbar_wait(kv_free, (kv_gen - 1) & 1);
          mbar_expect_tx(kv_full, Cfg::KV_BYTES);
          for (int j = 0; j < HPC; ++j) {
            const size_t tile = (static_cast<size_t>(b) * p.H + (g * HPC + j)) * Cfg::KV_TILE_BYTES;
            bulk_load_1d(kv + (2 * j) * Cfg::KV_TILE_BYTES, p.Kp + tile, Cfg::KV_TILE_BYTES, kv_full);
            bulk_load_1d(kv + (2 * j + 1) * Cfg::KV_TILE_BYTES, p.Vp + tile, Cfg::KV_TILE_BYTES, kv_full);
          }
          ++kv_gen;
          prev_bg = bg;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== projection MMA issuer: Q(unit) = X Wq^T  (M=128, N=160, K=C) =====================
    constexpr uint32_t idesc_q = umma_idesc_bf16(A3_BM, A3_BN);
    uint32_t it = 0;
    int i = 0;
    for (int u = u0; u < u1; ++u, ++i) {
      const int slot = i & 1;
      if (i >= 2) mbar_wait(&slot_free[slot], ((i >> 1) - 1) & 1);
      tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % A3_STAGES;
        const uint32_t ph = (it / A3_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint8_t* a_src = smem + s * A3_STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(a_src);
          const uint64_t dw = umma_desc_sw128(a_src + A3_A_BYTES);
#pragma unroll
          for (int k = 0; k < A3_BK / 16; ++k)
            umma_bf16_ss(tmem + slot * A3_BN, da + 2 * k, dw + 2 * k, idesc_q, (kb | k) != 0);
          umma_commit(&empty[s]);
          if (kb == kblocks - 1) umma_commit(&q_full[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===================== attention MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(A3_BM, A3_KEYS);
    constexpr uint32_t idesc_o = umma_idesc_bf16(A3_BM, (D == 160) ? 80 : D_PAD);
    uint32_t n = 0;                 // global head counter of this CTA
    uint32_t kv_gen = 0;
    int prev_bg = -1;
    int i = 0;
    for (int u = u0; u < u1; ++u, ++i) {
      const int slot = i & 1;
      const uint32_t tslot = tmem + slot * A3_BN;
      const int bg = u / p.MT;
      mbar_wait(&q_ready[slot], (i >> 1) & 1);
      if (bg != prev_bg) {
        mbar_wait(kv_full, kv_gen & 1);
        ++kv_gen;
        prev_bg = bg;
      }
      tc_fence_after();
      auto issue_qk = [&](int j, uint32_t nn) {
        const uint32_t sbuf = tmem + ((D != 160 && (nn & 1)) ? Cfg::TM_SBUF1 : Cfg::TM_SBUF0);
        const uint32_t k_tile = smem_u32(kv + (2 * j) * Cfg::KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D_PAD / 16; ++k) {
          const uint64_t db = umma_desc(k_tile + k * 2 * (A3_KEYS * 16), A3_KEYS * 16, 128, UMMA_LAYOUT_NONE);
          umma_bf16_ts(sbuf, tslot + j * Cfg::QB_COLS + k * 8, db, idesc_s, k != 0);
        }
        umma_commit(&s_full[nn & 1]);
      };
      if (lane == 0) issue_qk(0, n);
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < HPC; ++j) {
        const uint32_t nn = n + j;
        const uint32_t w = nn & 1;
        if (j + 1 < HPC) {
          // the S/P buffer of head nn+1 was last used by head nn-1, whose PV has already been issued (in order)
          if (lane == 0) issue_qk(j + 1, nn + 1);
          __syncwarp();
        }
        mbar_wait(&p_ready[w], (nn >> 1) & 1);
        if (nn >= 1) mbar_wait(&o_free[(nn - 1) & 1], ((nn - 1) >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sbuf = tmem + ((D != 160 && w) ? Cfg::TM_SBUF1 : Cfg::TM_SBUF0);
          const uint32_t v_tile = smem_u32(kv + (2 * j + 1) * Cfg::KV_TILE_BYTES);
          if constexpr (D == 160) {
#pragma unroll
            for (int k = 0; k < A3_KEYS / 16; ++k) {
              const uint64_t db = umma_desc(v_tile + k * 2 * (D_PAD * 16), D_PAD * 16, 128, UMMA_LAYOUT_NONE);
              umma_bf16_ts(tslot + Cfg::TM_O_OFF, sbuf + k * 8, db, idesc_o, k != 0);
            }
#pragma unroll
            for (int k = 0; k < A3_KEYS / 16; ++k) {
              const uint64_t db = umma_desc(v_tile + 80 * 16 + k * 2 * (D_PAD * 16), D_PAD * 16, 128, UMMA_LAYOUT_NONE);
              umma_bf16_ts(tmem + Cfg::TM_OHI, sbuf + k * 8, db, idesc_o, k != 0);
            }
          } else {
#pragma unroll
            for (int k = 0; k < A3_KEYS / 16; ++k) {
              const uint64_t db = umma_desc(v_tile + k * 2 * (D_PAD * 16), D_PAD * 16, 128, UMMA_LAYOUT_NONE);
              umma_bf16_ts(tslot + Cfg::TM_O_OFF, sbuf + k * 8, db, idesc_o, k != 0);
            }
          }
          umma_commit(&o_full[w]);
          if (j == HPC - 1 && u + 1 < u1 && (u + 1) / p.MT != bg) umma_commit(kv_free);
        }
        __syncwarp();
      }
      n += HPC;
    }
  }
  } else {
    // ===================== softmax groups (warps 4..7 and 8..11): one thread per query row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const int Lt = p.Lt;
    const int Li = p.Li;
    const float cs = p.scale_log2e;
    uint32_t n = 0;
    int i = 0;
    for (int u = u0; u < u1; ++u, ++i) {
      const int slot = i & 1;
      const uint32_t tslot = tlane + slot * A3_BN;
      const int bg = u / p.MT;
      const int mt = u - bg * p.MT;
      const int b = bg / p.G;
      const int g = bg - b * p.G;
      const int m0 = mt * A3_BM;

      // ---- Q: fp32 accumulator -> packed bf16 in place.  Group wg converts columns [80 wg, 80 wg + 80). ----
      mbar_wait(&q_full[slot], (i >> 1) & 1);
      tc_fence_after();
      {
        uint32_t a[32], b2[32], c16[16], o[48];
        tmem_ld_x32(tslot + wg * 80, a);
        tmem_ld_x32(tslot + wg * 80 + 32, b2);
        tmem_ld_x16(tslot + wg * 80 + 64, c16);
        tmem_ld_wait();
        named_bar_sync(1, 256);                     // every fp32 column has been read before any is overwritten
        if constexpr (D == 40) {
          // two heads: 20 data words + 4 zero words each (dims 40..47 pad the K = 48 contraction)
          uint32_t t[40];
          pack_pairs3<32>(a, t);
          pack_pairs3<32>(b2, t + 16);
          pack_pairs3<16>(c16, t + 32);
#pragma unroll
          for (int k = 0; k < 20; ++k) { o[k] = t[k]; o[24 + k] = t[20 + k]; }
#pragma unroll
          for (int k = 20; k < 24; ++k) { o[k] = 0u; o[24 + k] = 0u; }
          tmem_st_x16(tslot + wg * 48, o);
          tmem_st_x16(tslot + wg * 48 + 16, o + 16);
          tmem_st_x16(tslot + wg * 48 + 32, o + 32);
        } else {
          pack_pairs3<32>(a, o);
          pack_pairs3<32>(b2, o + 16);
          pack_pairs3<16>(c16, o + 32);
          tmem_st_x16(tslot + wg * 40, o);
          tmem_st_x16(tslot + wg * 40 + 16, o + 16);
          tmem_st_x8(tslot + wg * 40 + 32, o + 32);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_ready[slot]);

      const bool row_ok = (m0 + row) < p.S;
      __nv_bfloat16* orow = p.O + (static_cast<size_t>(b) * p.S + (m0 + row)) * p.C + g * A3_BN;

#pragma unroll 1
      for (int j = 0; j < HPC; ++j) {
        const uint32_t nn = n + j;
        if (static_cast<int>(nn & 1) != wg) continue;
        const uint32_t par = (nn >> 1) & 1;
        const uint32_t sbuf = tlane + ((D != 160 && wg) ? Cfg::TM_SBUF1 : Cfg::TM_SBUF0);
        mbar_wait(&s_full[wg], par);
        tc_fence_after();
        float s[A3_KEYS];
#pragma unroll
        for (int c = 0; c < A3_KEYS / 32; ++c) {
          uint32_t v[32];
          tmem_ld_x32(sbuf + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) s[c * 32 + k] = __uint_as_float(v[k]);
        }
        if (Lt < 64) {
#pragma unroll
          for (int k = 0; k < 64; ++k) s[k] = (k < Lt) ? s[k] : -INFINITY;
        }
#pragma unroll
        for (int k = 64; k < A3_IMG_OFF; ++k) s[k] = (k < Lt) ? s[k] : -INFINITY;
#pragma unroll
        for (int k = A3_IMG_OFF; k < A3_KEYS; ++k) s[k] = (k - A3_IMG_OFF < Li) ? s[k] : -INFINITY;
        float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
        for (int k = 4; k < A3_IMG_OFF; ++k) m4[k & 3] = fmaxf(m4[k & 3], s[k]);
        const float mt_ = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        float i4[4] = {s[A3_IMG_OFF], s[A3_IMG_OFF + 1], s[A3_IMG_OFF + 2], s[A3_IMG_OFF + 3]};
#pragma unroll
        for (int k = A3_IMG_OFF + 4; k < A3_KEYS; ++k) i4[k & 3] = fmaxf(i4[k & 3], s[k]);
        const float mi_ = fmaxf(fmaxf(i4[0], i4[1]), fmaxf(i4[2], i4[3]));
        const float mts = mt_ * cs, mis = mi_ * cs;
        float l4[4] = {0.f, 0.f, 0.f, 0.f}, li4[2] = {0.f, 0.f};
#pragma unroll
        for (int kc = 0; kc < A3_IMG_OFF / 8; ++kc) {
          if (kc * 8 < Lt) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float e = fast_exp2(fmaf(s[kc * 8 + k], cs, -mts));
              s[kc * 8 + k] = e;
              l4[k & 3] += e;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) s[kc * 8 + k] = 0.f;
          }
        }
#pragma unroll
        for (int kc = A3_IMG_OFF / 8; kc < A3_KEYS / 8; ++kc) {
          if ((kc - A3_IMG_OFF / 8) * 8 < Li) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float e = fast_exp2(fmaf(s[kc * 8 + k], cs, -mis));
              s[kc * 8 + k] = e;
              li4[k & 1] += e;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) s[kc * 8 + k] = 0.f;
          }
        }
        const float lt = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        const float li = li4[0] + li4[1];
        const float at = p.w_text / lt;
        const float ai = p.w_img / li;
        if (p.stats != nullptr && row_ok) {
          const size_t idx = ((static_cast<size_t>(b) * p.H + (g * HPC + j)) * p.S + (m0 + row));
          reinterpret_cast<float4*>(p.stats)[idx] = make_float4(mts, lt, mis, li);
        }
        float ft, fi, oscale;
        if (p.w_text != 0.f) { ft = 1.f; fi = ai / at; oscale = at; }
        else                 { ft = 0.f; fi = 1.f;     oscale = ai; }
        uint32_t pk[A3_KEYS / 2];
#pragma unroll
        for (int k = 0; k < A3_IMG_OFF / 2; ++k) pk[k] = (ft == 1.f) ? pack_bf16x2(s[2 * k], s[2 * k + 1]) : 0u;
#pragma unroll
        for (int k = A3_IMG_OFF / 2; k < A3_KEYS / 2; ++k) pk[k] = pack_bf16x2(s[2 * k] * fi, s[2 * k + 1] * fi);
        tmem_st_x16(sbuf, pk);
        tmem_st_x16(sbuf + 16, pk + 16);
        tmem_st_x16(sbuf + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[wg]);

        // ---- O_j: TMEM -> registers -> * row scale -> bf16 -> HBM (this thread's row, D*2 contiguous bytes) ----
        mbar_wait(&o_full[wg], par);
        tc_fence_after();
        __nv_bfloat16* dst = orow + j * D;
        auto emit = [&](const uint32_t* v, int col0, int ncols) {
          if (!row_ok) return;
#pragma unroll
          for (int c = 0; c < ncols / 8; ++c) {
            const uint32_t* w8 = v + c * 8;
            st_global_v4(dst + col0 + c * 8,
                         pack_bf16x2(__uint_as_float(w8[0]) * oscale, __uint_as_float(w8[1]) * oscale),
                         pack_bf16x2(__uint_as_float(w8[2]) * oscale, __uint_as_float(w8[3]) * oscale),
                         pack_bf16x2(__uint_as_float(w8[4]) * oscale, __uint_as_float(w8[5]) * oscale),
                         pack_bf16x2(__uint_as_float(w8[6]) * oscale, __uint_as_float(w8[7]) * oscale));
          }
        };
        const uint32_t o_addr = tslot + Cfg::TM_O_OFF;
        if constexpr (D == 40) {
          uint32_t a[32], c8[8];
          tmem_ld_x32(o_addr, a);
          tmem_ld_x8(o_addr + 32, c8);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&o_free[wg]);
          if (j == HPC - 1) mbar_arrive(&slot_free[slot]);
          emit(a, 0, 32);
          emit(c8, 32, 8);
        } else if constexpr (D == 80) {
          uint32_t a[32], b2[32], c16[16];
          tmem_ld_x32(o_addr, a);
          tmem_ld_x32(o_addr + 32, b2);
          tmem_ld_x16(o_addr + 64, c16);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&o_free[wg]);
          if (j == HPC - 1) mbar_arrive(&slot_free[slot]);
          emit(a, 0, 32);
          emit(b2, 32, 32);
          emit(c16, 64, 16);
        } else {
          const uint32_t ohi = tlane + Cfg::TM_OHI;
          {
            uint32_t a[32], b2[32], c16[16];
            tmem_ld_x32(o_addr, a);
            tmem_ld_x32(o_addr + 32, b2);
            tmem_ld_x16(o_addr + 64, c16);
            tmem_ld_wait();
            emit(a, 0, 32);
            emit(b2, 32, 32);
            emit(c16, 64, 16);
          }
          {
            uint32_t a[32], b2[32], c16[16];
            tmem_ld_x32(ohi, a);
            tmem_ld_x32(ohi + 32, b2);
            tmem_ld_x16(ohi + 64, c16);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_free[wg]);
            mbar_arrive(&slot_free[slot]);
            emit(a, 80, 32);
            emit(b2, 112, 32);
            emit(c16, 144, 16);
          }
        }
      }
      n += HPC;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <int D>
static int launch_attn3(const CUtensorMap& tmX, const CUtensorMap& tmWq, const Attn3Params& p, cudaStream_t stream) {
  using Cfg = Attn3Cfg<D>;
  auto kern = dual_attn_fwd_persistent_kernel<D>;
  static bool attr_done = false;
  if (!attr_done) {
    PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const long long sms = sm_count();
  const int grid = static_cast<int>(p.units < sms ? p.units : sms);
  kern<<<grid, A3_THREADS, Cfg::SMEM_BYTES, stream>>>(tmX, tmWq, p);
  PV_LAUNCHED();
  return PV_OK;
}

int dual_attn_core_bf16_persistent(const void* X, const void* Wq, const void* Kp, const void* Vp, void* O, float* stats,
                                   int B, int S, int C, int H, int Lt, int Li, float w_text, float w_img,
                                   cudaStream_t stream) {
  PV_REQUIRE(B > 0 && S > 0 && H > 0 && C % H == 0, "bad shape B=%d S=%d C=%d H=%d", B, S, C, H);
  const int d = C / H;
  PV_REQUIRE(d == 40 || d == 80 || d == 160, "head_dim %d unsupported (40/80/160)", d);
  PV_REQUIRE(C % A3_BN == 0 && C % A3_BK == 0, "C=%d must be a multiple of 320", C);
  PV_REQUIRE(Lt >= 1 && Lt <= A3_IMG_OFF && Li >= 1 && Li <= A3_KEYS - A3_IMG_OFF,
             "need 1 <= Lt <= %d and 1 <= Li <= %d (Lt=%d Li=%d)", A3_IMG_OFF, A3_KEYS - A3_IMG_OFF, Lt, Li);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(Kp) |
              reinterpret_cast<uintptr_t>(Vp) | reinterpret_cast<uintptr_t>(O)) % 16 == 0, "pointers must be 16-byte aligned");
  CUtensorMap tmX, tmWq;
  if (make_tmap_3d(&tmX, X, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, A3_BK, A3_BM, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmWq, Wq, 2, C, C, 1, C * 2ull, static_cast<uint64_t>(C) * C * 2, A3_BK, A3_BN, 1, Swz::B128)) return PV_ERR_CUDA;
  Attn3Params p;
  p.Kp = static_cast<const uint8_t*>(Kp);
  p.Vp = static_cast<const uint8_t*>(Vp);
  p.O = static_cast<__nv_bfloat16*>(O);
  p.stats = stats;
  p.S = S; p.C = C; p.H = H; p.Lt = Lt; p.Li = Li;
  p.G = C / A3_BN;
  p.MT = (S + A3_BM - 1) / A3_BM;
  p.units = static_cast<long long>(B) * p.G * p.MT;
  PV_REQUIRE(p.units < (1ll << 30), "too many work units");
  p.w_text = w_text; p.w_img = w_img;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(d));
  switch (d) {
    case 40: return launch_attn3<40>(tmX, tmWq, p, stream);
    case 80: return launch_attn3<80>(tmX, tmWq, p, stream);
    default: return launch_attn3<160>(tmX, tmWq, p, stream);
  }
}

}  // namespace pv
