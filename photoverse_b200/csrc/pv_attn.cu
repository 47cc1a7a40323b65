// photoverse_b200 -- fused Q-projection + dual-branch cross-attention for sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces, in ONE kernel, reference models/attention_processor.py:297 (to_q), :307-322 (text SDPA),
// :400-407 (image SDPA) and :411-420 (branch fusion):
//
//   O[b, rows, g*160 : (g+1)*160] = sum over the heads h of group g of
//        P_h V_h ,   P_h = [ w_text * softmax(Q_h K_text,h^T / sqrt(d)) | w_img * softmax(Q_h K_img,h^T / sqrt(d)) ]
//   with Q = X Wq^T computed on the fly (never written to HBM).
//
// Work decomposition: CTA = (128 query rows of one sample) x (one group of heads covering 160 channels:
// 4 heads of d=40, 2 heads of d=80 or 1 head of d=160 -- the three SD-1.5 attn2 families).
//
//   warp 0      TMA producer: X tile [128x64] + Wq rows [160x64] per K-block (128B swizzle, 4-stage ring);
//               packed K / V^T tiles of the group's heads by 1-D bulk copies (pre-laid-out UMMA images)
//   warp 1      MMA issuer (one thread): Q-proj (N=160) -> TMEM; per head QK^T (N=96 keys) -> TMEM S,
//               PV (N=d_pad, K=96) -> TMEM O; software-pipelined across heads with double-buffered S/P/O
//   warps 2..5  one thread per query row: drain Q (fp32 -> bf16 core-matrix tiles in smem), per head the
//               segment softmax straight out of TMEM (all <=96 keys are resident: no online rescaling),
//               branch weights and 1/rowsum folded into P (bf16, smem), drain O -> bf16 -> TMA store.
//
// Key layout inside the 96 padded key slots: [0,Lt) text (Lt <= 80), [80,80+Li) image (Li <= 16), the rest is
// zero padding that the softmax masks to -inf (segment membership of a column is static).
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int AT_BM = 128;          // query rows per CTA
constexpr int AT_BN = 160;          // channels (heads * head_dim) per CTA
constexpr int AT_BK = 64;
constexpr int AT_KEYS = PV_KEYS_PAD;  // 96
constexpr int AT_IMG_OFF = PV_IMG_KEY_OFFSET;  // 80: first image-key slot
constexpr int AT_STAGES = 4;
constexpr int AT_THREADS = 192;
constexpr int AT_A_BYTES = AT_BM * AT_BK * 2;              // 16384
constexpr int AT_W_BYTES = AT_BN * AT_BK * 2;              // 20480
constexpr int AT_STAGE_BYTES = AT_A_BYTES + AT_W_BYTES;    // 36864
constexpr int AT_RING_BYTES = AT_STAGES * AT_STAGE_BYTES;  // 147456

template <int D>
struct AttnCfg {
  static constexpr int HPC = AT_BN / D;                     // heads per CTA
  static constexpr int D_PAD = (D + 15) / 16 * 16;          // 48 / 80 / 160
  static constexpr int NKC = D_PAD / 8;                     // 16-byte K-chunks per Q/K row
  static constexpr int DCH = D / 8;                         // valid chunks per head
  static constexpr int KV_TILE_BYTES = AT_KEYS * D_PAD * 2; // one K tile or one V^T tile
  static constexpr int KV_BYTES = HPC * 2 * KV_TILE_BYTES;
  static constexpr int Q_TILE_BYTES = AT_BM * D_PAD * 2;    // per head
  static constexpr int P_TILE_BYTES = AT_BM * AT_KEYS * 2;  // 24576
  static constexpr int NPBUF = HPC > 1 ? 2 : 1;
  static constexpr int OST_BYTES = AT_BM * AT_BN * 2;       // 40960 (4 warps x [32 x 160] bf16)
  // after the projection main loop the stage ring is dead and is re-used for Q / P / O staging tiles
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_P = OFF_Q + HPC * Q_TILE_BYTES;
  static constexpr int OFF_OST = OFF_P + NPBUF * P_TILE_BYTES;
  static_assert(OFF_OST + OST_BYTES <= AT_RING_BYTES, "epilogue tiles must fit in the dead stage ring");
  static constexpr int OFF_KV = AT_RING_BYTES;
  static constexpr int OFF_BAR = OFF_KV + KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // TMEM columns
  static constexpr uint32_t TM_Q = 0;
  static constexpr uint32_t TM_S0 = 160;
  static constexpr uint32_t TM_S1 = 256;
  static constexpr uint32_t TM_O0 = 352;
  static constexpr uint32_t TM_O1 = 352 + D_PAD;
  static_assert(HPC == 1 || TM_O1 + D_PAD <= 512, "TMEM budget");
  static_assert(TM_O0 + D_PAD <= 512, "TMEM budget");
};

struct AttnParams {
  const uint8_t* Kp;     // [B][H] packed K tiles
  const uint8_t* Vp;     // [B][H] packed V^T tiles
  float* stats;          // optional [B][H][S][4]
  int S, H, Lt, Li;
  float w_text, w_img;
  float scale_log2e;     // (1/sqrt(d)) * log2(e)
};

// one thread = one query row: convert 8 fp32 accumulator columns to one 16-byte bf16 chunk
__device__ __forceinline__ void store_chunk8(uint8_t* dst, const uint32_t* v) {
  st_shared_v4(dst, pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1])),
               pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3])),
               pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5])),
               pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7])));
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
dual_attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                             const __grid_constant__ CUtensorMap tmO, const AttnParams p, int C) {
  using Cfg = AttnCfg<D>;
  constexpr int HPC = Cfg::HPC;
  constexpr int D_PAD = Cfg::D_PAD;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kv = smem + Cfg::OFF_KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                       // [AT_STAGES]
  uint64_t* empty = full + AT_STAGES;          // [AT_STAGES]
  uint64_t* kv_full = empty + AT_STAGES;       // 1
  uint64_t* q_full = kv_full + 1;              // 1  (MMA -> row threads: Q accumulator complete)
  uint64_t* q_ready = q_full + 1;              // 1  (row threads -> MMA: bf16 Q tiles in smem)
  uint64_t* s_full = q_ready + 1;              // [2]
  uint64_t* p_ready = s_full + 2;              // [2]
  uint64_t* o_full = p_ready + 2;              // [2]
  uint64_t* o_free = o_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x;                    // head group
  const int m0 = blockIdx.y * AT_BM;           // first query row inside the sample
  const int b = blockIdx.z;
  const int kblocks = C / AT_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < AT_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // K / V^T tiles of this CTA's heads: contiguous pre-packed images, one bulk copy each
      mbar_expect_tx(kv_full, Cfg::KV_BYTES);
      for (int j = 0; j < HPC; ++j) {
        const size_t tile = (static_cast<size_t>(b) * p.H + (g * HPC + j)) * Cfg::KV_TILE_BYTES;
        bulk_load_1d(kv + (2 * j) * Cfg::KV_TILE_BYTES, p.Kp + tile, Cfg::KV_TILE_BYTES, kv_full);
        bulk_load_1d(kv + (2 * j + 1) * Cfg::KV_TILE_BYTES, p.Vp + tile, Cfg::KV_TILE_BYTES, kv_full);
      }
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % AT_STAGES;
        const uint32_t ph = (kb / AT_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* a_dst = smem + s * AT_STAGE_BYTES;
        mbar_expect_tx(&full[s], AT_STAGE_BYTES);
        tma_load_3d(a_dst, &tmX, &full[s], kb * AT_BK, m0, b);
        tma_load_3d(a_dst + AT_A_BYTES, &tmWq, &full[s], kb * AT_BK, g * AT_BN, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ---- Q = X Wq^T  (M=128, N=160, K=C) ----
    {
      constexpr uint32_t idesc_q = umma_idesc_bf16(AT_BM, AT_BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % AT_STAGES;
        const uint32_t ph = (kb / AT_STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint8_t* a_src = smem + s * AT_STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(a_src);
          const uint64_t dw = umma_desc_sw128(a_src + AT_A_BYTES);
#pragma unroll
          for (int k = 0; k < AT_BK / 16; ++k)
            umma_bf16_ss(tmem + Cfg::TM_Q, da + 2 * k, dw + 2 * k, idesc_q, (kb | k) != 0);
          umma_commit(&empty[s]);
          if (kb == kblocks - 1) umma_commit(q_full);
        }
        __syncwarp();
      }
    }
    // ---- attention over the group's heads ----
    mbar_wait(kv_full, 0);
    mbar_wait(q_ready, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_KEYS);   // S = Q_h K_h^T : N = 96 keys
    constexpr uint32_t idesc_o = umma_idesc_bf16(AT_BM, D_PAD);     // O = P V_h     : N = d_pad
    // no-swizzle K-major core-matrix tiles: LBO = rows*16 (next 16-byte K chunk), SBO = 128 (next 8 rows)
    auto issue_qk = [&](int j) {
      const uint32_t q_tile = smem_u32(smem + Cfg::OFF_Q + j * Cfg::Q_TILE_BYTES);
      const uint32_t k_tile = smem_u32(kv + (2 * j) * Cfg::KV_TILE_BYTES);
      const uint32_t s_col = (j & 1) ? Cfg::TM_S1 : Cfg::TM_S0;
#pragma unroll
      for (int k = 0; k < D_PAD / 16; ++k) {
        const uint64_t da = umma_desc(q_tile + k * 2 * (AT_BM * 16), AT_BM * 16, 128, UMMA_LAYOUT_NONE);
        const uint64_t db = umma_desc(k_tile + k * 2 * (AT_KEYS * 16), AT_KEYS * 16, 128, UMMA_LAYOUT_NONE);
        umma_bf16_ss(tmem + s_col, da, db, idesc_s, k != 0);
      }
      umma_commit(&s_full[j & 1]);
    };
    auto issue_pv = [&](int j) {
      const uint32_t p_tile = smem_u32(smem + Cfg::OFF_P + (j % Cfg::NPBUF) * Cfg::P_TILE_BYTES);
      const uint32_t v_tile = smem_u32(kv + (2 * j + 1) * Cfg::KV_TILE_BYTES);
      const uint32_t o_col = (j & 1) ? Cfg::TM_O1 : Cfg::TM_O0;
#pragma unroll
      for (int k = 0; k < AT_KEYS / 16; ++k) {
        const uint64_t da = umma_desc(p_tile + k * 2 * (AT_BM * 16), AT_BM * 16, 128, UMMA_LAYOUT_NONE);
        const uint64_t db = umma_desc(v_tile + k * 2 * (D_PAD * 16), D_PAD * 16, 128, UMMA_LAYOUT_NONE);
        umma_bf16_ss(tmem + o_col, da, db, idesc_o, k != 0);
      }
      umma_commit(&o_full[j & 1]);
    };
    if (lane == 0) issue_qk(0);
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < HPC; ++j) {
      if (j + 1 < HPC) {
        // S[(j+1)&1] was last read for head j-1; its p_ready was already observed below
        if (lane == 0) issue_qk(j + 1);
        __syncwarp();
      }
      mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
      if (j >= 2) mbar_wait(&o_free[j & 1], ((j - 2) >> 1) & 1);   // O buffer drained by the row threads
      tc_fence_after();
      if (lane == 0) issue_pv(j);
      __syncwarp();
    }
  } else {
    // ===================== row threads (warps 2..5): one thread per query row =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;                       // row inside the 128-row tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;

    // ---- drain Q: TMEM fp32 -> bf16 K-major core-matrix tiles (chunk kc of row r at kc*2048 + r*16) ----
    mbar_wait(q_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < AT_BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_addr + Cfg::TM_Q + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int gc = c * 4 + ch;                       // global 8-column chunk 0..19
        const int j = gc / Cfg::DCH;
        const int kc = gc % Cfg::DCH;
        store_chunk8(smem + Cfg::OFF_Q + j * Cfg::Q_TILE_BYTES + kc * (AT_BM * 16) + row * 16, &v[ch * 8]);
      }
    }
    if constexpr (Cfg::NKC > Cfg::DCH) {                 // d=40: zero the padding chunk (columns 40..47)
#pragma unroll
      for (int j = 0; j < HPC; ++j)
        for (int kc = Cfg::DCH; kc < Cfg::NKC; ++kc)
          st_shared_v4(smem + Cfg::OFF_Q + j * Cfg::Q_TILE_BYTES + kc * (AT_BM * 16) + row * 16, 0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(q_ready);

    uint8_t* ost = smem + Cfg::OFF_OST + q * (32 * AT_BN * 2);   // this warp's [32 x 160] bf16 staging
    const int Lt = p.Lt;
    const int Li = p.Li;
    const float cs = p.scale_log2e;

    // O_j: TMEM fp32 -> * row scale -> bf16 -> staging.  TMEM is read in the widest shapes that tile D.
    auto drain_o = [&](int j, float oscale) {
      mbar_wait(&o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t o_addr = tmem + lane_addr + ((j & 1) ? Cfg::TM_O1 : Cfg::TM_O0);
      uint8_t* dst = ost + lane * (AT_BN * 2) + j * (D * 2);
      auto emit = [&](const uint32_t* v, int col0, int n) {
#pragma unroll
        for (int c = 0; c < n / 8; ++c) {
          const uint32_t* w = v + c * 8;
          st_shared_v4(dst + (col0 / 8 + c) * 16,
                       pack_bf16x2(__uint_as_float(w[0]) * oscale, __uint_as_float(w[1]) * oscale),
                       pack_bf16x2(__uint_as_float(w[2]) * oscale, __uint_as_float(w[3]) * oscale),
                       pack_bf16x2(__uint_as_float(w[4]) * oscale, __uint_as_float(w[5]) * oscale),
                       pack_bf16x2(__uint_as_float(w[6]) * oscale, __uint_as_float(w[7]) * oscale));
        }
      };
      if constexpr (D == 40) {
        uint32_t a[32], c8[8];
        tmem_ld_x32(o_addr, a);
        tmem_ld_x8(o_addr + 32, c8);
        tmem_ld_wait();
        emit(a, 0, 32);
        emit(c8, 32, 8);
      } else if constexpr (D == 80) {
        uint32_t a[32], b2[32], c16[16];
        tmem_ld_x32(o_addr, a);
        tmem_ld_x32(o_addr + 32, b2);
        tmem_ld_x16(o_addr + 64, c16);
        tmem_ld_wait();
        emit(a, 0, 32);
        emit(b2, 32, 32);
        emit(c16, 64, 16);
      } else {
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t a[32];
          tmem_ld_x32(o_addr + c * 32, a);
          tmem_ld_wait();
          emit(a, c * 32, 32);
        }
      }
      tc_fence_before();
      mbar_arrive(&o_free[j & 1]);
    };

    float oscale_prev = 0.f;
#pragma unroll 1
    for (int j = 0; j < HPC; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_col = (j & 1) ? Cfg::TM_S1 : Cfg::TM_S0;
      float s[AT_KEYS];
#pragma unroll
      for (int c = 0; c < AT_KEYS / 32; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_addr + s_col + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(v[i]);
      }
      // Static segments: columns [0,80) are the text keys, [80,96) the image keys; padding slots -> -inf.
      if (Lt < 64) {                                   // uncommon short prompts (warp-uniform)
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] = (k < Lt) ? s[k] : -INFINITY;
      }
#pragma unroll
      for (int k = 64; k < AT_IMG_OFF; ++k) s[k] = (k < Lt) ? s[k] : -INFINITY;
#pragma unroll
      for (int k = AT_IMG_OFF; k < AT_KEYS; ++k) s[k] = (k - AT_IMG_OFF < Li) ? s[k] : -INFINITY;
      // segment maxima (4-way trees for ILP)
      float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int k = 4; k < AT_IMG_OFF; ++k) m4[k & 3] = fmaxf(m4[k & 3], s[k]);
      const float mt = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float i4[4] = {s[AT_IMG_OFF], s[AT_IMG_OFF + 1], s[AT_IMG_OFF + 2], s[AT_IMG_OFF + 3]};
#pragma unroll
      for (int k = AT_IMG_OFF + 4; k < AT_KEYS; ++k) i4[k & 3] = fmaxf(i4[k & 3], s[k]);
      const float mi = fmaxf(fmaxf(i4[0], i4[1]), fmaxf(i4[2], i4[3]));
      const float mts = mt * cs, mis = mi * cs;
      // exponentials: one FFMA + one MUFU.EX2 per key; 8-key chunks that are entirely padding are skipped
      float l4[4] = {0.f, 0.f, 0.f, 0.f}, li4[2] = {0.f, 0.f};
#pragma unroll
      for (int kc = 0; kc < AT_IMG_OFF / 8; ++kc) {
        if (kc * 8 < Lt) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float e = fast_exp2(fmaf(s[kc * 8 + i], cs, -mts));
            s[kc * 8 + i] = e;
            l4[i & 3] += e;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) s[kc * 8 + i] = 0.f;
        }
      }
#pragma unroll
      for (int kc = AT_IMG_OFF / 8; kc < AT_KEYS / 8; ++kc) {
        if ((kc - AT_IMG_OFF / 8) * 8 < Li) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float e = fast_exp2(fmaf(s[kc * 8 + i], cs, -mis));
            s[kc * 8 + i] = e;
            li4[i & 1] += e;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) s[kc * 8 + i] = 0.f;
        }
      }
      const float lt = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      const float li = li4[0] + li4[1];
      const float at = p.w_text / lt;                  // branch weight / row sum, text segment
      const float ai = p.w_img / li;                   // image segment (Li >= 1 is enforced by the host)
      if (p.stats != nullptr && m0 + row < p.S) {
        const size_t idx = ((static_cast<size_t>(b) * p.H + (g * HPC + j)) * p.S + (m0 + row));
        reinterpret_cast<float4*>(p.stats)[idx] = make_float4(mts, lt, mis, li);
      }
      // P = [e_text * ft | e_img * fi], O row scaled by `oscale` afterwards: the dominant segment stays unscaled
      // (ft or fi == 1), saving one multiply per key.  w_text == 0 (image-only fusion branch) flips the roles.
      float ft, fi, oscale;
      if (p.w_text != 0.f) { ft = 1.f; fi = ai / at; oscale = at; }
      else                 { ft = 0.f; fi = 1.f;     oscale = ai; }
      // P tile (bf16, K-major core matrices): chunk kc (8 keys) of row r at kc*2048 + r*16
      uint8_t* ptile = smem + Cfg::OFF_P + (j % Cfg::NPBUF) * Cfg::P_TILE_BYTES + row * 16;
      if (ft == 1.f) {
#pragma unroll
        for (int kc = 0; kc < AT_IMG_OFF / 8; ++kc) {
          const float* e = &s[kc * 8];
          st_shared_v4(ptile + kc * (AT_BM * 16), pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]),
                       pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        }
      } else {
#pragma unroll
        for (int kc = 0; kc < AT_IMG_OFF / 8; ++kc) st_shared_v4(ptile + kc * (AT_BM * 16), 0, 0, 0, 0);
      }
#pragma unroll
      for (int kc = AT_IMG_OFF / 8; kc < AT_KEYS / 8; ++kc) {
        const float* e = &s[kc * 8];
        st_shared_v4(ptile + kc * (AT_BM * 16), pack_bf16x2(e[0] * fi, e[1] * fi), pack_bf16x2(e[2] * fi, e[3] * fi),
                     pack_bf16x2(e[4] * fi, e[5] * fi), pack_bf16x2(e[6] * fi, e[7] * fi));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_ready[j & 1]);
      if (j >= 1) drain_o(j - 1, oscale_prev);
      oscale_prev = oscale;
    }
    drain_o(HPC - 1, oscale_prev);

    // ---- O tile -> HBM: each warp stores its own [32 rows x 160 channels] slab; TMA clips rows >= S ----
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(&tmO, ost, g * AT_BN, m0 + q * 32, b);
      bulk_commit();
      bulk_wait_read<0>();
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int D>
static int launch_attn(const CUtensorMap& tmX, const CUtensorMap& tmWq, const CUtensorMap& tmO, const AttnParams& p,
                       int B, int S, int C, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  auto kern = dual_attn_fwd_tcgen05_kernel<D>;
  static bool attr_done = false;
  if (!attr_done) {
    PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(C / AT_BN, (S + AT_BM - 1) / AT_BM, B);
  kern<<<grid, AT_THREADS, Cfg::SMEM_BYTES, stream>>>(tmX, tmWq, tmO, p, C);
  PV_LAUNCHED();
  return PV_OK;
}

int64_t attn_kv_tile_bytes(int d) { return static_cast<int64_t>(AT_KEYS) * ((d + 15) / 16 * 16) * 2; }

// X:[B,S,C] bf16, Wq:[C,C] bf16, Kp/Vp packed tiles, O:[B,S,C] bf16
int dual_attn_core_bf16(const void* X, const void* Wq, const void* Kp, const void* Vp, void* O, float* stats, int B,
                        int S, int C, int H, int Lt, int Li, float w_text, float w_img, cudaStream_t stream) {
  PV_REQUIRE(B > 0 && S > 0 && H > 0 && C % H == 0, "bad shape B=%d S=%d C=%d H=%d", B, S, C, H);
  const int d = C / H;
  PV_REQUIRE(d == 40 || d == 80 || d == 160, "head_dim %d unsupported (40/80/160)", d);
  PV_REQUIRE(C % AT_BN == 0 && C % AT_BK == 0, "C=%d must be a multiple of 320", C);
  PV_REQUIRE(Lt >= 1 && Lt <= AT_IMG_OFF && Li >= 1 && Li <= AT_KEYS - AT_IMG_OFF,
             "need 1 <= Lt <= %d and 1 <= Li <= %d (Lt=%d Li=%d)", AT_IMG_OFF, AT_KEYS - AT_IMG_OFF, Lt, Li);
  PV_REQUIRE(B <= 65535 && (S + AT_BM - 1) / AT_BM <= 65535, "grid too large");
  PV_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(Kp) |
              reinterpret_cast<uintptr_t>(Vp) | reinterpret_cast<uintptr_t>(O)) % 16 == 0, "pointers must be 16-byte aligned");
  CUtensorMap tmX, tmWq, tmO;
  if (make_tmap_3d(&tmX, X, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, AT_BK, AT_BM, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmWq, Wq, 2, C, C, 1, C * 2ull, static_cast<uint64_t>(C) * C * 2, AT_BK, AT_BN, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmO, O, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, AT_BN, 32, 1, Swz::None)) return PV_ERR_CUDA;
  AttnParams p;
  p.Kp = static_cast<const uint8_t*>(Kp);
  p.Vp = static_cast<const uint8_t*>(Vp);
  p.stats = stats;
  p.S = S; p.H = H; p.Lt = Lt; p.Li = Li;
  p.w_text = w_text; p.w_img = w_img;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(d));
  switch (d) {
    case 40: return launch_attn<40>(tmX, tmWq, tmO, p, B, S, C, stream);
    case 80: return launch_attn<80>(tmX, tmWq, tmO, p, B, S, C, stream);
    default: return launch_attn<160>(tmX, tmWq, tmO, p, B, S, C, stream);
  }
}

}  // namespace pv
