// photoverse_b200 -- fused Q-projection + dual-branch cross-attention, variant 2 ("TS": operands in TMEM).
//
// Same math and work decomposition as pv_attn.cu (CTA = 128 query rows x 160 channels = 4/2/1 heads), but the two
// attention contractions take their A operand from TENSOR MEMORY instead of shared memory:
//
//   Q   : fp32 accumulator -> bf16, packed two per 32-bit column, written back IN PLACE with tcgen05.st
//         -> S_h = Q_h K_h^T   is tcgen05.mma with A = TMEM (Q_h), B = smem (packed K_h tile)
//   P   : softmax output bf16, written over the first 48 columns of its own S buffer
//         -> O_h = P_h V_h     is tcgen05.mma with A = TMEM (P_h), B = smem (packed V_h^T tile)
//   O   : TMEM -> registers -> scaled -> bf16 -> straight to HBM (each thread owns one row: D*2 contiguous bytes)
//
// No Q / P / O staging in shared memory => the CTA needs only the TMA ring and the K/V tiles (<= 110 KB) and 256
// TMEM columns, so TWO CTAs are resident per SM for head dims 40 and 80: one CTA's softmax (MUFU/FMA pipes)
// overlaps the other's projection / attention MMAs (tensor pipe) and TMA traffic.  d=160 keeps one CTA per SM
// (its O accumulator alone is 160 columns) with a deeper ring.
//
// TMEM column map (fp32 columns):
//            Q acc      Q bf16 (in place)   O acc          S acc / P bf16 (in place)
//   d=40    [0,160)     4 x 24 = [0,96)     [96,144)       [160,256) / [160,208)
//   d=80    [0,160)     2 x 40 = [0,80)     [80,160)       [160,256) / [160,208)
//   d=160   [0,160)     1 x 80 = [0,80)     [256,416)      [160,256) / [160,208)      (512 columns allocated)
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int A2_BM = 128;
constexpr int A2_BN = 160;
constexpr int A2_BK = 64;
constexpr int A2_KEYS = PV_KEYS_PAD;
constexpr int A2_IMG_OFF = PV_IMG_KEY_OFFSET;
constexpr int A2_THREADS = 192;
constexpr int A2_A_BYTES = A2_BM * A2_BK * 2;
constexpr int A2_W_BYTES = A2_BN * A2_BK * 2;
constexpr int A2_STAGE_BYTES = A2_A_BYTES + A2_W_BYTES;

template <int D>
struct Attn2Cfg {
  static constexpr int HPC = A2_BN / D;
  static constexpr int D_PAD = (D + 15) / 16 * 16;
  static constexpr int QB_COLS = D_PAD / 2;                       // packed bf16 columns per head
  static constexpr bool TWO_CTA = (D != 160);
  static constexpr int STAGES = TWO_CTA ? 2 : 4;
  static constexpr int NKV = (D == 40) ? 2 : 1;                   // K/V tile-pair buffers
  static constexpr int KV_TILE_BYTES = A2_KEYS * D_PAD * 2;
  static constexpr int OFF_KV = STAGES * A2_STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_KV + NKV * 2 * KV_TILE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(!TWO_CTA || SMEM_BYTES <= 113 * 1024, "two CTAs per SM need <= 113 KB each");
  static constexpr uint32_t TMEM_COLS = TWO_CTA ? 256 : 512;
  static constexpr uint32_t TM_Q = 0;
  static constexpr uint32_t TM_S = 160;
  static constexpr uint32_t TM_O = (D == 160) ? 256 : HPC * QB_COLS;
  static_assert(TM_O + D_PAD <= (D == 160 ? 512 : 160), "O accumulator must fit beside the packed Q");
};

struct Attn2Params {
  const uint8_t* Kp;
  const uint8_t* Vp;
  __nv_bfloat16* O;        // [B,S,C]
  float* stats;            // optional [B,H,S,4]
  int S, C, H, Lt, Li;
  float w_text, w_img, scale_log2e;
};

template <int N>
__device__ __forceinline__ void pack_pairs(const uint32_t* v, uint32_t* out) {     // N fp32 -> N/2 bf16x2
#pragma unroll
  for (int i = 0; i < N / 2; ++i) out[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
}

template <int D>
__global__ void __launch_bounds__(A2_THREADS, (D != 160) ? 2 : 1)
dual_attn_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                        const Attn2Params p) {
  using Cfg = Attn2Cfg<D>;
  constexpr int HPC = Cfg::HPC;
  constexpr int D_PAD = Cfg::D_PAD;
  constexpr int NKV = Cfg::NKV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kv = smem + Cfg::OFF_KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = full + Cfg::STAGES;        // [STAGES]
  uint64_t* kv_full = empty + Cfg::STAGES;     // [2]
  uint64_t* kv_empty = kv_full + 2;            // [2]
  uint64_t* q_full = kv_empty + 2;
  uint64_t* q_ready = q_full + 1;
  uint64_t* s_full = q_ready + 1;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint64_t* o_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x;
  const int m0 = blockIdx.y * A2_BM;
  const int b = blockIdx.z;
  const int kblocks = p.C / A2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_ready, 128);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      auto load_kv = [&](int j) {
        const int buf = j % NKV;
        const size_t tile = (static_cast<size_t>(b) * p.H + (g * HPC + j)) * Cfg::KV_TILE_BYTES;
        mbar_expect_tx(&kv_full[buf], 2 * Cfg::KV_TILE_BYTES);
        bulk_load_1d(kv + (2 * buf) * Cfg::KV_TILE_BYTES, p.Kp + tile, Cfg::KV_TILE_BYTES, &kv_full[buf]);
        bulk_load_1d(kv + (2 * buf + 1) * Cfg::KV_TILE_BYTES, p.Vp + tile, Cfg::KV_TILE_BYTES, &kv_full[buf]);
      };
      for (int j = 0; j < NKV && j < HPC; ++j) load_kv(j);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* a_dst = smem + s * A2_STAGE_BYTES;
        mbar_expect_tx(&full[s], A2_STAGE_BYTES);
        tma_load_3d(a_dst, &tmX, &full[s], kb * A2_BK, m0, b);
        tma_load_3d(a_dst + A2_A_BYTES, &tmWq, &full[s], kb * A2_BK, g * A2_BN, 0);
      }
      for (int j = NKV; j < HPC; ++j) {          // later heads re-use a K/V buffer once its PV has completed
        mbar_wait(&kv_empty[j % NKV], ((j / NKV) - 1) & 1);
        load_kv(j);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      constexpr uint32_t idesc_q = umma_idesc_bf16(A2_BM, A2_BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint8_t* a_src = smem + s * A2_STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(a_src);
          const uint64_t dw = umma_desc_sw128(a_src + A2_A_BYTES);
#pragma unroll
          for (int k = 0; k < A2_BK / 16; ++k)
            umma_bf16_ss(tmem + Cfg::TM_Q, da + 2 * k, dw + 2 * k, idesc_q, (kb | k) != 0);
          umma_commit(&empty[s]);
          if (kb == kblocks - 1) umma_commit(q_full);
        }
        __syncwarp();
      }
    }
    mbar_wait(q_ready, 0);
    tc_fence_after();
    constexpr uint32_t idesc_s = umma_idesc_bf16(A2_BM, A2_KEYS);
    constexpr uint32_t idesc_o = umma_idesc_bf16(A2_BM, D_PAD);
#pragma unroll 1
    for (int j = 0; j < HPC; ++j) {
      const int buf = j % NKV;
      mbar_wait(&kv_full[buf], (j / NKV) & 1);
      if (j >= 1) mbar_wait(o_full, (j - 1) & 1);          // PV(j-1) has consumed P_{j-1}, which aliases S
      tc_fence_after();
      if (lane == 0) {
        // S = Q_j K_j^T : A = packed bf16 Q_j in TMEM (8 columns per K=16), B = K tile (no-swizzle core matrices).
        const uint32_t k_tile = smem_u32(kv + (2 * buf) * Cfg::KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D_PAD / 16; ++k) {
          const uint64_t db = umma_desc(k_tile + k * 2 * (A2_KEYS * 16), A2_KEYS * 16, 128, UMMA_LAYOUT_NONE);
          umma_bf16_ts(tmem + Cfg::TM_S, tmem + Cfg::TM_Q + j * Cfg::QB_COLS + k * 8, db, idesc_s, k != 0);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, j & 1);
      if (j >= 1) mbar_wait(o_free, (j - 1) & 1);          // O accumulator drained by the row threads
      tc_fence_after();
      if (lane == 0) {
        const uint32_t v_tile = smem_u32(kv + (2 * buf + 1) * Cfg::KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < A2_KEYS / 16; ++k) {
          const uint64_t db = umma_desc(v_tile + k * 2 * (D_PAD * 16), D_PAD * 16, 128, UMMA_LAYOUT_NONE);
          umma_bf16_ts(tmem + Cfg::TM_O, tmem + Cfg::TM_S + k * 8, db, idesc_o, k != 0);
        }
        umma_commit(o_full);
        if (j + NKV < HPC) umma_commit(&kv_empty[buf]);
      }
      __syncwarp();
    }
  } else {
    // ===================== row threads (warps 2..5): one thread per query row == TMEM lane =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);

    // ---- Q: fp32 accumulator -> packed bf16, in place (head j: columns [j*D, j*D+D) -> [j*QB, j*QB + D/2)) ----
    mbar_wait(q_full, 0);
    tc_fence_after();
#pragma unroll
    for (int j = 0; j < HPC; ++j) {
      const uint32_t src = tlane + Cfg::TM_Q + j * D;
      const uint32_t dst = tlane + Cfg::TM_Q + j * Cfg::QB_COLS;
      if constexpr (D == 40) {
        uint32_t a[32], c8[8], o[24];
        tmem_ld_x32(src, a);
        tmem_ld_x8(src + 32, c8);
        tmem_ld_wait();
        pack_pairs<32>(a, o);
        pack_pairs<8>(c8, o + 16);
        o[20] = o[21] = o[22] = o[23] = 0u;              // dims 40..47: zero padding of the K=48 contraction
        tmem_st_x16(dst, o);
        tmem_st_x8(dst + 16, o + 16);
      } else if constexpr (D == 80) {
        uint32_t a[32], b2[32], c16[16], o[40];
        tmem_ld_x32(src, a);
        tmem_ld_x32(src + 32, b2);
        tmem_ld_x16(src + 64, c16);
        tmem_ld_wait();
        pack_pairs<32>(a, o);
        pack_pairs<32>(b2, o + 16);
        pack_pairs<16>(c16, o + 32);
        tmem_st_x16(dst, o);
        tmem_st_x16(dst + 16, o + 16);
        tmem_st_x8(dst + 32, o + 32);
      } else {
        uint32_t o[80];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          uint32_t a[32];
          tmem_ld_x32(src + c * 32, a);
          tmem_ld_wait();
          pack_pairs<32>(a, o + c * 16);
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) tmem_st_x16(dst + c * 16, o + c * 16);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(q_ready);

    const int Lt = p.Lt;
    const int Li = p.Li;
    const float cs = p.scale_log2e;
    const bool row_ok = (m0 + row) < p.S;
    __nv_bfloat16* orow = p.O + (static_cast<size_t>(b) * p.S + (m0 + row)) * p.C + g * A2_BN;

#pragma unroll 1
    for (int j = 0; j < HPC; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float s[A2_KEYS];
#pragma unroll
      for (int c = 0; c < A2_KEYS / 32; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tlane + Cfg::TM_S + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(v[i]);
      }
      if (Lt < 64) {
#pragma unroll
        for (int k = 0; k < 64; ++k) s[k] = (k < Lt) ? s[k] : -INFINITY;
      }
#pragma unroll
      for (int k = 64; k < A2_IMG_OFF; ++k) s[k] = (k < Lt) ? s[k] : -INFINITY;
#pragma unroll
      for (int k = A2_IMG_OFF; k < A2_KEYS; ++k) s[k] = (k - A2_IMG_OFF < Li) ? s[k] : -INFINITY;
      float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
      for (int k = 4; k < A2_IMG_OFF; ++k) m4[k & 3] = fmaxf(m4[k & 3], s[k]);
      const float mt = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float i4[4] = {s[A2_IMG_OFF], s[A2_IMG_OFF + 1], s[A2_IMG_OFF + 2], s[A2_IMG_OFF + 3]};
#pragma unroll
      for (int k = A2_IMG_OFF + 4; k < A2_KEYS; ++k) i4[k & 3] = fmaxf(i4[k & 3], s[k]);
      const float mi = fmaxf(fmaxf(i4[0], i4[1]), fmaxf(i4[2], i4[3]));
      const float mts = mt * cs, mis = mi * cs;
      float l4[4] = {0.f, 0.f, 0.f, 0.f}, li4[2] = {0.f, 0.f};
#pragma unroll
      for (int kc = 0; kc < A2_IMG_OFF / 8; ++kc) {
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
      for (int kc = A2_IMG_OFF / 8; kc < A2_KEYS / 8; ++kc) {
        if ((kc - A2_IMG_OFF / 8) * 8 < Li) {
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
      const float at = p.w_text / lt;
      const float ai = p.w_img / li;
      if (p.stats != nullptr && row_ok) {
        const size_t idx = ((static_cast<size_t>(b) * p.H + (g * HPC + j)) * p.S + (m0 + row));
        reinterpret_cast<float4*>(p.stats)[idx] = make_float4(mts, lt, mis, li);
      }
      float ft, fi, oscale;
      if (p.w_text != 0.f) { ft = 1.f; fi = ai / at; oscale = at; }
      else                 { ft = 0.f; fi = 1.f;     oscale = ai; }
      // P (bf16 pairs) over the first 48 columns of the S buffer
      uint32_t pk[A2_KEYS / 2];
#pragma unroll
      for (int i = 0; i < A2_IMG_OFF / 2; ++i) pk[i] = (ft == 1.f) ? pack_bf16x2(s[2 * i], s[2 * i + 1]) : 0u;
#pragma unroll
      for (int i = A2_IMG_OFF / 2; i < A2_KEYS / 2; ++i) pk[i] = pack_bf16x2(s[2 * i] * fi, s[2 * i + 1] * fi);
      tmem_st_x16(tlane + Cfg::TM_S, pk);
      tmem_st_x16(tlane + Cfg::TM_S + 16, pk + 16);
      tmem_st_x16(tlane + Cfg::TM_S + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);

      // ---- O_j: TMEM -> registers -> * row scale -> bf16 -> HBM (this thread's row, D*2 contiguous bytes) ----
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      __nv_bfloat16* dst = orow + j * D;
      auto emit = [&](const uint32_t* v, int col0, int n) {
        if (!row_ok) return;
#pragma unroll
        for (int c = 0; c < n / 8; ++c) {
          const uint32_t* w = v + c * 8;
          st_global_v4(dst + col0 + c * 8,
                       pack_bf16x2(__uint_as_float(w[0]) * oscale, __uint_as_float(w[1]) * oscale),
                       pack_bf16x2(__uint_as_float(w[2]) * oscale, __uint_as_float(w[3]) * oscale),
                       pack_bf16x2(__uint_as_float(w[4]) * oscale, __uint_as_float(w[5]) * oscale),
                       pack_bf16x2(__uint_as_float(w[6]) * oscale, __uint_as_float(w[7]) * oscale));
        }
      };
      if constexpr (D == 40) {
        uint32_t a[32], c8[8];
        tmem_ld_x32(tlane + Cfg::TM_O, a);
        tmem_ld_x8(tlane + Cfg::TM_O + 32, c8);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(o_free);
        emit(a, 0, 32);
        emit(c8, 32, 8);
      } else if constexpr (D == 80) {
        uint32_t a[32], b2[32], c16[16];
        tmem_ld_x32(tlane + Cfg::TM_O, a);
        tmem_ld_x32(tlane + Cfg::TM_O + 32, b2);
        tmem_ld_x16(tlane + Cfg::TM_O + 64, c16);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(o_free);
        emit(a, 0, 32);
        emit(b2, 32, 32);
        emit(c16, 64, 16);
      } else {
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t a[32];
          tmem_ld_x32(tlane + Cfg::TM_O + c * 32, a);
          tmem_ld_wait();
          emit(a, c * 32, 32);
        }
        tc_fence_before();
        mbar_arrive(o_free);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem);
}

template <int D>
static int launch_attn2(const CUtensorMap& tmX, const CUtensorMap& tmWq, const Attn2Params& p, int B, int S, int C,
                        cudaStream_t stream) {
  using Cfg = Attn2Cfg<D>;
  auto kern = dual_attn_fwd_ts_kernel<D>;
  static bool attr_done = false;
  if (!attr_done) {
    PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(C / A2_BN, (S + A2_BM - 1) / A2_BM, B);
  kern<<<grid, A2_THREADS, Cfg::SMEM_BYTES, stream>>>(tmX, tmWq, p);
  PV_LAUNCHED();
  return PV_OK;
}

int dual_attn_core_bf16_ts(const void* X, const void* Wq, const void* Kp, const void* Vp, void* O, float* stats, int B,
                           int S, int C, int H, int Lt, int Li, float w_text, float w_img, cudaStream_t stream) {
  PV_REQUIRE(B > 0 && S > 0 && H > 0 && C % H == 0, "bad shape B=%d S=%d C=%d H=%d", B, S, C, H);
  const int d = C / H;
  PV_REQUIRE(d == 40 || d == 80 || d == 160, "head_dim %d unsupported (40/80/160)", d);
  PV_REQUIRE(C % A2_BN == 0 && C % A2_BK == 0, "C=%d must be a multiple of 320", C);
  PV_REQUIRE(Lt >= 1 && Lt <= A2_IMG_OFF && Li >= 1 && Li <= A2_KEYS - A2_IMG_OFF,
             "need 1 <= Lt <= %d and 1 <= Li <= %d (Lt=%d Li=%d)", A2_IMG_OFF, A2_KEYS - A2_IMG_OFF, Lt, Li);
  PV_REQUIRE(B <= 65535 && (S + A2_BM - 1) / A2_BM <= 65535, "grid too large");
  PV_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(Kp) |
              reinterpret_cast<uintptr_t>(Vp) | reinterpret_cast<uintptr_t>(O)) % 16 == 0, "pointers must be 16-byte aligned");
  CUtensorMap tmX, tmWq;
  if (make_tmap_3d(&tmX, X, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, A2_BK, A2_BM, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmWq, Wq, 2, C, C, 1, C * 2ull, static_cast<uint64_t>(C) * C * 2, A2_BK, A2_BN, 1, Swz::B128)) return PV_ERR_CUDA;
  Attn2Params p;
  p.Kp = static_cast<const uint8_t*>(Kp);
  p.Vp = static_cast<const uint8_t*>(Vp);
  p.O = static_cast<__nv_bfloat16*>(O);
  p.stats = stats;
  p.S = S; p.C = C; p.H = H; p.Lt = Lt; p.Li = Li;
  p.w_text = w_text; p.w_img = w_img;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(d));
  switch (d) {
    case 40: return launch_attn2<40>(tmX, tmWq, p, B, S, C, stream);
    case 80: return launch_attn2<80>(tmX, tmWq, p, B, S, C, stream);
    default: return launch_attn2<160>(tmX, tmWq, p, B, S, C, stream);
  }
}

}  // namespace pv
