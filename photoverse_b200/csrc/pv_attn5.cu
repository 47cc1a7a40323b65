// photoverse_b200 -- fused Q-projection + dual-branch cross-attention, variant 5: persistent CTA pairs with KEY-SPLIT
// softmax groups (16 softmax warps per CTA).
//
// Producer / issuer warps, cross-CTA protocol, shared-memory operands and TMEM maps are those of pv_attn4.cu.  What
// changes is the softmax side, which the device timeline of variant 4 showed to bound the C = 320 layers (a softmax
// group was busy ~2750 cycles per head, one thread per query row over all 96 key slots, and the S/P-buffer dependency
// P(n) -> PV(n) -> QK^T(n+2) put that latency on the critical path):
//   * each softmax group is now EIGHT warps: two warps per TMEM lane quarter, one per half of the key slots
//     (half 0: text keys 0..47; half 1: text keys 48..79 + the 16 image slots).  Each thread handles 48 logits of its
//     row: per-head latency halves, and 16 resident softmax warps (4 per scheduler) hide MUFU / TMEM latencies.
//   * the two halves of a row combine through ONE exchange of (local max, local sum) in shared memory and a 64-thread
//     named barrier; every P entry is then scaled by 2^(m_local - m) * w / l as it is packed to bf16, so the O
//     accumulator needs no row scaling when it is drained.
//   * the Q conversion and the O drain are split the same way (half the columns per warp); each warp stages and
//     TMA-stores its own column slab, so the drain needs no further synchronisation.
#include <type_traits>

#include "pv_common.cuh"
#include "pv_softmax.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int A5_BM = 128;
constexpr int A5_BN = 160;
constexpr int A5_BK = 64;
constexpr int A5_KEYS = PV_KEYS_PAD;
constexpr int A5_IMG_OFF = PV_IMG_KEY_OFFSET;
constexpr int A5_THREADS = 640;      // 0 TMA, 1 projection MMA (leader), 2 attention MMA (leader), 3 K/V relay, 4..19 softmax
constexpr int A5_A_BYTES = A5_BM * A5_BK * 2;
constexpr int A5_WH_BYTES = (A5_BN / 2) * A5_BK * 2;
constexpr int A5_KB_WSTAT = 5;
constexpr int A5_MAX_STAGES = 8;
static_assert(A5_KEYS == 96 && A5_IMG_OFF == 80, "the key-split softmax is written for 80 text + 16 image slots");

template <int D, bool WSTAT>
struct Attn5Cfg {
  static constexpr int HPC = A5_BN / D;
  static constexpr int D_PAD = (D + 15) / 16 * 16;
  static constexpr int NB = (D == 160) ? 80 : D_PAD;              // N of one PV MMA (pair-wide)
  static constexpr int NBLK = (D == 160) ? 2 : 1;
  static constexpr int KH_BYTES = (A5_KEYS / 2) * D_PAD * 2;       // 48 keys of a K tile
  static constexpr int VBLK_BYTES = (NB / 2) * A5_KEYS * 2;        // NB/2 dims of a V^T tile
  static constexpr int VH_BYTES = NBLK * VBLK_BYTES;
  static constexpr int KV_HEAD_BYTES = KH_BYTES + VH_BYTES;
  static constexpr int KV_BYTES = HPC * KV_HEAD_BYTES;
  static constexpr int STAGE_BYTES = WSTAT ? A5_A_BYTES : A5_A_BYTES + A5_WH_BYTES;
  // PSMEM (C = 320): P goes to shared memory instead of over S in tensor memory, so an S buffer is free again as soon as
  // its logits are in registers and QK^T(n+2) no longer waits for PV(n); 48 KB of the X ring pay for the two P tiles
  static constexpr bool PSMEM = WSTAT;
  static constexpr int STAGES = WSTAT ? 4 : (D == 40 ? 6 : 5);
  static constexpr int OFF_W = STAGES * STAGE_BYTES;
  static constexpr int W_RES_BYTES = WSTAT ? A5_KB_WSTAT * A5_WH_BYTES : 0;
  static constexpr int OFF_KV = OFF_W + W_RES_BYTES;
  // O staging, one slab of 32 rows per softmax warp: d = 40: half 0 drains 24 columns of a head, half 1 the other 16;
  // d >= 80: 40 columns per warp and store
  static constexpr int OW0 = (D == 40) ? 24 : 40;
  static constexpr int OW1 = (D == 40) ? 16 : 40;
  static constexpr int OST0_BYTES = 32 * OW0 * 2;
  static constexpr int OST1_BYTES = 32 * OW1 * 2;
  static constexpr int OFF_OST = OFF_KV + KV_BYTES;
  static constexpr int OFF_XCH = OFF_OST + 8 * (OST0_BYTES + OST1_BYTES);   // (max, sum) exchange [group][quarter][half][lane]
  static constexpr int XCH_BYTES = 2 * 2 * 4 * 2 * 32 * 8;   // double-buffered by head parity: with PSMEM a half may start its next head early
  static constexpr int P_BYTES = A5_BM * A5_KEYS * 2;                        // one group's P tile: 12 key chunks x 128 rows x 16 B
  static constexpr int OFF_P = OFF_XCH + XCH_BYTES;
  static constexpr int OFF_BAR = OFF_P + (PSMEM ? 2 * P_BYTES : 0);
  static constexpr int ALIGN_SLACK = PSMEM ? 512 : 1024;     // dynamic smem starts 1024-aligned in practice; checked at run time
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + ALIGN_SLACK;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(STAGES <= A5_MAX_STAGES, "barrier array");
  static constexpr uint32_t TM_SBUF0 = 320;
  static constexpr uint32_t TM_SBUF1 = 416;
  __host__ __device__ static constexpr uint32_t q_col(int j) { return D == 40 ? 40 * j + 16 : D == 80 ? 80 * j + 40 : 40; }
  __host__ __device__ static constexpr uint32_t o_col(int w) { return D == 40 ? 48 * w : D == 80 ? 80 * w : 0; }
};

struct Attn5Params {
  const uint8_t* Kp;
  const uint8_t* Vp;
  float* stats;            // optional [B,H,S,4]
  int S, C, H, Lt, Li;
  int G, MTP;              // head groups per sample (C/160), row-tile PAIRS per sample
  int V;                   // units per head group = B * MTP
  float w_text, w_img, scale_log2e;
  unsigned long long* trace;   // debug timeline (leader CTA 0 only)
  int trace_cap;
};

template <int D, bool LT77, bool WSTAT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(A5_THREADS, 1)
dual_attn_fwd_pair16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                            const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                            const Attn5Params p) {
  using Cfg = Attn5Cfg<D, WSTAT>;
  constexpr int HPC = Cfg::HPC;
  constexpr int D_PAD = Cfg::D_PAD;
  constexpr int nst = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (static_cast<int>(smem - smem_raw) > Cfg::ALIGN_SLACK) __trap();
  uint8_t* kv = smem + Cfg::OFF_KV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                        // [STAGES]  (leader) both CTAs' X (+W half) stages landed
  uint64_t* empty = full + A5_MAX_STAGES;       // [STAGES]  multicast commit
  uint64_t* kv_full = empty + A5_MAX_STAGES;    // 1  local: this CTA's K/V halves landed
  uint64_t* kv_both = kv_full + 1;              // 1  (leader) both CTAs' K/V halves landed            count 2
  uint64_t* kv_free = kv_both + 1;              // 1  multicast commit
  uint64_t* q_full = kv_free + 1;               // [2] multicast commit: Q accumulators of a slot complete
  uint64_t* q_ready = q_full + 2;               // [2] (leader) packed bf16 Q in both CTAs' TMEM         count 32
  uint64_t* slot_free = q_ready + 2;            // [2] (leader) slot consumed in both CTAs               count 32
  uint64_t* s_full = slot_free + 2;             // [2] multicast commit, per softmax group
  uint64_t* p_ready = s_full + 2;               // [2] (leader)                                          count 16
  uint64_t* o_full = p_ready + 2;               // [2] multicast commit
  uint64_t* w_full = o_full + 2;                // 1  (leader) both halves of the resident Wq slice landed
  uint64_t* s_free = w_full + 1;                // [2] (leader, PSMEM) S buffer read into registers by both CTAs    count 16
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int kblocks = p.C / A5_BK;
  // Static schedule: pair c serves head group g = c % G and a contiguous range [u0, u1) of that group's V = B * MTP
  // (sample, row-tile pair) units; CTA `rank` of the pair owns row tile 2 * pair + rank.
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int g = pair % p.G;
  const int r_pair = pair / p.G;
  const int npair_g = (npairs - g + p.G - 1) / p.G;
  const int u0 = static_cast<int>((static_cast<long long>(p.V) * r_pair) / npair_g);
  const int u1 = static_cast<int>((static_cast<long long>(p.V) * (r_pair + 1)) / npair_g);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
    for (int s = 0; s < A5_MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(kv_full, 1);
    mbar_init(kv_both, 2);
    mbar_init(kv_free, 1);
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_ready[i], 32);
      mbar_init(&slot_free[i], 32);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 16);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 16);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Everything above reads nothing a predecessor kernel may have written; the resident Wq slice (static weights) is
  // requested before the wait as well, so launch latency, TMEM allocation and that load overlap the previous kernel's tail.
  if constexpr (WSTAT) {
    if (warp == 0) {
      if (u0 < u1 && elect_one()) {
        const uint32_t bar = mapa_u32(smem_u32(w_full), 0);
        if (rank == 0) mbar_expect_tx(w_full, 2 * Cfg::W_RES_BYTES);
        for (int kb = 0; kb < A5_KB_WSTAT; ++kb)
          tma_load_3d_2sm(smem + Cfg::OFF_W + kb * A5_WH_BYTES, &tmWq, bar, kb * A5_BK, g * A5_BN + static_cast<int>(rank) * (A5_BN / 2), 0);
      }
      __syncwarp();
    }
  }
  pdl_wait();
  pdl_launch_dependents();

  // one elected lane per warp arrives on the LEADER CTA's copy of `bar` (after every lane's TMEM traffic is fenced)
  auto arrive_leader = [&](uint64_t* bar) {
    __syncwarp();
    if (elect_one()) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    __syncwarp();
  };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");   // 128*56 + 512*104 <= 640*96 (the launch allocation)
  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own X rows, own halves of Wq / K / V^T) =====================
    uint32_t it = 0;
    int prev_b = -1;
    uint32_t kv_gen = 0;
    for (int u = u0; u < u1; ++u) {
      const int b = u / p.MTP;
      const int mt = 2 * (u - b * p.MTP) + static_cast<int>(rank);
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % nst;
        const uint32_t ph = (it / nst) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
          tma_load_3d_2sm(a_dst, &tmX, bar, kb * A5_BK, mt * A5_BM, b);
          if constexpr (!WSTAT)
            tma_load_3d_2sm(a_dst + A5_A_BYTES, &tmWq, bar, kb * A5_BK, g * A5_BN + static_cast<int>(rank) * (A5_BN / 2), 0);
        }
        __syncwarp();
      }
      if (b != prev_b) {
        // this CTA's halves of the K / V^T tiles of (sample b, head group g): keys [48 rank, 48 rank + 48) of K,
        // dims [NB/2 rank, NB/2 rank + NB/2) of every N-block of V^T -- contiguous pieces of the packed UMMA images
        if (kv_gen > 0) mbar_wait(kv_free, (kv_gen - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(kv_full, Cfg::KV_BYTES);
          for (int j = 0; j < HPC; ++j) {
            const size_t tile = (static_cast<size_t>(b) * p.H + (g * HPC + j)) * (A5_KEYS * D_PAD * 2);
            uint8_t* kd = kv + j * Cfg::KV_HEAD_BYTES;
            for (int kc = 0; kc < D_PAD / 8; ++kc)
              bulk_load_1d(kd + kc * (48 * 16), p.Kp + tile + (static_cast<size_t>(kc) * A5_KEYS + 48 * rank) * 16, 48 * 16, kv_full);
            uint8_t* vd = kd + Cfg::KH_BYTES;
            for (int blk = 0; blk < Cfg::NBLK; ++blk)
              for (int kc = 0; kc < A5_KEYS / 8; ++kc)
                bulk_load_1d(vd + blk * Cfg::VBLK_BYTES + kc * ((Cfg::NB / 2) * 16),
                             p.Vp + tile + (static_cast<size_t>(kc) * D_PAD + blk * Cfg::NB + (Cfg::NB / 2) * rank) * 16,
                             (Cfg::NB / 2) * 16, kv_full);
          }
        }
        __syncwarp();
        ++kv_gen;
        prev_b = b;
      }
    }
  } else if (warp == 3) {
    // ===================== K/V relay: "my halves have landed" -> leader's kv_both =====================
    int prev_b = -1;
    uint32_t kv_gen = 0;
    for (int u = u0; u < u1; ++u) {
      const int b = u / p.MTP;
      if (b != prev_b) {
        mbar_wait(kv_full, kv_gen & 1);
        if (elect_one()) mbar_arrive_cluster(mapa_u32(smem_u32(kv_both), 0));
        __syncwarp();
        ++kv_gen;
        prev_b = b;
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== projection MMA issuer (leader): Q = X Wq^T for BOTH CTAs, M = 256, N = 160 =====================
    constexpr uint32_t idesc_q = umma_idesc_bf16(2 * A5_BM, A5_BN);
    A3Trace tr = a3_trace_init_raw(p.trace, p.trace_cap, 0);
    uint32_t it = 0;
    int i = 0;
    if constexpr (WSTAT) {
      if (u0 < u1) mbar_wait(w_full, 0);
    }
    for (int u = u0; u < u1; ++u, ++i) {
      const int slot = i & 1;
      if (i >= 2) mbar_wait(&slot_free[slot], ((i >> 1) - 1) & 1);
      tc_fence_after();
      a3_trace(tr, 10, i);
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % nst;
        const uint32_t ph = (it / nst) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        a3_trace(tr, 12, kb);
        if (elect_one()) {
          const uint8_t* a_src = smem + s * Cfg::STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(a_src);
          const uint64_t dw = umma_desc_sw128(WSTAT ? smem + Cfg::OFF_W + kb * A5_WH_BYTES : a_src + A5_A_BYTES);
#pragma unroll
          for (int k = 0; k < A5_BK / 16; ++k)
            umma_bf16_ss_2sm(tmem + slot * A5_BN, da + 2 * k, dw + 2 * k, idesc_q, (kb | k) != 0);
          umma_commit_2sm(&empty[s]);
          if (kb == kblocks - 1) umma_commit_2sm(&q_full[slot]);
        }
        __syncwarp();
      }
      a3_trace(tr, 11, i);
    }
    a3_trace_done_raw(p.trace, tr, 0);
  } else if (warp == 2 && rank == 0) {
    // ===================== attention MMA issuer (leader), flat loop over the heads (see pv_attn3.cu) =====================
    // This warp shares its scheduler with four softmax warps that saturate the MUFU and FMA pipes, so every instruction
    // here costs ~10 cycles of issue latency and sits on the critical path P(n) -> PV(n) -> QK^T(n+2): no integer
    // division (MUFU.RCP), descriptors as one add on a precomputed word.
    constexpr uint32_t idesc_s = umma_idesc_bf16(2 * A5_BM, A5_KEYS);
    constexpr uint32_t idesc_o = umma_idesc_bf16(2 * A5_BM, Cfg::NB);
    A3Trace tr = a3_trace_init_raw(p.trace, p.trace_cap, 1);
    const int nheads = (u1 - u0) * HPC;
    int issued_qk = 0;
    uint32_t kv_gen = 0;
    int q_units_ready = 0;
    // K/V of one sample serve MTP consecutive units: `kv_end` = first unit (relative to u0) the loaded tiles do not cover
    const int first_end = (u0 / p.MTP + 1) * p.MTP - u0;
    int kv_end = 0;
    auto kv_advance = [&]() {
      mbar_wait(kv_both, kv_gen & 1);
      ++kv_gen;
      kv_end = (kv_end == 0) ? first_end : kv_end + p.MTP;
    };
    // descriptor words: the smem address field is the low 14 bits (16-byte units) of the low word
    const uint64_t kdesc0 = umma_desc(smem_u32(kv), 48 * 16, 128, UMMA_LAYOUT_NONE);
    const uint64_t vdesc0 = umma_desc(smem_u32(kv + Cfg::KH_BYTES), (Cfg::NB / 2) * 16, 128, UMMA_LAYOUT_NONE);
    const uint64_t pdesc0 = umma_desc(smem_u32(smem + Cfg::OFF_P), A5_BM * 16, 128, UMMA_LAYOUT_NONE);
    auto unit_of = [&](int nn) { return nn / HPC; };
    auto issue_qk = [&](int nn) {
      const int i = unit_of(nn), j = nn - i * HPC;
      const uint32_t tslot = tmem + (i & 1) * A5_BN;
      const uint32_t sbuf = tmem + ((nn & 1) ? Cfg::TM_SBUF1 : Cfg::TM_SBUF0);
      const uint64_t kd = kdesc0 + static_cast<uint32_t>(j * (Cfg::KV_HEAD_BYTES >> 4));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D_PAD / 16; ++k)
          umma_bf16_ts_2sm(sbuf, tslot + Cfg::q_col(j) + k * 8, kd + k * ((2 * 48 * 16) >> 4), idesc_s, k != 0);
        umma_commit_2sm(&s_full[nn & 1]);
      }
      __syncwarp();
      a3_trace(tr, 20, nn);
    };
    if constexpr (Cfg::PSMEM) {
      // S buffers are released by the softmax groups as soon as the logits are in registers (s_free), P comes back
      // through shared memory: QK^T runs up to two heads ahead of the PV stream and never waits for a PV.
      auto qk_ready = [&](int n2, bool block, int pv_next) -> bool {
        const int i2 = unit_of(n2);
        if (n2 >= 2) {                                         // S buffer n2 & 1 : logits of head n2 - 2 consumed
          const uint32_t ph = ((n2 - 2) >> 1) & 1;
          if (block) mbar_wait(&s_free[n2 & 1], ph);
          else if (!mbar_test_wait(&s_free[n2 & 1], ph)) return false;
        }
        if (q_units_ready <= i2) {
          if (block) mbar_wait(&q_ready[i2 & 1], (i2 >> 1) & 1);
          else if (!mbar_test_wait(&q_ready[i2 & 1], (i2 >> 1) & 1)) return false;
          q_units_ready = i2 + 1;
        }
        if (i2 >= kv_end) {
          // the K/V tiles of the next sample replace the current ones after the last PV of this sample (kv_free)
          if (pv_next < i2 * HPC) return false;               // (block implies pv_next == n2 == i2 * HPC here)
          kv_advance();
        }
        return true;
      };
#pragma unroll 1
      for (int nn = 0; nn < nheads; ++nn) {
        const int i = unit_of(nn), j = nn - i * HPC;
        while (issued_qk < nheads && issued_qk <= nn + 2) {
          if (!qk_ready(issued_qk, issued_qk <= nn, nn)) break;
          tc_fence_after();
          issue_qk(issued_qk);
          ++issued_qk;
        }
        const uint32_t w = nn & 1;
        a3_trace(tr, 25, nn);
        mbar_wait(&p_ready[w], (nn >> 1) & 1);
        tc_fence_after();
        a3_trace(tr, 26, nn);
        const uint64_t pd = pdesc0 + static_cast<uint32_t>(w * (Cfg::P_BYTES >> 4));
        const uint64_t vd = vdesc0 + static_cast<uint32_t>(j * (Cfg::KV_HEAD_BYTES >> 4));
        if (elect_one()) {
          const uint32_t tslot = tmem + (i & 1) * A5_BN;
#pragma unroll
          for (int k = 0; k < A5_KEYS / 16; ++k)
            umma_bf16_ss_2sm(tslot + Cfg::o_col(w), pd + k * ((2 * A5_BM * 16) >> 4), vd + k * ((2 * (Cfg::NB / 2) * 16) >> 4),
                             idesc_o, k != 0);
          umma_commit_2sm(&o_full[w]);
          if (j == HPC - 1 && nn + 1 < nheads && i + 1 == kv_end) umma_commit_2sm(kv_free);
        }
        __syncwarp();
        a3_trace(tr, 21, nn);
      }
    } else {
#pragma unroll 1
    for (int nn = 0; nn < nheads; ++nn) {
      const int i = unit_of(nn), j = nn - i * HPC;
      if (issued_qk <= nn) {
        if (q_units_ready <= i) { mbar_wait(&q_ready[i & 1], (i >> 1) & 1); q_units_ready = i + 1; }
        if (i >= kv_end) kv_advance();
        tc_fence_after();
        issue_qk(nn);
        issued_qk = nn + 1;
      }
      if (nn + 1 < nheads && issued_qk == nn + 1) {
        const int i2 = unit_of(nn + 1);
        bool ok = (i2 == i);
        if (!ok && i2 < kv_end) {
          if (q_units_ready > i2) ok = true;
          else if (mbar_test_wait(&q_ready[i2 & 1], (i2 >> 1) & 1)) { ok = true; q_units_ready = i2 + 1; }
        }
        if (ok) {
          tc_fence_after();
          issue_qk(nn + 1);
          issued_qk = nn + 2;
        }
      }
      const uint32_t w = nn & 1;
      a3_trace(tr, 25, nn);
      mbar_wait(&p_ready[w], (nn >> 1) & 1);
      tc_fence_after();
      a3_trace(tr, 26, nn);
      const uint64_t vd = vdesc0 + static_cast<uint32_t>(j * (Cfg::KV_HEAD_BYTES >> 4));
      if (elect_one()) {
        const uint32_t tslot = tmem + (i & 1) * A5_BN;
        const uint32_t sbuf = tmem + (w ? Cfg::TM_SBUF1 : Cfg::TM_SBUF0);
#pragma unroll
        for (int blk = 0; blk < Cfg::NBLK; ++blk) {
#pragma unroll
          for (int k = 0; k < A5_KEYS / 16; ++k)
            umma_bf16_ts_2sm(tslot + Cfg::o_col(w) + blk * Cfg::NB, sbuf + k * 8,
                             vd + ((blk * Cfg::VBLK_BYTES + k * 2 * ((Cfg::NB / 2) * 16)) >> 4), idesc_o, k != 0);
        }
        umma_commit_2sm(&o_full[w]);
        if (j == HPC - 1 && nn + 1 < nheads && i + 1 == kv_end) umma_commit_2sm(kv_free);
      }
      __syncwarp();
      a3_trace(tr, 21, nn);
    }
    }
    a3_trace_done_raw(p.trace, tr, 1);
  }
  } else {
    // ===================== softmax groups (warps 4..11 and 12..19) of BOTH CTAs =====================
    // group wg: 8 warps = 4 TMEM lane quarters x 2 key halves; one thread per (query row, key half)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int sw = warp - 4;
    const int wg = sw >> 3;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    auto role = [&](auto half_tag) {
      constexpr int HF = decltype(half_tag)::value;
      constexpr int OWH = HF ? Cfg::OW1 : Cfg::OW0;
      const int row = q * 32 + lane;
      const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t sbuf = tlane + (wg ? Cfg::TM_SBUF1 : Cfg::TM_SBUF0);
      const uint32_t pair_bar = 1 + wg * 4 + q;      // named barrier of the two warps that share this group's row slab
      float2* xch = reinterpret_cast<float2*>(smem + Cfg::OFF_XCH) + ((wg * 4 + q) * 2) * 32 + lane;
      uint8_t* ost = smem + Cfg::OFF_OST +
                     (HF ? 8 * Cfg::OST0_BYTES + (wg * 4 + q) * Cfg::OST1_BYTES : (wg * 4 + q) * Cfg::OST0_BYTES);
      const CUtensorMap* tmOh = HF ? &tmO1 : &tmO0;
      const int Lt = p.Lt;
      const int Li = p.Li;
      const float cs = p.scale_log2e;
      A3Trace tr = a3_trace_init_raw(p.trace, p.trace_cap, 2 + wg);
      if (q != 0 || HF != 0) tr.base = nullptr;
      PendingO pend;
      pend.valid = false;

      // fp32 accumulator columns -> bf16 -> this warp's staging slab (rows = lanes)
      auto stage_store = [&](const uint32_t* v, int col0, int ncols) {
#pragma unroll
        for (int c = 0; c < ncols / 8; ++c) {
          uint32_t w4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            w4[k] = pack_bf16x2(__uint_as_float(v[c * 8 + 2 * k]), __uint_as_float(v[c * 8 + 2 * k + 1]));
          st_shared_v4(ost + lane * (OWH * 2) + (col0 + c * 8) * 2, w4[0], w4[1], w4[2], w4[3]);
        }
      };
      // this warp's column slab of a finished O accumulator (already normalised: the row scale is folded into P)
      auto drain = [&](const PendingO& po) {
        mbar_wait(&o_full[wg], po.parity);
        tc_fence_after();
        if constexpr (D == 40) {
          if (elect_one()) bulk_wait_read<0>();          // the previous TMA store of this warp has read the slab
          __syncwarp();
          if constexpr (HF == 0) {
            uint32_t a[16], c8[8];
            tmem_ld_x16(po.taddr, a);
            tmem_ld_x8(po.taddr + 16, c8);
            tmem_ld_wait();
            stage_store(a, 0, 16);
            stage_store(c8, 16, 8);
          } else {
            uint32_t a[16];
            tmem_ld_x16(po.taddr + 24, a);
            tmem_ld_wait();
            stage_store(a, 0, 16);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_3d(tmOh, ost, po.c0 + (HF ? 24 : 0), po.r0, po.b);
            bulk_commit();
          }
          __syncwarp();
        } else {
          // d = 80: columns [40 HF, 40 HF + 40) of the head;  d = 160: N-block HF (80 columns) as two slabs of 40
          constexpr int NCH = (D == 160) ? 2 : 1;
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const int coff = (D == 160) ? 80 * HF + 40 * ch : 40 * HF;
            if (elect_one()) bulk_wait_read<0>();
            __syncwarp();
            uint32_t a[32], c8[8];
            tmem_ld_x32(po.taddr + coff, a);
            tmem_ld_x8(po.taddr + coff + 32, c8);
            tmem_ld_wait();
            stage_store(a, 0, 32);
            stage_store(c8, 32, 8);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              tma_store_3d(tmOh, ost, po.c0 + coff, po.r0, po.b);
              bulk_commit();
            }
            __syncwarp();
          }
        }
      };

      // Q of unit iu: fp32 accumulator -> packed bf16, written inside columns this group has just read
      auto convert_unit = [&](int iu) {
        const int slot = iu & 1;
        const uint32_t tslot = tlane + slot * A5_BN;
        tc_fence_after();
        if constexpr (D == 40) {
          // this warp converts head wg + 2 HF : fp32 [40 j, 40 j + 40) -> bf16 [40 j + 16, 40 j + 40)
          const int j = wg + 2 * HF;
          uint32_t a[32], c8[8], o[24];
          tmem_ld_x32(tslot + 40 * j, a);
          tmem_ld_x8(tslot + 40 * j + 32, c8);
          tmem_ld_wait();
          pack_pairs3<32>(a, o);
          pack_pairs3<8>(c8, o + 16);
          o[20] = o[21] = o[22] = o[23] = 0u;          // dims 40..47 pad the K = 48 contraction
          tmem_st_x16(tslot + 40 * j + 16, o);
          tmem_st_x8(tslot + 40 * j + 32, o + 16);
        } else {
          // d = 80: group wg converts head wg; d = 160: dims [80 wg, 80 wg + 80) of the single head.  This warp takes 40
          // fp32 columns [80 wg + 40 HF, +40) -> 20 packed columns.  The destination of one half overlaps the source of
          // the other, so both halves load, meet at the pair barrier, then store.
          const uint32_t srcc = tslot + wg * 80 + 40 * HF;
          uint32_t a[32], c8[8], o[20];
          tmem_ld_x32(srcc, a);
          tmem_ld_x8(srcc + 32, c8);
          tmem_ld_wait();
          pack_pairs3<32>(a, o);
          pack_pairs3<8>(c8, o + 16);
          tc_fence_before();
          named_bar_sync(pair_bar, 64);
          tc_fence_after();
          const uint32_t dstc = tslot + ((D == 80) ? (wg * 80 + 40) : (40 + wg * 40)) + 20 * HF;
          tmem_st_x16(dstc, o);
          tmem_st_x4(dstc + 16, o + 16);
        }
        tmem_st_wait();
        tc_fence_before();
        arrive_leader(&q_ready[slot]);
      };
      int converted = 0;              // units [0, converted) of this CTA have had their Q converted by this warp

      int i = 0;
#pragma unroll 1
      for (int u = u0; u < u1; ++u, ++i) {
        const int slot = i & 1;
        const uint32_t tslot = tlane + slot * A5_BN;
        const int b = u / p.MTP;
        const int mt = 2 * (u - b * p.MTP) + static_cast<int>(rank);
        const int m0 = mt * A5_BM;

        if (converted <= i) {            // not converted ahead of time (see below): wait for the projection now
          mbar_wait(&q_full[slot], (i >> 1) & 1);
          a3_trace(tr, 30 + 10 * wg, i);
          convert_unit(i);
          converted = i + 1;
          a3_trace(tr, 31 + 10 * wg, i);
        }
        // this group's last head of the previous unit: its PV ran during the conversion above; draining it now returns
        // the previous slot to the projection pipeline as early as possible
        if constexpr (!Cfg::PSMEM) {
          if (pend.valid) {
            drain(pend);
            tc_fence_before();
            arrive_leader(&slot_free[pend.slot]);
            pend.valid = false;
            a3_trace(tr, 32 + 10 * wg, i);
          }
        }

        const bool row_ok = (m0 + row) < p.S;
        bool had_head = false;

#pragma unroll 1
        for (int j = 0; j < HPC; ++j) {
          const int nn = i * HPC + j;
          if ((nn & 1) != wg) continue;
          had_head = true;
          const uint32_t par = (nn >> 1) & 1;
          // d = 40: this group's previous head of the same unit.  Its PV was issued right after P was published and has
          // completed by the time S(nn) can arrive; it must leave TMEM before P(nn) is published (PV(nn) overwrites it).
          if constexpr (!Cfg::PSMEM) {
            if (pend.valid) {
              drain(pend);
              pend.valid = false;
            }
          }
          mbar_wait(&s_full[wg], par);
          tc_fence_after();
          a3_trace(tr, 33 + 10 * wg, nn);
          uint32_t sr[48];                         // this half's 48 logits (fp32 bits), later the exponentials
          tmem_ld32_raw(sbuf + 48 * HF, sr);
          tmem_ld16_raw(sbuf + 48 * HF + 32, sr + 32);
          tmem_ld_wait();
          if constexpr (Cfg::PSMEM) {               // the S buffer may be overwritten by QK^T(nn + 2) from here on
            tc_fence_before();
            arrive_leader(&s_free[wg]);
          }
          a3_trace(tr, 35 + 10 * wg, nn);

          const uint64_t cs2 = f2_pack(cs, cs);
          float mloc, lloc, mis = 0.f, li = 1.f;
          if constexpr (HF == 0) {
            // text keys 0..47 (all real keys when Lt == 77; otherwise a run-time mask)
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 48; ++c) {
              const float v = __uint_as_float(sr[c]);
              mx[c & 3] = fmaxf(mx[c & 3], LT77 ? v : (c < Lt ? v : -INFINITY));
            }
            mloc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * cs;
            const uint64_t nm2 = f2_pack(-mloc, -mloc);
            uint64_t lacc[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
            for (int k = 0; k < 24; ++k) {
              const int c = 2 * k;
              float a, b2;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), cs2, nm2), a, b2);
              a = fast_exp2(a);
              b2 = fast_exp2(b2);
              if constexpr (!LT77) {
                a = (c < Lt) ? a : 0.f;
                b2 = (c + 1 < Lt) ? b2 : 0.f;
              }
              lacc[k & 1] = f2_add(lacc[k & 1], f2_pack(a, b2));
              sr[c] = __float_as_uint(a);
              sr[c + 1] = __float_as_uint(b2);
            }
            float l0, l1, l2, l3;
            f2_unpack(lacc[0], l0, l1);
            f2_unpack(lacc[1], l2, l3);
            lloc = (l0 + l1) + (l2 + l3);
          } else {
            // local 0..31 = text keys 48..79 (real keys: < Lt; 77..79 are padding when Lt == 77), local 32..47 = image slots
            float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (LT77 && 48 + c >= 77) continue;
              const float v = __uint_as_float(sr[c]);
              mx[c & 3] = fmaxf(mx[c & 3], LT77 ? v : (48 + c < Lt ? v : -INFINITY));
            }
            float mi2[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 16; ++c)
              mi2[c & 1] = fmaxf(mi2[c & 1], (c < Li) ? __uint_as_float(sr[32 + c]) : -INFINITY);
            mloc = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * cs;
            mis = fmaxf(mi2[0], mi2[1]) * cs;
            const uint64_t nm2 = f2_pack(-mloc, -mloc), nmi2 = f2_pack(-mis, -mis);
            uint64_t lacc[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
            uint64_t iacc = f2_pack(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int c = 2 * k;
              if (LT77 && 48 + c >= 77) { sr[c] = 0u; sr[c + 1] = 0u; continue; }
              float a, b2;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), cs2, nm2), a, b2);
              a = fast_exp2(a);
              b2 = fast_exp2(b2);
              if constexpr (LT77) {
                if (48 + c + 1 >= 77) b2 = 0.f;
              } else {
                a = (48 + c < Lt) ? a : 0.f;
                b2 = (48 + c + 1 < Lt) ? b2 : 0.f;
              }
              lacc[k & 1] = f2_add(lacc[k & 1], f2_pack(a, b2));
              sr[c] = __float_as_uint(a);
              sr[c + 1] = __float_as_uint(b2);
            }
            if (Li > 8) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int c = 32 + 2 * k;
                float a, b2;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), cs2, nmi2), a, b2);
                a = (2 * k < Li) ? fast_exp2(a) : 0.f;
                b2 = (2 * k + 1 < Li) ? fast_exp2(b2) : 0.f;
                iacc = f2_add(iacc, f2_pack(a, b2));
                sr[c] = __float_as_uint(a);
                sr[c + 1] = __float_as_uint(b2);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = 32 + 2 * k;
                float a, b2;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), cs2, nmi2), a, b2);
                a = (2 * k < Li) ? fast_exp2(a) : 0.f;
                b2 = (2 * k + 1 < Li) ? fast_exp2(b2) : 0.f;
                iacc = f2_add(iacc, f2_pack(a, b2));
                sr[c] = __float_as_uint(a);
                sr[c + 1] = __float_as_uint(b2);
              }
#pragma unroll
              for (int c = 40; c < 48; ++c) sr[c] = 0u;
            }
            float l0, l1, l2, l3, li0, li1;
            f2_unpack(lacc[0], l0, l1);
            f2_unpack(lacc[1], l2, l3);
            f2_unpack(iacc, li0, li1);
            lloc = (l0 + l1) + (l2 + l3);
            li = li0 + li1;
          }

          a3_trace(tr, 36 + 10 * wg, nn);
          // combine the two halves of the text segment: m = max(m0, m1), l = l0 2^(m0 - m) + l1 2^(m1 - m)
          float2* xb = xch + par * (2 * 4 * 2 * 32);
          xb[HF * 32] = make_float2(mloc, lloc);
          named_bar_sync(pair_bar, 64);             // also: the other half has loaded its S columns (P overwrites them)
          const float2 ot = xb[(HF ^ 1) * 32];
          a3_trace(tr, 37 + 10 * wg, nn);
          const float mts = fmaxf(mloc, ot.x);
          const float fo = fast_exp2(mloc - mts);
          const float lt = lloc * fo + ot.y * fast_exp2(ot.x - mts);
          const float tf = fo * (p.w_text / lt);    // text entries: e * 2^(m_local - m) * w_text / l_text
          if constexpr (HF == 1) {
            if (p.stats != nullptr && row_ok) {
              const size_t idx = ((static_cast<size_t>(b) * p.H + (g * HPC + j)) * p.S + (m0 + row));
              reinterpret_cast<float4*>(p.stats)[idx] = make_float4(mts, lt, mis, li);
            }
          }
          if constexpr (Cfg::PSMEM) {
            // the previous head of this group (any unit): its PV completed long ago; it must have left TMEM -- and have
            // read its P tile -- before P(nn) is published
            if (pend.valid) {
              drain(pend);
              if (pend.last_of_unit) {
                tc_fence_before();
                arrive_leader(&slot_free[pend.slot]);
              }
              pend.valid = false;
            }
          }
          a3_trace(tr, 38 + 10 * wg, nn);
          // P (bf16 pairs): half 0 -> packed columns [0, 24), half 1 -> [24, 48) of the S buffer (PSMEM: key chunks
          // [6 HF, 6 HF + 6) of the group's P tile in shared memory, 16 B per row and chunk)
          {
            const uint64_t tf2 = f2_pack(tf, tf);
            const float ai = p.w_img / li;            // image entries: e * w_img / l_img   (half 1 only)
            const uint64_t ai2 = f2_pack(ai, ai);
            uint32_t pk[24];
#pragma unroll
            for (int k = 0; k < 24; ++k) {
              float a, b2;
              const uint64_t f2 = (HF == 1 && k >= 16) ? ai2 : tf2;
              f2_unpack(f2_mul(f2_pack(__uint_as_float(sr[2 * k]), __uint_as_float(sr[2 * k + 1])), f2), a, b2);
              pk[k] = pack_bf16x2(a, b2);
            }
            if constexpr (Cfg::PSMEM) {
              uint8_t* prow = smem + Cfg::OFF_P + wg * Cfg::P_BYTES + (6 * HF) * (A5_BM * 16) + row * 16;
#pragma unroll
              for (int c = 0; c < 6; ++c)
                st_shared_v4(prow + c * (A5_BM * 16), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            } else {
              tmem_st_x16(sbuf + 24 * HF, pk);
              tmem_st_x8(sbuf + 24 * HF + 16, pk + 16);
            }
          }
          if constexpr (Cfg::PSMEM) {
            fence_proxy_async_smem();               // generic-proxy stores -> visible to the tensor core's smem reads
          } else {
            tmem_st_wait();
            tc_fence_before();
          }
          arrive_leader(&p_ready[wg]);
          a3_trace(tr, 34 + 10 * wg, nn);
          // While the PV MMA of this head and the QK^T of the next one run: convert the NEXT unit's Q if its projection
          // has already completed, so that the attention issuer can look ahead across the unit boundary.  (For d >= 80
          // the conversion contains a pair barrier; each warp performs it exactly once between its exchange barriers of
          // unit i and unit i + 1 -- here or at the top of the unit loop -- so the two halves always pair up.)
          if (converted == i + 1 && u + 1 < u1 && mbar_test_wait(&q_full[(i + 1) & 1], ((i + 1) >> 1) & 1)) {
            convert_unit(i + 1);
            converted = i + 2;
          }

          pend.valid = true;
          pend.c0 = g * A5_BN + j * D;
          pend.r0 = m0 + q * 32;
          pend.b = b;
          pend.taddr = tslot + Cfg::o_col(wg);
          pend.oscale = 1.f;
          pend.parity = par;
          pend.slot = slot;
          pend.last_of_unit = (j + 2 >= HPC);
        }
        if (!had_head) {                                         // d = 160: the other group owns this unit's head
          arrive_leader(&slot_free[slot]);
          if (converted == i + 1 && u + 1 < u1 && mbar_test_wait(&q_full[(i + 1) & 1], ((i + 1) >> 1) & 1)) {
            convert_unit(i + 1);
            converted = i + 2;
          }
        }
      }
      if (pend.valid) drain(pend);
      if (elect_one()) bulk_wait_read<0>();
      __syncwarp();
      a3_trace_done_raw(p.trace, tr, 2 + wg);
    };
    if (((sw >> 2) & 1) == 0) role(std::integral_constant<int, 0>{});
    else role(std::integral_constant<int, 1>{});
  }


  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // neither CTA frees TMEM / exits while pair-wide MMAs or remote signals are in flight
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm<512>(tmem);
}

extern int g_opt_attn3_wstat;
extern unsigned long long* g_attn3_trace;
extern int g_attn3_trace_cap;

template <int D, bool LT77, bool WSTAT>
static int launch_attn5(const CUtensorMap& tmX, const CUtensorMap& tmWq, const CUtensorMap& tmO0, const CUtensorMap& tmO1,
                        const Attn5Params& p, long long unit_pairs, cudaStream_t stream) {
  using Cfg = Attn5Cfg<D, WSTAT>;
  auto kern = dual_attn_fwd_pair16_kernel<D, LT77, WSTAT>;
  static bool attr_done = false;
  if (!attr_done) {
    PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const long long max_pairs = sm_count() / 2;
  const int npairs = static_cast<int>(unit_pairs < max_pairs ? unit_pairs : max_pairs);
  PV_CUDA(launch_pdl(kern, dim3(2 * npairs), dim3(A5_THREADS), Cfg::SMEM_BYTES, stream, tmX, tmWq, tmO0, tmO1, p));
  PV_LAUNCHED();
  return PV_OK;
}

// Same contract as dual_attn_core_bf16_persistent (pv_attn3.cu); requires at least two row tiles per sample.
int dual_attn_core_bf16_pair16(const void* X, const void* Wq, const void* Kp, const void* Vp, void* O, float* stats, int B,
                             int S, int C, int H, int Lt, int Li, float w_text, float w_img, cudaStream_t stream) {
  PV_REQUIRE(B > 0 && S > 0 && H > 0 && C % H == 0, "bad shape B=%d S=%d C=%d H=%d", B, S, C, H);
  const int d = C / H;
  PV_REQUIRE(d == 40 || d == 80 || d == 160, "head_dim %d unsupported (40/80/160)", d);
  PV_REQUIRE(C % A5_BN == 0 && C % A5_BK == 0, "C=%d must be a multiple of 320", C);
  PV_REQUIRE(Lt >= 1 && Lt <= A5_IMG_OFF && Li >= 1 && Li <= A5_KEYS - A5_IMG_OFF,
             "need 1 <= Lt <= %d and 1 <= Li <= %d (Lt=%d Li=%d)", A5_IMG_OFF, A5_KEYS - A5_IMG_OFF, Lt, Li);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wq) | reinterpret_cast<uintptr_t>(Kp) |
              reinterpret_cast<uintptr_t>(Vp) | reinterpret_cast<uintptr_t>(O)) % 16 == 0, "pointers must be 16-byte aligned");
  CUtensorMap tmX, tmWq, tmO0, tmO1;
  if (make_tmap_3d(&tmX, X, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, A5_BK, A5_BM, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmWq, Wq, 2, C, C, 1, C * 2ull, static_cast<uint64_t>(C) * C * 2, A5_BK, A5_BN / 2, 1, Swz::B128)) return PV_ERR_CUDA;
  // O leaves through per-warp column slabs: d = 40: 24 + 16 columns of a head, d >= 80: 40 columns
  if (make_tmap_3d(&tmO0, O, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, d == 40 ? 24 : 40, 32, 1, Swz::None)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmO1, O, 2, C, S, B, C * 2ull, static_cast<uint64_t>(S) * C * 2, d == 40 ? 16 : 40, 32, 1, Swz::None)) return PV_ERR_CUDA;
  Attn5Params p;
  p.Kp = static_cast<const uint8_t*>(Kp);
  p.Vp = static_cast<const uint8_t*>(Vp);
  p.stats = stats;
  p.S = S; p.C = C; p.H = H; p.Lt = Lt; p.Li = Li;
  p.G = C / A5_BN;
  const int MT = (S + A5_BM - 1) / A5_BM;
  p.MTP = (MT + 1) / 2;
  p.V = B * p.MTP;
  const long long unit_pairs = static_cast<long long>(p.V) * p.G;
  PV_REQUIRE(unit_pairs < (1ll << 30), "too many work units");
  p.w_text = w_text; p.w_img = w_img;
  p.trace = g_attn3_trace;
  p.trace_cap = g_attn3_trace_cap;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(d));
  switch (d) {
    case 40:
      if (C == A5_KB_WSTAT * A5_BK && g_opt_attn3_wstat)
        return Lt == 77 ? launch_attn5<40, true, true>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream)
                        : launch_attn5<40, false, true>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream);
      return Lt == 77 ? launch_attn5<40, true, false>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream)
                      : launch_attn5<40, false, false>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream);
    case 80:
      return Lt == 77 ? launch_attn5<80, true, false>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream)
                      : launch_attn5<80, false, false>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream);
    default:
      return Lt == 77 ? launch_attn5<160, true, false>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream)
                      : launch_attn5<160, false, false>(tmX, tmWq, tmO0, tmO1, p, unit_pairs, stream);
  }
}

}  // namespace pv
