// photoverse_b200 -- the out projection (reference models/attention_processor.py:423) as a PHASE of the fused processor
// kernels: after a CTA pair has finished its attention units it runs, inside the same launch, the
//   Y[256 rows, 160 cols] = O[256 rows, C] * Wo[160 cols, C]^T + bias
// tiles of its own static (head group, unit range) share.  A tile's 256 rows of O were written (TMA stores) by the G
// pairs that own the row block's head groups; they announce a finished unit on a per-row-block counter in global memory
// and the tile's producer waits for that counter instead of for a kernel boundary.
//
// The pipeline is the persistent CTA-pair GEMM of pv_gemm3.cu (W slice split across the pair and resident for C <= 640,
// deep A ring, M = 256 cta_group::2 MMAs into double-buffered TMEM accumulators, bias / bf16 / TMA-store epilogue under
// the next tile's MMAs) with 3-D (channel, row, sample) coordinates so that tiles coincide with attention units.
#pragma once
#include "pv_common.cuh"

namespace pv {

constexpr int OP_BN = 160;
constexpr int OP_BK = 64;
constexpr int OP_A_BYTES = 128 * OP_BK * 2;
constexpr int OP_WH_BYTES = (OP_BN / 2) * OP_BK * 2;
constexpr int OP_MAX_STAGES = 8;
constexpr int OP_BAR_BYTES = (2 * OP_MAX_STAGES + 5) * 8;     // full[8] empty[8] acc_full[2] slot_free[2] w_full

template <int KB_RES, int NSTAGES>                            // resident K-blocks of the W half (0: streamed)
struct OutProjCfg {
  static constexpr bool WSTAT = KB_RES > 0;
  static constexpr int KB_RESIDENT = KB_RES;
  static constexpr int W_RES_BYTES = KB_RES * OP_WH_BYTES;
  static constexpr int STAGES = NSTAGES;
  static constexpr int STAGE_BYTES = WSTAT ? OP_A_BYTES : OP_A_BYTES + OP_WH_BYTES;
  static constexpr int OFF_W = STAGES * STAGE_BYTES;
  static constexpr int OFF_OST = OFF_W + W_RES_BYTES;
  static constexpr int OST_WARP_BYTES = 32 * 80 * 2;
  static constexpr int OFF_BIAS = OFF_OST + 8 * OST_WARP_BYTES;
  static constexpr int BYTES = OFF_BIAS + OP_BN * 4;
  static_assert(STAGES <= OP_MAX_STAGES, "barrier array");
};

struct OutProjArgs {
  const float* bias;       // [C] or nullptr
  unsigned int* sync;      // [2 * V] : ready[V] then seen[V]; all zero between launches (self-resetting)
  int G, MTP, C;
  int V;                   // units per column group = B * MTP
  int u0, u1;              // this pair's tiles: units [u0, u1) of column group g
  int g;
  unsigned int ready_target;   // arrivals that complete a row block: (draining warps per unit) x G
};

__device__ __forceinline__ void op_mbar_init(uint64_t* pb) {
  uint64_t* full = pb;
  uint64_t* empty = full + OP_MAX_STAGES;
  uint64_t* acc_full = empty + OP_MAX_STAGES;
  uint64_t* slot_free = acc_full + 2;
  uint64_t* w_full = slot_free + 2;
  for (int s = 0; s < OP_MAX_STAGES; ++s) {
    mbar_init(&full[s], 1);
    mbar_init(&empty[s], 1);
  }
  for (int i = 0; i < 2; ++i) {
    mbar_init(&acc_full[i], 1);
    mbar_init(&slot_free[i], 16);
  }
  mbar_init(w_full, 1);
}

// An attention unit's O rows are in global memory: publish.  Called by ONE lane of a warp after
// cp.async.bulk.wait_group (full completion, not .read) has covered the warp's stores of that unit.
__device__ __forceinline__ void op_signal_unit(unsigned int* sync, int u) {
  asm volatile("fence.proxy.async;" ::: "memory");          // async-proxy (TMA) writes -> ordered before the release below
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sync + u) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {             // groups complete AND their writes performed
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Runs the out-projection tiles of this pair.  Every thread of BOTH CTAs of the pair calls it after a CTA + cluster
// barrier that closed the attention phase (shared memory [0, Cfg::BYTES) and all 512 TMEM columns are free, the barriers
// at `pb` were initialised by op_mbar_init at kernel start and never used).  Roles: warp 0 producer (both CTAs),
// warp 1 MMA issuer (leader), warps 4..11 epilogue (both CTAs); other warps return at once.
template <typename Cfg>
__device__ __forceinline__ void outproj_phase(uint8_t* smem, uint64_t* pb, uint32_t tmem, const CUtensorMap* tmA,
                                              const CUtensorMap* tmW, const CUtensorMap* tmD, const OutProjArgs& a) {
  constexpr bool WSTAT = Cfg::WSTAT;
  constexpr int nst = Cfg::STAGES;
  uint64_t* full = pb;
  uint64_t* empty = full + OP_MAX_STAGES;
  uint64_t* acc_full = empty + OP_MAX_STAGES;
  uint64_t* slot_free = acc_full + 2;
  uint64_t* w_full = slot_free + 2;
  float* sbias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int kblocks = a.C / OP_BK;
  const int n0 = a.g * OP_BN;

  if (warp == 0) {
    // ===================== producer (both CTAs): resident W half, then the A (= O) tiles as they become ready ==========
    if constexpr (WSTAT) {
      if (a.u0 < a.u1 && elect_one()) {
        const uint32_t bar = mapa_u32(smem_u32(w_full), 0);
        if (rank == 0) mbar_expect_tx(w_full, 2 * Cfg::W_RES_BYTES);
        for (int kb = 0; kb < Cfg::KB_RESIDENT; ++kb)
          tma_load_3d_2sm(smem + Cfg::OFF_W + kb * OP_WH_BYTES, tmW, bar, kb * OP_BK, n0 + static_cast<int>(rank) * (OP_BN / 2), 0);
      }
      __syncwarp();
    }
    uint32_t it = 0;
    const unsigned int target = a.ready_target;
    unsigned int* ready = a.sync;
    unsigned int* seen = a.sync + a.V;
    for (int u = a.u0; u < a.u1; ++u) {
      const int b = u / a.MTP;
      const int mt = 2 * (u - b * a.MTP) + static_cast<int>(rank);
      if (lane == 0) {
        unsigned int v;
        uint32_t tries = 0;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ready + u) : "memory");
          if (v >= target) break;
          __nanosleep(100);
          if (++tries > 20000000u) __trap();      // ~2 s: a protocol bug must surface as an error, not hang the GPU
        }
        // 2 G producers (both CTAs of the G pairs that own this row block's column groups) look at ready[u]; the last
        // one to have seen it complete puts both words back to zero for the next launch
        unsigned int old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(seen + u) : "memory");
        if (old == static_cast<unsigned int>(2 * a.G - 1)) {
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ready + u), "r"(0u) : "memory");
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(seen + u), "r"(0u) : "memory");
        }
      }
      __syncwarp();
      asm volatile("fence.proxy.async;" ::: "memory");      // acquire above -> ordered before the TMA (async-proxy) reads
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % nst;
        const uint32_t ph = (it / nst) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
          tma_load_3d_2sm(a_dst, tmA, bar, kb * OP_BK, mt * 128, b);
          if constexpr (!WSTAT)
            tma_load_3d_2sm(a_dst + OP_A_BYTES, tmW, bar, kb * OP_BK, n0 + static_cast<int>(rank) * (OP_BN / 2), 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader): M = 256, N = 160 for both CTAs =====================
      constexpr uint32_t idesc = umma_idesc_bf16(256, OP_BN);
      uint32_t it = 0;
      int i = 0;
      if constexpr (WSTAT) {
        if (a.u0 < a.u1) mbar_wait(w_full, 0);
      }
      for (int u = a.u0; u < a.u1; ++u, ++i) {
        const int slot = i & 1;
        if (i >= 2) mbar_wait(&slot_free[slot], ((i >> 1) - 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % nst;
          const uint32_t ph = (it / nst) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint8_t* a_src = smem + s * Cfg::STAGE_BYTES;
            const uint64_t da = umma_desc_sw128(a_src);
            const uint64_t dw = umma_desc_sw128(WSTAT ? smem + Cfg::OFF_W + kb * OP_WH_BYTES : a_src + OP_A_BYTES);
#pragma unroll
            for (int k = 0; k < OP_BK / 16; ++k) umma_bf16_ss_2sm(tmem + slot * OP_BN, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
            umma_commit_2sm(&empty[s]);
            if (kb == kblocks - 1) umma_commit_2sm(&acc_full[slot]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue (both CTAs): group w handles columns [80 w, 80 w + 80) =====================
    const int w = (warp - 4) >> 2;
    const int q = warp & 3;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* ost = smem + Cfg::OFF_OST + (warp - 4) * Cfg::OST_WARP_BYTES;
    const float* bw = sbias + 80 * w;
    // this group's 80 bias values (the attention phase owned the shared memory until the barrier before this call)
    for (int c = (warp & 3) * 32 + lane; c < 80; c += 128) sbias[80 * w + c] = a.bias ? a.bias[n0 + 80 * w + c] : 0.f;
    named_bar_sync(9 + w, 128);
    int i = 0;
    for (int u = a.u0; u < a.u1; ++u, ++i) {
      const int slot = i & 1;
      const int b = u / a.MTP;
      const int m0 = (2 * (u - b * a.MTP) + static_cast<int>(rank)) * 128;
      mbar_wait(&acc_full[slot], (i >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32], r2[16];
      const uint32_t src = tlane + slot * OP_BN + 80 * w;
      tmem_ld_x32(src, r0);
      tmem_ld_x32(src + 32, r1);
      tmem_ld_x16(src + 64, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive_cluster(mapa_u32(smem_u32(&slot_free[slot]), 0));   // accumulator is in registers
      if (elect_one()) bulk_wait_read<0>();              // previous TMA store of this warp has read the staging tile
      __syncwarp();
      auto put = [&](const uint32_t* v, int col0, int ncols) {
#pragma unroll
        for (int c = 0; c < ncols / 8; ++c) {
          uint32_t w4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cc = col0 + c * 8 + 2 * k;
            w4[k] = pack_bf16x2(__uint_as_float(v[c * 8 + 2 * k]) + bw[cc], __uint_as_float(v[c * 8 + 2 * k + 1]) + bw[cc + 1]);
          }
          st_shared_v4(ost + lane * 160 + (col0 + c * 8) * 2, w4[0], w4[1], w4[2], w4[3]);
        }
      };
      put(r0, 0, 32);
      put(r1, 32, 32);
      put(r2, 64, 16);
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_store_3d(tmD, ost, n0 + 80 * w, m0 + q * 32, b);
        bulk_commit();
      }
      __syncwarp();
    }
    if (elect_one()) bulk_wait_read<0>();
    __syncwarp();
  }
}

}  // namespace pv
