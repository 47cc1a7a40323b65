// photoverse_b200 -- CTA-pair (cta_group::2) tcgen05 projection GEMM:  D = A * W^T + bias, bf16 operands.
//
// A cluster of two CTAs computes a 256 x BN output tile with ONE tcgen05.mma stream of M = 256 issued by the leader CTA:
//   * each CTA TMA-loads its own 128 rows of A and HALF (BN/2 rows) of the W tile -> per CTA 16 KB + BN/2 * 128 B of
//     TMA ingest and shared-memory operand reads per K-block instead of 16 KB + BN * 128 B.  The persistent attention
//     kernel's knock-out experiments show that per-SM TMA ingest (~44 B/clk), not the tensor pipe, bounds
//     the projections of this path, which is what the pair halves for the weight operand.
//   * both CTAs' loads signal the LEADER's `full` barrier (cp.async.bulk.tensor ... .cta_group::2 with a shared::cluster
//     barrier address); tcgen05.commit ... multicast::cluster releases the stage in both CTAs
//   * each CTA's TMEM receives its own 128 accumulator rows; the epilogue (bias, cast, swizzled smem, TMA store) is per CTA
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int G2_BK = 64;
constexpr int G2_THREADS = 192;
constexpr int G2_A_BYTES = 128 * G2_BK * 2;

template <int BN>
struct Gemm2Cfg {
  static constexpr int W_BYTES = (BN / 2) * G2_BK * 2;          // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = G2_A_BYTES + W_BYTES;
  static constexpr int STAGES = BN >= 256 ? 5 : 6;
  static constexpr int EPI_TILE_BYTES = 32 * 128;
  static constexpr int EPI_BYTES = 4 * 2 * EPI_TILE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = BN <= 128 ? 128 : 256;
};

template <int BN, bool OUT_F32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                          const __grid_constant__ CUtensorMap tmD, const float* __restrict__ bias, long long strideBias,
                          int N, int K, int w_batched) {
  using Cfg = Gemm2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + Cfg::EPI_BYTES);
  uint64_t* empty = full + Cfg::STAGES;
  uint64_t* accum_full = empty + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n0 = blockIdx.y * BN;
  const int m0 = (blockIdx.x >> 1) * 256 + static_cast<int>(rank) * 128;
  const int batch = blockIdx.z;
  const int kblocks = (K + G2_BK - 1) / G2_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % Cfg::STAGES;
      const uint32_t ph = (kb / Cfg::STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
        const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);          // the leader's barrier
        if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes land on it
        tma_load_3d_2sm(a_dst, &tmA, bar, kb * G2_BK, m0, batch);
        tma_load_3d_2sm(a_dst + G2_A_BYTES, &tmW, bar, kb * G2_BK, n0 + static_cast<int>(rank) * (BN / 2), w_batched ? batch : 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint8_t* a_src = smem + s * Cfg::STAGE_BYTES;
          const uint64_t da = umma_desc_sw128(a_src);
          const uint64_t dw = umma_desc_sw128(a_src + G2_A_BYTES);
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k) umma_bf16_ss_2sm(tmem_acc, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty[s]);
          if (kb == kblocks - 1) umma_commit_2sm(accum_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    uint8_t* stg = epi + q * 2 * Cfg::EPI_TILE_BYTES;
    const float* bias_b = bias ? bias + strideBias * batch : nullptr;
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int row = lane;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= N) break;
      uint32_t v[32];
      tmem_ld_x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      uint8_t* buf = stg + (c & 1) * Cfg::EPI_TILE_BYTES;
      if (c >= 2) {
        if (elect_one()) bulk_wait_read<1>();
        __syncwarp();
      }
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float b = 0.f;
        if (bias_b) {
          const int n = n0 + c * 32 + j;
          b = (n < N) ? __ldg(bias_b + n) : 0.f;
        }
        f[j] = __uint_as_float(v[j]) + b;
      }
      if constexpr (OUT_F32) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const int pch = ch ^ (row & 7);
          st_shared_v4(buf + row * 128 + pch * 16, __float_as_uint(f[ch * 4 + 0]), __float_as_uint(f[ch * 4 + 1]),
                       __float_as_uint(f[ch * 4 + 2]), __float_as_uint(f[ch * 4 + 3]));
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int pch = ch ^ ((row >> 1) & 3);
          st_shared_v4(buf + row * 64 + pch * 16, pack_bf16x2(f[ch * 8 + 0], f[ch * 8 + 1]),
                       pack_bf16x2(f[ch * 8 + 2], f[ch * 8 + 3]), pack_bf16x2(f[ch * 8 + 4], f[ch * 8 + 5]),
                       pack_bf16x2(f[ch * 8 + 6], f[ch * 8 + 7]));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_store_3d(&tmD, buf, n0 + c * 32, m0 + q * 32, batch);
        bulk_commit();
      }
      __syncwarp();
    }
    if (elect_one()) bulk_wait_read<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // neither CTA may free TMEM / exit while the pair's MMAs or signals are in flight
  tc_fence_after();
  if (warp == 1) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_acc);
}

template <int BN, bool OUT_F32>
static int launch_g2(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmD, const float* bias,
                     long long strideBias, long long M, int N, int K, int batch, int w_batched, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  auto kern = gemm2_bf16_tcgen05_kernel<BN, OUT_F32>;
  static bool attr_done = false;
  if (!attr_done) {
    PV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid(static_cast<unsigned>(2 * ((M + 255) / 256)), (N + BN - 1) / BN, batch);
  kern<<<grid, G2_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmW, tmD, bias, strideBias, N, K, w_batched);
  PV_LAUNCHED();
  return PV_OK;
}

// same contract as gemm_bf16 (pv_gemm.cu); requires N % 16 == 0 handled by TMA clipping (W half boxes may be partly OOB)
int gemm2_bf16(const void* A, const void* W, const float* bias, void* D, bool out_f32, long long M, long long N, long long K,
               long long batch, long long lda, long long ldw, long long ldd, long long strideA, long long strideW,
               long long strideBias, long long strideD, cudaStream_t stream) {
  const int oe = out_f32 ? 4 : 2;
  const int w_batched = strideW != 0;
  const int bn = (N % 256 == 0) ? 256 : (N % 160 == 0) ? 160 : (N > 160 ? 256 : 160);
  PV_REQUIRE(2 * ((M + 255) / 256) <= 2147483647LL && batch <= 65535, "grid too large");
  CUtensorMap tmA, tmW, tmD;
  if (make_tmap_3d(&tmA, A, 2, K, M, batch, lda * 2, (batch > 1 ? strideA : lda * M) * 2, G2_BK, 128, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmW, W, 2, K, N, w_batched ? batch : 1, ldw * 2, (w_batched ? strideW : ldw * N) * 2, G2_BK, bn / 2, 1, Swz::B128))
    return PV_ERR_CUDA;
  if (make_tmap_3d(&tmD, D, oe, N, M, batch, ldd * oe, (batch > 1 ? strideD : ldd * M) * oe, 32, 32, 1, out_f32 ? Swz::B128 : Swz::B64))
    return PV_ERR_CUDA;
  if (bn == 256)
    return out_f32 ? launch_g2<256, true>(tmA, tmW, tmD, bias, strideBias, M, (int)N, (int)K, (int)batch, w_batched, stream)
                   : launch_g2<256, false>(tmA, tmW, tmD, bias, strideBias, M, (int)N, (int)K, (int)batch, w_batched, stream);
  return out_f32 ? launch_g2<160, true>(tmA, tmW, tmD, bias, strideBias, M, (int)N, (int)K, (int)batch, w_batched, stream)
                 : launch_g2<160, false>(tmA, tmW, tmD, bias, strideBias, M, (int)N, (int)K, (int)batch, w_batched, stream);
}

}  // namespace pv
