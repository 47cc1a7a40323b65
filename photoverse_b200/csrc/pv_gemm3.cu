// photoverse_b200 -- persistent CTA-pair (cta_group::2) projection GEMM for the out projection of the processor
// (reference models/attention_processor.py:423):   D[M, N] = A[M, K] * W[N, K]^T + bias,  bf16 in / bf16 out,
// N % 160 == 0, K % 64 == 0.
//
// It is the projection pipeline of pv_attn4.cu with a plain epilogue instead of the attention: a pair of CTAs walks
// 256-row x 160-column output tiles of ONE column group g (so the pair's W slice never changes),
//   * the W slice [160 x K] is split across the pair (80 rows each) and, for K <= 640, stays RESIDENT in shared memory:
//     only A is streamed (16 KB per K-block per CTA) through a deep TMA ring;  K = 1280 streams 10 KB W halves as well
//   * ONE tcgen05.mma stream of M = 256 issued by the leader into double-buffered TMEM accumulators (2 x 160 columns):
//     the MMAs of tile i+1 run under the epilogue of tile i
//   * epilogue warps (2 x 4 per CTA): TMEM -> +bias -> bf16 -> per-warp staging tile -> TMA store
// The single-CTA kernel of pv_gemm.cu spends most of a K = 320 tile in its prologue / epilogue (5 K-blocks) and is bound
// by TMA ingest (36 KB per K-block); this one has neither problem.
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int G3_BN = 160;
constexpr int G3_BK = 64;
constexpr int G3_THREADS = 384;          // warp 0 TMA, warp 1 MMA (leader), warps 4..11 epilogue
constexpr int G3_A_BYTES = 128 * G3_BK * 2;
constexpr int G3_WH_BYTES = (G3_BN / 2) * G3_BK * 2;
constexpr int G3_MAX_STAGES = 8;

template <int KB_RES>                    // resident K-blocks of the W half (0: streamed)
struct Gemm3Cfg {
  static constexpr bool WSTAT = KB_RES > 0;
  static constexpr int STAGE_BYTES = WSTAT ? G3_A_BYTES : G3_A_BYTES + G3_WH_BYTES;
  static constexpr int W_RES_BYTES = KB_RES * G3_WH_BYTES;
  static constexpr int OST_WARP_BYTES = 32 * 80 * 2;
  static constexpr int STAGES = KB_RES == 5 ? 8 : KB_RES == 10 ? 5 : 7;
  static constexpr int OFF_W = STAGES * STAGE_BYTES;
  static constexpr int OFF_OST = OFF_W + W_RES_BYTES;
  static constexpr int OFF_BIAS = OFF_OST + 8 * OST_WARP_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + G3_BN * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int KB_RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G3_THREADS, 1)
gemm3_pair_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                             const __grid_constant__ CUtensorMap tmD, const float* __restrict__ bias, int G, int V, int K) {
  using Cfg = Gemm3Cfg<KB_RES>;
  constexpr bool WSTAT = Cfg::WSTAT;
  constexpr int nst = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sbias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);   // [STAGES] (leader) both CTAs' stages landed
  uint64_t* empty = full + G3_MAX_STAGES;                              // [STAGES] multicast commit
  uint64_t* acc_full = empty + G3_MAX_STAGES;                          // [2] multicast commit
  uint64_t* slot_free = acc_full + 2;                                  // [2] (leader) count 16
  uint64_t* w_full = slot_free + 2;                                    // 1   (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int kblocks = K / G3_BK;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int g = pair % G;
  const int r_pair = pair / G;
  const int npair_g = (npairs - g + G - 1) / G;
  const int u0 = static_cast<int>((static_cast<long long>(V) * r_pair) / npair_g);
  const int u1 = static_cast<int>((static_cast<long long>(V) * (r_pair + 1)) / npair_g);
  const int n0 = g * G3_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < G3_MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&slot_free[i], 16);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  for (int c = threadIdx.x; c < G3_BN; c += G3_THREADS) sbias[c] = bias ? bias[n0 + c] : 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // Everything above (and the resident-W load below, issued before the wait) reads only static weights: it overlaps the
  // tail of the previous kernel in the stream.  A, D are touched only after the predecessor has completed.
  if constexpr (WSTAT) {
    if (warp == 0) {
      if (u0 < u1 && elect_one()) {
        const uint32_t bar = mapa_u32(smem_u32(w_full), 0);
        if (rank == 0) mbar_expect_tx(w_full, 2 * Cfg::W_RES_BYTES);
        for (int kb = 0; kb < KB_RES; ++kb)
          tma_load_3d_2sm(smem + Cfg::OFF_W + kb * G3_WH_BYTES, &tmW, bar, kb * G3_BK, n0 + static_cast<int>(rank) * (G3_BN / 2), 0);
      }
      __syncwarp();
    }
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    uint32_t it = 0;
    for (int u = u0; u < u1; ++u) {
      const int m0 = (2 * u + static_cast<int>(rank)) * 128;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % nst;
        const uint32_t ph = (it / nst) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + s * Cfg::STAGE_BYTES;
          const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
          tma_load_3d_2sm(a_dst, &tmA, bar, kb * G3_BK, m0, 0);
          if constexpr (!WSTAT)
            tma_load_3d_2sm(a_dst + G3_A_BYTES, &tmW, bar, kb * G3_BK, n0 + static_cast<int>(rank) * (G3_BN / 2), 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader): M = 256, N = 160 for both CTAs =====================
    constexpr uint32_t idesc = umma_idesc_bf16(256, G3_BN);
    uint32_t it = 0;
    int i = 0;
    if constexpr (WSTAT) {
      if (u0 < u1) mbar_wait(w_full, 0);
    }
    for (int u = u0; u < u1; ++u, ++i) {
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
          const uint64_t dw = umma_desc_sw128(WSTAT ? smem + Cfg::OFF_W + kb * G3_WH_BYTES : a_src + G3_A_BYTES);
#pragma unroll
          for (int k = 0; k < G3_BK / 16; ++k) umma_bf16_ss_2sm(tmem + slot * G3_BN, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty[s]);
          if (kb == kblocks - 1) umma_commit_2sm(&acc_full[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): group w handles columns [80 w, 80 w + 80) =====================
    const int w = (warp - 4) >> 2;
    const int q = warp & 3;
    const uint32_t tlane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* ost = smem + Cfg::OFF_OST + (warp - 4) * Cfg::OST_WARP_BYTES;
    const float* bw = sbias + 80 * w;
    int i = 0;
    for (int u = u0; u < u1; ++u, ++i) {
      const int slot = i & 1;
      const int m0 = (2 * u + static_cast<int>(rank)) * 128;
      mbar_wait(&acc_full[slot], (i >> 1) & 1);
      tc_fence_after();
      uint32_t a[32], b2[32], c16[16];
      const uint32_t src = tlane + slot * G3_BN + 80 * w;
      tmem_ld_x32(src, a);
      tmem_ld_x32(src + 32, b2);
      tmem_ld_x16(src + 64, c16);
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
      put(a, 0, 32);
      put(b2, 32, 32);
      put(c16, 64, 16);
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        tma_store_3d(&tmD, ost, n0 + 80 * w, m0 + q * 32, 0);
        bulk_commit();
      }
      __syncwarp();
    }
    if (elect_one()) bulk_wait_read<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm<512>(tmem);
}

template <int KB_RES>
static int launch_g3(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmD, const float* bias, int G, int V,
                     int K, cudaStream_t stream) {
  using Cfg = Gemm3Cfg<KB_RES>;
  auto kern = gemm3_pair_persistent_kernel<KB_RES>;
  PV_CUDA(set_max_smem_once(kern, Cfg::SMEM_BYTES));
  const long long units = static_cast<long long>(G) * V;
  const long long max_pairs = sm_count() / 2;
  const int npairs = static_cast<int>(units < max_pairs ? units : max_pairs);
  PV_CUDA(launch_pdl(kern, dim3(2 * npairs), dim3(G3_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmW, tmD, bias, G, V, K));
  PV_LAUNCHED();
  return PV_OK;
}

// D[M,N] (bf16, row stride ldd) = A[M,K] (lda) * W[N,K]^T (ldw) + bias ; N % 160 == 0, K % 64 == 0
int gemm3_bf16(const void* A, const void* W, const float* bias, void* D, long long M, long long N, long long K, long long lda,
               long long ldw, long long ldd, cudaStream_t stream) {
  PV_REQUIRE(M > 0 && N % G3_BN == 0 && K % G3_BK == 0 && N > 0 && K > 0, "gemm3: need N %% 160 == 0 and K %% 64 == 0 (N=%lld K=%lld)", N, K);
  PV_REQUIRE((lda * 2) % 16 == 0 && (ldw * 2) % 16 == 0 && (ldd * 2) % 16 == 0, "row strides must be 16-byte multiples");
  CUtensorMap tmA, tmW, tmD;
  if (make_tmap_3d(&tmA, A, 2, K, M, 1, lda * 2, lda * M * 2, G3_BK, 128, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmW, W, 2, K, N, 1, ldw * 2, ldw * N * 2, G3_BK, G3_BN / 2, 1, Swz::B128)) return PV_ERR_CUDA;
  if (make_tmap_3d(&tmD, D, 2, N, M, 1, ldd * 2, ldd * M * 2, 80, 32, 1, Swz::None)) return PV_ERR_CUDA;
  const int G = static_cast<int>(N / G3_BN);
  const long long V = (M + 255) / 256;
  PV_REQUIRE(V * G < (1ll << 30), "too many tiles");
  if (K == 320) return launch_g3<5>(tmA, tmW, tmD, bias, G, static_cast<int>(V), static_cast<int>(K), stream);
  if (K == 640) return launch_g3<10>(tmA, tmW, tmD, bias, G, static_cast<int>(V), static_cast<int>(K), stream);
  return launch_g3<0>(tmA, tmW, tmD, bias, G, static_cast<int>(V), static_cast<int>(K), stream);
}

}  // namespace pv
