// photoverse_b200 -- both LoRA factor gradients of one projection in ONE pass over the activations.
//
// peft 0.10.0 lora.Linear (train.py:348-354):  y = W x + s B (A x)   =>   with G = dL/dy [M, out], X [M, in]:
//     T = X A^T  [M, r]        U = G B  [M, r]
//     dB = s G^T T  [out, r]   dA = s U^T X  [r, in]
// Round 1 ran this as two skinny GEMMs, a transpose and two split-M weight-gradient kernels with their reductions (8
// launches per projection, ~380 per training step, X and G read four times).  Here a block walks a contiguous range of
// 32-row tiles: a tile of X and G is staged once in shared memory (bf16 / fp32 as given), T and U are formed by warp-level
// dot products (A, B^T resident in shared memory), and every thread carries its columns' slices of dA and dB in registers
// across all tiles of the block.  The per-block partials are summed in a fixed order by reduce_partials_kernel
// (deterministic two-pass reduction, like every other reduction of the training path).  HBM-bound: X and G are read once.
// Ranks up to 16; the shipped recipe's rank 128 (prepare_dataset_and_train.sh:2) is tensor-core work and goes through
// pv_linear_fwd / pv_linear_bwd_weight instead.
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

template <typename T> __device__ __forceinline__ float lb_ld(const T* p);
template <> __device__ __forceinline__ float lb_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float lb_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

constexpr int LB_ROWS = 32;          // rows per tile (16 / 8 when the 32-row tiles of wide layers do not fit shared memory)
constexpr int LB_THREADS = 256;
constexpr int LB_MAXCOLS = 5;        // columns per thread: in, out <= 1280

// out[i] = alpha * sum_z part[z][i], z in a FIXED order (bit-reproducible): a block owns 32 outputs, its 8 warps sum
// interleaved slices of the z range (independent loads in flight instead of one serial chain of `splits` L2 round trips),
// and warp 0 adds the 8 slice sums in warp order.
__global__ void __launch_bounds__(256)
lora_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int splits, float alpha) {
  __shared__ float red[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = static_cast<long long>(blockIdx.x) * 32 + lane;
  float s = 0.f;
  if (i < n) {
    float s0 = 0.f, s1 = 0.f;
    int z = warp;
    for (; z + 8 < splits; z += 16) {
      s0 += part[static_cast<size_t>(z) * n + i];
      s1 += part[static_cast<size_t>(z + 8) * n + i];
    }
    if (z < splits) s0 += part[static_cast<size_t>(z) * n + i];
    s = s0 + s1;
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && i < n) {
    float t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += red[w][lane];
    out[i] = alpha * t;
  }
}

// part[block] = [ dA partial (R x in) | dB partial (out x R) ]
template <int R, typename T>
__global__ void __launch_bounds__(LB_THREADS)
lora_wgrad_partial_kernel(const T* __restrict__ X, const T* __restrict__ G, const float* __restrict__ A,
                          const float* __restrict__ Bm, float* __restrict__ part, long long M, int in_f, int out_f, int r,
                          long long ldx, long long ldg, long long tiles_per_block, int rows) {
  extern __shared__ __align__(16) uint8_t smem_lb[];
  float* As = reinterpret_cast<float*>(smem_lb);                 // [R][in_f]     (rows >= r zero)
  float* Bs = As + static_cast<size_t>(R) * in_f;                // [R][out_f]    B^T (rows >= r zero)
  float* Ts = Bs + static_cast<size_t>(R) * out_f;               // [LB_ROWS][R]
  float* Us = Ts + LB_ROWS * R;                                  // [LB_ROWS][R]
  T* Xs = reinterpret_cast<T*>(Us + LB_ROWS * R);                // [rows][in_f]
  T* Gs = Xs + static_cast<size_t>(rows) * in_f;                 // [rows][out_f]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < R * in_f; i += LB_THREADS) {
    const int j = i / in_f, c = i - j * in_f;
    As[i] = j < r ? A[static_cast<size_t>(j) * in_f + c] : 0.f;
  }
  for (int i = threadIdx.x; i < R * out_f; i += LB_THREADS) {
    const int j = i / out_f, c = i - j * out_f;
    Bs[i] = j < r ? Bm[static_cast<size_t>(c) * r + j] : 0.f;
  }
  float da[LB_MAXCOLS][R], db[LB_MAXCOLS][R];
#pragma unroll
  for (int k = 0; k < LB_MAXCOLS; ++k)
#pragma unroll
    for (int j = 0; j < R; ++j) { da[k][j] = 0.f; db[k][j] = 0.f; }

  const long long tile0 = static_cast<long long>(blockIdx.x) * tiles_per_block;
  for (long long tt = 0; tt < tiles_per_block; ++tt) {
    const long long m0 = (tile0 + tt) * rows;
    if (m0 >= M) break;
    __syncthreads();                                             // previous tile fully consumed (and As / Bs visible)
    // ---- stage the tile (rows past M: zeros) ----
    for (int i = threadIdx.x; i < rows * in_f; i += LB_THREADS) {
      const int rr = i / in_f, c = i - rr * in_f;
      Xs[i] = (m0 + rr < M) ? X[(m0 + rr) * ldx + c] : T(0.f);
    }
    for (int i = threadIdx.x; i < rows * out_f; i += LB_THREADS) {
      const int rr = i / out_f, c = i - rr * out_f;
      Gs[i] = (m0 + rr < M) ? G[(m0 + rr) * ldg + c] : T(0.f);
    }
    __syncthreads();
    // ---- T = X A^T, U = G B : warp w owns rows w, w + 8, w + 16, w + 24 ----
#pragma unroll 1
    for (int rr = warp; rr < rows; rr += LB_THREADS / 32) {
      float t[R], u[R];
#pragma unroll
      for (int j = 0; j < R; ++j) { t[j] = 0.f; u[j] = 0.f; }
      for (int c = lane; c < in_f; c += 32) {
        const float x = lb_ld(Xs + static_cast<size_t>(rr) * in_f + c);
#pragma unroll
        for (int j = 0; j < R; ++j) t[j] = fmaf(x, As[j * in_f + c], t[j]);
      }
      for (int c = lane; c < out_f; c += 32) {
        const float g = lb_ld(Gs + static_cast<size_t>(rr) * out_f + c);
#pragma unroll
        for (int j = 0; j < R; ++j) u[j] = fmaf(g, Bs[j * out_f + c], u[j]);
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          t[j] += __shfl_xor_sync(0xffffffffu, t[j], o);
          u[j] += __shfl_xor_sync(0xffffffffu, u[j], o);
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) { Ts[rr * R + j] = t[j]; Us[rr * R + j] = u[j]; }
      }
    }
    __syncthreads();
    // ---- dA[j][c] += sum_rows U[row][j] X[row][c] ; dB[c][j] += sum_rows G[row][c] T[row][j] ----
#pragma unroll
    for (int k = 0; k < LB_MAXCOLS; ++k) {
      const int c = threadIdx.x + k * LB_THREADS;
      if (c < in_f) {
#pragma unroll 4
        for (int rr = 0; rr < rows; ++rr) {
          const float x = lb_ld(Xs + static_cast<size_t>(rr) * in_f + c);
#pragma unroll
          for (int j = 0; j < R; ++j) da[k][j] = fmaf(Us[rr * R + j], x, da[k][j]);
        }
      }
      if (c < out_f) {
#pragma unroll 4
        for (int rr = 0; rr < rows; ++rr) {
          const float g = lb_ld(Gs + static_cast<size_t>(rr) * out_f + c);
#pragma unroll
          for (int j = 0; j < R; ++j) db[k][j] = fmaf(g, Ts[rr * R + j], db[k][j]);
        }
      }
    }
  }
  // ---- this block's partials ----
  float* pa = part + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(r) * in_f + static_cast<size_t>(out_f) * r);
  float* pb = pa + static_cast<size_t>(r) * in_f;
#pragma unroll
  for (int k = 0; k < LB_MAXCOLS; ++k) {
    const int c = threadIdx.x + k * LB_THREADS;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (j < r) {
        if (c < in_f) pa[static_cast<size_t>(j) * in_f + c] = da[k][j];
        if (c < out_f) pb[static_cast<size_t>(c) * r + j] = db[k][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bf16: the same four products on the tensor cores (warp-level mma.sync.m16n8k16 -- n = 8 IS the LoRA rank tile; the
// products are far too skinny for tcgen05 tiles).  64-row tiles; X / G tiles, A and B^T live in shared memory as bf16:
//   phase 1  warps 0-3: T = X A^T   warps 4-7: U = G B     (16 rows per warp, fp32 accumulate, rounded to bf16 like the
//            activations of the GEMM route)
//   phase 2  dB^T-free forms dB[out, r] += G^T T and dA^T[in, r] += X^T U: A operands are TRANSPOSED ldmatrix loads of the
//            row-major G / X tiles, B operands transposed loads of T / U; the 16-row output tiles (out / 16 + in / 16 of
//            them) are dealt round-robin to the 8 warps and stay in registers across all tiles of the block.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lb_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void lb_ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void lb_ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void lb_ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void lb_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int LM_ROWS = 64;          // rows per tile (32 / 16 for wide layers: lora_mma_rows)
constexpr int LM_MAXT = 10;          // 16-row output tiles per warp and matrix: in, out <= 1280 -> 80 tiles / 8 warps

// NT = rank tiles of 8 (rank <= 8: 1, <= 16: 2); in_f, out_f multiples of 16
template <int NT>
__global__ void __launch_bounds__(LB_THREADS)
lora_wgrad_mma_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ G, const float* __restrict__ A,
                      const float* __restrict__ Bm, float* __restrict__ part, long long M, int in_f, int out_f, int r,
                      long long ldx, long long ldg, long long tiles_per_block, int rows) {
  constexpr int R = 8 * NT;
  constexpr int PT = R + 8;                                      // pitch of the T / U tiles (elements)
  extern __shared__ __align__(16) uint8_t smem_lb[];
  const int pin = in_f + 8, pout = out_f + 8;                    // row pitches: odd multiples of 16 bytes
  __nv_bfloat16* Xs = reinterpret_cast<__nv_bfloat16*>(smem_lb);             // [64][pin]
  __nv_bfloat16* Gs = Xs + rows * pin;                                         // [rows][pout]
  __nv_bfloat16* As = Gs + rows * pout;                                        // [R][pin]    A   (n = rank, k = in)
  __nv_bfloat16* Bs = As + R * pin;                                            // [R][pout]   B^T (n = rank, k = out)
  __nv_bfloat16* Ts = Bs + R * pout;                                           // [64][PT]
  __nv_bfloat16* Us = Ts + LM_ROWS * PT;                                       // [64][PT]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  for (int i = threadIdx.x; i < R * in_f; i += LB_THREADS) {
    const int j = i / in_f, c = i - j * in_f;
    As[j * pin + c] = __float2bfloat16_rn(j < r ? A[static_cast<size_t>(j) * in_f + c] : 0.f);
  }
  for (int i = threadIdx.x; i < R * out_f; i += LB_THREADS) {
    const int j = i / out_f, c = i - j * out_f;
    Bs[j * pout + c] = __float2bfloat16_rn(j < r ? Bm[static_cast<size_t>(c) * r + j] : 0.f);
  }
  const int mt_b = out_f / 16, mt_a = in_f / 16;                 // output tiles of dB and of dA^T
  float accb[LM_MAXT][NT][4], acca[LM_MAXT][NT][4];
#pragma unroll
  for (int i = 0; i < LM_MAXT; ++i)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { accb[i][n][e] = 0.f; acca[i][n][e] = 0.f; }

  const long long tile0 = static_cast<long long>(blockIdx.x) * tiles_per_block;
  for (long long tt = 0; tt < tiles_per_block; ++tt) {
    const long long m0 = (tile0 + tt) * rows;
    if (m0 >= M) break;
    __syncthreads();
    // ---- stage the tile: 16-byte loads, rows past M zero ----
    for (int i = threadIdx.x; i < rows * (in_f / 8); i += LB_THREADS) {
      const int rr = i / (in_f / 8), c = (i - rr * (in_f / 8)) * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m0 + rr < M) v = *reinterpret_cast<const uint4*>(X + (m0 + rr) * ldx + c);
      *reinterpret_cast<uint4*>(Xs + rr * pin + c) = v;
    }
    for (int i = threadIdx.x; i < rows * (out_f / 8); i += LB_THREADS) {
      const int rr = i / (out_f / 8), c = (i - rr * (out_f / 8)) * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (m0 + rr < M) v = *reinterpret_cast<const uint4*>(G + (m0 + rr) * ldg + c);
      *reinterpret_cast<uint4*>(Gs + rr * pout + c) = v;
    }
    __syncthreads();
    // ---- phase 1: T (warps 0-3) / U (warps 4-7), 16 rows per warp ----
    if ((warp & 3) * 16 < rows) {
      const bool is_u = warp >= 4;
      const int r0 = (warp & 3) * 16;
      const __nv_bfloat16* src = is_u ? Gs : Xs;
      const __nv_bfloat16* wsm = is_u ? Bs : As;
      const int pitch = is_u ? pout : pin;
      const int ksteps = (is_u ? out_f : in_f) / 16;
      float acc[NT][4];
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
      const uint32_t a_addr = smem_u32(src + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * pitch + (lane >> 4) * 8);
      const uint32_t b_addr = smem_u32(wsm + (lane & 7) * pitch + ((lane >> 3) & 1) * 8);     // x2: lanes 0-15 address the two 8x8
      for (int k = 0; k < ksteps; ++k) {
        uint32_t a0, a1, a2, a3;
        lb_ldsm_x4(a_addr + k * 32, a0, a1, a2, a3);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          uint32_t b0, b1;
          lb_ldsm_x2(b_addr + (n * 8 * pitch + k * 16) * 2, b0, b1);
          lb_mma(acc[n], a0, a1, a2, a3, b0, b1);
        }
      }
      __nv_bfloat16* dst = is_u ? Us : Ts;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        *reinterpret_cast<uint32_t*>(dst + (r0 + g) * PT + n * 8 + 2 * t) = pack_bf16x2(acc[n][0], acc[n][1]);
        *reinterpret_cast<uint32_t*>(dst + (r0 + g + 8) * PT + n * 8 + 2 * t) = pack_bf16x2(acc[n][2], acc[n][3]);
      }
    }
    __syncthreads();
    // ---- phase 2: dB += G^T T, dA^T += X^T U (k = the tile's 64 rows) ----
    {
      // A (16 m x 16 k) = transposed load of the [row = k][col = m] tile; B (16 k x 8 n) = transposed load of [row = k][n]
      const int a_row = (lane & 7) + (lane >> 4) * 8, a_col = ((lane >> 3) & 1) * 8;
      const int b_row = (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int i = 0; i < LM_MAXT; ++i) {
        const int mt = warp + 8 * i;
        if (mt < mt_b) {
          const uint32_t ga = smem_u32(Gs + a_row * pout + mt * 16 + a_col);
          const uint32_t tb = smem_u32(Ts + b_row * PT);
          for (int k = 0; k < rows / 16; ++k) {
            uint32_t a0, a1, a2, a3;
            lb_ldsm_x4_t(ga + (k * 16 * pout) * 2, a0, a1, a2, a3);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              uint32_t b0, b1;
              lb_ldsm_x2_t(tb + (k * 16 * PT + n * 8) * 2, b0, b1);
              lb_mma(accb[i][n], a0, a1, a2, a3, b0, b1);
            }
          }
        }
        if (mt < mt_a) {
          const uint32_t xa = smem_u32(Xs + a_row * pin + mt * 16 + a_col);
          const uint32_t ub = smem_u32(Us + b_row * PT);
          for (int k = 0; k < rows / 16; ++k) {
            uint32_t a0, a1, a2, a3;
            lb_ldsm_x4_t(xa + (k * 16 * pin) * 2, a0, a1, a2, a3);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
              uint32_t b0, b1;
              lb_ldsm_x2_t(ub + (k * 16 * PT + n * 8) * 2, b0, b1);
              lb_mma(acca[i][n], a0, a1, a2, a3, b0, b1);
            }
          }
        }
      }
    }
  }
  // ---- partials: [ dA (r x in) | dB (out x r) ]; accumulator (row g / g + 8 of the 16-row tile, rank columns 2 t, 2 t + 1) ----
  float* pa = part + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(r) * in_f + static_cast<size_t>(out_f) * r);
  float* pb = pa + static_cast<size_t>(r) * in_f;
#pragma unroll
  for (int i = 0; i < LM_MAXT; ++i) {
    const int mt = warp + 8 * i;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = mt * 16 + g + (e >= 2 ? 8 : 0);
        const int j = n * 8 + 2 * t + (e & 1);
        if (j < r) {
          if (mt < mt_b) pb[static_cast<size_t>(row) * r + j] = accb[i][n][e];
          if (mt < mt_a) pa[static_cast<size_t>(j) * in_f + row] = acca[i][n][e];
        }
      }
    }
  }
}

static size_t lora_mma_smem(int in_f, int out_f, int R, int rows) {
  return (static_cast<size_t>(rows + R) * (in_f + 8 + out_f + 8) + 2 * LM_ROWS * (R + 8)) * 2;
}
static int lora_mma_rows(int in_f, int out_f, int R) {           // largest row tile that fits shared memory; 0: none
  for (int rows = LM_ROWS; rows >= 16; rows /= 2)
    if (lora_mma_smem(in_f, out_f, R, rows) <= 227u * 1024u) return rows;
  return 0;
}
static bool lora_mma_ok(int in_f, int out_f, int r, long long ldx, long long ldg) {
  return in_f % 16 == 0 && out_f % 16 == 0 && in_f <= 16 * 8 * LM_MAXT && out_f <= 16 * 8 * LM_MAXT && ldx % 8 == 0 && ldg % 8 == 0 &&
         lora_mma_rows(in_f, out_f, r <= 8 ? 8 : 16) > 0;
}

static size_t lora_smem(int in_f, int out_f, int R, int rows, int esize) {
  return (static_cast<size_t>(R) * (in_f + out_f) + 2 * LB_ROWS * R) * 4 + static_cast<size_t>(rows) * (in_f + out_f) * esize;
}
static int lora_rows(int in_f, int out_f, int r, int esize) {      // largest row tile that fits; 0: none
  const int R = r <= 8 ? 8 : 16;
  for (int rows = LB_ROWS; rows >= 8; rows /= 2)
    if (lora_smem(in_f, out_f, R, rows, esize) <= 227u * 1024u) return rows;
  return 0;
}
static long long lora_blocks(long long M, int rows) {
  const long long tiles = (M + rows - 1) / rows;
  const long long want = 2ll * sm_count();                       // two resident blocks per SM keep HBM busy
  return tiles < want ? tiles : want;
}

bool lora_bwd_supported(int in_f, int out_f, int r) {
  return r >= 1 && r <= 16 && in_f >= 1 && out_f >= 1 && in_f <= LB_MAXCOLS * LB_THREADS && out_f <= LB_MAXCOLS * LB_THREADS &&
         lora_rows(in_f, out_f, r, 4) > 0;
}

long long lora_bwd_ws_bytes(long long M, int in_f, int out_f, int r) {
  // sized for the smaller (bf16) row tile count bound: blocks <= 2 x SMs either way
  return 2ll * sm_count() * (static_cast<long long>(r) * in_f + static_cast<long long>(out_f) * r) * 4;
}

template <int R, typename T>
static int launch_lora(const void* X, const void* G, const float* A, const float* Bm, float scaling, float* dAB, void* ws,
                       long long M, int in_f, int out_f, int r, long long ldx, long long ldg, cudaStream_t stream) {
  auto kern = lora_wgrad_partial_kernel<R, T>;
  const int rows = lora_rows(in_f, out_f, r, static_cast<int>(sizeof(T)));
  PV_REQUIRE(rows > 0, "shared memory budget (in=%d out=%d)", in_f, out_f);
  const size_t smem = lora_smem(in_f, out_f, R, rows, static_cast<int>(sizeof(T)));
  PV_CUDA(set_max_smem_once(kern, static_cast<int>(smem)));
  const long long nb = lora_blocks(M, rows);
  const long long tiles = (M + rows - 1) / rows;
  const long long tpb = (tiles + nb - 1) / nb;
  float* part = static_cast<float*>(ws);
  kern<<<static_cast<unsigned>(nb), LB_THREADS, smem, stream>>>(static_cast<const T*>(X), static_cast<const T*>(G), A, Bm, part, M, in_f,
                                                                out_f, r, ldx, ldg, tpb, rows);
  PV_LAUNCHED();
  const long long n = static_cast<long long>(r) * in_f + static_cast<long long>(out_f) * r;
  lora_reduce_kernel<<<static_cast<unsigned>((n + 31) / 32), 256, 0, stream>>>(part, dAB, n, static_cast<int>(nb), scaling);
  PV_LAUNCHED();
  return PV_OK;
}

// dAB = [ dA (r x in) | dB (out x r) ] fp32, contiguous.
int lora_bwd(bool bf16, const void* X, const void* G, const float* A, const float* Bm, float scaling, float* dAB, void* ws,
             long long M, int in_f, int out_f, int r, long long ldx, long long ldg, cudaStream_t stream) {
  PV_REQUIRE(M > 0 && lora_bwd_supported(in_f, out_f, r), "lora_bwd: need 1 <= r <= 16 and in, out <= %d (in=%d out=%d r=%d)",
             LB_MAXCOLS * LB_THREADS, in_f, out_f, r);
  PV_REQUIRE(ws != nullptr, "workspace required (pv_lora_bwd_ws_bytes)");
  if (bf16 && lora_mma_ok(in_f, out_f, r, ldx, ldg) &&
      (reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(G)) % 16 == 0) {
    const int R = r <= 8 ? 8 : 16;
    const int rows = lora_mma_rows(in_f, out_f, R);
    const size_t smem = lora_mma_smem(in_f, out_f, R, rows);
    const long long tiles = (M + rows - 1) / rows;
    // two resident blocks per SM (one stages its tile while the other computes) where registers and shared memory allow
    const long long want = static_cast<long long>(sm_count()) * ((R == 8 && smem <= 112u * 1024u) ? 2 : 1);
    const long long nb = tiles < want ? tiles : want;
    const long long tpb = (tiles + nb - 1) / nb;
    float* part = static_cast<float*>(ws);
    if (R == 8) {
      PV_CUDA(set_max_smem_once(lora_wgrad_mma_kernel<1>, static_cast<int>(smem)));
      lora_wgrad_mma_kernel<1><<<static_cast<unsigned>(nb), LB_THREADS, smem, stream>>>(
          static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(G), A, Bm, part, M, in_f, out_f, r, ldx, ldg, tpb, rows);
    } else {
      PV_CUDA(set_max_smem_once(lora_wgrad_mma_kernel<2>, static_cast<int>(smem)));
      lora_wgrad_mma_kernel<2><<<static_cast<unsigned>(nb), LB_THREADS, smem, stream>>>(
          static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(G), A, Bm, part, M, in_f, out_f, r, ldx, ldg, tpb, rows);
    }
    PV_LAUNCHED();
    const long long n = static_cast<long long>(r) * in_f + static_cast<long long>(out_f) * r;
    lora_reduce_kernel<<<static_cast<unsigned>((n + 31) / 32), 256, 0, stream>>>(part, dAB, n, static_cast<int>(nb), scaling);
    PV_LAUNCHED();
    return PV_OK;
  }
  if (r <= 8)
    return bf16 ? launch_lora<8, __nv_bfloat16>(X, G, A, Bm, scaling, dAB, ws, M, in_f, out_f, r, ldx, ldg, stream)
                : launch_lora<8, float>(X, G, A, Bm, scaling, dAB, ws, M, in_f, out_f, r, ldx, ldg, stream);
  return bf16 ? launch_lora<16, __nv_bfloat16>(X, G, A, Bm, scaling, dAB, ws, M, in_f, out_f, r, ldx, ldg, stream)
              : launch_lora<16, float>(X, G, A, Bm, scaling, dAB, ws, M, in_f, out_f, r, ldx, ldg, stream);
}

}  // namespace pv
