// photoverse_b200 -- HBM-bound epilogues of the denoise loop that CALLS the hot path (the UNet evaluation
// of reference models/infer.py:103-116, diffusers' UNet2DConditionModel around the 16 attn2 layers), inference only.
//
//  * group_norm_nhwc : y = [SiLU](GroupNorm(x) * gamma + beta) on a channels-last activation [B, HW, C] (bf16 in / out,
//                      fp32 statistics).  The stock path of a channels-last backbone is five launches per norm --
//                      NHWC -> NCHW copy, row moments, normalise, SiLU, NCHW -> NHWC copy for the convolution that
//                      follows, 9 bytes of traffic per element where 3 suffice -- and was 16 % of the GPU time of a
//                      generation step (profiles/r02_bench_kernel_shares.csv).  Two launches here: per-(sample, pixel
//                      chunk) partial sums, then normalise + affine + SiLU; the second pass re-reads x from L2.
//                      Optional per-(sample, channel) addend `add[b, c]` applied on the fly in both passes: the bias of
//                      the convolution that produced x plus the time-embedding projection (ResnetBlock2D: conv1 ->
//                      + temb -> norm2), which the stock path adds in two more elementwise launches.
//  * add_bias_nhwc   : out = a + b + bias[c] -- residual sum of a block with the bias of its last convolution.
//  * layer_norm      : bf16 LayerNorm of [rows, C] (C <= 1280), one warp per row, statistics in registers.
//  * geglu           : y = h[:, :N] * gelu(h[:, N:]) (exact erf form) in one pass over the [M, 2N] projection.
//
// Every thread owns one 16-byte vector of 8 channels for the whole kernel (block = (C / 8) x k threads, k pixel lanes), so
// loads and stores are coalesced 16-byte accesses along C and the per-channel affine terms live in registers.  All
// reductions have a fixed order (bit-reproducible).
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

constexpr int GN_THREADS_MAX = 512;
constexpr int GN_MAX_G = 64;
constexpr int GN_UNROLL = 4;

__device__ __forceinline__ void bf16x8_unpack(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 bf16x8_pack(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&p);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Partial sums of (x - K_g) and (x - K_g)^2 per (sample, pixel chunk, group); K_g = x[b, 0, first channel of g] keeps
// E[x^2] - E[x]^2 away from cancellation when |mean| >> std.   part: [B][chunks][G][2] fp32.
__global__ void __launch_bounds__(GN_THREADS_MAX)
gn_stats_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ addv, float* __restrict__ part,
                     long long HW, int C, int G, int rows_per_cta, int k) {
  extern __shared__ float gn_smem[];               // [2][k][C]
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int pl = threadIdx.x / vecs;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const __nv_bfloat16* xb = x + static_cast<size_t>(b) * HW * C;
  // d = (x + add) - (x0 + add0) = x + (add - add0 - x0): one subtraction per element either way
  float shift[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j, c0 = (c / cpg) * cpg;
    shift[j] = __bfloat162float(xb[c0]);
    if (addv != nullptr) shift[j] += addv[static_cast<size_t>(b) * C + c0] - addv[static_cast<size_t>(b) * C + c];
  }
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long p1 = p0 + rows_per_cta < HW ? p0 + rows_per_cta : HW;
  const uint4* xv = reinterpret_cast<const uint4*>(xb);
  // GN_UNROLL independent 16-byte loads in flight per thread (one load per iteration left the kernels latency-bound:
  // 27 % of the DRAM rate under ncu)
  for (long long p = p0 + pl; p < p1; p += static_cast<long long>(GN_UNROLL) * k) {
    uint4 raw[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const long long pp = p + static_cast<long long>(u) * k;
      if (pp < p1) raw[u] = __ldg(xv + pp * vecs + v);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      if (p + static_cast<long long>(u) * k < p1) {
        float f[8];
        bf16x8_unpack(raw[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[j] - shift[j];
          s[j] += d;
          q[j] = fmaf(d, d, q[j]);
        }
      }
    }
  }
  float* ss = gn_smem;
  float* sq = gn_smem + k * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ss[pl * C + v * 8 + j] = s[j];
    sq[pl * C + v * 8 + j] = q[j];
  }
  __syncthreads();
  // channel c: sum over the k pixel lanes (fixed order), then group g: sum over its channels (fixed order)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, d = 0.f;
    for (int l = 0; l < k; ++l) {
      a += ss[l * C + c];
      d += sq[l * C + c];
    }
    ss[c] = a;                                     // lane 0's row is re-used for the per-channel totals: thread c only
    sq[c] = d;                                     // reads column c of every lane and writes column c of lane 0
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {       // (blocks of a tiny activation can be narrower than G)
    float a = 0.f, d = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += ss[c];
      d += sq[c];
    }
    float* o = part + ((static_cast<size_t>(b) * gridDim.x + blockIdx.x) * G + g) * 2;
    o[0] = a;
    o[1] = d;
  }
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS_MAX)
gn_apply_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ addv, const float* __restrict__ part,
                     const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                     float* __restrict__ save_stats, long long HW, int C, int G, int stat_chunks, int rows_per_cta, int k,
                     float eps) {
  __shared__ float s_mean[GN_MAX_G], s_rstd[GN_MAX_G];
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int pl = threadIdx.x / vecs;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const __nv_bfloat16* xb = x + static_cast<size_t>(b) * HW * C;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {       // (blocks of a tiny activation can be narrower than G)
    float a = 0.f, d = 0.f;
    const float* pp = part + (static_cast<size_t>(b) * stat_chunks * G + g) * 2;
    for (int c = 0; c < stat_chunks; ++c) {
      a += pp[static_cast<size_t>(c) * G * 2];
      d += pp[static_cast<size_t>(c) * G * 2 + 1];
    }
    const float n = static_cast<float>(HW) * static_cast<float>(cpg);
    const float m = a / n;
    const float var = fmaxf(d / n - m * m, 0.f);
    s_mean[g] = __bfloat162float(xb[g * cpg]) + (addv != nullptr ? addv[static_cast<size_t>(b) * C + g * cpg] : 0.f) + m;
    s_rstd[g] = rsqrtf(var + eps);
    if (save_stats != nullptr && blockIdx.x == 0) {          // (mean, rstd) of (sample, group) for the backward pass
      save_stats[(static_cast<size_t>(b) * G + g) * 2] = s_mean[g];
      save_stats[(static_cast<size_t>(b) * G + g) * 2 + 1] = s_rstd[g];
    }
  }
  __syncthreads();
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    sc[j] = __ldg(gamma + c) * s_rstd[g];
    const float a = addv != nullptr ? addv[static_cast<size_t>(b) * C + c] : 0.f;
    sf[j] = fmaf(a - s_mean[g], sc[j], __ldg(beta + c));
  }
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long p1 = p0 + rows_per_cta < HW ? p0 + rows_per_cta : HW;
  const uint4* xv = reinterpret_cast<const uint4*>(xb);
  uint4* yv = reinterpret_cast<uint4*>(y + static_cast<size_t>(b) * HW * C);
  for (long long p = p0 + pl; p < p1; p += static_cast<long long>(GN_UNROLL) * k) {
    uint4 raw[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const long long pp = p + static_cast<long long>(u) * k;
      if (pp < p1) raw[u] = __ldg(xv + pp * vecs + v);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const long long pp = p + static_cast<long long>(u) * k;
      if (pp < p1) {
        float f[8];
        bf16x8_unpack(raw[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(f[j], sc[j], sf[j]);
          if (SILU) t = __fdividef(t, 1.f + __expf(-t));
          f[j] = t;
        }
        yv[pp * vecs + v] = bf16x8_pack(f);
      }
    }
  }
}

// Launch geometry from (HW, C) alone -- not from the batch or the device -- so that a sample's statistics are summed in the
// same order whatever batch it is generated in (generation shards by sample across GPUs and must not depend on the
// shard size): one CTA per ~64 KB of a sample's activation (32 KB: 35 % slower at the 64 x 64 maps).
static void gn_geometry(long long B, long long HW, int C, int* k, int* chunks, int* rows_per_cta) {
  (void)B;
  const int vecs = C / 8;
  int kk = GN_THREADS_MAX / vecs;
  if (kk < 1) kk = 1;
  if (kk > 16) kk = 16;
  if (kk > HW) kk = static_cast<int>(HW);
  long long rows = (65536 + 2ll * C - 1) / (2ll * C);
  rows = (rows + kk - 1) / kk * kk;
  if (rows > HW) rows = (HW + kk - 1) / kk * kk;
  *k = kk;
  *rows_per_cta = static_cast<int>(rows);
  *chunks = static_cast<int>((HW + rows - 1) / rows);
}

long long group_norm_nhwc_ws_bytes(long long B, long long HW, int C, int G) {
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % 8 != 0 || C % G != 0 || G > GN_MAX_G || C / 8 > GN_THREADS_MAX) return -1;
  int k, chunks, rows;
  gn_geometry(B, HW, C, &k, &chunks, &rows);
  return B * chunks * G * 2ll * static_cast<long long>(sizeof(float));
}

int group_norm_nhwc(const void* x, const float* add_bc, const float* gamma, const float* beta, void* y, float* save_stats,
                    void* ws, long long B, long long HW, int C, int G, float eps, bool silu, cudaStream_t stream) {
  PV_REQUIRE(group_norm_nhwc_ws_bytes(B, HW, C, G) >= 0,
             "need C %% 8 == 0, C %% G == 0, G <= %d, C <= %d (B=%lld HW=%lld C=%d G=%d)", GN_MAX_G, 8 * GN_THREADS_MAX, B, HW, C, G);
  PV_REQUIRE(B <= 65535, "batch too large for one launch (B=%lld)", B);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0, "x / y must be 16-byte aligned");
  int k, chunks, rows;
  gn_geometry(B, HW, C, &k, &chunks, &rows);
  const int threads = (C / 8) * k;
  const size_t smem = 2ull * k * C * sizeof(float);
  auto stats = gn_stats_nhwc_kernel;
  if (smem > 48 * 1024) PV_CUDA(set_max_smem_once(stats, static_cast<int>(smem)));
  stats<<<dim3(chunks, static_cast<unsigned>(B)), threads, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), add_bc,
                                                                            static_cast<float*>(ws), HW, C, G, rows, k);
  PV_LAUNCHED();
  if (silu)
    gn_apply_nhwc_kernel<true><<<dim3(chunks, static_cast<unsigned>(B)), threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), add_bc, static_cast<const float*>(ws), gamma, beta, static_cast<__nv_bfloat16*>(y),
        save_stats, HW, C, G, chunks, rows, k, eps);
  else
    gn_apply_nhwc_kernel<false><<<dim3(chunks, static_cast<unsigned>(B)), threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), add_bc, static_cast<const float*>(ws), gamma, beta, static_cast<__nv_bfloat16*>(y),
        save_stats, HW, C, G, chunks, rows, k, eps);
  PV_LAUNCHED();
  return PV_OK;
}

// ---- backward of y = [SiLU](GroupNorm(x + add) * gamma + beta) with respect to x (training step: the backbone is frozen,
// train.py:348-370, so only the input gradient is needed).  With xh = (x + add - mean) * rstd, z = xh * gamma + beta,
// dz = dy * silu'(z):   dx = rstd * (dz * gamma - mean_g(dz * gamma) - xh * mean_g(dz * gamma * xh)).
// Same two-pass shape as the forward: per-(sample, pixel chunk, group) partial sums, then the element pass.
struct GnBwdChan {
  float g[8], bt[8], xs[8], xo[8];       // gamma, beta, rstd_g, (add_c - mean_g) * rstd_g
};
__device__ __forceinline__ void gn_bwd_chan(GnBwdChan& ch, const float* __restrict__ stats, const float* __restrict__ addv,
                                            const float* __restrict__ gamma, const float* __restrict__ beta, int b, int v,
                                            int C, int G, int cpg) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    const float mean = stats[(static_cast<size_t>(b) * G + g) * 2], rstd = stats[(static_cast<size_t>(b) * G + g) * 2 + 1];
    const float a = addv != nullptr ? addv[static_cast<size_t>(b) * C + c] : 0.f;
    ch.g[j] = __ldg(gamma + c);
    ch.bt[j] = __ldg(beta + c);
    ch.xs[j] = rstd;
    ch.xo[j] = (a - mean) * rstd;
  }
}
template <bool SILU>
__device__ __forceinline__ float gn_dz(float dy, float z) {
  if (!SILU) return dy;
  const float sg = __fdividef(1.f, 1.f + __expf(-z));
  return dy * sg * fmaf(z, 1.f - sg, 1.f);
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS_MAX)
gn_bwd_stats_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ addv, const __nv_bfloat16* __restrict__ dy,
                         const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ part, long long HW, int C, int G, int rows_per_cta, int k) {
  extern __shared__ float gn_smem[];               // [2][k][C]
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int pl = threadIdx.x / vecs;
  const int b = blockIdx.y;
  const int cpg = C / G;
  GnBwdChan ch;
  gn_bwd_chan(ch, stats, addv, gamma, beta, b, v, C, G, cpg);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long p1 = p0 + rows_per_cta < HW ? p0 + rows_per_cta : HW;
  const uint4* xv = reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * HW * C);
  const uint4* dv = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(b) * HW * C);
  for (long long p = p0 + pl; p < p1; p += 2ll * k) {
    uint4 rx[2], rd[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long pp = p + static_cast<long long>(u) * k;
      if (pp < p1) {
        rx[u] = __ldg(xv + pp * vecs + v);
        rd[u] = __ldg(dv + pp * vecs + v);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p + static_cast<long long>(u) * k < p1) {
        float f[8], d[8];
        bf16x8_unpack(rx[u], f);
        bf16x8_unpack(rd[u], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaf(f[j], ch.xs[j], ch.xo[j]);
          const float t = gn_dz<SILU>(d[j], fmaf(xh, ch.g[j], ch.bt[j])) * ch.g[j];
          s[j] += t;
          q[j] = fmaf(t, xh, q[j]);
        }
      }
    }
  }
  float* ss = gn_smem;
  float* sq = gn_smem + k * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ss[pl * C + v * 8 + j] = s[j];
    sq[pl * C + v * 8 + j] = q[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, d = 0.f;
    for (int l = 0; l < k; ++l) {
      a += ss[l * C + c];
      d += sq[l * C + c];
    }
    ss[c] = a;
    sq[c] = d;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {       // (blocks of a tiny activation can be narrower than G)
    float a = 0.f, d = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      a += ss[c];
      d += sq[c];
    }
    float* o = part + ((static_cast<size_t>(b) * gridDim.x + blockIdx.x) * G + g) * 2;
    o[0] = a;
    o[1] = d;
  }
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS_MAX)
gn_bwd_apply_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ addv, const __nv_bfloat16* __restrict__ dy,
                         const float* __restrict__ stats, const float* __restrict__ part, const float* __restrict__ gamma,
                         const float* __restrict__ beta, __nv_bfloat16* __restrict__ dx, long long HW, int C, int G,
                         int stat_chunks, int rows_per_cta, int k) {
  __shared__ float s_m1[GN_MAX_G], s_m2[GN_MAX_G];
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs;
  const int pl = threadIdx.x / vecs;
  const int b = blockIdx.y;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {       // (blocks of a tiny activation can be narrower than G)
    float a = 0.f, d = 0.f;
    const float* pp = part + (static_cast<size_t>(b) * stat_chunks * G + g) * 2;
    for (int c = 0; c < stat_chunks; ++c) {
      a += pp[static_cast<size_t>(c) * G * 2];
      d += pp[static_cast<size_t>(c) * G * 2 + 1];
    }
    const float n = static_cast<float>(HW) * static_cast<float>(cpg);
    s_m1[g] = a / n;
    s_m2[g] = d / n;
  }
  GnBwdChan ch;
  gn_bwd_chan(ch, stats, addv, gamma, beta, b, v, C, G, cpg);
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (v * 8 + j) / cpg;
    m1[j] = s_m1[g];
    m2[j] = s_m2[g];
  }
  const long long p0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long p1 = p0 + rows_per_cta < HW ? p0 + rows_per_cta : HW;
  const uint4* xv = reinterpret_cast<const uint4*>(x + static_cast<size_t>(b) * HW * C);
  const uint4* dv = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(b) * HW * C);
  uint4* ov = reinterpret_cast<uint4*>(dx + static_cast<size_t>(b) * HW * C);
  for (long long p = p0 + pl; p < p1; p += 2ll * k) {
    uint4 rx[2], rd[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long pp = p + static_cast<long long>(u) * k;
      if (pp < p1) {
        rx[u] = __ldg(xv + pp * vecs + v);
        rd[u] = __ldg(dv + pp * vecs + v);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long pp = p + static_cast<long long>(u) * k;
      if (pp < p1) {
        float f[8], d[8];
        bf16x8_unpack(rx[u], f);
        bf16x8_unpack(rd[u], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaf(f[j], ch.xs[j], ch.xo[j]);
          const float t = gn_dz<SILU>(d[j], fmaf(xh, ch.g[j], ch.bt[j])) * ch.g[j];
          f[j] = ch.xs[j] * (t - m1[j] - xh * m2[j]);
        }
        ov[pp * vecs + v] = bf16x8_pack(f);
      }
    }
  }
}

int group_norm_nhwc_bwd(const void* x, const float* add_bc, const void* dy, const float* stats, const float* gamma,
                        const float* beta, void* dx, void* ws, long long B, long long HW, int C, int G, bool silu,
                        cudaStream_t stream) {
  PV_REQUIRE(group_norm_nhwc_ws_bytes(B, HW, C, G) >= 0,
             "need C %% 8 == 0, C %% G == 0, G <= %d, C <= %d (B=%lld HW=%lld C=%d G=%d)", GN_MAX_G, 8 * GN_THREADS_MAX, B, HW, C, G);
  PV_REQUIRE(B <= 65535, "batch too large for one launch (B=%lld)", B);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) % 16 == 0,
             "x / dy / dx must be 16-byte aligned");
  int k, chunks, rows;
  gn_geometry(B, HW, C, &k, &chunks, &rows);
  const int threads = (C / 8) * k;
  const size_t smem = 2ull * k * C * sizeof(float);
  const dim3 grid(chunks, static_cast<unsigned>(B));
  const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* dd = static_cast<const __nv_bfloat16*>(dy);
  if (silu) {
    gn_bwd_stats_nhwc_kernel<true><<<grid, threads, smem, stream>>>(xx, add_bc, dd, stats, gamma, beta, static_cast<float*>(ws), HW, C, G, rows, k);
    PV_LAUNCHED();
    gn_bwd_apply_nhwc_kernel<true><<<grid, threads, 0, stream>>>(xx, add_bc, dd, stats, static_cast<const float*>(ws), gamma, beta,
                                                                 static_cast<__nv_bfloat16*>(dx), HW, C, G, chunks, rows, k);
  } else {
    gn_bwd_stats_nhwc_kernel<false><<<grid, threads, smem, stream>>>(xx, add_bc, dd, stats, gamma, beta, static_cast<float*>(ws), HW, C, G, rows, k);
    PV_LAUNCHED();
    gn_bwd_apply_nhwc_kernel<false><<<grid, threads, 0, stream>>>(xx, add_bc, dd, stats, static_cast<const float*>(ws), gamma, beta,
                                                                  static_cast<__nv_bfloat16*>(dx), HW, C, G, chunks, rows, k);
  }
  PV_LAUNCHED();
  return PV_OK;
}

// out[r, c] = a[r, c] + b[r, c] + bias[c]   (rows x C dense, C % 8 == 0; out may alias a or b)
__global__ void __launch_bounds__(256)
add_bias_nhwc_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, const float* __restrict__ bias, __nv_bfloat16* out,
                     long long rows, int C) {
  const int vecs = C >> 3;
  const long long total = rows * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vecs);
    float fa[8], fb[8];
    bf16x8_unpack(reinterpret_cast<const uint4*>(a)[i], fa);
    bf16x8_unpack(reinterpret_cast<const uint4*>(b)[i], fb);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * v), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * v + 1);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] = fa[j] + (fb[j] + bb[j]);
    reinterpret_cast<uint4*>(out)[i] = bf16x8_pack(fa);
  }
}

int add_bias_nhwc(const void* a, const void* b, const float* bias, void* out, long long rows, int C, cudaStream_t stream) {
  PV_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, "need C %% 8 == 0 (rows=%lld C=%d)", rows, C);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out) |
              reinterpret_cast<uintptr_t>(bias)) % 16 == 0, "pointers must be 16-byte aligned");
  const long long total = rows * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (blocks > cap) blocks = cap;
  add_bias_nhwc_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
                                                                         bias, static_cast<__nv_bfloat16*>(out), rows, C);
  PV_LAUNCHED();
  return PV_OK;
}

// LayerNorm over the last dimension, bf16 in / out, fp32 affine: one warp per row, the row held in registers (NV 16-byte
// vectors per lane), mean and variance as two passes over the registers.
// res != nullptr: the residual sum that precedes the norm rides along -- s = x + res is written to `sum_out` (rounded to
// bf16 like the stand-alone add) and the norm is taken of the rounded sum (BasicTransformerBlock: x = attn(..) + x; norm(x)).
template <int NV>
__global__ void __launch_bounds__(256)
layer_norm_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                       __nv_bfloat16* __restrict__ sum_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                       __nv_bfloat16* __restrict__ y, long long rows, int C, float eps) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
  float f[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < vecs) {
      bf16x8_unpack(__ldg(xr + idx), f[i]);
      if (res != nullptr) {
        float r[8];
        bf16x8_unpack(__ldg(reinterpret_cast<const uint4*>(res + row * C) + idx), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i][j] += r[j];
        const uint4 packed = bf16x8_pack(f[i]);
        reinterpret_cast<uint4*>(sum_out + row * C)[idx] = packed;
        bf16x8_unpack(packed, f[i]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 + lane < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        ss = fmaf(d, d, ss);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < vecs) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * idx), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * idx + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((f[i][j] - mean) * rstd, gg[j], bb[j]);
      yr[idx] = bf16x8_pack(o);
    }
  }
}

int layer_norm_bf16(const void* x, const void* res, void* sum_out, const float* gamma, const float* beta, void* y, long long rows,
                    int C, float eps, cudaStream_t stream) {
  PV_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 8 * 32 * 5, "need C %% 8 == 0 and C <= 1280 (rows=%lld C=%d)", rows, C);
  PV_REQUIRE((res == nullptr) == (sum_out == nullptr), "res and sum_out go together");
  PV_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
              reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(sum_out)) % 16 == 0,
             "pointers must be 16-byte aligned");
  const int nv = (C / 8 + 31) / 32;
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rr = static_cast<const __nv_bfloat16*>(res);
  __nv_bfloat16* so = static_cast<__nv_bfloat16*>(sum_out);
  __nv_bfloat16* yy = static_cast<__nv_bfloat16*>(y);
  switch (nv) {
    case 1: layer_norm_bf16_kernel<1><<<blocks, 256, 0, stream>>>(xx, rr, so, gamma, beta, yy, rows, C, eps); break;
    case 2: layer_norm_bf16_kernel<2><<<blocks, 256, 0, stream>>>(xx, rr, so, gamma, beta, yy, rows, C, eps); break;
    case 3: layer_norm_bf16_kernel<3><<<blocks, 256, 0, stream>>>(xx, rr, so, gamma, beta, yy, rows, C, eps); break;
    case 4: layer_norm_bf16_kernel<4><<<blocks, 256, 0, stream>>>(xx, rr, so, gamma, beta, yy, rows, C, eps); break;
    default: layer_norm_bf16_kernel<5><<<blocks, 256, 0, stream>>>(xx, rr, so, gamma, beta, yy, rows, C, eps); break;
  }
  PV_LAUNCHED();
  return PV_OK;
}

// Input gradient of the LayerNorm above (frozen affine): dx = rstd * (dy * gamma - mean(dy * gamma) - xh * mean(dy * gamma * xh)).
// One warp per row with x and dy in registers; mean / rstd are recomputed from the row (nothing but x is kept by the
// forward pass).
template <int NV>
__global__ void __launch_bounds__(256)
layer_norm_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ gamma,
                           __nv_bfloat16* __restrict__ dx, long long rows, int C, float eps) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int vecs = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
  const uint4* dr = reinterpret_cast<const uint4*>(dy + row * C);
  float f[NV][8], t[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < vecs) {
      bf16x8_unpack(__ldg(xr + idx), f[i]);
      bf16x8_unpack(__ldg(dr + idx), t[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i * 32 + lane < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        ss = fmaf(d, d, ss);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < vecs) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[i][j] = (f[i][j] - mean) * rstd;          // xh
        t[i][j] *= gg[j];                           // dy * gamma
        s1 += t[i][j];
        s2 = fmaf(t[i][j], f[i][j], s2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 / static_cast<float>(C), m2 = s2 / static_cast<float>(C);
  uint4* outr = reinterpret_cast<uint4*>(dx + row * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = i * 32 + lane;
    if (idx < vecs) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (t[i][j] - m1 - f[i][j] * m2);
      outr[idx] = bf16x8_pack(o);
    }
  }
}

int layer_norm_bwd_bf16(const void* x, const void* dy, const float* gamma, void* dx, long long rows, int C, float eps,
                        cudaStream_t stream) {
  PV_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 8 * 32 * 5, "need C %% 8 == 0 and C <= 1280 (rows=%lld C=%d)", rows, C);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
              reinterpret_cast<uintptr_t>(gamma)) % 16 == 0, "pointers must be 16-byte aligned");
  const int nv = (C / 8 + 31) / 32;
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* dd = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* oo = static_cast<__nv_bfloat16*>(dx);
  switch (nv) {
    case 1: layer_norm_bwd_bf16_kernel<1><<<blocks, 256, 0, stream>>>(xx, dd, gamma, oo, rows, C, eps); break;
    case 2: layer_norm_bwd_bf16_kernel<2><<<blocks, 256, 0, stream>>>(xx, dd, gamma, oo, rows, C, eps); break;
    case 3: layer_norm_bwd_bf16_kernel<3><<<blocks, 256, 0, stream>>>(xx, dd, gamma, oo, rows, C, eps); break;
    case 4: layer_norm_bwd_bf16_kernel<4><<<blocks, 256, 0, stream>>>(xx, dd, gamma, oo, rows, C, eps); break;
    default: layer_norm_bwd_bf16_kernel<5><<<blocks, 256, 0, stream>>>(xx, dd, gamma, oo, rows, C, eps); break;
  }
  PV_LAUNCHED();
  return PV_OK;
}

// y[m, n] = h[m, n] * gelu(h[m, N + n]);  h: [M, 2N] (row stride ldh), y: [M, N] dense; N % 8 == 0.
__global__ void __launch_bounds__(256)
geglu_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ y, long long M, int N, long long ldh) {
  const int vecs = N >> 3;
  const long long total = M * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / vecs;
    const int v = static_cast<int>(i - m * vecs);
    const uint4* row = reinterpret_cast<const uint4*>(h + m * ldh);
    float a[8], g[8];
    bf16x8_unpack(__ldg(row + v), a);
    bf16x8_unpack(__ldg(row + vecs + v), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // torch computes gelu(gate) in fp32 and rounds it to bf16 before the multiplication: same rounding points
      const float ge = __bfloat162float(__float2bfloat16_rn(0.5f * g[j] * (1.f + erff(g[j] * 0.70710678118654752f))));
      a[j] *= ge;
    }
    reinterpret_cast<uint4*>(y + m * N)[v] = bf16x8_pack(a);
  }
}

int geglu(const void* h, void* y, long long M, int N, long long ldh, cudaStream_t stream) {
  PV_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && ldh >= 2ll * N && ldh % 8 == 0, "need N %% 8 == 0 and ldh >= 2N, ldh %% 8 == 0 (M=%lld N=%d ldh=%lld)", M, N, ldh);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(y)) % 16 == 0, "h / y must be 16-byte aligned");
  const long long total = M * (N / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (blocks > cap) blocks = cap;
  geglu_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(h), static_cast<__nv_bfloat16*>(y), M, N, ldh);
  PV_LAUNCHED();
  return PV_OK;
}

// dh = [dy * gelu(g) | dy * a * gelu'(g)] for h = [a | g]:  gelu'(g) = Phi(g) + g * phi(g).
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dh,
                 long long M, int N, long long ldh) {
  const int vecs = N >> 3;
  const long long total = M * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / vecs;
    const int v = static_cast<int>(i - m * vecs);
    const uint4* row = reinterpret_cast<const uint4*>(h + m * ldh);
    float a[8], g[8], d[8];
    bf16x8_unpack(__ldg(row + v), a);
    bf16x8_unpack(__ldg(row + vecs + v), g);
    bf16x8_unpack(__ldg(reinterpret_cast<const uint4*>(dy + m * N) + v), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cdf = 0.5f * (1.f + erff(g[j] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * g[j] * g[j]);
      const float da = d[j] * (g[j] * cdf);
      const float dg = d[j] * a[j] * fmaf(g[j], pdf, cdf);
      a[j] = da;
      g[j] = dg;
    }
    uint4* orow = reinterpret_cast<uint4*>(dh + m * 2 * N);
    orow[v] = bf16x8_pack(a);
    orow[vecs + v] = bf16x8_pack(g);
  }
}

int geglu_bwd(const void* h, const void* dy, void* dh, long long M, int N, long long ldh, cudaStream_t stream) {
  PV_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && ldh >= 2ll * N && ldh % 8 == 0, "need N %% 8 == 0 and ldh >= 2N, ldh %% 8 == 0 (M=%lld N=%d ldh=%lld)", M, N, ldh);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dh)) % 16 == 0,
             "h / dy / dh must be 16-byte aligned");
  const long long total = M * (N / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (blocks > cap) blocks = cap;
  geglu_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(dy),
                                                                     static_cast<__nv_bfloat16*>(dh), M, N, ldh);
  PV_LAUNCHED();
  return PV_OK;
}

}  // namespace pv
