// photoverse_b200 -- the training objective of the path in one reduction (reference train.py:509-535):
//     loss = MSE(noise_pred, noise) + w_text * mean|concept_text_embeddings| + w_vis * mean(||V_ip||)
// (:516 L_mse, :509 L_text, :512-513 L_vis over the stacked `to_v_ip_norm` side outputs of the 16 processors,
//  :535 weights 0.01 / 0.001).  The reference runs ~10 small elementwise / reduction kernels for it; here one kernel
// produces per-block partial sums of the three terms and a second one adds them in a fixed order (deterministic), and the
// backward is one elementwise kernel for d noise_pred, d concept and d ||V_ip||.
#include "pv_common.cuh"
#include "pv_host.h"
#include "../../include/photoverse_b200.h"

namespace pv {

template <typename T> __device__ __forceinline__ float ls_ld(const T* p);
template <> __device__ __forceinline__ float ls_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ls_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void ls_st(T* p, float v);
template <> __device__ __forceinline__ void ls_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void ls_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int LS_BLOCKS = 296;

__device__ __forceinline__ float ls_block_sum(float v, float* red) {      // fixed-order block reduction (256 threads)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
  }
  __syncthreads();
  return s;
}

template <typename T>
__global__ void __launch_bounds__(256)
loss_partial_kernel(const T* __restrict__ pred, const T* __restrict__ target, long long n, const T* __restrict__ concept, long long m,
                    const T* __restrict__ vnorm, long long k, float* __restrict__ part) {
  __shared__ float red[8];
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float a = 0.f, b = 0.f, c = 0.f;
  for (long long i = i0; i < n; i += stride) { const float d = ls_ld(pred + i) - ls_ld(target + i); a = fmaf(d, d, a); }
  for (long long i = i0; i < m; i += stride) b += fabsf(ls_ld(concept + i));
  for (long long i = i0; i < k; i += stride) c += ls_ld(vnorm + i);
  a = ls_block_sum(a, red);
  b = ls_block_sum(b, red);
  c = ls_block_sum(c, red);
  if (threadIdx.x == 0) { part[3 * blockIdx.x] = a; part[3 * blockIdx.x + 1] = b; part[3 * blockIdx.x + 2] = c; }
}

// out = {loss, l_mse, l_text, l_vis}
__global__ void loss_final_kernel(const float* __restrict__ part, int nblocks, long long n, long long m, long long k, float w_text,
                                  float w_vis, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = 0; i < nblocks; ++i) { a += part[3 * i]; b += part[3 * i + 1]; c += part[3 * i + 2]; }
  const float l_mse = a / static_cast<float>(n), l_text = m > 0 ? b / static_cast<float>(m) : 0.f,
              l_vis = k > 0 ? c / static_cast<float>(k) : 0.f;
  out[0] = l_mse + w_text * l_text + w_vis * l_vis;
  out[1] = l_mse;
  out[2] = l_text;
  out[3] = l_vis;
}

template <typename T>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const T* __restrict__ pred, const T* __restrict__ target, long long n, const T* __restrict__ concept, long long m,
                long long k, float w_text, float w_vis, const float* __restrict__ gloss, T* __restrict__ d_pred,
                T* __restrict__ d_concept, T* __restrict__ d_vnorm) {
  const float g = gloss[0];
  const float cp = 2.f * g / static_cast<float>(n), cc = m > 0 ? w_text * g / static_cast<float>(m) : 0.f,
              cv = k > 0 ? w_vis * g / static_cast<float>(k) : 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (long long i = i0; i < n; i += stride) ls_st(d_pred + i, cp * (ls_ld(pred + i) - ls_ld(target + i)));
  for (long long i = i0; i < m; i += stride) {
    const float v = ls_ld(concept + i);
    ls_st(d_concept + i, v > 0.f ? cc : (v < 0.f ? -cc : 0.f));            // d|x| = sign(x), 0 at 0 (torch.abs backward)
  }
  for (long long i = i0; i < k; i += stride) ls_st(d_vnorm + i, cv);
}

long long train_loss_ws_bytes() { return LS_BLOCKS * 3 * 4; }

int train_loss_fwd(bool bf16, const void* pred, const void* target, long long n, const void* concept, long long m, const void* vnorm,
                   long long k, float w_text, float w_vis, float* out4, void* ws, cudaStream_t stream) {
  PV_REQUIRE(n > 0 && m >= 0 && k >= 0 && ws != nullptr, "bad arguments");
  float* part = static_cast<float*>(ws);
  if (bf16)
    loss_partial_kernel<__nv_bfloat16><<<LS_BLOCKS, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), static_cast<const __nv_bfloat16*>(target), n,
                                                                      static_cast<const __nv_bfloat16*>(concept), m,
                                                                      static_cast<const __nv_bfloat16*>(vnorm), k, part);
  else
    loss_partial_kernel<float><<<LS_BLOCKS, 256, 0, stream>>>(static_cast<const float*>(pred), static_cast<const float*>(target), n,
                                                              static_cast<const float*>(concept), m, static_cast<const float*>(vnorm), k, part);
  PV_LAUNCHED();
  loss_final_kernel<<<1, 32, 0, stream>>>(part, LS_BLOCKS, n, m, k, w_text, w_vis, out4);
  PV_LAUNCHED();
  return PV_OK;
}

int train_loss_bwd(bool bf16, const void* pred, const void* target, long long n, const void* concept, long long m, long long k,
                   float w_text, float w_vis, const float* gloss, void* d_pred, void* d_concept, void* d_vnorm, cudaStream_t stream) {
  PV_REQUIRE(n > 0 && gloss != nullptr, "bad arguments");
  if (bf16)
    loss_bwd_kernel<__nv_bfloat16><<<LS_BLOCKS * 2, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(pred), static_cast<const __nv_bfloat16*>(target), n,
                                                                      static_cast<const __nv_bfloat16*>(concept), m, k, w_text, w_vis, gloss,
                                                                      static_cast<__nv_bfloat16*>(d_pred), static_cast<__nv_bfloat16*>(d_concept),
                                                                      static_cast<__nv_bfloat16*>(d_vnorm));
  else
    loss_bwd_kernel<float><<<LS_BLOCKS * 2, 256, 0, stream>>>(static_cast<const float*>(pred), static_cast<const float*>(target), n,
                                                              static_cast<const float*>(concept), m, k, w_text, w_vis, gloss,
                                                              static_cast<float*>(d_pred), static_cast<float*>(d_concept), static_cast<float*>(d_vnorm));
  PV_LAUNCHED();
  return PV_OK;
}

}  // namespace pv
