// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand source and smem layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include "../../photoverse_b200/csrc/pv_common.cuh"
using namespace pv;

template <int N, int MODE>   // MODE 0: SS, SW128 operands   1: TS (A in TMEM), B no-swizzle core matrices   2: SS no-swizzle
__global__ void k(long long* out, int reps, int nblk) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    long long t0 = 0, t1 = 0;
    for (int r = 0; r < 2; ++r) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < reps; ++i) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (MODE == 0) {
              const uint64_t da = umma_desc_sw128(smem + (i % nblk) * 16384) + 2 * kk;
              const uint64_t db = umma_desc_sw128(smem + 49152 + (i % nblk) * 0) + 2 * kk;
              umma_bf16_ss(tmem + 256, da, db, idesc, 1);
            } else if (MODE == 1) {
              const uint64_t db = umma_desc(smem_u32(smem) + kk * 2 * (N * 16), N * 16, 128, UMMA_LAYOUT_NONE);
              umma_bf16_ts(tmem + 256, tmem + kk * 8, db, idesc, 1);
            } else {
              const uint64_t da = umma_desc(smem_u32(smem) + 49152 + kk * 2 * (128 * 16), 128 * 16, 128, UMMA_LAYOUT_NONE);
              const uint64_t db = umma_desc(smem_u32(smem) + kk * 2 * (N * 16), N * 16, 128, UMMA_LAYOUT_NONE);
              umma_bf16_ss(tmem + 256, da, db, idesc, 1);
            }
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, r & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N, int MODE>
void run(long long* d, const char* name, int grid) {
  const int reps = 256;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, MODE><<<grid, 128, 100 * 1024>>>(d, reps, 3);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-4s N=%3d grid=%3d : %7.1f cycles / MMA (ideal %5.1f)  %s\n", name, N, grid, double(mx) / (reps * 4), N / 2.0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  for (int grid : {1, 148}) {
    run<32, 0>(d, "SS", grid); run<48, 0>(d, "SS", grid); run<64, 0>(d, "SS", grid); run<96, 0>(d, "SS", grid);
    run<128, 0>(d, "SS", grid); run<160, 0>(d, "SS", grid); run<192, 0>(d, "SS", grid); run<256, 0>(d, "SS", grid);
    run<48, 1>(d, "TS", grid); run<80, 1>(d, "TS", grid); run<96, 1>(d, "TS", grid); run<160, 1>(d, "TS", grid);
    run<48, 2>(d, "SSn", grid); run<96, 2>(d, "SSn", grid); run<160, 2>(d, "SSn", grid);
  }
  return 0;
}
