// Issue-rate microbenchmark of the instructions a softmax inner loop is made of (sm_100a):
//   ex2.approx.ftz.f32 / .f16x2 / .bf16x2, cvt.rn.{f16x2,bf16x2}.f32, fma.rn.f32x2, max.f32 (2 and 3 inputs).
// Prints cycles per warp-instruction per SM sub-partition (4 warps resident per scheduler, 8 independent chains each).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sfu_rate sfu_rate.cu ; run: ./sfu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;

template <int OP>
__global__ void rate_kernel(unsigned* out, long long* cycles, float seed) {
  unsigned r[CH];
  float f[CH], g[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    f[i] = seed * (threadIdx.x + i + 1) * 1e-3f;
    g[i] = -f[i];
    r[i] = __float_as_uint(f[i]) & 0x3bff3bffu;
  }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 3) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(g[i])); f[i] = __uint_as_float(r[i] | 0x30000000u); }
      if (OP == 4) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(g[i])); f[i] = __uint_as_float(r[i] | 0x30000000u); }
      if (OP == 5) {
        unsigned long long a, b;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f[i]), "f"(g[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(a));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(g[i]) : "l"(a));
      }
      if (OP == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
      if (OP == 7) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(g[(i + 1) % CH]));
      if (OP == 8) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[i]));
      if (OP == 9) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(r[i]));
      if (OP == 10) asm volatile("add.rn.f16x2 %0, %0, %0;" : "+r"(r[i]));
    }
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) acc ^= r[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, unsigned* out, long long* cyc) {
  const int blocks = 148, threads = 512;                         // 16 warps per SM = 4 per scheduler
  rate_kernel<OP><<<blocks, threads>>>(out, cyc, 1.25f);
  rate_kernel<OP><<<blocks, threads>>>(out, cyc, 1.25f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < blocks; ++i) mean += h[i];
  mean /= blocks;
  const double per = mean / (double(ITERS) * CH * 4);            // warp-instructions per scheduler
  printf("%-28s %7.2f cycles per warp-instruction per scheduler  (%5.1f lanes/clk/SM)\n", name, per, 128.0 / per);
}

int main() {
  unsigned* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("ex2.approx.ftz.f32", out, cyc);
  run<1>("ex2.approx.f16x2", out, cyc);
  run<2>("ex2.approx.ftz.bf16x2", out, cyc);
  run<3>("cvt.rn.f16x2.f32 (+LOP)", out, cyc);
  run<4>("cvt.rn.bf16x2.f32 (+LOP)", out, cyc);
  run<5>("fma.rn.f32x2", out, cyc);
  run<6>("max.f32 (2 inputs)", out, cyc);
  run<7>("max.f32 (3 inputs)", out, cyc);
  run<8>("fma.rn.f32", out, cyc);
  run<9>("fma.rn.f16x2", out, cyc);
  run<10>("add.rn.f16x2", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
