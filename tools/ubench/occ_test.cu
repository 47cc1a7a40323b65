// Does a kernel that allocates tensor memory get more than one CTA per SM?  (occupancy API + actual co-residency)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int COLS, int MODE>
__global__ void __launch_bounds__(384, 2) k(unsigned* out, int spin) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  if (MODE >= 1) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (MODE >= 2 && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  __syncthreads();
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (threadIdx.x == 0) out[blockIdx.x] = smid;
  __syncthreads();
  if (MODE >= 1 && threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(COLS) : "memory");
}

template <typename K>
void test(const char* name, K kern, unsigned* out) {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 384, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms1, ms2;
  kern<<<148, 384>>>(out, 2000000);
  cudaDeviceSynchronize();
  cudaEventRecord(e0); kern<<<148, 384>>>(out, 2000000); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms1, e0, e1);
  cudaEventRecord(e0); kern<<<296, 384>>>(out, 2000000); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms2, e0, e1);
  printf("%-40s occupancy API %d   148 CTAs %.3f ms   296 CTAs %.3f ms  -> %s   (%s)\n", name, nb, ms1, ms2,
         ms2 < 1.5f * ms1 ? "2 CTAs co-resident" : "serialised", cudaGetErrorString(cudaGetLastError()));
}

int main() {
  unsigned* out;
  cudaMalloc(&out, 4096);
  test("no tensor memory", k<256, 0>, out);
  test("tcgen05.alloc 256 columns", k<256, 1>, out);
  test("tcgen05.alloc 256 + commit", k<256, 2>, out);
  test("tcgen05.alloc 128 columns", k<128, 1>, out);
  return 0;
}
