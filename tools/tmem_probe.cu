// Empirical fragment layout of tcgen05.ld.16x256b (which TMEM lane / column lands in which thread / register):
// every warp fills its 32 lanes x 64 columns with lane*1000 + column through 32x32b stores, then reads them back
// with 16x256b loads.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(uint32_t* out /*[4 warps][2 halves][2 reps][32 threads][4 regs]*/) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) {
    const uint32_t v = (uint32_t)((warp * 32 + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int half = 0; half < 2; ++half)
    for (int rep = 0; rep < 2; ++rep) {
      uint32_t r0, r1, r2, r3;
      const uint32_t a = base + ((uint32_t)(half * 16) << 16) + (uint32_t)(rep * 8);
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t* o = out + ((((warp * 2 + half) * 2 + rep) * 32 + lane) * 4);
      o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
    }
  // x2 form: registers 4..7 = the next 8 columns?
  if (warp == 0) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[4 * 2 * 2 * 32 * 4 + lane * 8 + i] = r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(64u) : "memory");
}

int main() {
  const int n = 4 * 2 * 2 * 32 * 4 + 32 * 8;
  uint32_t* d; cudaMalloc(&d, n * 4);
  cudaMemset(d, 0xff, n * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[4 * 2 * 2 * 32 * 4 + 32 * 8];
  cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
  for (int warp = 0; warp < 2; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int rep = 0; rep < 2; ++rep) {
        printf("warp %d lane-half %d col-rep %d: thread -> (lane,col) x4\n", warp, half, rep);
        for (int t = 0; t < 32; ++t) {
          const uint32_t* o = h + ((((warp * 2 + half) * 2 + rep) * 32 + t) * 4);
          printf("  t%02d: (%u,%u) (%u,%u) (%u,%u) (%u,%u)\n", t, o[0] / 1000, o[0] % 1000, o[1] / 1000, o[1] % 1000, o[2] / 1000, o[2] % 1000, o[3] / 1000, o[3] % 1000);
        }
      }
  printf("x2 form, warp 0:\n");
  for (int t = 0; t < 32; ++t) {
    const uint32_t* o = h + 4 * 2 * 2 * 32 * 4 + t * 8;
    printf("  t%02d:", t);
    for (int i = 0; i < 8; ++i) printf(" (%u,%u)", o[i] / 1000, o[i] % 1000);
    printf("\n");
  }
  return 0;
}
