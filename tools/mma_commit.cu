// Microbenchmark (study tool): does anything the MMA-issuing thread does at a unit boundary of the persistent block
// kernel drain the tensor pipe?  75 MMAs (M=128, N=64, K=16) per unit, then a variant-specific boundary sequence.
//   0 nothing, 1 one tcgen05.commit, 2 two commits, 3 two commits + tcgen05.fence::after_thread_sync,
//   4 two commits + mbarrier test_wait on an already completed barrier + fence, 5 commit only every 4th unit,
//   6 as 4 but the two test_waits are issued in the middle of the unit and only consumed at the boundary,
//   7 two commits + two volatile shared-memory flag loads at the boundary (a helper thread would own the mbarrier waits),
//   8 one commit + one flag load,
//   9 variant 4 split over TWO issuing threads (warps 0 and 1), even / odd units, each with its own TMEM buffer
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
template <int V>
__global__ void __launch_bounds__(128, 1) k(long long* out, int units) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile uint32_t flags[4];
  constexpr int P = 30, N = 64;
  for (int i = threadIdx.x; i < (24 * 1024 + 25 * 2 * 64 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { flags[0] = 1000000; flags[1] = 1000000; }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[7])) : "memory");   // bars[7]: phase 0 complete
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (V == 9 ? (threadIdx.x == 0 || threadIdx.x == 32) : threadIdx.x == 0) {
    const int me = threadIdx.x >> 5;
    const uint64_t a_base = smem_desc(smem_u32(smem) + (2 * P + 2) * 16, 22 * P * 16, P * 16), b_base = smem_desc(smem_u32(smem + 24 * 1024), N * 16, 128);
    const long long t0 = clock64();
    for (int u = (V == 9 ? me : 0); u < units; u += (V == 9 ? 2 : 1)) {
      const uint32_t d0 = tmem + (u & 1) * 3 * N;
      bool pre0 = false, pre1 = false;
#pragma unroll
      for (int tap = 0; tap < 25; ++tap) {
        if (V == 6 && tap == 12) { pre0 = test_wait(&bars[7], 0); pre1 = test_wait(&bars[7], 0); }
        const int dy = tap / 5 - 2, dx = tap % 5 - 2;
#pragma unroll
        for (int tl = 0; tl < 3; ++tl) mma(d0 + tl * N, a_base + (uint64_t)(int64_t)((dy * P + dx) + tl * 8), b_base + (uint64_t)(tap * 2 * N), instr_desc(128, N), tap != 0);
      }
      if (V == 1) commit(&bars[u & 1]);
      if ((V >= 2 && V <= 4) || V == 6 || V == 9) { commit(&bars[u & 1]); commit(&bars[2 + (u & 1)]); }
      if (V == 7) { commit(&bars[u & 1]); commit(&bars[2 + (u & 1)]); while ((int)flags[0] < u) { } while ((int)flags[1] < u) { } asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (V == 8) { commit(&bars[u & 1]); while ((int)flags[0] < u) { } asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (V == 6) { if (!pre0) { while (!test_wait(&bars[7], 0)) { } } if (!pre1) { while (!test_wait(&bars[7], 0)) { } } asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (V == 5 && (u & 3) == 3) commit(&bars[0]);
      if (V == 3 || V == 4 || V == 9) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (V == 4 || V == 9) { while (!test_wait(&bars[7], 0)) { } }
    }
    commit(&bars[6 - me]);
    while (!test_wait(&bars[6 - me], 0)) { }
    out[blockIdx.x * 2 + me] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int V> void run() {
  long long* d; cudaMalloc(&d, 2 * 148 * 8); cudaMemset(d, 0, 2 * 148 * 8);
  const int smem = 24 * 1024 + 25 * 2 * 64 * 16;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { k<V><<<148, 128, smem>>>(d, 26); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("variant %d failed: %s\n", V, cudaGetErrorString(e)); return; } }
  long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const long long tt = h[0] > h[1] ? h[0] : h[1];
  printf("variant %d: %.0f cycles per unit of 75 MMAs (%.1f per MMA)\n", V, tt / 26.0, tt / 26.0 / 75.0);
  cudaFree(d);
}
int main() { run<0>(); run<1>(); run<2>(); run<3>(); run<4>(); run<5>(); run<6>(); run<7>(); run<8>(); run<9>(); return 0; }
