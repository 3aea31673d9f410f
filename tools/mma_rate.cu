// Microbenchmark (study tool, not product code): issue rate of tcgen05.mma kind::f16, M=128 per CTA, K=16, operands in
// shared memory in the un-swizzled K-major layout the block kernels use (core matrix = 8 rows x 16 B).
//   mode 0: cta_group::1                 (A 4 KB + B N*32 B of shared memory per MMA)
//   mode 1: cta_group::2 (CTA pair)      (A 4 KB + B N*16 B per CTA per MMA, M = 256 per instruction)
//   mode 2: cta_group::1 .ws, B operand kept in the collector buffer over runs of `reuse` MMAs
//   mode 3: cta_group::1 kind::i8 (s8 x s8 -> s32, K = 32 per instruction: the same two 16-byte chunks per row as a
//           16-bit K = 16 MMA, i.e. twice the operations for the same operand bytes) - the int8 peak denominator
// Prints cycles per MMA for each N.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/mma_rate tools/mma_rate.cu
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(uint32_t M, uint32_t N, bool i8 = false) {
  return ((i8 ? 2u : 1u) << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <int MODE>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, int cstate) {
  if (MODE == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (MODE == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (MODE == 3)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else {
    if (cstate == 0)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if (cstate == 1)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  }
}

template <int MODE, int N, int reuse>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int P = 30;                                  // staged patch pitch in pixels (block 1 persistent unit)
  uint8_t* s_a = smem;                                   // 2 chunks x 22 x 30 x 16 B
  uint8_t* s_b = smem + 24 * 1024;                       // 25 taps x 2 chunks x N x 16 B
  for (int i = threadIdx.x; i < (24 * 1024 + 25 * 2 * 256 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (MODE == 1) rank = cg::this_cluster().block_rank();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (MODE == 1) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MODE == 1) cg::this_cluster().sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long cyc = 0;
  if (threadIdx.x == 0 && rank == 0) {
    constexpr int NB = MODE == 1 ? N / 2 : N;            // B rows held by this CTA
    constexpr uint32_t IDESC = instr_desc(MODE == 1 ? 256 : 128, N, MODE == 3);
    const uint32_t a0 = smem_u32(s_a) + (2 * P + 2) * 16, b0 = smem_u32(s_b);
    constexpr int G = 512 / N < 4 ? 512 / N : 4;         // accumulator tiles in flight (TMEM columns)
    const long long t0 = clock64();
    int cnt = 0;
    const uint64_t a_base = smem_desc(a0, 22 * P * 16, P * 16), b_base = smem_desc(b0, NB * 16, 128);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int tap = 0; tap < 25; ++tap) {
        const int dy = tap / 5 - 2, dx = tap % 5 - 2;
        const uint64_t bd = b_base + (uint64_t)(tap * (2 * NB));
#pragma unroll
        for (int tl = 0; tl < G; ++tl) {
          const uint64_t ad = a_base + (uint64_t)(int64_t)((dy * P + dx) + tl * 8);
          const int cstate = (tl % reuse == 0) ? 0 : ((tl % reuse == reuse - 1 || tl == G - 1) ? 2 : 1);
          mma<MODE>(tmem + tl * N, ad, bd, IDESC, tap != 0, reuse == 1 ? 0 : cstate);
          ++cnt;
        }
      }
    }
    if (MODE == 1)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    cyc = clock64() - t0;
    out[blockIdx.x * 2] = cyc;
    out[blockIdx.x * 2 + 1] = cnt;
  } else if (MODE == 1 && threadIdx.x == 0) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MODE == 1) cg::this_cluster().sync();
  if (warp == 0) {
    if (MODE == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int MODE, int N, int reuse = 1>
void run(const char* name, int, int grid) {
  long long* d;
  cudaMalloc(&d, 4096 * sizeof(long long));
  cudaMemset(d, 0, 4096 * sizeof(long long));
  const int smem = 24 * 1024 + 25 * 2 * 256 * 16;
  cudaFuncSetAttribute(rate_kernel<MODE, N, reuse>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MODE == 1 ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<MODE, N, reuse>, d, 40);
    if (e != cudaSuccess) { printf("%s N=%d launch failed: %s\n", name, N, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s N=%d failed: %s\n", name, N, cudaGetErrorString(e)); return; }
  }
  long long h[4096];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int stride = MODE == 1 ? 4 : 2;
  double worst = 0, best = 1e30;
  for (int b = 0; b < grid / (MODE == 1 ? 2 : 1); ++b) {
    const double c = (double)h[b * stride] / (double)h[b * stride + 1];
    worst = c > worst ? c : worst; best = c < best ? c : best;
  }
  const double ideal = 128.0 * N / 256.0 / (MODE == 1 ? 1 : 1);
  printf("%-22s N=%3d reuse=%d grid=%3d: %.1f .. %.1f cycles per MMA per CTA-slot (M=%d), ideal %.0f\n", name, N, reuse, grid, best, worst,
         MODE == 1 ? 256 : 128, ideal);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    const int g2 = grid == 1 ? 2 : 148;
    run<0, 64>("cta_group::1", 1, grid); run<0, 96>("cta_group::1", 1, grid); run<0, 128>("cta_group::1", 1, grid); run<0, 256>("cta_group::1", 1, grid);
    run<1, 64>("cta_group::2", 1, g2); run<1, 96>("cta_group::2", 1, g2); run<1, 128>("cta_group::2", 1, g2); run<1, 256>("cta_group::2", 1, g2);
    run<3, 64>("cta_group::1 kind::i8", 1, grid); run<3, 96>("cta_group::1 kind::i8", 1, grid); run<3, 128>("cta_group::1 kind::i8", 1, grid); run<3, 256>("cta_group::1 kind::i8", 1, grid);
    run<2, 64, 4>("cta_group::1 .ws", 4, grid); run<2, 128, 4>("cta_group::1 .ws", 4, grid); run<2, 64, 1>("cta_group::1 .ws", 1, grid);
  }
  return 0;
}
