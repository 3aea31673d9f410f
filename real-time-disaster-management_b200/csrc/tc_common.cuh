// sm_100a building blocks for the tensor-core kernels: mbarrier, 1-D bulk copy (TMA engine, UBLKCP),
// tcgen05 TMEM allocation / MMA / commit / load, and the shared-memory + instruction descriptors.
//
// Shared-memory operand layout used everywhere here is the un-swizzled K-major "interleaved" canonical
// layout: a core matrix is 8 rows x 16 bytes stored contiguously (row stride 16 B); the next 8 rows are
// SBO bytes further, the next 16 bytes of K are LBO bytes further.  Because rows are 16 B apart, a
// convolution tap is just a different start address on the same staged image (see tc_block.cuh).
#pragma once
#include "common.cuh"

namespace ernet {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Soft watchdog.  A wait that exceeds ~0.15 s records {tag, block, aux} in a global debug buffer, raises a
// CTA-wide abort flag in shared memory, and returns false; every other wait in the CTA then returns false
// as soon as it sees the flag, the roles fall through to the common exit, and the kernel terminates
// normally (a __trap() here was observed to wedge the process instead of surfacing an error).  The host
// reads the buffer with ernet_debug_device_status().
__device__ unsigned int g_tc_status[8];   // [0] = number of timeouts, [1] = first tag, [2] = blockIdx.x, [3] = aux
__device__ unsigned int* g_tc_host_flag;  // mapped pinned host word (ernet_abi.cu): raised on a timeout so that the host's
                                          // synchronising entry points report an error instead of wrong probabilities
__device__ __forceinline__ void tc_raise_timeout(uint32_t tag, uint32_t aux) {
  if (atomicAdd(&g_tc_status[0], 1u) == 0) { g_tc_status[1] = tag; g_tc_status[2] = blockIdx.x; g_tc_status[3] = aux; }
  unsigned int* hp = g_tc_host_flag;
  if (hp) { *reinterpret_cast<volatile unsigned int*>(hp) = 1u; __threadfence_system(); }
}

// Non-blocking phase test.  The waits below poll with it: a suspended mbarrier.try_wait was measured (in-kernel
// timeline, tools/timeline.py) to resume only at its ~10 us time limit in some producer waits.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag, uint32_t tag,
                                          uint32_t aux = 0) {
  if (mbar_test_wait(bar, parity)) return true;
  const long long t0 = clock64();
  int spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if ((++spins & 63) == 0) {
      if (*abort_flag) return false;
      if (clock64() - t0 > 300000000LL) {
        tc_raise_timeout(tag, aux);
        *abort_flag = 1u;
        return false;
      }
    }
  }
  return true;
}

// Same contract, but the wait SUSPENDS in hardware (mbarrier.try_wait) instead of polling: for warps that are not on the
// critical path of a kernel whose bottleneck is the CUDA cores (tc_dblock.cuh) - a dozen polling epilogue warps were
// measured to take half of the issue slots there.
__device__ __forceinline__ bool mbar_wait_suspend(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag, uint32_t tag,
                                                  uint32_t aux = 0) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*abort_flag) return false;
    if (clock64() - t0 > 300000000LL) {
      tc_raise_timeout(tag, aux);
      *abort_flag = 1u;
      return false;
    }
  }
  return true;
}

// The epilogue warps of the persistent block kernels wait for their accumulators with the SUSPENDING form (default) instead
// of polling: in blocks 2 and 3 eight warps wait ~80 % of the time, and their polling cost issue slots next to the MMA
// issuer and power - measured over 2000 steps under the power cap: 0.1990 -> 0.1975 ms per step, SM clock 1875 -> 1931 MHz.
// ERNET_EPI_SUSPEND=0 restores polling (device-wide switch, read once per kernel).
__device__ unsigned int g_epi_suspend = 1u;
__device__ __forceinline__ bool mbar_wait_epi(uint64_t* bar, uint32_t parity, volatile uint32_t* abort_flag, uint32_t tag, uint32_t aux, bool suspend) {
  return suspend ? mbar_wait_suspend(bar, parity, abort_flag, tag, aux) : mbar_wait(bar, parity, abort_flag, tag, aux);
}

// Optional in-kernel timeline (study builds only: -DERNET_TIMELINE).  Slot layout: [cta < 148][unit < 32][8 stamps].
#ifdef ERNET_TIMELINE
__device__ unsigned long long g_timeline[3 * 148 * 32 * 8];
#define ERNET_TL(k, j) do { if (blockIdx.x < 148 && (k) < 32) g_timeline[((tl_kernel * 148 + blockIdx.x) * 32 + (k)) * 8 + (j)] = (unsigned long long)clock64(); } while (0)
// whole-chain view on the global timer: per kernel {first CTA entry, first return from the PDL wait, last CTA exit, last CTA entry}
__device__ unsigned long long g_chain[8][4];
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ERNET_CHAIN_ENTRY(kid) do { if (threadIdx.x == 0) { const unsigned long long t_ = ::ernet::tc::gtime_ns(); atomicMin(&::ernet::tc::g_chain[kid][0], t_); atomicMax(&::ernet::tc::g_chain[kid][3], t_); } } while (0)
#define ERNET_CHAIN_WAITED(kid) do { if (threadIdx.x == 0) atomicMin(&::ernet::tc::g_chain[kid][1], ::ernet::tc::gtime_ns()); } while (0)
#define ERNET_CHAIN_EXIT(kid) do { if (threadIdx.x == 0) atomicMax(&::ernet::tc::g_chain[kid][2], ::ernet::tc::gtime_ns()); } while (0)
#else
#define ERNET_TL(k, j) do { } while (0)
#define ERNET_CHAIN_ENTRY(kid) do { } while (0)
#define ERNET_CHAIN_WAITED(kid) do { } while (0)
#define ERNET_CHAIN_EXIT(kid) do { } while (0)
#endif

// One lane of a converged warp (elect.sync): keeps the surrounding code warp-uniform so that descriptors
// live in uniform registers instead of going through R2UR + retry loops.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- async proxy ------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes: multiple of 16).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues.
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 columns of fp32/int32 accumulators -> 32 registers per thread (lane = TMEM lane).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// K-major, no swizzle.  Fields (cute::UMMA::SmemDescriptor): [0,14) start>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version=1 (Blackwell), [61,64) layout type = 0 (SWIZZLE_NONE).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Split form for hot loops: the low word is (start>>4) | LBO field, so moving the start address by a
// multiple of 16 bytes is one 32-bit add of (bytes>>4); the high word (SBO, version) is loop-invariant.
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint64_t desc_make(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// cute::UMMA::InstrDescriptor: [4,6) D format (1 = f32, 2 = s32), [7,10) A format, [10,13) B format
// (kind::f16: 0 = f16, 1 = bf16; kind::i8: 0 = u8, 1 = s8), [15] A major (0 = K), [16] B major (0 = K),
// [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t instr_desc(uint32_t d_fmt, uint32_t ab_fmt, uint32_t M, uint32_t N) {
  return (d_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- the 25 distinct taps of an ACFF block ------------------------------------------------------
// Union of the three dilated 3x3 stencils of acff.py:25-30 expressed as offsets from the OUTPUT
// coordinate: d=1 -> {0,1,2}, d=2 -> {-1,1,3}, d=3 -> {-2,1,4}; (1,1) is shared by all three.
// Sorted by (dy, dx); pack_tc.py builds the weight images in the same order.
// Arithmetic form (compile-time foldable in unrolled loops): rows of the table are
//   dy=-2:{-2,1,4} dy=-1:{-1,1,3} dy=0:{0,1,2} dy=1:{-2..4} dy=2:{0,1,2} dy=3:{-1,1,3} dy=4:{-2,1,4}
__host__ __device__ constexpr int tap_dy(int t) { return t < 3 ? -2 : t < 6 ? -1 : t < 9 ? 0 : t < 16 ? 1 : t < 19 ? 2 : t < 22 ? 3 : 4; }
__host__ __device__ constexpr int tap_dx(int t) {
  return t < 3 ? -2 + 3 * t : t < 6 ? -1 + 2 * (t - 3) : t < 9 ? (t - 6) : t < 16 ? (t - 9) - 2 : t < 19 ? (t - 16)
         : t < 22 ? -1 + 2 * (t - 19) : -2 + 3 * (t - 22);
}
__device__ constexpr int8_t kTapDy[25] = {-2, -2, -2, -1, -1, -1, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4};
__device__ constexpr int8_t kTapDx[25] = {-2, 1, 4, -1, 1, 3, 0, 1, 2, -2, -1, 0, 1, 2, 3, 4, 0, 1, 2, -1, 1, 3, -2, 1, 4};

}  // namespace tc
}  // namespace ernet
