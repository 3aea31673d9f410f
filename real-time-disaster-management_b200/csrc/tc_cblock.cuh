// CTA-pair version of the persistent fused ACFF block kernel (math and layouts: tc_block.cuh, schedule: tc_pblock.cuh).
//
// Two CTAs of a cluster (one SM each) issue ONE tcgen05.mma.cta_group::2 of M = 256: 128 output pixels from each
// CTA's own staged input patch against a weight tile that is split across the pair - each CTA keeps only N/2 rows of
// it in shared memory.  That is what makes blocks 2 and 3 fit the machine:
//   * block 2 (64 -> 96 channels, 25 taps): the folded weights are 307 KB; the 154 KB half stays RESIDENT in each CTA
//     instead of being streamed from L2 for every pair of tiles (measured: the streaming kernel ran at 93 cycles per
//     MMA against 56 for the MMA itself), and the operand fetch per MMA drops from 7 KB to 5.5 KB of shared memory
//     per SM (measured with tools/mma_rate.cu: 56.2 -> 49.5 cycles at N = 96);
//   * block 3 (96 -> 128, 614 KB of weights): each CTA streams only its half per image.
// K runs OUTER: a unit's input arrives as NC/2 separate TMA boxes of 16 channels (ring of NSLOT slots), and all 25
// taps of one 16-channel slice are issued before the next slice is needed, so a slot is free again after 25*GX MMAs
// and a ~70 KB ring is enough to keep loads ahead of the tensor pipe.  Weights are fetched as bundles of 5 taps of
// one slice, [5 taps][2 chunks][N/2][16 B], by tensor-map TMA (resident: every bundle once; streamed: ring).
//
// Pair protocol (rank 0 = leader, issues every MMA):
//   in_full / w_full   live in the LEADER; both CTAs' TMA loads complete_tx on them (cp.async.bulk.tensor
//                      .cta_group::2 with the leader's barrier address), the leader's producers post expect_tx
//   in_empty / w_empty / acc_full   one copy per CTA, signalled by tcgen05.commit ... multicast::cluster
//   acc_empty          lives in the leader, 8 local + 8 remote arrivals (epilogue warps of both CTAs)
// Roles per CTA: warp 0 input producer, warp 1 TMEM + (leader) MMA issuer, warp 2 weight producer, warps 3-10 epilogue,
// warp 11 zero halo of the output images.
#pragma once
#include "tc_pblock.cuh"

namespace ernet {
namespace tc {

constexpr int kCThreads = 384;   // 12 warps

template <int NC_, int N_, int HIN_, int HU_, int GX_, int NSLOT_, bool WRES_, int WSTAGES_, bool POOL_ = true,
          bool ACT_ = true, int NREAL_ = N_>
struct CCfg {
  static constexpr int NC = NC_, N = N_, HIN = HIN_, HU = HU_, GX = GX_, NSLOT = NSLOT_, WSTAGES = WSTAGES_;
  static constexpr bool WRES = WRES_, POOL = POOL_, ACT = ACT_;
  static constexpr int TAPS = 25, NREAL = NREAL_, NH = N / 2, TG = 5, NTG = TAPS / TG;
  static constexpr int WELEM = NH * 4 > 256 ? 8 : 4;                   // element size of the weight tensor map (box rows <= 256 elements)
  static constexpr int WP = HIN + 3;
  // A pooled block uses HU = HIN - 3 rows/cols, so its taps reach at most the ONE halo row/col that P8 stores after the
  // image and a box may be clamped to the tensor.  An un-pooled block (HU = HIN - 2, ErNET blocks 4-6) reaches one row/col
  // further: its box keeps the full 22 x (8 GX + 6) footprint and the TMA unit zero-fills what lies outside the tensor.
  static constexpr int BW = (POOL_ && (8 * GX + 6) >= WP) ? WP : (8 * GX + 6);
  static constexpr int BH = (POOL_ && 22 >= WP) ? WP : 22;
  static constexpr int CHUNK_BYTES = BH * BW * 16;
  static constexpr int SLOT_BYTES = 2 * CHUNK_BYTES;                   // one K step: 16 channels of the patch
  static constexpr int KS = NC / 2;
  static constexpr int TR = (HU + 15) / 16, TCOLS = (HU + 7) / 8;
  static constexpr int UX = (TCOLS + GX - 1) / GX;
  static constexpr int UNITS_PER_IMG = TR * UX;
  static constexpr int TAPW_BYTES = 2 * NH * 16;                       // one tap of one K step, this CTA's half
  static constexpr int WB_BYTES = TG * TAPW_BYTES;                     // weight bundle
  static constexpr int NBUNDLE = KS * NTG;                             // bundles per unit
  static constexpr int W_SMEM = (WRES ? NBUNDLE : WSTAGES) * WB_BYTES;
  static constexpr int OUT_H = POOL ? HU / 2 : HU, OP = OUT_H + 3;
  static constexpr int OFF_W = NSLOT * SLOT_BYTES;
  static constexpr int OFF_BAR = (OFF_W + W_SMEM + 15) / 16 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 16;
  // every unit issues GX tiles: a tile past the last tile column only multiplies TMA zero fill and is masked in the
  // epilogue (x >= HU), so both halves of a pair-unit always issue the same MMAs
  static_assert(NC % 2 == 0 && N % 32 == 0 && N <= 256, "operand shape");
  static_assert(2 * GX * N <= 512, "two TMEM accumulator buffers");
  static_assert(SLOT_BYTES % 128 == 0 && WB_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(NSLOT <= 16 && WSTAGES <= 16, "barrier arrays");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// ---- cluster helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA box loads whose completion is posted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {   // same warp of both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair when every MMA issued so far has completed
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <class Cfg, int KIND, int OUT>
__global__ void __launch_bounds__(kCThreads, 1)
acff_cblock_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ EpiParams<Cfg::N> par, uint16_t* __restrict__ out, int batch) {
  constexpr int N = Cfg::N, NH = Cfg::NH, GX = Cfg::GX, NSLOT = Cfg::NSLOT, BW = Cfg::BW, OP = Cfg::OP, KS = Cfg::KS;
  constexpr bool BF16 = KIND == KIND_BF16;
  // kind::i8: 16 int8 channels per 16-byte chunk, K step 32 = the same two chunks per MMA; s8 x s8 -> s32
  constexpr uint32_t IDESC = KIND == KIND_I8 ? instr_desc(2u, 1u, 256u, (uint32_t)N) : instr_desc(1u, BF16 ? 1u : 0u, 256u, (uint32_t)N);
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? Cfg::NREAL / 16 : Cfg::NREAL / 8;
  constexpr int tl_kernel = Cfg::NC == 2 ? 0 : Cfg::NC == 8 ? 1 : 2;   // timeline slot (study builds)
  (void)tl_kernel;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* in_full = bars;            // [16] leader
  uint64_t* in_empty = bars + 16;      // [16] per CTA
  uint64_t* w_full = bars + 32;        // [16] leader (resident weights: entry 0 only)
  uint64_t* w_empty = bars + 48;       // [16] per CTA
  uint64_t* acc_full = bars + 60;      // [2]  per CTA   (w_empty uses at most 12 entries)
  uint64_t* acc_empty = bars + 62;     // [2]  leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_BAR + 512);
  volatile uint32_t* abort_flag = tmem_slot + 1;
  static_assert(Cfg::WSTAGES <= 12, "w_empty entries");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int total_units = batch * Cfg::UNITS_PER_IMG;
  const int pair_units = (total_units + 1) >> 1;
  const int nk = pair_units > pair ? (pair_units - pair + npairs - 1) / npairs : 0;   // pair-units this pair processes
  // unit of this CTA in round k (the last pair-unit may have no second half: that CTA recomputes the last unit, no stores)
  auto unit_of = [&](int k, bool& dup) { const int u = 2 * (k * npairs + pair) + (int)rank; dup = u >= total_units; return dup ? total_units - 1 : u; };

  ERNET_CHAIN_ENTRY(tl_kernel + 1);
  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    for (int i = 0; i < 16; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); mbar_init(&w_full[i], 1); }
    for (int i = 0; i < 12; ++i) mbar_init(&w_empty[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // barriers of both CTAs are initialised before anything is posted on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (threadIdx.x == 0) ERNET_TL(31, 6);

  if (warp == 0) {
    // ------------------------------------------------------------------ input producer: one box per (unit, K step)
    pdl_wait();
    ERNET_CHAIN_WAITED(tl_kernel + 1);
    if (lane == 0) {
      int q = 0;
      for (int k = 0; k < nk; ++k) {
        bool dup;
        const int u = unit_of(k, dup);
        const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
        const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
        bool ok = true;
        for (int ks = 0; ks < KS; ++ks, ++q) {
          const int sl = q % NSLOT, use = q / NSLOT;
          if (use > 0 && !mbar_wait(&in_empty[sl], (use - 1) & 1, abort_flag, 0x700u, q)) { ok = false; break; }
          if (ks == 0) ERNET_TL(k, 0);
          if (leader) mbar_expect_tx(&in_full[sl], 2 * Cfg::SLOT_BYTES);
          tma2_load_4d(smem + sl * Cfg::SLOT_BYTES, &tmap_in, ux * GX * 8 * 4, ty * 16, 2 * ks, img, mapa_u32(smem_u32(&in_full[sl]), 0));
        }
        if (!ok) break;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ weight producer (constants: no pdl_wait)
    if (lane == 0) {
      if (Cfg::WRES) {
        // one barrier per K step (NTG bundles): the first MMAs of the CTA pair start when the first slice of the resident
        // weights has landed, not after all of them (154 KB per CTA in block 2: ~2 us at the head of every launch)
        if (leader)
          for (int ks = 0; ks < KS; ++ks) mbar_expect_tx(&w_full[ks], 2 * Cfg::NTG * Cfg::WB_BYTES);
        for (int b = 0; b < Cfg::NBUNDLE; ++b)
          tma2_load_3d(s_w + b * Cfg::WB_BYTES, &tmap_w, (int)rank * NH * (16 / Cfg::WELEM), 2 * (b / Cfg::NTG), (b % Cfg::NTG) * Cfg::TG,
                       mapa_u32(smem_u32(&w_full[b / Cfg::NTG]), 0));
      } else {
        int it = 0;
        bool ok = true;
        for (int k = 0; k < nk && ok; ++k)
          for (int b = 0; b < Cfg::NBUNDLE; ++b, ++it) {
            const int s = it % Cfg::WSTAGES, use = it / Cfg::WSTAGES;
            if (use > 0 && !mbar_wait(&w_empty[s], (use - 1) & 1, abort_flag, 0x701u, it)) { ok = false; break; }
            if (leader) mbar_expect_tx(&w_full[s], 2 * Cfg::WB_BYTES);
            tma2_load_3d(s_w + s * Cfg::WB_BYTES, &tmap_w, (int)rank * NH * (16 / Cfg::WELEM), 2 * (b / Cfg::NTG), (b % Cfg::NTG) * Cfg::TG,
                         mapa_u32(smem_u32(&w_full[s]), 0));
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      bool ok = true;
      static_assert(!Cfg::WRES || KS <= 16, "one weight barrier per K step");
      const uint32_t in_addr = smem_u32(smem), w_addr = smem_u32(s_w);
      constexpr uint32_t A_HI = desc_hi(BW * 16), B_HI = desc_hi(128);
      const uint32_t w_lo0 = desc_lo(w_addr, NH * 16);
      int ws = 0, q = 0;
      uint32_t wphase = 0;
      for (int k = 0; k < nk && ok; ++k) {
        constexpr int ntile = GX;
        const int buf = k & 1, use = k >> 1;
        if (use > 0) ok = mbar_wait(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x704u, k);
        ERNET_TL(k, 2);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * GX * N);
        for (int ks = 0; ks < KS && ok; ++ks, ++q) {
          const int sl = q % NSLOT;
          ok = mbar_wait(&in_full[sl], (q / NSLOT) & 1, abort_flag, 0x703u, q);
          if (ks == 0) ERNET_TL(k, 1);
          if (ok && Cfg::WRES && k == 0) ok = mbar_wait(&w_full[ks], 0, abort_flag, 0x702u, ks);    // resident weights of this K step
          if (!ok) break;
          tc_fence_after();
          const uint32_t a_lo0 = desc_lo(in_addr + sl * Cfg::SLOT_BYTES + (uint32_t)((2 * BW + 2) * 16), Cfg::CHUNK_BYTES);
#pragma unroll
          for (int tg = 0; tg < Cfg::NTG; ++tg) {
            uint32_t b_lo;
            if (Cfg::WRES) {
              b_lo = w_lo0 + (uint32_t)((ks * Cfg::NTG + tg) * (Cfg::WB_BYTES >> 4));
            } else {
              ok = mbar_wait(&w_full[ws], wphase, abort_flag, 0x705u, q * 8 + tg);
              if (!ok) break;
              tc_fence_after();
              b_lo = w_lo0 + (uint32_t)(ws * (Cfg::WB_BYTES >> 4));
            }
#pragma unroll
            for (int t = 0; t < Cfg::TG; ++t) {
              const int tap = tg * Cfg::TG + t;
              const uint32_t toff = (uint32_t)(tap_dy(tap) * BW + tap_dx(tap));
              const uint64_t bd = desc_make(b_lo + (uint32_t)(t * (Cfg::TAPW_BYTES >> 4)), B_HI);
#pragma unroll
              for (int tl = 0; tl < GX; ++tl)
                if (tl < ntile) {
                  if (KIND == KIND_I8) mma2_i8(d0 + tl * N, desc_make(a_lo0 + toff + (uint32_t)(tl * 8), A_HI), bd, IDESC, (ks | tap) != 0 ? 1u : 0u);
                  else mma2_f16(d0 + tl * N, desc_make(a_lo0 + toff + (uint32_t)(tl * 8), A_HI), bd, IDESC, (ks | tap) != 0 ? 1u : 0u);
                }
            }
            if (!Cfg::WRES) {
              mma2_commit(&w_empty[ws]);
              if (++ws == Cfg::WSTAGES) { ws = 0; wphase ^= 1; }
            }
          }
          if (ok) mma2_commit(&in_empty[sl]);
        }
        if (ok) mma2_commit(&acc_full[buf]);
        ERNET_TL(k, 3);
      }
    }
    __syncwarp();
  } else if (warp == 11) {
    // ------------------------------------------------------------------ zero halo of the output images this CTA starts
    if (OUT != OUT_NHWC) {
      pdl_wait();
      constexpr int BORDER = 3 * OP + (OP - 3) * 3;
      for (int k = 0; k < nk; ++k) {
        bool dup;
        const int u = unit_of(k, dup);
        const int img = u / Cfg::UNITS_PER_IMG;
        if (dup || u - img * Cfg::UNITS_PER_IMG != 0) continue;
        uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)img * OUT_CHUNKS * OP * OP;
        for (int i = lane; i < OUT_CHUNKS * BORDER; i += 32) {
          const int ch = i / BORDER, kk = i - ch * BORDER;
          int rr, cc;
          if (kk < 3 * OP) { rr = kk / OP; cc = kk - rr * OP; if (rr == 2) rr = OP - 1; }
          else { const int k2 = kk - 3 * OP; rr = 2 + k2 / 3; cc = k2 % 3; if (cc == 2) cc = OP - 1; }
          oimg[(ch * OP + rr) * OP + cc] = make_uint4(0, 0, 0, 0);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 3..10)
    const int q4 = warp & 3;
    const int ehalf = (warp - 3) >> 2;
    const int rr = 4 * q4 + (lane >> 3), cc = lane & 7;
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int qsel = (xodd ? 2 : 0) + (yodd ? 1 : 0);
    pdl_wait();                                             // stores below must not overtake the previous kernel's readers
    const bool susp = g_epi_suspend != 0;
    for (int k = 0; k < nk; ++k) {
      bool dup;
      const int u = unit_of(k, dup);
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      constexpr int ntile = GX;
      const int buf = k & 1, use = k >> 1;
      if (!mbar_wait_epi(&acc_full[buf], use & 1, abort_flag, 0x800u + warp, k, susp)) break;
      if (threadIdx.x == 96) ERNET_TL(k, 4);
      tc_fence_after();
      for (int tl = ehalf; tl < ntile; tl += 2) {
        const int y = ty * 16 + rr, x = (ux * GX + tl) * 8 + cc;
        const bool valid = (y < Cfg::HU) && (x < Cfg::HU) && !dup;
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * GX * N + tl * N);
        epilogue_tile<Cfg, KIND, OUT>(par, tbase, y, x, valid, xodd, yodd, qsel, out, img);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&acc_empty[buf]), 0));
      if (threadIdx.x == 96) ERNET_TL(k, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's shared memory and barriers stay alive until both CTAs are done
  ERNET_CHAIN_EXIT(tl_kernel + 1);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
// Input map: same tensor as tc_pblock.cuh, box = one K step (2 chunks) of the patch.
template <class Cfg>
inline int make_cinput_map(CUtensorMap* map, const void* base, int batch) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)Cfg::WP * 4, (cuuint64_t)Cfg::WP, (cuuint64_t)Cfg::NC, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)Cfg::WP * 16, (cuuint64_t)Cfg::WP * Cfg::WP * 16, (cuuint64_t)Cfg::NC * Cfg::WP * Cfg::WP * 16};
  const cuuint32_t box[4] = {(cuuint32_t)Cfg::BW * 4, (cuuint32_t)Cfg::BH, 2, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled (input) failed with CUresult %d", (int)r);
  return ERNET_OK;
}
// Weight map over the packed image [tap][chunk][N][16 B]: box = (N/2 rows x 16 B, 2 chunks, 5 taps).
template <class Cfg>
inline int make_weight_map(CUtensorMap* map, const void* wimg) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)Cfg::N * (16 / Cfg::WELEM), (cuuint64_t)Cfg::NC, (cuuint64_t)Cfg::TAPS};
  const cuuint64_t strides[2] = {(cuuint64_t)Cfg::N * 16, (cuuint64_t)Cfg::NC * Cfg::N * 16};
  const cuuint32_t box[3] = {(cuuint32_t)Cfg::NH * (16 / Cfg::WELEM), 2, (cuuint32_t)Cfg::TG};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, Cfg::WELEM == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(wimg), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed with CUresult %d", (int)r);
  return ERNET_OK;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// CTA pairs that can be co-resident (one CTA per SM; GPCs with an odd SM count leave one SM without a partner).
template <class Cfg, int KIND, int OUT>
inline int max_pairs(int num_sms) {
  static int cached = -1;
  if (cached < 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_sms / 2 * 2); cfg.blockDim = dim3(kCThreads); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, acff_cblock_kernel<Cfg, KIND, OUT>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms / 2; }
    cached = n < num_sms / 2 ? n : num_sms / 2;
  }
  return cached;
}

template <class Cfg, int KIND, int OUT>
inline int launch_acff_cblock(const void* in, const void* wimg, const EpiParams<Cfg::N>& par, void* out, int batch, int num_sms,
                              cudaStream_t stream) {
  CUtensorMap map_in, map_w;
  int rc = make_cinput_map<Cfg>(&map_in, in, batch);
  if (rc) return rc;
  if ((rc = make_weight_map<Cfg>(&map_w, wimg))) return rc;
  const int pair_units = (batch * Cfg::UNITS_PER_IMG + 1) / 2;
  const int pairs = pair_units < max_pairs<Cfg, KIND, OUT>(num_sms) ? pair_units : max_pairs<Cfg, KIND, OUT>(num_sms);
  ERNET_CUDA(launch_pdl_pair(acff_cblock_kernel<Cfg, KIND, OUT>, dim3(2 * pairs), dim3(kCThreads), Cfg::SMEM_BYTES, stream, map_in, map_w, par,
                             static_cast<uint16_t*>(out), batch));
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT>
inline int set_cblock_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(acff_cblock_kernel<Cfg, KIND, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  return ERNET_OK;
}

// Pair configurations: GX tiles per CTA per unit, input slots (K steps in flight), weight residency / ring depth (bundles).
using CBlock1 = CCfg<2, 64, 69, 66, 3, 4, true, 1>;              // 26 KB of weights resident per CTA + 4 x 21 KB slots (one K step per unit)
using CBlock2 = CCfg<8, 96, 33, 30, 2, 5, true, 1>;             // 154 KB of weights resident per CTA + 5 x 15 KB slots
using CBlock3 = CCfg<12, 128, 15, 12, 2, 12, false, 9>;         // 12 x 10 KB slots (two units) + 9 x 10 KB weight bundles
// Small batches (real-time use: one frame at a time, aider-predict.py / real-time-inference.py): ONE tile per CTA and unit, so
// that an image's tiles spread over twice as many CTA pairs and the serial MMA chain of a launch halves (block 3 at batch 1:
// both tiles of the image on one pair, 150 instead of 300 MMAs deep).  Same MMAs per output in the same order: bit-identical.
// Used while all pair-units still fit one round (kSmallBatch2 / kSmallBatch3 images); beyond that the two-tile units win
// (half the weight traffic per image).
using CBlock2S = CCfg<8, 96, 33, 30, 1, 5, true, 1>;
using CBlock3S = CCfg<12, 128, 15, 12, 1, 12, false, 9>;
constexpr int kSmallBatch2 = 18, kSmallBatch3 = 74;
// Squeeze_RedConv: ACFF2 without the pool (conv_red2 sits between it and pool2), ACFF3 on 48 input channels
using CBlock2R = CCfg<8, 96, 33, 30, 2, 5, true, 1, /*POOL*/ false>;
using CBlock3R = CCfg<6, 128, 15, 12, 2, 9, false, 9>;
// int8 engine (chunks of 16 channels): both weight halves are resident (77 KB / 154 KB per CTA)
// baseline ErNET (240x240 inputs: 119 -> 117/58 -> 56/28 -> 26/13 -> 11 -> 9 -> 7); block 1 runs on tc_pblock.cuh
using EBlock2 = CCfg<8, 96, 58, 56, 2, 5, true, 1>;
using EBlock3 = CCfg<12, 128, 28, 26, 2, 6, false, 9>;
using EBlock4 = CCfg<16, 128, 13, 11, 2, 8, false, 9, /*POOL*/ false>;
using EBlock5 = CCfg<16, 128, 11, 9, 2, 8, false, 9, /*POOL*/ false>;
using EBlock6 = CCfg<16, 256, 9, 7, 1, 8, false, 6, /*POOL*/ false>;
using CBlock2Q = CCfg<4, 96, 33, 30, 2, 8, true, 1>;
using CBlock3Q = CCfg<6, 128, 15, 12, 2, 6, true, 1>;
// int8 Squeeze_RedConv: ACFF2 un-pooled with an fp16 output (conv_red2 follows as a 16-bit 1-tap instance that writes the
// int8 pool2 tensor), ACFF3 on 48 real + 16 zero int8 channels
using CBlock2RQ = CCfg<4, 96, 33, 30, 2, 6, true, 1, /*POOL*/ false>;
using CBlock3RQ = CCfg<4, 128, 15, 12, 2, 8, true, 1>;

}  // namespace tc
}  // namespace ernet
