// Persistent version of the fused ACFF block kernel (see tc_block.cuh for the math and the data layouts).
//
// What changes is the schedule.  tc_block.cuh runs one image per CTA: nothing overlaps the image load, blocks
// whose accumulators fill TMEM cannot overlap their epilogue, and a batch of 256 on 148 SMs pays 2 full waves.
// Here one CTA per SM loops over small UNITS - GX horizontally adjacent 16x8 tiles of one image:
//   * the unit's input patch, (16+6) x (8*GX+6) pixels x all channel chunks, is fetched by ONE TMA tensor-map
//     box copy (cp.async.bulk.tensor.5d, SASS UTMALDG) from the padded P8/P16 image into a ring of NSTAGE
//     buffers - the box lands as [chunk][row][col][16 B], i.e. already the UMMA operand layout with pitch = box
//     width, and parts of the box outside the tensor are zero-filled by the TMA unit;
//   * the MMA warp ping-pongs two TMEM accumulator buffers of GX x N columns, so the epilogue of unit k runs
//     under the MMAs of unit k+1, and the TMA of unit k+2 runs under both;
//   * weights are loaded once per CTA (block 1, conv_red2) or streamed through the ring per unit;
//   * units are dealt round-robin to the CTAs: 256 images x 15 units over 148 SMs leaves < 4 % imbalance.
// Roles: warp 0 input producer, warp 1 MMA issuer, warp 2 weight-ring producer, 8-12 epilogue warps, then one warp that writes the
// zero halo of the output images (kept off the producer's critical path: ~10 k cycles per image when it was inline), and the
// second MMA issuer (resident-weight configurations).  Anything the issuing thread waits on between two MMAs - an
// mbarrier test, a shared-memory load, even tcgen05.commit - empties the tensor pipe for 50-200 cycles (tools/
// mma_commit.cu: 4060 instead of 3744 cycles per unit of 75 MMAs); with even units issued by warp 1 and odd units by
// warp 12 one thread's boundary stalls sit under the other thread's MMAs (3736 cycles per unit).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encode function is fetched through the runtime)

#include "tc_block.cuh"

namespace ernet {
namespace tc {

// PAIR_: the input has ONE real 16-byte chunk per pixel (16 int8 channels, or the 8 real 16-bit channels of RedConv's
// stem) where an MMA K step wants two.  Instead of multiplying a zero chunk, the second chunk of the A descriptor is
// pointed at the SAME staged chunk shifted by another tap (LBO = byte distance between the two taps' windows): one
// MMA does two taps, 13 instead of 25 per tile.  Weights: [13 pairs][tap A chunk, tap B chunk][N][16 B].
//
// TAIL_: HU = 16 q + 2 leaves two output rows for a fifth row of 16 x 8 tiles that would be 12.5 % useful (block 1: 9 of
// its 45 tiles per image).  Instead ONE tail unit per image covers both rows over the whole width with two tiles whose M
// rows run ALONG a row: its box is 8 rows x 128 pixels (pitch 128 pixels, zero fill right of the image), the A
// descriptor's SBO is 128 bytes (the next core matrix = the next 8 pixels of the same row), tile 0 = row HU-2, tile 1 =
// row HU-1.  Same 25 shifted windows, same weights; TMEM lane = x, so the y half of the 2x2 pool is a max of the two
// tiles' accumulators in one thread and the x half one exchange with lane^1.  38 instead of 45 tiles per image.
template <int NC_, int N_, int HIN_, int HU_, int GX_, int NSTAGE_, bool WRES_, int WSTAGES_,
          bool POOL_ = true, int TAPS_ = 25, bool ACT_ = true, int NREAL_ = N_, bool PAIR_ = false, bool TAIL_ = false>
struct PCfg {
  static constexpr bool PAIR = PAIR_, TAIL = TAIL_;
  static constexpr int NPAIR = 13;
  // epilogue warps per TMEM lane quarter: one per tile of the unit (up to 3), so that with tap pairing (13 MMAs per
  // tile) the epilogue of a unit still finishes inside the unit's MMA time
  static constexpr int EPW = GX_ <= 3 ? GX_ : 2;
  static constexpr int NEPI = 4 * EPW;
  static constexpr int WARP_HALO = 3 + NEPI, WARP_MMA2 = 4 + NEPI;
  static constexpr int THREADS = 32 * (5 + NEPI);
  static constexpr int NC = NC_, N = N_, HIN = HIN_, HU = HU_, GX = GX_, NSTAGE = NSTAGE_, WSTAGES = WSTAGES_;
  static constexpr bool WRES = WRES_, POOL = POOL_, ACT = ACT_;
  static constexpr int TAPS = TAPS_, NREAL = NREAL_;
  static constexpr int WP = HIN + 3;                                   // padded image pitch / height
  static constexpr int HALO = TAPS_ == 1 ? 0 : 1;                      // a 1x1 convolution needs no neighbours: the box is the tiles
  static constexpr int BW = (8 * GX + 6 * HALO) < WP ? (8 * GX + 6 * HALO) : WP;     // box width  (pixels)
  static constexpr int BH = (16 + 6 * HALO) < WP ? (16 + 6 * HALO) : WP;             // box height (16 output rows + 6)
  static constexpr int CHUNK_BYTES = BH * BW * 16;
  static constexpr int STAGE_BYTES = (PAIR ? 1 : NC) * CHUNK_BYTES;          // bytes one box delivers
  static constexpr int TBW = 128, TBH = 8;                                  // tail box: pixels per row, rows (2 output rows + 6)
  static constexpr int TAIL_CHUNK_BYTES = TBH * TBW * 16, TAIL_STAGE_BYTES = (PAIR ? 1 : NC) * TAIL_CHUNK_BYTES;
  static constexpr int STAGE_MAX = (TAIL && TAIL_STAGE_BYTES > STAGE_BYTES) ? TAIL_STAGE_BYTES : STAGE_BYTES;
  static constexpr int STAGE_STRIDE = (STAGE_MAX + 127) / 128 * 128;        // TMA destinations are 128-byte aligned
  static constexpr int TR = TAIL ? HU / 16 : (HU + 15) / 16, TCOLS = (HU + 7) / 8;   // rows of regular tiles
  static constexpr int UX = (TCOLS + GX - 1) / GX;                     // units per tile row
  static constexpr int UNITS_PER_IMG = TR * UX + (TAIL ? 1 : 0);
  static_assert(!TAIL || (HU % 16 == 2 && HU <= TBW && POOL_ && TAPS_ == 25 && GX_ >= 2 && WRES_), "tail unit: two rows left over, pooled 25-tap block");
  static constexpr int KSTEPS = NC / 2;
  static constexpr int TAP_BYTES = NC * N * 16;
  static constexpr int W_BYTES = (PAIR ? NPAIR : TAPS) * TAP_BYTES;
  static_assert(!PAIR || (NC == 2 && TAPS == 25 && WRES_), "tap pairing: one real chunk of two, resident weights");
  static constexpr int W_SMEM = WRES ? W_BYTES : WSTAGES * TAP_BYTES;
  static constexpr int OUT_H = POOL ? HU / 2 : HU, OP = OUT_H + 3;
  static constexpr int OFF_W = NSTAGE * STAGE_STRIDE;
  static constexpr int OFF_BAR = (OFF_W + W_SMEM + 15) / 16 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static_assert(NC % 2 == 0 && N % 32 == 0 && N <= 256, "operand shape");
  static_assert(2 * GX * N <= 512, "two TMEM accumulator buffers");
  static_assert(NSTAGE <= 4 && WSTAGES <= 8, "barrier arrays");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

constexpr int kPThreads = 544;   // upper bound: 17 warps (Cfg::THREADS is what a configuration launches)
#ifndef ERNET_PBLOCK_MINB
#define ERNET_PBLOCK_MINB 1
#endif

// cp.async.bulk.tensor.4d: box of the tensor described by `tmap` at coordinates (c0..c3), completion on `bar`.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

template <class Cfg, int KIND, int OUT>
__global__ void __launch_bounds__(kPThreads, ERNET_PBLOCK_MINB)
acff_pblock_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_tail,
                   const uint16_t* __restrict__ wimg, const __grid_constant__ EpiParams<Cfg::N> par, uint16_t* __restrict__ out, int batch) {
  constexpr int N = Cfg::N, GX = Cfg::GX, NSTAGE = Cfg::NSTAGE, BW = Cfg::BW, OP = Cfg::OP;
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr uint32_t IDESC = KIND == KIND_I8 ? instr_desc(2u, 1u, 128u, (uint32_t)N) : instr_desc(1u, BF16 ? 1u : 0u, 128u, (uint32_t)N);
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? Cfg::NREAL / 16 : Cfg::NREAL / 8;
  constexpr int tl_kernel = Cfg::NC == 2 ? 0 : Cfg::NC == 8 ? 1 : 2;   // timeline slot (study builds)
  (void)tl_kernel;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_w = bars;             // [1]  resident weights
  uint64_t* in_full = bars + 1;       // [4]
  uint64_t* in_empty = bars + 5;      // [4]
  uint64_t* w_full = bars + 9;        // [8]
  uint64_t* w_empty = bars + 17;      // [8]
  uint64_t* acc_full = bars + 25;     // [2]
  uint64_t* acc_empty = bars + 27;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  volatile uint32_t* abort_flag = tmem_slot + 1;
  volatile uint32_t* turn = tmem_slot + 2;     // units whose MMAs are (nearly) all issued: hand-over between the two issuers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = batch * Cfg::UNITS_PER_IMG;
  ERNET_CHAIN_ENTRY(1);

  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    *turn = 0u;
    mbar_init(bar_w, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], Cfg::NEPI); }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
    if (Cfg::TAIL) tma_prefetch_desc(&tmap_tail);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (threadIdx.x == 0) ERNET_TL(31, 6);

  if (warp == 0) {
    // ------------------------------------------------------------------ input producer (+ halo of the output)
    if (lane == 0 && Cfg::WRES) {          // constants first: not ordered after the previous kernel (PDL)
      mbar_expect_tx(bar_w, Cfg::W_BYTES);
      bulk_g2s(s_w, wimg, Cfg::W_BYTES, bar_w);
    }
    pdl_wait();
    ERNET_CHAIN_WAITED(1);
    int k = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      if (lane == 0) {
        const int st = k % NSTAGE, use = k / NSTAGE;
        if (use > 0 && !mbar_wait(&in_empty[st], (use - 1) & 1, abort_flag, 0x500u, k)) break;
        ERNET_TL(k, 0);
        if (Cfg::TAIL && ty == Cfg::TR) {     // tail unit: rows HU-4 .. HU+3 of the image over the whole width (u64 elements: 2 per pixel)
          mbar_expect_tx(&in_full[st], Cfg::TAIL_STAGE_BYTES);
          tma_load_4d(smem + st * Cfg::STAGE_STRIDE, &tmap_tail, 0, Cfg::TR * 16, 0, img, &in_full[st]);
        } else {
          mbar_expect_tx(&in_full[st], Cfg::STAGE_BYTES);
          tma_load_4d(smem + st * Cfg::STAGE_STRIDE, &tmap_in, (ux * GX * 8 + 2 * (1 - Cfg::HALO)) * 4, ty * 16 + 2 * (1 - Cfg::HALO), 0, img, &in_full[st]);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ weight-ring producer (streamed weights)
    if (!Cfg::WRES && lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        bool ok = true;
        for (int tap = 0; tap < Cfg::TAPS && ok; ++tap, ++it) {
          const int s = it % Cfg::WSTAGES, use = it / Cfg::WSTAGES;
          if (use > 0) ok = mbar_wait(&w_empty[s], (use - 1) & 1, abort_flag, 0x501u, it);
          if (!ok) break;
          mbar_expect_tx(&w_full[s], Cfg::TAP_BYTES);
          bulk_g2s(s_w + s * Cfg::TAP_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)tap * Cfg::TAP_BYTES, Cfg::TAP_BYTES, &w_full[s]);
        }
        if (!ok) break;
      }
    }
  } else if (warp == 1 || warp == Cfg::WARP_MMA2) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 takes units k = 0, 2, .. and warp 12
    // the odd ones when the weights are resident (each unit has its own TMEM buffer k & 1 and input stage, so the two
    // streams are independent); with streamed weights the ring is consumed in order by warp 1 alone
    constexpr int NISSUE = Cfg::WRES ? 2 : 1;
    const int me = warp == 1 ? 0 : 1;
    if (me < NISSUE && elect_one()) {
      bool ok = true;
      if (Cfg::WRES) ok = mbar_wait(bar_w, 0, abort_flag, 0x502u);
      const uint32_t in_addr = smem_u32(smem), w_addr = smem_u32(s_w);
      constexpr uint32_t A_HI = desc_hi(BW * 16), B_HI = desc_hi(128);
      constexpr uint32_t A_KSTEP = (2 * Cfg::CHUNK_BYTES) >> 4, B_KSTEP = (2 * N * 16) >> 4;
      constexpr int TAP_UNROLL = Cfg::KSTEPS == 1 ? Cfg::TAPS : 1;
      const uint32_t w_lo0 = desc_lo(w_addr, N * 16);
      int ws = 0;
      uint32_t wphase = 0;
      int k = me;
      for (int u = blockIdx.x + me * (int)gridDim.x; u < total_units && ok; u += NISSUE * (int)gridDim.x, k += NISSUE) {
        const int r = u % Cfg::UNITS_PER_IMG, ux = r % Cfg::UX;
        const bool tail = Cfg::TAIL && r == Cfg::UNITS_PER_IMG - 1;
        const int ntile = tail ? 2 : min(GX, Cfg::TCOLS - ux * GX);
        const int st = k % NSTAGE, buf = k & 1, use = k >> 1;
        ok = mbar_wait(&in_full[st], (k / NSTAGE) & 1, abort_flag, 0x503u, k);
        ERNET_TL(k, 1);
        if (ok && use > 0) ok = mbar_wait(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x504u, k);
        ERNET_TL(k, 2);
        if (!ok) break;
        if (NISSUE > 1) {          // start when the other issuer is within a few MMAs of the end of unit k - 1 (timing only:
          while (*turn < (uint32_t)k) { if (*abort_flag) { ok = false; break; } }   // the two units share no data)
          if (!ok) break;
        }
        tc_fence_after();
        // tile tl of the unit: output origin = box origin + (2, 2 + 8*tl)
        const uint32_t a_lo0 = desc_lo(in_addr + st * Cfg::STAGE_STRIDE + (uint32_t)(Cfg::HALO * (2 * BW + 2) * 16), Cfg::CHUNK_BYTES);
        const uint32_t d0 = tmem_base + (uint32_t)(buf * GX * N);
        if (Cfg::TAIL && tail) {
          // tail unit: M rows run along an image row (SBO = 128 bytes), pitch TBW pixels, tile tl = output row HU - 2 + tl
          constexpr int TBW = Cfg::TBW;
          constexpr uint32_t A_HI_T = desc_hi(128);
          const uint32_t t_lo0 = desc_lo(in_addr + st * Cfg::STAGE_STRIDE + (uint32_t)((2 * TBW + 2) * 16), Cfg::TAIL_CHUNK_BYTES);
          if constexpr (Cfg::PAIR) {
#pragma unroll
            for (int pr = 0; pr < Cfg::NPAIR; ++pr) {
              const int tapA = pr == 0 ? 0 : 2 * pr - 1, tapB = pr == 0 ? 1 : 2 * pr;
              const int offA = tap_dy(tapA) * TBW + tap_dx(tapA), offB = tap_dy(tapB) * TBW + tap_dx(tapB);
              const uint32_t a_lo = ((t_lo0 & 0x3FFFu) + (uint32_t)offA) | ((uint32_t)(offB - offA) << 16);
              const uint32_t b_lo = w_lo0 + (uint32_t)(pr * (Cfg::TAP_BYTES >> 4));
              if (NISSUE > 1 && pr == Cfg::NPAIR - 2) *turn = (uint32_t)(k + 1);
#pragma unroll
              for (int tl = 0; tl < 2; ++tl) {
                const uint64_t ad = desc_make(a_lo + (uint32_t)(tl * TBW), A_HI_T), bd = desc_make(b_lo, B_HI);
                if (KIND == KIND_I8) mma_i8(d0 + tl * N, ad, bd, IDESC, pr != 0 ? 1u : 0u);
                else                 mma_f16(d0 + tl * N, ad, bd, IDESC, pr != 0 ? 1u : 0u);
              }
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < Cfg::TAPS; ++tap) {
              const uint32_t b_lo = w_lo0 + (uint32_t)(tap * (Cfg::TAP_BYTES >> 4));
              const uint32_t toff = (uint32_t)(tap_dy(tap) * TBW + tap_dx(tap));
              if (NISSUE > 1 && tap == Cfg::TAPS - 4) *turn = (uint32_t)(k + 1);
#pragma unroll
              for (int tl = 0; tl < 2; ++tl) {
#pragma unroll
                for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
                  const uint64_t ad = desc_make(t_lo0 + toff + (uint32_t)(tl * TBW) + ks * ((2 * Cfg::TAIL_CHUNK_BYTES) >> 4), A_HI_T);
                  const uint64_t bd = desc_make(b_lo + ks * B_KSTEP, B_HI);
                  if (KIND == KIND_I8) mma_i8(d0 + tl * N, ad, bd, IDESC, (tap | ks) != 0 ? 1u : 0u);
                  else                 mma_f16(d0 + tl * N, ad, bd, IDESC, (tap | ks) != 0 ? 1u : 0u);
                }
              }
            }
          }
        } else
        if constexpr (Cfg::PAIR) {
#pragma unroll
          for (int pr = 0; pr < Cfg::NPAIR; ++pr) {
            // pair 0 = tap 0 alone (its partner chunk has zero weights), pair p = taps 2p-1, 2p; windows ascend with the tap index
            const int tapA = pr == 0 ? 0 : 2 * pr - 1, tapB = pr == 0 ? 1 : 2 * pr;
            const int offA = tap_dy(tapA) * BW + tap_dx(tapA), offB = tap_dy(tapB) * BW + tap_dx(tapB);
            const uint32_t a_lo = ((a_lo0 & 0x3FFFu) + (uint32_t)offA) | ((uint32_t)(offB - offA) << 16);
            const uint32_t b_lo = w_lo0 + (uint32_t)(pr * (Cfg::TAP_BYTES >> 4));
            if (NISSUE > 1 && pr == Cfg::NPAIR - 2) *turn = (uint32_t)(k + 1);
#pragma unroll
            for (int tl = 0; tl < GX; ++tl) {
              if (tl < ntile) {
                const uint64_t ad = desc_make(a_lo + (uint32_t)(tl * 8), A_HI), bd = desc_make(b_lo, B_HI);
                if (KIND == KIND_I8) mma_i8(d0 + tl * N, ad, bd, IDESC, pr != 0 ? 1u : 0u);
                else                 mma_f16(d0 + tl * N, ad, bd, IDESC, pr != 0 ? 1u : 0u);
              }
            }
          }
        } else
#pragma unroll TAP_UNROLL
        for (int tap = 0; tap < Cfg::TAPS; ++tap) {
          uint32_t b_lo;
          if (Cfg::WRES) {
            b_lo = w_lo0 + (uint32_t)(tap * (Cfg::TAP_BYTES >> 4));
          } else {
            ok = mbar_wait(&w_full[ws], wphase, abort_flag, 0x505u, k * 32 + tap);
            if (!ok) break;
            tc_fence_after();
            b_lo = w_lo0 + (uint32_t)(ws * (Cfg::TAP_BYTES >> 4));
          }
          const uint32_t toff = Cfg::TAPS == 1 ? 0u : (uint32_t)(tap_dy(tap) * BW + tap_dx(tap));
          if (NISSUE > 1 && tap == (Cfg::TAPS > 4 ? Cfg::TAPS - 4 : 0)) *turn = (uint32_t)(k + 1);
#pragma unroll
          for (int tl = 0; tl < GX; ++tl) {
            if (tl < ntile) {
#pragma unroll
              for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
                const uint64_t ad = desc_make(a_lo0 + toff + (uint32_t)(tl * 8) + ks * A_KSTEP, A_HI);
                const uint64_t bd = desc_make(b_lo + ks * B_KSTEP, B_HI);
                if (KIND == KIND_I8) mma_i8(d0 + tl * N, ad, bd, IDESC, (tap | ks) != 0 ? 1u : 0u);
                else                 mma_f16(d0 + tl * N, ad, bd, IDESC, (tap | ks) != 0 ? 1u : 0u);
              }
            }
          }
          if (!Cfg::WRES) {
            mma_commit(&w_empty[ws]);
            if (++ws == Cfg::WSTAGES) { ws = 0; wphase ^= 1; }
          }
        }
        if (ok) { mma_commit(&in_empty[st]); mma_commit(&acc_full[buf]); }
        ERNET_TL(k, 3);
      }
    }
    __syncwarp();
  } else if (warp == Cfg::WARP_HALO) {
    // ------------------------------------------------------------------ zero halo of the output images this CTA starts
    if (OUT != OUT_NHWC) {
      pdl_wait();
      constexpr int BORDER = 3 * OP + (OP - 3) * 3;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int img = u / Cfg::UNITS_PER_IMG;
        if (u - img * Cfg::UNITS_PER_IMG != 0) continue;
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
    const int ehalf = (warp - 3) >> 2;           // which tiles of the unit this warp takes: ehalf, ehalf + EPW, ..
    const int rr = 4 * q4 + (lane >> 3), cc = lane & 7;
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int qsel = (xodd ? 2 : 0) + (yodd ? 1 : 0);
    const bool susp = g_epi_suspend != 0;
    int k = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      const bool tail = Cfg::TAIL && ty == Cfg::TR;
      const int ntile = tail ? 0 : min(GX, Cfg::TCOLS - ux * GX);
      const int buf = k & 1, use = k >> 1;
      if (!mbar_wait_epi(&acc_full[buf], use & 1, abort_flag, 0x600u + warp, k, susp)) break;
      if (threadIdx.x == 96) ERNET_TL(k, 4);
      tc_fence_after();
      if constexpr (Cfg::TAIL) {
        if (tail) {                                // lane = x, tiles 0 / 1 = the two rows; the EPW warps of a lane quarter share the
          const int x = 32 * q4 + lane;            // blocks of 32 columns (one warp alone made the tail unit epilogue-bound)
          const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * GX * N);
          epilogue_tail<Cfg, KIND, OUT>(par, tbase, x, x < Cfg::HU, xodd, out, img, ehalf, Cfg::EPW);
        }
      }
      for (int tl = ehalf; tl < ntile; tl += Cfg::EPW) {
        const int y = ty * 16 + rr, x = (ux * GX + tl) * 8 + cc;
        const bool valid = (y < Cfg::HU) && (x < Cfg::HU);
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * GX * N + tl * N);
        epilogue_tile<Cfg, KIND, OUT>(par, tbase, y, x, valid, xodd, yodd, qsel, out, img);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (threadIdx.x == 96) ERNET_TL(k, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ERNET_TL(31, 7);
  ERNET_CHAIN_EXIT(1);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map of a padded P8 / P16 activation tensor (batch, NC, WP, WP x 16 bytes): the pixel and byte axes are fused
// into one inner axis of 32-bit words so that a box row is one contiguous run of BW*16 bytes (a 16-byte inner box
// makes the TMA unit issue one request per pixel).  Box = (BW*4 words, BH rows, NC chunks, 1 image).
template <class Cfg>
inline int make_input_map(CUtensorMap* map, const void* base, int batch) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)Cfg::WP * 4, (cuuint64_t)Cfg::WP, (cuuint64_t)Cfg::NC, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)Cfg::WP * 16, (cuuint64_t)Cfg::WP * Cfg::WP * 16,
                                 (cuuint64_t)Cfg::NC * Cfg::WP * Cfg::WP * 16};
  const cuuint32_t box[4] = {(cuuint32_t)Cfg::BW * 4, (cuuint32_t)Cfg::BH, (cuuint32_t)(Cfg::PAIR ? 1 : Cfg::NC), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ERNET_OK;
}

// Tail box (PCfg::TAIL): 8 rows x 128 pixels, 64-bit elements (two per pixel) so that the 2048-byte row stays within the
// 256-element box limit; columns right of the padded image are zero-filled by the TMA unit.
template <class Cfg>
inline int make_tail_map(CUtensorMap* map, const void* base, int batch) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)Cfg::WP * 2, (cuuint64_t)Cfg::WP, (cuuint64_t)Cfg::NC, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)Cfg::WP * 16, (cuuint64_t)Cfg::WP * Cfg::WP * 16,
                                 (cuuint64_t)Cfg::NC * Cfg::WP * Cfg::WP * 16};
  const cuuint32_t box[4] = {(cuuint32_t)Cfg::TBW * 2, (cuuint32_t)Cfg::TBH, (cuuint32_t)(Cfg::PAIR ? 1 : Cfg::NC), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled (tail box) failed with CUresult %d", (int)r);
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT>
inline int launch_acff_pblock(const void* in, const void* wimg, const EpiParams<Cfg::N>& par, void* out, int batch, int num_sms,
                              cudaStream_t stream) {
  CUtensorMap map, map_tail;
  int rc = make_input_map<Cfg>(&map, in, batch);
  if (rc) return rc;
  if (Cfg::TAIL) { if ((rc = make_tail_map<Cfg>(&map_tail, in, batch))) return rc; }
  else map_tail = map;
  const int total = batch * Cfg::UNITS_PER_IMG;
  const int grid = total < num_sms ? total : num_sms;
  ERNET_CUDA(launch_pdl(acff_pblock_kernel<Cfg, KIND, OUT>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, map, map_tail,
                        static_cast<const uint16_t*>(wimg), par, static_cast<uint16_t*>(out), batch));
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT>
inline int set_pblock_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(acff_pblock_kernel<Cfg, KIND, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  return ERNET_OK;
}

// Persistent configurations (same math as CfgBlock*): GX tiles per unit, input ring depth, weight residency.
using PBlock1 = PCfg<2, 64, 69, 66, 3, 3, true, 1>;            // 15 units / image, 21 KB boxes, weights resident
// default for the 16-bit Squeeze_ErNET: rows 64, 65 as one tail unit - 12 + 1 units, 38 instead of 45 tiles per image
using PBlock1T = PCfg<2, 64, 69, 66, 3, 3, true, 1, true, 25, true, 64, false, /*TAIL*/ true>;
using PBlock2 = PCfg<8, 96, 33, 30, 2, 2, false, 8>;           //  4 units / image, 62 KB boxes
using PBlock3 = PCfg<12, 128, 15, 12, 2, 2, false, 4>;         //  1 unit  / image, box = whole padded image
// Squeeze_RedConv: conv_red2 (96 -> 48, 1x1, bias only) + 2x2 pool as a 1-tap instance: 2 units / image, 98 KB boxes
// block 1 with one real input chunk (int8 Squeeze_ErNET, 16-bit Squeeze_RedConv): 13 two-tap MMAs per tile
using PBlock1P = PCfg<2, 64, 69, 66, 3, 3, true, 1, true, 25, true, 64, /*PAIR*/ true>;
using PBlock1PT = PCfg<2, 64, 69, 66, 3, 3, true, 1, true, 25, true, 64, /*PAIR*/ true, /*TAIL*/ true>;
using EBlock1 = PCfg<2, 64, 119, 116, 3, 3, true, 1>;          // ErNET block 1: 40 units / image
using PRed2R = PCfg<12, 64, 30, 30, 4, 2, true, 1, /*POOL*/ true, /*TAPS*/ 1, /*ACT*/ false, /*NREAL*/ 48>;
// int8 engine: same instance writing the int8 pool2 tensor as 4 chunks of 16 channels (48 real + 16 exact zeros)
using PRed2RQ = PCfg<12, 64, 30, 30, 4, 2, true, 1, /*POOL*/ true, /*TAPS*/ 1, /*ACT*/ false, /*NREAL*/ 64>;

}  // namespace tc
}  // namespace ernet
