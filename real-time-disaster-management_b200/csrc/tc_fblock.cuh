// Fused eval transform + conv1 + ACFF block 1: ONE persistent kernel in which the CUDA-core work of the transform
// (dataloaders/aider.py:421-426, model/squeeze_ernet.py:11,25 - ingest_fast.cuh) runs UNDER the tensor-core work of block 1
// (model/acff.py:37-59 + pool1 - tc_pblock.cuh) on the same SM, instead of in a kernel of its own in front of it.
//
// Why.  The stand-alone transform + conv1 kernel is bound by CUDA-core instruction issue (47 us per 256 frames with the
// tensor pipe idle) and block 1 is bound by the tensor pipe (25 MMAs of N = 64 per tile, 59-64 us, about half of its issue
// slots idle).  The two want different halves of the SM.  Running them as two
// co-resident CTAs was measured in round 1 and lost (the block scheduler packs CUDA-core CTAs until the epilogue warps
// starve); here the split is explicit: every CTA carries kHelperWarps "helper" warps that do nothing but transform
// bands of frames, next to the usual block-1 roles (TMA producer, two MMA issuers, 12 epilogue warps, halo warp).
//
// Data flow.  Helper warps of CTA c take the bands c, c + grid, c + 2 grid, .. (band = kStemBand conv1 rows of one frame,
// image-major order) and write the stem tensor to global memory (it stays in L2) exactly as the stand-alone kernel does;
// when a band is complete one thread publishes it: __threadfence + atomicAdd(ready[img]).  The TMA producer of ANY CTA,
// before it loads the first patch of a unit of image img, waits until ready[img] == bands per image (ld.acquire.gpu),
// orders the async proxy after it (fence.proxy.async) and issues the load.  Helpers never wait for consumers and all CTAs
// of the grid are co-resident (grid <= #SMs, one CTA per SM), so the only waits are forward in image order: no deadlock.
// Consumers run ~one band behind the helpers; the kernel takes max(transform, block 1) + the first band (~6 us).
// Measured (B200, bf16, 256 frames): 118-122 us against 49 + 64 us for the two kernels - parity, not a win: the helpers are
// the bottleneck (15 warps carry 54 % of the kernel's 59 M warp instructions at 1.2 eligible warps per scheduler; the
// stand-alone transform kernel needs 28 warps per SM to issue at 71 %).  Hence opt-in (ernet_set_fuse_ingest); DESIGN.md 5a.
// The flags re-arm themselves: every unit's producer counts itself in taken[img]; the 15th resets both words, so a
// replayed CUDA graph (constant kernel arguments) and the next launch find zeros.  (A watchdog abort leaves them dirty:
// the host clears them when it reports the timeout.)
#pragma once
#include "ingest_fast.cuh"
#include "tc_pblock.cuh"

namespace ernet {
namespace tc {

constexpr int kHelperWarps = 15;                       // 480 threads: 3 row groups x 140 columns in the horizontal pass; 32 warps x 64 registers
constexpr int kHelperThreads = 32 * kHelperWarps;
constexpr int kStemBands = (69 + kStemBand - 1) / kStemBand;

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
constexpr int kTabInts = kCrop * (1 + 5 + 1 + 1 + 5);     // xmin, kx[5], ymin, ylen, ky[5] of the 140 output rows / columns

template <class Cfg>
struct FCfg {
  static constexpr int THREADS = kHelperThreads + Cfg::THREADS;
  static constexpr int OFF_HBAR = Cfg::OFF_BAR + 256;                              // helpers' bulk-copy mbarrier
  static constexpr int OFF_TAB = OFF_HBAR + 16;                                    // resize tables (staged once: the helper
  static constexpr int OFF_SF = OFF_TAB + kTabInts * 4;                            // warps cannot hide global-load latency)
  static constexpr int OFF_INGEST = (OFF_SF + (int)sizeof(StemFrag) + 127) / 128 * 128;   // FastGeom region of the helpers
  static_assert(OFF_SF % 16 == 0, "StemFrag is read with 16-byte loads");
  static_assert(Cfg::WRES && Cfg::TAPS == 25 && Cfg::POOL && Cfg::N == 64, "block 1 of the Squeeze models");
};

// Cfg/KIND/OUT: block 1 instance (PBlock1 / PBlock1P);  T/CS/FSOUT: the transform + conv1 instance that feeds it.
template <class Cfg, int KIND, int OUT, typename T, int CS, int FSOUT>
__global__ void __launch_bounds__(FCfg<Cfg>::THREADS, 1)
ingest_block1_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint16_t* __restrict__ wimg,
                     const __grid_constant__ EpiParams<Cfg::N> par, uint16_t* __restrict__ out, int batch,
                     const uint8_t* __restrict__ frames, const uint8_t* __restrict__ frames_end, int H, int W, int bgr,
                     const int* __restrict__ xmin, const int* __restrict__ kx, const int* __restrict__ ymin, const int* __restrict__ ylen,
                     const int* __restrict__ ky, const StemFrag* __restrict__ sf, const __grid_constant__ FastGeom geo,
                     const __grid_constant__ StemQ sq, int zero_chunk1, void* __restrict__ stem,
                     uint32_t* __restrict__ ready, uint32_t* __restrict__ taken) {
  using F = FCfg<Cfg>;
  constexpr int N = Cfg::N, GX = Cfg::GX, NSTAGE = Cfg::NSTAGE, BW = Cfg::BW, OP = Cfg::OP;
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr uint32_t IDESC = KIND == KIND_I8 ? instr_desc(2u, 1u, 128u, (uint32_t)N) : instr_desc(1u, BF16 ? 1u : 0u, 128u, (uint32_t)N);
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? Cfg::NREAL / 16 : Cfg::NREAL / 8;
  constexpr int tl_kernel = 0;
  (void)tl_kernel;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_w = bars;             // [1]  resident weights
  uint64_t* in_full = bars + 1;       // [4]
  uint64_t* in_empty = bars + 5;      // [4]
  uint64_t* acc_full = bars + 25;     // [2]
  uint64_t* acc_empty = bars + 27;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 29);
  volatile uint32_t* abort_flag = tmem_slot + 1;
  volatile uint32_t* turn = tmem_slot + 2;
  uint64_t* hbar = reinterpret_cast<uint64_t*>(smem + F::OFF_HBAR);

  const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = hw_warp - kHelperWarps;               // block-1 role index (negative: helper warp)
  const int total_units = batch * Cfg::UNITS_PER_IMG;
  ERNET_CHAIN_ENTRY(1);

  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    *turn = 0u;
    mbar_init(bar_w, 1);
    mbar_init(hbar, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], Cfg::NEPI); }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp < 0) {
    // ------------------------------------------------------------------ helper warps: transform + conv1, band by band
    // constants first (not ordered after the previous kernel): resize tables and conv1 fragments -> shared memory
    int* s_tab = reinterpret_cast<int*>(smem + F::OFF_TAB);
    StemFrag* s_sf = reinterpret_cast<StemFrag*>(smem + F::OFF_SF);
    int* s_xmin = s_tab; int* s_kx = s_tab + kCrop; int* s_ymin = s_kx + 5 * kCrop; int* s_ylen = s_ymin + kCrop; int* s_ky = s_ylen + kCrop;
    for (int i = threadIdx.x; i < kCrop; i += kHelperThreads) { s_xmin[i] = __ldg(xmin + i); s_ymin[i] = __ldg(ymin + i); s_ylen[i] = __ldg(ylen + i); }
    for (int i = threadIdx.x; i < 5 * kCrop; i += kHelperThreads) { s_kx[i] = __ldg(kx + i); s_ky[i] = __ldg(ky + i); }
    for (int i = threadIdx.x; i < (int)(sizeof(StemFrag) / 16); i += kHelperThreads)
      reinterpret_cast<uint4*>(s_sf)[i] = __ldg(reinterpret_cast<const uint4*>(sf) + i);
    BandSyncNamed<kHelperThreads>::sync();
    pdl_wait();                                          // frames / the stem buffer belong to earlier work of the stream
    ERNET_CHAIN_WAITED(1);
    const int nitems = batch * kStemBands;
    uint8_t* fsm = smem + F::OFF_INGEST;
    uint32_t uses = 0;                                   // completed phases of hbar
    int item = blockIdx.x;
    BandCopy cur{};
    bool issued = false;
    if (item < nitems) {
      cur = band_geometry<1>(item / kStemBands, item % kStemBands, frames, frames_end, H, W, s_ymin, s_ylen);
      issued = cur.bulk;
      if (cur.bulk && threadIdx.x == 0) band_issue_bulk(fsm, hbar, cur);
    }
    for (; item < nitems; item += gridDim.x) {
      const int b = item / kStemBands, band = item - b * kStemBands;
      const int nitem = item + (int)gridDim.x;
      BandCopy nxt{};
      const bool has_next = nitem < nitems;
      if (has_next) nxt = band_geometry<1>(nitem / kStemBands, nitem % kStemBands, frames, frames_end, H, W, s_ymin, s_ylen);
      // the raw rows are dead once the horizontal pass is done: the NEXT band's bulk copy runs under phases 2 and 3
      auto prefetch = [&]() {
        if (has_next && nxt.bulk && threadIdx.x == 0) { fence_proxy_async(); band_issue_bulk(fsm, hbar, nxt); }
      };
      ingest_stem5_band<T, CS, FSOUT, kHelperThreads, BandSyncNamed<kHelperThreads>, 1>(
          fsm, geo, hbar, uses, cur, issued, threadIdx.x, b, band, frames, frames_end, W, bgr, s_xmin, s_kx, s_ymin, s_ky, s_sf, sq,
          zero_chunk1, stem, prefetch);
      if (cur.bulk) ++uses;
      BandSyncNamed<kHelperThreads>::sync();             // every helper thread's stores of the band are issued (and its buffers are free)
      if (threadIdx.x == 0) red_release_gpu_add(ready + b, 1u);    // one release for the whole group (cumulativity over the barrier)
      cur = nxt;
      issued = nxt.bulk;
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ input producer
    if (lane == 0) {
      mbar_expect_tx(bar_w, Cfg::W_BYTES);               // constants first: not ordered after the previous kernel (PDL)
      bulk_g2s(s_w, wimg, Cfg::W_BYTES, bar_w);
    }
    pdl_wait();
    int k = 0, seen_img = -1;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      if (lane == 0) {
        const int st = k % NSTAGE, use = k / NSTAGE;
        // waits of warps that are off the critical path SUSPEND (mbarrier.try_wait) instead of polling: a polling warp is
        // always eligible and takes issue slots from the helper warps, which is what this kernel exists to avoid
        if (use > 0 && !mbar_wait_suspend(&in_empty[st], (use - 1) & 1, abort_flag, 0x500u, k)) break;
        if (img != seen_img) {                           // every band of the image published?
          if (ld_acquire_gpu(ready + img) < (uint32_t)kStemBands) {
            const long long t0 = clock64();
            bool ok = true;
            while (ld_acquire_gpu(ready + img) < (uint32_t)kStemBands) {
              __nanosleep(200);
              if (*abort_flag) { ok = false; break; }
              if (clock64() - t0 > 300000000LL) { tc_raise_timeout(0x5f0u, (uint32_t)img); *abort_flag = 1u; ok = false; break; }
            }
            if (!ok) break;
          }
          seen_img = img;
        }
        fence_proxy_async_global();                      // the TMA engine reads what the helpers' generic stores published
        if (atomicAdd(taken + img, 1u) == (uint32_t)(Cfg::UNITS_PER_IMG - 1)) {   // last unit of the image: re-arm its flags
          taken[img] = 0u;
          ready[img] = 0u;
        }
        mbar_expect_tx(&in_full[st], Cfg::STAGE_BYTES);
        tma_load_4d(smem + st * Cfg::STAGE_STRIDE, &tmap_in, (ux * GX * 8) * 4, ty * 16, 0, img, &in_full[st]);
      }
    }
  } else if (warp == 1 || warp == Cfg::WARP_MMA2) {
    // ------------------------------------------------------------------ MMA issuers (even / odd units), as in tc_pblock.cuh
    const int me = warp == 1 ? 0 : 1;
    if (elect_one()) {
      // every wait of this kernel suspends or backs off: block 1 is not its bottleneck (the helpers are), and a spinning
      // issuer warp costs its scheduler up to a quarter of the issue slots the helpers need
      bool ok = mbar_wait_suspend(bar_w, 0, abort_flag, 0x502u);
      const uint32_t in_addr = smem_u32(smem), w_addr = smem_u32(s_w);
      constexpr uint32_t A_HI = desc_hi(BW * 16), B_HI = desc_hi(128);
      const uint32_t w_lo0 = desc_lo(w_addr, N * 16);
      int k = me;
      for (int u = blockIdx.x + me * (int)gridDim.x; u < total_units && ok; u += 2 * (int)gridDim.x, k += 2) {
        const int r = u % Cfg::UNITS_PER_IMG, ux = r % Cfg::UX;
        const int ntile = min(GX, Cfg::TCOLS - ux * GX);
        const int st = k % NSTAGE, buf = k & 1, use = k >> 1;
        ok = mbar_wait_suspend(&in_full[st], (k / NSTAGE) & 1, abort_flag, 0x503u, k);
        if (ok && use > 0) ok = mbar_wait_suspend(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x504u, k);
        if (!ok) break;
        while (*turn < (uint32_t)k) { __nanosleep(32); if (*abort_flag) { ok = false; break; } }
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_lo0 = desc_lo(in_addr + st * Cfg::STAGE_STRIDE + (uint32_t)((2 * BW + 2) * 16), Cfg::CHUNK_BYTES);
        const uint32_t d0 = tmem_base + (uint32_t)(buf * GX * N);
        if constexpr (Cfg::PAIR) {
#pragma unroll
          for (int pr = 0; pr < Cfg::NPAIR; ++pr) {
            const int tapA = pr == 0 ? 0 : 2 * pr - 1, tapB = pr == 0 ? 1 : 2 * pr;
            const int offA = tap_dy(tapA) * BW + tap_dx(tapA), offB = tap_dy(tapB) * BW + tap_dx(tapB);
            const uint32_t a_lo = ((a_lo0 & 0x3FFFu) + (uint32_t)offA) | ((uint32_t)(offB - offA) << 16);
            const uint32_t b_lo = w_lo0 + (uint32_t)(pr * (Cfg::TAP_BYTES >> 4));
            if (pr == Cfg::NPAIR - 2) *turn = (uint32_t)(k + 1);
#pragma unroll
            for (int tl = 0; tl < GX; ++tl) {
              if (tl < ntile) {
                const uint64_t ad = desc_make(a_lo + (uint32_t)(tl * 8), A_HI), bd = desc_make(b_lo, B_HI);
                if (KIND == KIND_I8) mma_i8(d0 + tl * N, ad, bd, IDESC, pr != 0 ? 1u : 0u);
                else                 mma_f16(d0 + tl * N, ad, bd, IDESC, pr != 0 ? 1u : 0u);
              }
            }
          }
        } else {
#pragma unroll
          for (int tap = 0; tap < Cfg::TAPS; ++tap) {
            const uint32_t b_lo = w_lo0 + (uint32_t)(tap * (Cfg::TAP_BYTES >> 4));
            const uint32_t toff = (uint32_t)(tap_dy(tap) * BW + tap_dx(tap));
            if (tap == Cfg::TAPS - 4) *turn = (uint32_t)(k + 1);
#pragma unroll
            for (int tl = 0; tl < GX; ++tl) {
              if (tl < ntile) {
                const uint64_t ad = desc_make(a_lo0 + toff + (uint32_t)(tl * 8), A_HI);
                const uint64_t bd = desc_make(b_lo, B_HI);
                if (KIND == KIND_I8) mma_i8(d0 + tl * N, ad, bd, IDESC, tap != 0 ? 1u : 0u);
                else                 mma_f16(d0 + tl * N, ad, bd, IDESC, tap != 0 ? 1u : 0u);
              }
            }
          }
        }
        mma_commit(&in_empty[st]);
        mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp == Cfg::WARP_HALO) {
    // ------------------------------------------------------------------ zero halo of the output images this CTA starts
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
  } else if (warp >= 3 && warp < 3 + Cfg::NEPI) {
    // ------------------------------------------------------------------ epilogue (one TMEM load in flight: 72 registers)
    const int q4 = hw_warp & 3;                          // TMEM lane quarter = hardware warp id mod 4
    const int ehalf = (warp - 3) >> 2;
    const int rr = 4 * q4 + (lane >> 3), cc = lane & 7;
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int qsel = (xodd ? 2 : 0) + (yodd ? 1 : 0);
    pdl_wait();                                          // stores below must not overtake the previous kernel's readers
    int k = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      const int ntile = min(GX, Cfg::TCOLS - ux * GX);
      const int buf = k & 1, use = k >> 1;
      if (!mbar_wait_suspend(&acc_full[buf], use & 1, abort_flag, 0x600u + warp, k)) break;
      tc_fence_after();
      for (int tl = ehalf; tl < ntile; tl += Cfg::EPW) {
        const int y = ty * 16 + rr, x = (ux * GX + tl) * 8 + cc;
        const bool valid = (y < Cfg::HU) && (x < Cfg::HU);
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * GX * N + tl * N);
        epilogue_tile<Cfg, KIND, OUT, false>(par, tbase, y, x, valid, xodd, yodd, qsel, out, img);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  ERNET_CHAIN_EXIT(1);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
template <class Cfg>
inline bool fused_fits(const FastGeom& g) { return FCfg<Cfg>::OFF_INGEST + g.total <= 227 * 1024; }

template <class Cfg, int KIND, int OUT, typename T, int CS, int FSOUT>
inline int launch_ingest_block1(const IngestTables& t, const FastGeom& g, const uint8_t* frames, int batch, int bgr, const StemFrag* sf,
                                const StemQ& q, bool zero_chunk1, void* stem, const void* wimg, const EpiParams<Cfg::N>& par,
                                void* out, int num_sms, uint32_t* ready, uint32_t* taken, cudaStream_t stream) {
  CUtensorMap map;
  int rc = make_input_map<Cfg>(&map, stem, batch);
  if (rc) return rc;
  const int total = batch * Cfg::UNITS_PER_IMG;
  const int grid = total < num_sms ? total : num_sms;    // every CTA must be resident: consumers wait for other CTAs' helpers
  ERNET_CUDA(launch_pdl(ingest_block1_kernel<Cfg, KIND, OUT, T, CS, FSOUT>, dim3(grid), dim3(FCfg<Cfg>::THREADS),
                        (size_t)(FCfg<Cfg>::OFF_INGEST + g.total), stream, map, static_cast<const uint16_t*>(wimg), par,
                        static_cast<uint16_t*>(out), batch, frames, frames + (size_t)batch * t.H * t.W * 3, t.H, t.W, bgr, t.d_xmin, t.d_kx,
                        t.d_ymin, t.d_ylen, t.d_ky, sf, g, q, zero_chunk1 ? 1 : 0, stem, ready, taken));
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT, typename T, int CS, int FSOUT>
inline int set_fblock_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(ingest_block1_kernel<Cfg, KIND, OUT, T, CS, FSOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return ERNET_OK;
}

}  // namespace tc
}  // namespace ernet
