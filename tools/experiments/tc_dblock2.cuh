// EXPERIMENT (round 2), measured and rejected - kept outside the product build.  Wiring that was used: block 1 writes pool1
// as fp16 (an OUT_P8H output format of epilogue_tile), run_chunk_tc launches launch_acff_dblock<DBlock2, K16, OUT_P8> in place
// of the CTA-pair kernel, constants built by build_dblock from the layer-wise tensors (T_PW_W / T_DW_W / T_DW_B of block 2),
// epilogue bias = fused_conv.bias; mbar_wait_backoff = mbar_test_wait polling with __nanosleep(32).
// RESULT on B200 (bf16, 256 frames): parity green (23 logits / intermediate tests; logit error against the fp64 reference
// 8.2e-3 instead of 1.41e-2 for the 25-tap bf16 form), but 62 us against 54 us.  The depthwise stage is bound by SHARED-MEMORY
// BANDWIDTH, not by issue slots: a thread reads 90 pixel vectors (LDS.128) + 30 weight vectors for its 4 pixels x 8 channels x
// 27 taps - 1.2 uses per loaded vector - i.e. ~3.8 k cycles of the 128 B/clk shared-memory pipe per unit before any
// conflict, against 2.7 k cycles of instruction issue.  A 4 x 4 register tile would cut the traffic 3.5x but needs 64
// accumulator registers per branch under a 72-register cap.  Suspended vs polling waits made no difference (62 us both).
// ACFF block with the depthwise trio on the CUDA cores and ONLY the 1x1 convolution on the tensor cores
// (model/acff.py:25-35,46-53), persistent, sm_100a.  Used for block 2 of Squeeze_ErNET (64 -> 96 channels).
//
// Why block 2.  The 25-tap dense form (tc_block.cuh) trades the depthwise stage for 8.3x more tensor-core work: block 2
// issues 25 x 4 K-steps = 100 MMAs of N = 96 per tile and is bound by them (54 us per 256 images, 82 % tensor-pipe
// activity on 8x inflated work).  Its depthwise trio is only 27 x 64 MACs per pixel - 864 packed-fp16 FMAs - and its
// epilogue is light (96 channels on 30 x 30 pixels, a third of block 1's), so here the CUDA cores do the depthwise stage
// straight into the UMMA A operand in shared memory (the concat of acff.py:46 is the K order and never leaves the SM) and
// the tensor pipe runs the REAL 1x1 convolution: K = 192, 12 MMAs per tile.  Round 1 tried the same form on block 1 (16 -> 64,
// tools/experiments/tc_dblock.cuh) and lost: there the epilogue alone is 5.4 k warp instructions per unit.
//
// Per unit (one 16 x 8 tile of output pixels; 8 units per image; units dealt round-robin to one CTA per SM):
//   warp 0        TMA box load of the input patch (22 x 15 pixels x 8 chunks of 8 fp16 channels) into a 2-stage ring
//   warps 2-17    depthwise: two sets of eight warps (even / odd units), warp = 8-channel chunk, thread = 4 consecutive
//                 pixels x 8 channels, one dilation at a time; the columns of a patch row are loaded once per branch
//                 (LDS.128), tap weights are broadcast loads; results go into the A operand
//                 [k-chunk = branch * 8 + chunk][128 rows][16 B], double buffered
//   warp 1        12 tcgen05.mma (M = 128, N = 96, K = 16) per unit, accumulators double buffered in TMEM
//   warps 20-27   epilogue (two sets of four, even / odd units): bias, LeakyReLU, BN, 16-bit, 2x2 max-pool by register
//                 exchange, P8 store (tc_block.cuh epilogue_tile, one TMEM load in flight)
//   warp 18       zero halo of the output images (19 idle)
// Every wait polls with a short sleep (mbar_wait_backoff): the kernel lives on CUDA-core issue slots, a spinning warp would
// take them, and a suspended mbarrier.try_wait was measured to wake up late.
// Arithmetic: fp16 whatever the engine - block 1 writes pool1 as fp16 for this consumer (OUT_P8H), the 9-tap sums run in
// packed fp16 (|pool1| stays below ~1e3 with the shipped weights, far inside fp16 range; each HFMA2 rounds to 11 bits),
// the 1x1 conv runs kind::f16 with fp16 weights and fp32 accumulation, and the output is rounded once to the engine's type.
#pragma once
#include "tc_pblock.cuh"

namespace ernet {
namespace tc {

template <int NC_, int N_, int HIN_, int HU_>
struct DCfg {
  static constexpr int NC = NC_, C = 8 * NC_, N = N_, NREAL = N_, HIN = HIN_, HU = HU_, GX = 1, NSTAGE = 2;
  static constexpr bool POOL = true, ACT = true;
  static constexpr int WP = HIN + 3, BW = 8 + 7, BH = 22;          // 15-pixel pitch: the 16-byte loads of a half-warp spread over the banks
  static constexpr int CHUNK_BYTES = BH * BW * 16, STAGE_BYTES = NC * CHUNK_BYTES, STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
  static constexpr int TR = (HU + 15) / 16, TCOLS = (HU + 7) / 8, UX = TCOLS, UNITS_PER_IMG = TR * UX;
  static constexpr int KC = 3 * NC;                                // k-chunks of the 1x1 conv: [branch][chunk]
  static constexpr int A_TILE = KC * 128 * 16;                     // one tile of the A operand
  static constexpr int WF_BYTES = KC * N * 16;                     // fused_conv weights [k-chunk][N][8] fp16
  static constexpr int DWW_BYTES = 27 * C * 2, DWB_BYTES = 3 * C * 2;
  static constexpr int CONST_BYTES = WF_BYTES + DWW_BYTES + DWB_BYTES;
  static constexpr int OUT_H = HU / 2, OP = OUT_H + 3;
  static constexpr int OFF_A = NSTAGE * STAGE_STRIDE;
  static constexpr int OFF_WF = OFF_A + 2 * A_TILE;
  static constexpr int OFF_DW = OFF_WF + WF_BYTES;
  static constexpr int OFF_BAR = (OFF_DW + DWW_BYTES + DWB_BYTES + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int DW_SET = NC, DW_WARPS = 2 * DW_SET;         // one warp per 8-channel chunk, two sets (even / odd units)
  static constexpr int WARP_DW0 = 2, WARP_HALO = WARP_DW0 + DW_WARPS, WARP_EPI0 = (WARP_HALO + 1 + 3) / 4 * 4, EPI_WARPS = 8;
  static constexpr int THREADS = 32 * (WARP_EPI0 + EPI_WARPS);
  static_assert(OFF_A % 128 == 0 && A_TILE % 128 == 0 && CONST_BYTES % 16 == 0, "alignment");
  static_assert(2 * N <= 512 && N % 32 == 0, "two TMEM accumulator buffers");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(THREADS <= 1024, "block size");
};
using DBlock2 = DCfg<8, 96, 33, 30>;      // ACFF2 of Squeeze_ErNET: 64 -> 96, 33x33 -> 30x30 used -> pool 15x15

// constants of the kernel, built once per handle (build_dblock): fp16 images; the 1x1 bias goes in the epilogue parameters
template <class Cfg>
struct DBlockConsts {
  uint16_t wf[Cfg::WF_BYTES / 2];        // [k-chunk][n][8]: fused_conv.weight[n][k], k = branch*C + c
  uint16_t dww[27 * Cfg::C];             // [branch*9 + ky*3 + kx][c]
  uint16_t dwb[3 * Cfg::C];              // [branch][c]
};

__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <class Cfg, int KIND, int OUT>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
acff_dblock_kernel(const __grid_constant__ CUtensorMap tmap_in, const DBlockConsts<Cfg>* __restrict__ consts,
                   const __grid_constant__ EpiParams<Cfg::N> par, uint16_t* __restrict__ out, int batch) {
  constexpr int N = Cfg::N, NC = Cfg::NC, NSTAGE = Cfg::NSTAGE, BW = Cfg::BW, OP = Cfg::OP;
  constexpr uint32_t IDESC = instr_desc(1u, 0u, 128u, (uint32_t)N);          // f16 x f16 -> f32
  constexpr int OUT_CHUNKS = Cfg::NREAL / 8;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_a = smem + Cfg::OFF_A;
  uint8_t* s_wf = smem + Cfg::OFF_WF;
  const uint4* s_dww = reinterpret_cast<const uint4*>(smem + Cfg::OFF_DW);                      // [27][NC] x 16 B
  const uint4* s_dwb = reinterpret_cast<const uint4*>(smem + Cfg::OFF_DW + Cfg::DWW_BYTES);    // [3][NC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* in_full = bars;           // [2]
  uint64_t* in_empty = bars + 2;      // [2]  DW_SET depthwise warps
  uint64_t* a_full = bars + 4;        // [2]  DW_SET depthwise warps
  uint64_t* a_empty = bars + 6;       // [2]  MMA commit
  uint64_t* acc_full = bars + 8;      // [2]  MMA commit
  uint64_t* acc_empty = bars + 10;    // [2]  4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  volatile uint32_t* abort_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = batch * Cfg::UNITS_PER_IMG;
  ERNET_CHAIN_ENTRY(2);

  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], Cfg::DW_SET);
      mbar_init(&a_full[i], Cfg::DW_SET); mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  // constants -> shared memory (independent of the previous kernel)
  for (int i = threadIdx.x; i < Cfg::CONST_BYTES / 16; i += Cfg::THREADS)
    reinterpret_cast<uint4*>(s_wf)[i] = __ldg(reinterpret_cast<const uint4*>(consts) + i);
  fence_proxy_async();                       // the weight image is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ input producer
    pdl_wait();
    ERNET_CHAIN_WAITED(2);
    if (lane == 0) {
      int k = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
        const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
        const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
        const int st = k % NSTAGE, use = k / NSTAGE;
        if (use > 0 && !mbar_wait_backoff(&in_empty[st], (use - 1) & 1, abort_flag, 0x900u, k)) break;
        mbar_expect_tx(&in_full[st], Cfg::STAGE_BYTES);
        tma_load_4d(smem + st * Cfg::STAGE_STRIDE, &tmap_in, ux * 8 * 4, ty * 16, 0, img, &in_full[st]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: KC / 2 K steps per unit
    if (elect_one()) {
      const uint32_t a_addr = smem_u32(s_a), w_addr = smem_u32(s_wf);
      constexpr uint32_t AB_HI = desc_hi(128);
      bool ok = true;
      int k = 0;
      for (int u = blockIdx.x; u < total_units && ok; u += gridDim.x, ++k) {
        const int buf = k & 1, use = k >> 1;
        ok = mbar_wait_backoff(&a_full[buf], use & 1, abort_flag, 0x901u, k);
        if (ok && use > 0) ok = mbar_wait_backoff(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x902u, k);
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_lo = desc_lo(a_addr + (uint32_t)(buf * Cfg::A_TILE), 128 * 16);
#pragma unroll
        for (int ks = 0; ks < Cfg::KC / 2; ++ks)
          mma_f16(tmem_base + (uint32_t)(buf * N), desc_make(a_lo + (uint32_t)(ks * ((2 * 128 * 16) >> 4)), AB_HI),
                  desc_make(desc_lo(w_addr + (uint32_t)(ks * 2 * N * 16), N * 16), AB_HI), IDESC, ks != 0 ? 1u : 0u);
        mma_commit(&a_empty[buf]);
        mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp >= Cfg::WARP_DW0 && warp < Cfg::WARP_DW0 + Cfg::DW_WARPS) {
    // ------------------------------------------------------------------ depthwise trio on the CUDA cores (packed fp16)
    const int v = (warp - Cfg::WARP_DW0) % Cfg::DW_SET, set = (warp - Cfg::WARP_DW0) / Cfg::DW_SET;   // chunk (warp-uniform: broadcast weight loads)
    const int ly = lane >> 1, lx0 = (lane & 1) * 4;       // strip of 4 pixels in the 16 x 8 tile
    int k = set;
    for (int u = blockIdx.x + set * (int)gridDim.x; u < total_units; u += 2 * (int)gridDim.x, k += 2) {
      const int st = k % NSTAGE, buf = k & 1, use = k >> 1;
      if (!mbar_wait_backoff(&in_full[st], (k / NSTAGE) & 1, abort_flag, 0x903u + v, k)) break;
      if (use > 0 && !mbar_wait_backoff(&a_empty[buf], (use - 1) & 1, abort_flag, 0x920u + v, k)) break;
      const uint4* patch = reinterpret_cast<const uint4*>(smem + st * Cfg::STAGE_STRIDE + v * Cfg::CHUNK_BYTES) + ly * BW + lx0;
      uint4* arow = reinterpret_cast<uint4*>(s_a + buf * Cfg::A_TILE) + ly * 8 + lx0;
      // one dilation at a time: 16 accumulators live instead of 48 (the three branches share only the row dy = 1)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int dil = d + 1;
        uint32_t acc[4][4];
        const uint4 b = s_dwb[d * NC + v];
#pragma unroll
        for (int px = 0; px < 4; ++px) { acc[px][0] = b.x; acc[px][1] = b.y; acc[px][2] = b.z; acc[px][3] = b.w; }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int ry = 2 + ky * dil - (dil - 1);          // patch row of this tap row relative to the output row
          uint4 xr[10];
#pragma unroll
          for (int c = 0; c < 10; ++c) xr[c] = patch[ry * BW + c];        // unused columns are dropped by the compiler
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint4 w = s_dww[(d * 9 + ky * 3 + kx) * NC + v];
            const int col = 2 + kx * dil - (dil - 1);
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              acc[px][0] = hfma2_u32(xr[px + col].x, w.x, acc[px][0]);
              acc[px][1] = hfma2_u32(xr[px + col].y, w.y, acc[px][1]);
              acc[px][2] = hfma2_u32(xr[px + col].z, w.z, acc[px][2]);
              acc[px][3] = hfma2_u32(xr[px + col].w, w.w, acc[px][3]);
            }
          }
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) arow[(d * NC + v) * 128 + px] = make_uint4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&in_empty[st]);           // this warp has read everything it needs from the stage
      fence_proxy_async();                                  // generic-proxy writes of A -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
    }
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
  } else if (warp >= Cfg::WARP_EPI0) {
    // ------------------------------------------------------------------ epilogue: set 0 takes the even units, set 1 the odd ones
    const int q4 = warp & 3, set = (warp - Cfg::WARP_EPI0) >> 2;
    const int rr = 4 * q4 + (lane >> 3), cc = lane & 7;
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int qsel = (xodd ? 2 : 0) + (yodd ? 1 : 0);
    pdl_wait();                                             // stores below must not overtake the previous kernel's readers
    int k = set;
    for (int u = blockIdx.x + set * (int)gridDim.x; u < total_units; u += 2 * (int)gridDim.x, k += 2) {
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      const int buf = k & 1, use = k >> 1;
      if (!mbar_wait_backoff(&acc_full[buf], use & 1, abort_flag, 0xa00u + warp, k)) break;
      tc_fence_after();
      const int y = ty * 16 + rr, x = ux * 8 + cc;
      const bool valid = (y < Cfg::HU) && (x < Cfg::HU);
      const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * N);
      epilogue_tile<Cfg, KIND, OUT, false>(par, tbase, y, x, valid, xodd, yodd, qsel, out, img);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  ERNET_CHAIN_EXIT(2);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
// fp32 tensors of the layer-wise path (blob_format.h) -> fp16 images of this kernel.
//   pw_w [3C][N] (k = branch*C + c major), dw_w [3][9][C], dw_b [3][C]
template <class Cfg>
inline void build_dblock(const float* pw_w, const float* dw_w, const float* dw_b, DBlockConsts<Cfg>* out) {
  auto h16 = [](float f) { __half h = __float2half_rn(f); uint16_t b; memcpy(&b, &h, 2); return b; };
  for (int kc = 0; kc < Cfg::KC; ++kc)
    for (int n = 0; n < Cfg::N; ++n)
      for (int e = 0; e < 8; ++e) out->wf[(kc * Cfg::N + n) * 8 + e] = h16(pw_w[(kc * 8 + e) * Cfg::N + n]);
  for (int i = 0; i < 27 * Cfg::C; ++i) out->dww[i] = h16(dw_w[i]);
  for (int i = 0; i < 3 * Cfg::C; ++i) out->dwb[i] = h16(dw_b[i]);
}

template <class Cfg>
inline int make_dinput_map(CUtensorMap* map, const void* base, int batch) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)Cfg::WP * 4, (cuuint64_t)Cfg::WP, (cuuint64_t)Cfg::NC, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)Cfg::WP * 16, (cuuint64_t)Cfg::WP * Cfg::WP * 16, (cuuint64_t)Cfg::NC * Cfg::WP * Cfg::WP * 16};
  const cuuint32_t box[4] = {(cuuint32_t)Cfg::BW * 4, (cuuint32_t)Cfg::BH, (cuuint32_t)Cfg::NC, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT>
inline int launch_acff_dblock(const void* in, const DBlockConsts<Cfg>* consts, const EpiParams<Cfg::N>& par, void* out, int batch, int num_sms,
                              cudaStream_t stream) {
  CUtensorMap map;
  int rc = make_dinput_map<Cfg>(&map, in, batch);
  if (rc) return rc;
  const int total = batch * Cfg::UNITS_PER_IMG;
  const int grid = total < num_sms ? total : num_sms;
  ERNET_CUDA(launch_pdl(acff_dblock_kernel<Cfg, KIND, OUT>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, map, consts, par,
                        static_cast<uint16_t*>(out), batch));
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT>
inline int set_dblock_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(acff_dblock_kernel<Cfg, KIND, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  return ERNET_OK;
}

}  // namespace tc
}  // namespace ernet
