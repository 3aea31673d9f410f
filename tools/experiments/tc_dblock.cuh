// ACFF block 1 with the depthwise stage on the CUDA cores and ONLY the 1x1 convolution on the tensor cores
// (model/acff.py:25-35,46-53), persistent, sm_100a.
//
// The 25-tap dense form (tc_block.cuh) costs 25 MMAs of N = 64 per tile, and at N = 64 an MMA is bound by the fetch of
// its A operand (48.6 cycles instead of 32, tools/mma_rate.cu): 3655 cycles per unit of three tiles, 59 us per 256
// images - the longest kernel of the step.  Block 1 has only 16 input channels, so its depthwise trio is cheap in
// packed fp16 arithmetic (27 HFMA2 x 4 per pixel and 8-channel chunk), and what is left for the tensor pipe is the
// real 1x1 convolution: K = 48, three MMAs per tile.  Per unit:
//   warp 0        TMA box load of the input patch (22 x 30 pixels x 2 chunks of 8 fp16 channels) into a 3-stage ring
//   warps 2-13    depthwise (two sets of six, even / odd units): one thread = 4 consecutive pixels x 8 channels, one
//                 dilation at a time; the columns of a patch row are loaded once per branch (LDS.128), tap weights are broadcast
//                 loads; results go straight into the UMMA A operand [tile][k-chunk = branch*2 + chunk][128 rows][16 B]
//                 (the concat of acff.py:46 is the K order), double buffered
//   warp 1        9 tcgen05.mma (M = 128, N = 64, K = 16) per unit, accumulators double buffered in TMEM
//   warps 16-27   epilogue as in tc_pblock.cuh: bias, LeakyReLU, BN, 16-bit, 2x2 max-pool by register exchange, P8 store
//   warp 14       zero halo of the output images (15 idle)
// STATUS: experiment, off by default (ERNET_DW_BLOCK1=1 enables it).  Results match the oracle to the same tolerance as
// the 25-tap kernel, but it takes 82 us per 256 images against 59 us.  ncu: the kernel is bound by instruction ISSUE, not
// by shared memory or latency - 14.6 k warp instructions per unit at ~2.9 per cycle: 2.6 k HFMA2 + ~1 k loads/stores for
// the depthwise stage, ~5.4 k for the twelve epilogue warps (bias / LeakyReLU / BN / pack / pool exchange: free while the
// tensor pipe was the bottleneck, not free here), ~2 k of barrier polling.  With the polling gone the floor is ~2.5-3 k
// cycles per unit against 3.65 k for the 25-tap form: a 20 % gain on this kernel at best, so the MMA-bound form stays.
// Internally fp16 whatever the engine: the stem tensor is written as fp16 (the transform+conv1 kernel packs to fp16 for
// this consumer), products and the 9-tap sums stay far inside fp16 range (|x| < 8, |w| < 1), the 1x1 conv runs
// kind::f16 with fp32 accumulation, and the output is rounded once to the engine's type (bf16 / fp16).
#pragma once
#include "tc_pblock.cuh"

namespace ernet {
namespace tc {

struct DCfg1 {                                                     // block 1 of Squeeze_ErNET: 16 -> 64, 69x69 -> 66x66 -> pool
  static constexpr int NC = 2, C = 16, N = 64, NREAL = 64, HIN = 69, HU = 66, GX = 3, NSTAGE = 3;
  static constexpr bool POOL = true, ACT = true;
  static constexpr int WP = HIN + 3, BW = 8 * GX + 7, BH = 22;   // 31-pixel pitch: the depthwise warps' 16-byte loads (lane = row / half row) hit 8 distinct banks groups
  static constexpr int CHUNK_BYTES = BH * BW * 16, STAGE_BYTES = NC * CHUNK_BYTES, STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
  static constexpr int TR = (HU + 15) / 16, TCOLS = (HU + 7) / 8, UX = TCOLS / GX, UNITS_PER_IMG = TR * UX;
  static constexpr int KC = 3 * NC;                                // k-chunks of the 1x1 conv: [branch][chunk]
  static constexpr int A_TILE = KC * 128 * 16, A_BUF = GX * A_TILE;
  static constexpr int WF_BYTES = KC * N * 16;                     // fused_conv weights [k-chunk][N][8] fp16
  static constexpr int DWW_BYTES = 27 * C * 2, DWB_BYTES = 3 * C * 2;
  static constexpr int OUT_H = HU / 2, OP = OUT_H + 3;
  static constexpr int OFF_A = NSTAGE * STAGE_STRIDE;
  static constexpr int OFF_WF = OFF_A + 2 * A_BUF;
  static constexpr int OFF_DW = OFF_WF + WF_BYTES;
  static constexpr int OFF_BAR = (OFF_DW + DWW_BYTES + DWB_BYTES + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  // two sets of 6 depthwise warps: set 0 takes the even units, set 1 the odd ones (own A buffer and TMEM buffer each), so
  // that twice as many warps hide each other's shared-memory latency without loading anything twice
  static constexpr int DW_SET = 6, DW_WARPS = 2 * DW_SET, EPI_WARPS = 12;
  static constexpr int WARP_DW0 = 2, WARP_EPI0 = 16, WARP_HALO = 14, THREADS = 32 * 28;
  static_assert(TCOLS % GX == 0, "whole units");
  static_assert(OFF_A % 128 == 0 && A_TILE % 128 == 0, "alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// constants of the kernel, built once per handle (build_dblock1): fp16 images + the 1x1 bias in the epilogue parameters
struct DBlock1Consts {
  uint16_t wf[DCfg1::WF_BYTES / 2];      // [k-chunk][n][8]: fused_conv.weight[n][k], k = branch*16 + c
  uint16_t dww[27 * 16];                 // [branch*9 + ky*3 + kx][c]
  uint16_t dwb[3 * 16];                  // [branch][c]
};

__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int KIND, int OUT>
__global__ void __launch_bounds__(DCfg1::THREADS, 1)
acff_dblock1_kernel(const __grid_constant__ CUtensorMap tmap_in, const DBlock1Consts* __restrict__ consts,
                    const __grid_constant__ EpiParams<64> par, uint16_t* __restrict__ out, int batch) {
  using Cfg = DCfg1;
  constexpr int N = Cfg::N, GX = Cfg::GX, NSTAGE = Cfg::NSTAGE, BW = Cfg::BW, OP = Cfg::OP;
  constexpr uint32_t IDESC = instr_desc(1u, 0u, 128u, (uint32_t)N);          // f16 x f16 -> f32
  constexpr int OUT_CHUNKS = Cfg::NREAL / 8;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_a = smem + Cfg::OFF_A;
  uint8_t* s_wf = smem + Cfg::OFF_WF;
  const uint4* s_dww = reinterpret_cast<const uint4*>(smem + Cfg::OFF_DW);                      // [27][2 chunks] x 16 B
  const uint4* s_dwb = reinterpret_cast<const uint4*>(smem + Cfg::OFF_DW + Cfg::DWW_BYTES);    // [3][2 chunks]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* in_full = bars;           // [4]
  uint64_t* in_empty = bars + 4;      // [4]  6 depthwise warps
  uint64_t* a_full = bars + 8;        // [2]  6 depthwise warps
  uint64_t* a_empty = bars + 10;      // [2]  MMA commit
  uint64_t* acc_full = bars + 12;     // [2]  MMA commit
  uint64_t* acc_empty = bars + 14;    // [2]  12 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  volatile uint32_t* abort_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = batch * Cfg::UNITS_PER_IMG;
  ERNET_CHAIN_ENTRY(1);

  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    for (int i = 0; i < 4; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], Cfg::DW_SET); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], Cfg::DW_SET); mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], Cfg::EPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // constants -> shared memory (independent of the previous kernel)
  for (int i = threadIdx.x; i < (Cfg::WF_BYTES + Cfg::DWW_BYTES + Cfg::DWB_BYTES) / 16; i += Cfg::THREADS)
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
    ERNET_CHAIN_WAITED(1);
    if (lane == 0) {
      int k = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
        const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
        const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
        const int st = k % NSTAGE, use = k / NSTAGE;
        if (use > 0 && !mbar_wait(&in_empty[st], (use - 1) & 1, abort_flag, 0x900u, k)) break;
        mbar_expect_tx(&in_full[st], Cfg::STAGE_BYTES);
        tma_load_4d(smem + st * Cfg::STAGE_STRIDE, &tmap_in, ux * GX * 8 * 4, ty * 16, 0, img, &in_full[st]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: 3 tiles x 3 K steps per unit
    if (elect_one()) {
      const uint32_t a_addr = smem_u32(s_a), w_addr = smem_u32(s_wf);
      constexpr uint32_t AB_HI = desc_hi(128);
      bool ok = true;
      int k = 0;
      for (int u = blockIdx.x; u < total_units && ok; u += gridDim.x, ++k) {
        const int buf = k & 1, use = k >> 1;
        ok = mbar_wait_suspend(&a_full[buf], use & 1, abort_flag, 0x901u, k);
        if (ok && use > 0) ok = mbar_wait_suspend(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x902u, k);
        if (!ok) break;
        tc_fence_after();
#pragma unroll
        for (int tl = 0; tl < GX; ++tl) {
          const uint32_t a_lo = desc_lo(a_addr + (uint32_t)(buf * Cfg::A_BUF + tl * Cfg::A_TILE), 128 * 16);
#pragma unroll
          for (int ks = 0; ks < Cfg::KC / 2; ++ks)
            mma_f16(tmem_base + (uint32_t)(buf * GX * N + tl * N), desc_make(a_lo + (uint32_t)(ks * ((2 * 128 * 16) >> 4)), AB_HI),
                    desc_make(desc_lo(w_addr + (uint32_t)(ks * 2 * N * 16), N * 16), AB_HI), IDESC, ks != 0 ? 1u : 0u);
        }
        mma_commit(&a_empty[buf]);
        mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp >= Cfg::WARP_DW0 && warp < Cfg::WARP_DW0 + Cfg::DW_WARPS) {
    // ------------------------------------------------------------------ depthwise trio on the CUDA cores (fp16)
    const int dwarp = (warp - Cfg::WARP_DW0) % Cfg::DW_SET, set = (warp - Cfg::WARP_DW0) / Cfg::DW_SET;
    const int v = dwarp / 3, tl = dwarp % 3;              // 8-channel chunk (warp-uniform: broadcast weight loads), tile of the unit
    const int ly = lane >> 1, lx0 = (lane & 1) * 4;       // strip of 4 pixels in the 16 x 8 tile
    int k = set;
    for (int u = blockIdx.x + set * (int)gridDim.x; u < total_units; u += 2 * (int)gridDim.x, k += 2) {
      const int st = k % NSTAGE, buf = k & 1, use = k >> 1;
      if (!mbar_wait_suspend(&in_full[st], (k / NSTAGE) & 1, abort_flag, 0x903u + dwarp, k)) break;
      if (use > 0 && !mbar_wait_suspend(&a_empty[buf], (use - 1) & 1, abort_flag, 0x910u + dwarp, k)) break;
      const uint4* patch = reinterpret_cast<const uint4*>(smem + st * Cfg::STAGE_STRIDE + v * Cfg::CHUNK_BYTES) + ly * BW + 8 * tl + lx0;
      uint4* arow = reinterpret_cast<uint4*>(s_a + buf * Cfg::A_BUF + tl * Cfg::A_TILE) + ly * 8 + lx0;
      // one dilation at a time: 16 accumulators live instead of 48 (the three branches share only the row dy = 1)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int dil = d + 1;
        uint32_t acc[4][4];
        const uint4 b = s_dwb[d * 2 + v];
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
            const uint4 w = s_dww[(d * 9 + ky * 3 + kx) * 2 + v];
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
        for (int px = 0; px < 4; ++px) arow[(d * 2 + v) * 128 + px] = make_uint4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
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
    // ------------------------------------------------------------------ epilogue (warps 16..27): one warp per (lane quarter, tile)
    const int q4 = warp & 3;
    const int tl = (warp - Cfg::WARP_EPI0) >> 2;
    const int rr = 4 * q4 + (lane >> 3), cc = lane & 7;
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int qsel = (xodd ? 2 : 0) + (yodd ? 1 : 0);
    int k = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
      const int img = u / Cfg::UNITS_PER_IMG, r = u - img * Cfg::UNITS_PER_IMG;
      const int ty = r / Cfg::UX, ux = r - ty * Cfg::UX;
      const int buf = k & 1, use = k >> 1;
      if (!mbar_wait_suspend(&acc_full[buf], use & 1, abort_flag, 0xa00u + warp, k)) break;
      tc_fence_after();
      const int y = ty * 16 + rr, x = (ux * GX + tl) * 8 + cc;
      const bool valid = (y < Cfg::HU) && (x < Cfg::HU);
      const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * GX * N + tl * N);
      epilogue_tile<Cfg, KIND, OUT>(par, tbase, y, x, valid, xodd, yodd, qsel, out, img);
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
// fp32 tensors of the layer-wise path (blob_format.h) -> fp16 images of this kernel.
//   pw_w [48][64] (k = branch*16 + c major), dw_w [3][9][16], dw_b [3][16]
inline void build_dblock1(const float* pw_w, const float* dw_w, const float* dw_b, DBlock1Consts* out) {
  auto h16 = [](float f) { __half h = __float2half_rn(f); uint16_t b; memcpy(&b, &h, 2); return b; };
  for (int kc = 0; kc < DCfg1::KC; ++kc)
    for (int n = 0; n < 64; ++n)
      for (int e = 0; e < 8; ++e) out->wf[(kc * 64 + n) * 8 + e] = h16(pw_w[(kc * 8 + e) * 64 + n]);
  for (int i = 0; i < 27 * 16; ++i) out->dww[i] = h16(dw_w[i]);
  for (int i = 0; i < 3 * 16; ++i) out->dwb[i] = h16(dw_b[i]);
}

inline int make_dinput_map(CUtensorMap* map, const void* base, int batch) {
  using Cfg = DCfg1;
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

template <int KIND, int OUT>
inline int launch_acff_dblock1(const void* in, const DBlock1Consts* consts, const EpiParams<64>& par, void* out, int batch, int num_sms,
                               cudaStream_t stream) {
  CUtensorMap map;
  int rc = make_dinput_map(&map, in, batch);
  if (rc) return rc;
  const int total = batch * DCfg1::UNITS_PER_IMG;
  const int grid = total < num_sms ? total : num_sms;
  ERNET_CUDA(launch_pdl(acff_dblock1_kernel<KIND, OUT>, dim3(grid), dim3(DCfg1::THREADS), DCfg1::SMEM_BYTES, stream, map, consts, par,
                        static_cast<uint16_t*>(out), batch));
  return ERNET_OK;
}

template <int KIND, int OUT>
inline int set_dblock1_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(acff_dblock1_kernel<KIND, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, DCfg1::SMEM_BYTES));
  return ERNET_OK;
}

}  // namespace tc
}  // namespace ernet
