// fp32 ACFF depthwise trio (model/acff.py:25-30,46) for the narrow first blocks (C = 8 / 16), TMA-staged.
//
// The register-tile kernel (simt_layers.cuh) is bound by memory latency at these channel counts: 100 scalar loads
// per thread are all the memory-level parallelism a warp has, two CTAs of 128-register threads fit per SM, and while
// they compute nothing is in flight (block 1: 84 us = 3.5 TB/s).  Here one CTA = one 24x24 tile of output pixels of one
// image: a single cp.async.bulk.tensor box copy brings the 30x30xC halo tile into shared memory - out-of-image
// coordinates are zero-filled by the TMA unit, which IS the conv padding 0/1/2 of the three branches - while the
// threads fetch their tap weights; three CTAs are resident per SM, so two tiles are always in flight under the one being
// computed.  The compute phase is the register-tile one (one channel x 4x4 patch per thread, 27 weights in registers,
// 6.25 shared-memory words per output, bias / ky-major / kx-minor FMA order): bit-identical to the other two kernels.
#pragma once
#include "simt_layers.cuh"
#include "tc_pblock.cuh"   // tma_load_4d, get_encode_fn

namespace ernet {

// C here is the number of channels ONE CTA stages (the whole pixel for C <= 64, a 64- or 32-channel chunk of it for the
// wide maps of the detector's add-fusion blocks and of block 3: the box is then {CC, BW, BH, 1} at channel offset chunk*CC).
template <int C, int TS = 24, int PXW = 4>
struct DwTmaCfg {
  static constexpr int TH = TS, TW = TS, PX = PXW, PY = 4;
  static constexpr int BH = TH + 6, BW = TW + 6;
  static constexpr int PATCHES = (TH / PY) * (TW / PX);          // 36 (4-wide patches), 48 (3-wide)
  static constexpr int ITEMS = PATCHES * C;                      // a thread takes items tid, tid + NT, ...: same channel each time
  static constexpr int NT = 192;                                 // a multiple of C: a thread keeps its channel
  static constexpr uint32_t BOX_BYTES = BH * BW * C * 4;         // 57,600 / 28,800
  static constexpr size_t SMEM = BOX_BYTES + 128;                // + alignment slack
};

template <int CT /*channels of the tensor*/, int C /*channels per CTA*/, bool ADD, int MINB, int NT, int TS, int PXW>
__global__ void __launch_bounds__(NT, MINB)
acff_dw_tma_kernel(const __grid_constant__ CUtensorMap tmap, int out_h, int out_w, int tiles_x, int tiles_y,
                   const float* __restrict__ w /*[3][9][C]*/, const float* __restrict__ bias /*[3][C]*/,
                   float* __restrict__ out) {
  using Cfg = DwTmaCfg<C, TS, PXW>;
  constexpr int PX = Cfg::PX, PY = Cfg::PY, BW = Cfg::BW;
  constexpr int OC = ADD ? CT : 3 * CT;
  constexpr int CHUNKS = CT / C;
  extern __shared__ uint8_t dwt_smem_raw[];
  float* tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dwt_smem_raw) + 127) & ~(uintptr_t)127);   // [BH][BW][C]
  __shared__ __align__(8) uint64_t full;
  int t = blockIdx.x;
  const int chunk = t % CHUNKS; t /= CHUNKS;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int x0 = tx * Cfg::TW, y0 = ty * Cfg::TH;
  if (threadIdx.x == 0) {
    tc::mbar_init(&full, 1);
    tc::fence_mbar_init();
    tc::mbar_expect_tx(&full, Cfg::BOX_BYTES);
    tc::tma_load_4d(tile, &tmap, chunk * C, x0 - 2, y0 - 2, b, &full);
  }
  const int c = threadIdx.x % C;                     // channel within the staged chunk
  const int cg = chunk * C + c;                      // channel of the tensor
  float wr[27], bv[3];
#pragma unroll
  for (int i = 0; i < 27; ++i) wr[i] = __ldg(w + i * CT + cg);
#pragma unroll
  for (int d = 0; d < 3; ++d) bv[d] = __ldg(bias + d * CT + cg);
  __syncthreads();                                   // the barrier initialisation is visible to every waiter
  while (!tc::mbar_try_wait(&full, 0)) {}

  float* ob = out + (size_t)b * out_h * out_w * OC + cg;
  for (int item = threadIdx.x; item < Cfg::ITEMS; item += NT) {
    const int patch = item / C;
    const int lx0 = (patch % (Cfg::TW / PX)) * PX, ly0 = (patch / (Cfg::TW / PX)) * PY;
    const int ox0 = x0 + lx0, oy0 = y0 + ly0;
    if (ox0 >= out_w || oy0 >= out_h) continue;
    float acc[3][PY][PX];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int py = 0; py < PY; ++py)
#pragma unroll
        for (int px = 0; px < PX; ++px) acc[d][py][px] = bv[d];
    const float* p0 = tile + ((size_t)ly0 * BW + lx0) * C + c;
#pragma unroll
    for (int ry = 0; ry < PY + 6; ++ry) {
      float xr[PX + 6];
#pragma unroll
      for (int cx = 0; cx < PX + 6; ++cx) xr[cx] = p0[(ry * BW + cx) * C];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int dil = d + 1;
#pragma unroll
        for (int py = 0; py < PY; ++py) {
          const int tt = ry - 2 - py + (dil - 1);
          if (tt < 0 || tt % dil != 0 || tt / dil > 2) continue;
          const int ky = tt / dil;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int col = 2 + kx * dil - (dil - 1);
#pragma unroll
            for (int px = 0; px < PX; ++px) acc[d][py][px] = fmaf(xr[px + col], wr[d * 9 + ky * 3 + kx], acc[d][py][px]);
          }
        }
      }
    }
    const bool whole = oy0 + PY <= out_h && ox0 + PX <= out_w;
    float* o0 = ob + ((size_t)oy0 * out_w + ox0) * OC;
#pragma unroll
    for (int py = 0; py < PY; ++py) {
      if (!whole && oy0 + py >= out_h) break;
      float* orow = o0 + (size_t)py * out_w * OC;
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        if (!whole && ox0 + px >= out_w) break;
        if constexpr (ADD) {
          orow[px * OC] = (acc[0][py][px] + acc[1][py][px]) + acc[2][py][px];
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) orow[px * OC + d * CT] = acc[d][py][px];
        }
      }
    }
  }
}

template <int CT, bool ADD, int MINB = 3, int NT = 192, int TS = 24, int C = CT, int PXW = 4>
inline int launch_acff_dw_tma_c(const float* x, int batch, int H, int W, int out_h, int out_w, const float* w,
                                const float* bias, float* out, cudaStream_t stream) {
  using Cfg = DwTmaCfg<C, TS, PXW>;
  static_assert(TS % PXW == 0, "tile width is a whole number of patches");
  tc::EncodeTiledFn enc = tc::get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if (reinterpret_cast<uintptr_t>(x) & 15) return -1;                      // TMA needs a 16-byte aligned base: use the other kernel
  CUtensorMap map;
  static_assert(CT % C == 0 && NT % C == 0, "chunking");
  const cuuint64_t dims[4] = {(cuuint64_t)CT, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)CT * 4, (cuuint64_t)W * CT * 4, (cuuint64_t)H * W * CT * 4};
  const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::BH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled (depthwise input) failed with CUresult %d", (int)r);
  // per device and cheap: set on every launch rather than tracking which devices have seen it
  ERNET_CUDA(cudaFuncSetAttribute(acff_dw_tma_kernel<CT, C, ADD, MINB, NT, TS, PXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  const int tiles_x = (out_w + Cfg::TW - 1) / Cfg::TW, tiles_y = (out_h + Cfg::TH - 1) / Cfg::TH;
  const long long grid = (long long)batch * tiles_x * tiles_y * (CT / C);
  if (grid > 0x7fffffffLL) return fail(ERNET_ERR_INVALID_ARG, "depthwise: batch too large for one launch");
  acff_dw_tma_kernel<CT, C, ADD, MINB, NT, TS, PXW><<<(unsigned)grid, NT, Cfg::SMEM, stream>>>(map, out_h, out_w, tiles_x, tiles_y, w, bias, out);
  ERNET_LAUNCH_CHECK("acff_dw_tma_kernel");
  return ERNET_OK;
}

// -1: this shape is not served by the TMA kernel (channel counts other than 8 / 16 / 64, or a map smaller than one tile).
// Measured on B200 (tools/dw_bench, B = 256, L2 flushed; fraction of the measured 6.55 TB/s copy peak):
//   C=16 69->67  (Squeeze_ErNET block 1)  24x24 tiles, 192 threads x 3 CTAs/SM   59 us  5.07 TB/s  77 %   (register tile 83 us)
//   C=16 119->117 (ErNET block 1)         same                                   149 us  6.09 TB/s  93 %   (238 us)
//   C=64 33->31  (block 2)                8x8 tiles, 128 threads x 4 CTAs/SM      49 us  5.28 TB/s  81 %   (64 us)
//   C=8  69->67  (Squeeze_RedConv block 1) 24x24 tiles, 192 threads x 3 CTAs/SM   38 us  3.95 TB/s  60 %   (63 us; 32-byte pixels)
inline int launch_acff_dw_tma(const float* x, int batch, int H, int W, int C, int out_h, int out_w, const float* w,
                              const float* bias, float* out, cudaStream_t stream) {
  // 3-pixel-wide patches at C = 8 / 16: the half- (quarter-) warps that share a load instruction then sit 48 (24) floats
  // apart instead of 64 (32) and hit disjoint banks (4-wide patches: 2-way / 4-way conflicts on every tile read)
  if (C == 16 && out_h >= 24 && out_w >= 24) return launch_acff_dw_tma_c<16, false, 3, 192, 24, 16, 3>(x, batch, H, W, out_h, out_w, w, bias, out, stream);
  if (C == 8 && out_h >= 24 && out_w >= 24) return launch_acff_dw_tma_c<8, false, 3, 192, 24, 8, 3>(x, batch, H, W, out_h, out_w, w, bias, out, stream);
  if (C == 64 && out_h >= 16 && out_w >= 16) return launch_acff_dw_tma_c<64, false, 4, 128, 8>(x, batch, H, W, out_h, out_w, w, bias, out, stream);
  return -1;
}

// Add-fusion flavour (detector, yolov3/models.py:302): 64-channel chunks of the 128- / 256-channel maps, 8x8 tiles.
inline int launch_acff_add_dw_tma(const float* x, int batch, int H, int W, int C, int out_h, int out_w, const float* w,
                                  const float* bias, float* out, cudaStream_t stream) {
  if (out_h < 16 || out_w < 16) return -1;
  if (C == 64) return launch_acff_dw_tma_c<64, true, 4, 128, 8, 64>(x, batch, H, W, out_h, out_w, w, bias, out, stream);
  if (C == 128) return launch_acff_dw_tma_c<128, true, 4, 128, 8, 64>(x, batch, H, W, out_h, out_w, w, bias, out, stream);
  if (C == 256) return launch_acff_dw_tma_c<256, true, 4, 128, 8, 64>(x, batch, H, W, out_h, out_w, w, bias, out, stream);
  return -1;
}

}  // namespace ernet
