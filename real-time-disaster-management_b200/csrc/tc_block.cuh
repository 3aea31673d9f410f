// Fused ACFF block on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
//   ACFF(x) = BN(LeakyReLU(W_f . cat(dw_1(x), dw_2(x), dw_3(x)) + b_f))      model/acff.py:37-59
//
// The three dilated depthwise convs and the 1x1 conv are both linear with nothing in between, so the
// block is ONE dense convolution over the 25 distinct taps of the three stencils:
//   z[n,y,x] = b_eff[n] + sum_{tap} sum_c W_eff[n,tap,c] * X[c, y+dy(tap), x+dx(tap)]
//   W_eff[n,tap,c] = sum_{d : tap in stencil d} W_f[n, d*C + c] * w_d[c,ky,kx],  b_eff = b_f + W_f . cat(b_d)
// (packer: pack_tc.py).  That turns the bandwidth-bound depthwise stage + concat into tensor-core work
// with zero CUDA-core instructions: the input image is staged ONCE in shared memory in the "P8" layout
// [channel chunk of 8][padded row][padded col][8 ch] (16 B per pixel-chunk), which is exactly the
// un-swizzled K-major UMMA operand layout with 16 B row stride - so each tap is the same staged image
// read through a descriptor whose start address is shifted by (dy*P + dx)*16 B, the zero padding of
// acff.py:25-30 is the zero halo of the staged image, and an MMA tile of 128 rows is a 16x8 block of
// output pixels (SBO = one image row).
//
// Warp roles (320 threads, one CTA per SM):
//   warp 0    producer: one 1-D bulk copy (UBLKCP) of the whole P8 image(s); weight images either
//             resident (block 1) or streamed tap by tap through a small ring; then zero-fills the halo
//             of the output tensor
//   warp 1    allocates TMEM, one lane issues tcgen05.mma (M=128, N=Cout, K=16) for
//             group-of-tiles x 25 taps x C/16 k-steps, commits to mbarriers
//   warps 2-9 epilogue (two warps per TMEM lane quarter, alternating tiles): tcgen05.ld -> +b_eff ->
//             LeakyReLU -> BN affine (per-channel constants come from the kernel-parameter constant bank,
//             so they are instruction operands, not loads) -> bf16/fp16 -> 2x2 max-pool by exchanging
//             register halves with the x- and y-neighbour lanes (a warp owns 4 rows x 8 cols of the
//             tile) -> one 16-byte store per lane per 32 channels in the next block's P8 layout (or NHWC)
#pragma once
#include "tc_common.cuh"

namespace ernet {
namespace tc {

// POOL_: 2x2 max-pool in the epilogue (else every pixel of the used region is stored).
// TAPS_: 25 = full ACFF stencil union, 1 = a plain 1x1 convolution (conv_red2, squeeze_ernet_redconv.py:16,33).
// ACT_:  LeakyReLU + BN affine in the epilogue (else bias only).
// NREAL_: output channels that exist (N may be padded up to a multiple of 32 with zero weights).
template <int NC_, int N_, int HIN_, int HU_, int IMGS_, int G_, int NBUF_, bool WRES_, int WSTAGES_,
          bool POOL_ = true, int TAPS_ = 25, bool ACT_ = true, int NREAL_ = N_>
struct BlockCfg {
  static constexpr int NC = NC_, N = N_, HIN = HIN_, HU = HU_, IMGS = IMGS_, G = G_, NBUF = NBUF_, WSTAGES = WSTAGES_;
  static constexpr bool WRES = WRES_;
  static constexpr bool POOL = POOL_, ACT = ACT_;
  static constexpr int TAPS = TAPS_, NREAL = NREAL_;
  static constexpr int P = HIN + 3;                       // padded pitch: cols -2 .. HIN
  static constexpr int CHUNK_BYTES = P * P * 16;
  static constexpr int IMG_BYTES = NC * CHUNK_BYTES;
  static constexpr int IN_BYTES = IMGS * IMG_BYTES;
  static constexpr int TR = (HU + 15) / 16, TCOLS = (HU + 7) / 8;
  static constexpr int TILES_PER_IMG = TR * TCOLS;
  static constexpr int T = IMGS * TILES_PER_IMG;
  static constexpr int NG = (T + G - 1) / G;
  static constexpr int KSTEPS = NC / 2;
  static constexpr int TAP_BYTES = NC * N * 16;
  static constexpr int W_BYTES = TAPS * TAP_BYTES;
  static constexpr int W_SMEM = WRES ? W_BYTES : WSTAGES * TAP_BYTES;
  static constexpr int OUT_H = POOL ? HU / 2 : HU, OP = OUT_H + 3;
  static constexpr int OFF_W = IN_BYTES;
  static constexpr int OFF_BAR = (OFF_W + W_SMEM + 15) / 16 * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  static_assert(NC % 2 == 0, "K step is 16 channels");
  static_assert(N % 32 == 0 && N <= 256, "N must be a multiple of 32");
  static_assert(G * N * NBUF <= 512, "TMEM has 512 columns");
  static_assert(HU % 2 == 0, "pooling needs an even used size");
  static_assert(TAPS == 25 || TAPS == 1, "tap set");
  static_assert(NREAL % 8 == 0 && NREAL <= N, "real output channels");
  static_assert(IN_BYTES % 16 == 0 && TAP_BYTES % 16 == 0, "bulk copies and un-swizzled descriptors need 16-byte alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// Per-output-channel epilogue constants, passed by value so they sit in the kernel-parameter constant bank.
template <int N>
struct EpiParams {
  float bias[N];    // b_eff
  float scale[N];   // BN gamma / sqrt(var + eps)
  float shift[N];   // BN beta - mean * scale
  float deq[N];     // int8 only: s_w[n] (int32 accumulator -> real value; the per-input-channel activation
                    // scales are folded into the quantised weights)
  float out_inv[N]; // int8 output only: 1 / (real value of one int8 step of output channel n)
};

// operand kind and output format of a block kernel instance
enum : int { KIND_F16 = 0, KIND_BF16 = 1, KIND_I8 = 2 };
enum : int { OUT_P8 = 0,      // 16-bit, (B, N/8, OP, OP, 8) with zero halo: next block's staged image
             OUT_NHWC = 1,    // 16-bit, (B, OUT_H, OUT_H, N): input of the CUDA-core ACFF4 kernels
             OUT_P16 = 2 };   // int8,   (B, N/16, OP, OP, 16) with zero halo: next int8 block's staged image

constexpr int kBlockThreads = 320;   // producer + MMA + 8 epilogue warps

// Epilogue of one 16x8 output tile for one warp (32 lanes = 4 rows x 8 cols of the tile): TMEM -> bias
// [-> LeakyReLU -> BN affine] -> 16-bit / int8 -> [2x2 max-pool via register-half exchange] -> store.
// `tbase` = TMEM address of the tile's first column for this warp's lane quarter; (y, x) = this lane's output
// pixel; `img` = global image index.
// PREFETCH: keep the next 32 columns' TMEM load in flight under the arithmetic of the current ones (64 data registers);
// off in the fused transform + block-1 kernel, whose 27 warps leave 72 registers per thread.
template <class Cfg, int KIND, int OUT, bool PREFETCH = true>
__device__ __forceinline__ void epilogue_tile(const EpiParams<Cfg::N>& par, uint32_t tbase, int y, int x, bool valid,
                                        bool xodd, bool yodd, int qsel, uint16_t* __restrict__ out, int img) {
  constexpr int N = Cfg::N, NREAL = Cfg::NREAL, OP = Cfg::OP, OUT_H = Cfg::OUT_H;
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? NREAL / 16 : NREAL / 8;
  const int py = y >> 1, px = x >> 1;
  const int img0 = img, im = 0;              // (names used by the body below)
  uint32_t v[PREFETCH ? 2 : 1][32];
  if (PREFETCH) tmem_ld32(tbase, v[0]);
#pragma unroll
  for (int cb = 0; cb < N / 32; ++cb) {
    if (!PREFETCH) tmem_ld32(tbase + cb * 32, v[0]);
    tmem_ld_wait();                                     // block cb has landed
    if (PREFETCH && cb + 1 < N / 32) tmem_ld32(tbase + (cb + 1) * 32, v[(cb + 1) & 1]);   // prefetch the next 32 columns
    float yv[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = cb * 32 + j;
      float z = KIND == KIND_I8 ? fmaf(__int2float_rn((int)v[PREFETCH ? (cb & 1) : 0][j]), par.deq[n], par.bias[n])
                                : __uint_as_float(v[PREFETCH ? (cb & 1) : 0][j]) + par.bias[n];
      if (Cfg::ACT) {
        z = fmaxf(z, 0.01f * z);                         // LeakyReLU(0.01), acff.py:33
        z = fmaf(z, par.scale[n], par.shift[n]);         // eval BatchNorm, acff.py:34
      }
      yv[j] = z;
    }
    // 2x2 max-pool (squeeze_ernet.py:13).  x-neighbour = lane^1, y-neighbour = lane^8.  Each step the
    // lane keeps one half of its channels, ships the other half to the neighbour, and maxes what it
    // receives, so the lane ends with exactly the 8 channels it stores.  Rounding / quantisation is
    // monotone, so it commutes with the max and is done first (fewer registers to exchange).
    if (OUT == OUT_P16) {
      uint32_t pk[8];                                    // 32 channels as packed int8
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t w = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int q = __float2int_rn(yv[4 * j + e] * par.out_inv[cb * 32 + 4 * j + e]);
          q = max(-127, min(127, q));
          w |= ((uint32_t)q & 0xffu) << (8 * e);
        }
        pk[j] = w;
      }
      uint32_t m1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t keep = xodd ? pk[j + 4] : pk[j];
        const uint32_t send = xodd ? pk[j] : pk[j + 4];
        m1[j] = __vmaxs4(keep, __shfl_xor_sync(0xffffffffu, send, 1));
      }
      uint32_t m2[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t keep = yodd ? m1[j + 2] : m1[j];
        const uint32_t send = yodd ? m1[j] : m1[j + 2];
        m2[j] = __vmaxs4(keep, __shfl_xor_sync(0xffffffffu, send, 8));
      }
      if (valid && (cb * 2 + (qsel >> 1)) * 16 < NREAL) {   // 8 channels = half of a 16-channel chunk
        const int ch = cb * 2 + (qsel >> 1);
        uint2* oimg = reinterpret_cast<uint2*>(out) + ((size_t)(img0 + im) * OUT_CHUNKS * OP * OP) * 2;
        oimg[((size_t)(ch * OP + py + 2) * OP + px + 2) * 2 + (qsel & 1)] = make_uint2(m2[0], m2[1]);
      }
    } else {
      uint32_t pk[16];                                   // 32 channels as packed bf16x2 / half2
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(yv[2 * j], yv[2 * j + 1]); pk[j] = *reinterpret_cast<uint32_t*>(&h); }
        else      { __half2 h = __floats2half2_rn(yv[2 * j], yv[2 * j + 1]);            pk[j] = *reinterpret_cast<uint32_t*>(&h); }
      }
      if (!Cfg::POOL) {                                  // no pooling: the lane stores its own pixel
        if (valid) {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int ch = cb * 4 + qq;
            if (ch * 8 < NREAL) {
              const uint4 o4 = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
              if (OUT == OUT_P8) {
                uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)(img0 + im) * OUT_CHUNKS * OP * OP;
                oimg[(ch * OP + y + 2) * OP + x + 2] = o4;
              } else {
                uint16_t* o = out + ((size_t)((img0 + im) * OUT_H + y) * OUT_H + x) * NREAL + ch * 8;
                *reinterpret_cast<uint4*>(o) = o4;
              }
            }
          }
        }
        continue;
      }
      uint32_t m1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t keep = xodd ? pk[j + 8] : pk[j];
        const uint32_t send = xodd ? pk[j] : pk[j + 8];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        if (BF16) { __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&keep), *reinterpret_cast<const __nv_bfloat162*>(&recv)); m1[j] = *reinterpret_cast<uint32_t*>(&r); }
        else      { __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&keep), *reinterpret_cast<const __half2*>(&recv)); m1[j] = *reinterpret_cast<uint32_t*>(&r); }
      }
      uint32_t m2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t keep = yodd ? m1[j + 4] : m1[j];
        const uint32_t send = yodd ? m1[j] : m1[j + 4];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 8);
        if (BF16) { __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&keep), *reinterpret_cast<const __nv_bfloat162*>(&recv)); m2[j] = *reinterpret_cast<uint32_t*>(&r); }
        else      { __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&keep), *reinterpret_cast<const __half2*>(&recv)); m2[j] = *reinterpret_cast<uint32_t*>(&r); }
      }
      if (valid && (cb * 4 + qsel) * 8 < NREAL) {
        const uint4 o4 = make_uint4(m2[0], m2[1], m2[2], m2[3]);
        const int ch = cb * 4 + qsel;
        if (OUT == OUT_P8) {
          uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)(img0 + im) * OUT_CHUNKS * OP * OP;
          oimg[(ch * OP + py + 2) * OP + px + 2] = o4;
        } else {
          uint16_t* o = out + ((size_t)((img0 + im) * OUT_H + py) * OUT_H + px) * NREAL + ch * 8;
          *reinterpret_cast<uint4*>(o) = o4;
        }
      }
    }
  }
      }

// Epilogue of the tail unit of tc_pblock.cuh (PCfg::TAIL): TMEM lane = x, the accumulators of output rows HU-2 and HU-1
// are tiles 0 and 1 (tbase, tbase + N).  Same arithmetic per element as epilogue_tile (bias, LeakyReLU, BN, rounding to
// the output type, then the max - rounding is monotone), so the results are bit-identical to the 16 x 8 tiling; the y half
// of the 2x2 pool is a max of the two tiles in this thread, the x half one exchange with lane^1, after which the lane owns
// 16 of the 32 channels of the block of columns: two 8-channel chunks (P8) or one 16-channel chunk (P16).
template <class Cfg, int KIND, int OUT>
__device__ __forceinline__ void epilogue_tail(const EpiParams<Cfg::N>& par, uint32_t tbase, int x, bool valid, bool xodd,
                                              uint16_t* __restrict__ out, int img, int cb0, int cb_step) {
  constexpr int N = Cfg::N, NREAL = Cfg::NREAL, OP = Cfg::OP;
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? NREAL / 16 : NREAL / 8;
  static_assert(OUT == OUT_P8 || OUT == OUT_P16, "the tail unit writes padded P8 / P16 images");
  const int py = (Cfg::HU - 2) >> 1, px = x >> 1;
#pragma unroll 1
  for (int cb = cb0; cb < N / 32; cb += cb_step) {          // blocks of 32 columns are dealt to the EPW warps of a lane quarter
    uint32_t v[2][32];
    tmem_ld32(tbase + cb * 32, v[0]);
    tmem_ld32(tbase + N + cb * 32, v[1]);
    tmem_ld_wait();
    uint32_t pk[2][16];                                  // per row: 32 channels as 16-bit pairs (P8) / int8 quads (P16: 8 used)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float yv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = cb * 32 + j;
        float z = KIND == KIND_I8 ? fmaf(__int2float_rn((int)v[r][j]), par.deq[n], par.bias[n]) : __uint_as_float(v[r][j]) + par.bias[n];
        if (Cfg::ACT) {
          z = fmaxf(z, 0.01f * z);
          z = fmaf(z, par.scale[n], par.shift[n]);
        }
        yv[j] = z;
      }
      if (OUT == OUT_P16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t w = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int q = __float2int_rn(yv[4 * j + e] * par.out_inv[cb * 32 + 4 * j + e]);
            q = max(-127, min(127, q));
            w |= ((uint32_t)q & 0xffu) << (8 * e);
          }
          pk[r][j] = w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(yv[2 * j], yv[2 * j + 1]); pk[r][j] = *reinterpret_cast<uint32_t*>(&h); }
          else      { __half2 h = __floats2half2_rn(yv[2 * j], yv[2 * j + 1]);            pk[r][j] = *reinterpret_cast<uint32_t*>(&h); }
        }
      }
    }
    if (OUT == OUT_P16) {
      uint32_t m[8], m1[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = __vmaxs4(pk[0][j], pk[1][j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t keep = xodd ? m[j + 4] : m[j];
        const uint32_t send = xodd ? m[j] : m[j + 4];
        m1[j] = __vmaxs4(keep, __shfl_xor_sync(0xffffffffu, send, 1));
      }
      const int ch = cb * 2 + (xodd ? 1 : 0);
      if (valid && ch * 16 < NREAL) {
        uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)img * OUT_CHUNKS * OP * OP;
        oimg[(ch * OP + py + 2) * OP + px + 2] = make_uint4(m1[0], m1[1], m1[2], m1[3]);
      }
    } else {
      uint32_t m[16], m1[8];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (BF16) { __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&pk[0][j]), *reinterpret_cast<const __nv_bfloat162*>(&pk[1][j])); m[j] = *reinterpret_cast<uint32_t*>(&r); }
        else      { __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&pk[0][j]), *reinterpret_cast<const __half2*>(&pk[1][j])); m[j] = *reinterpret_cast<uint32_t*>(&r); }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t keep = xodd ? m[j + 8] : m[j];
        const uint32_t send = xodd ? m[j] : m[j + 8];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        if (BF16) { __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&keep), *reinterpret_cast<const __nv_bfloat162*>(&recv)); m1[j] = *reinterpret_cast<uint32_t*>(&r); }
        else      { __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&keep), *reinterpret_cast<const __half2*>(&recv)); m1[j] = *reinterpret_cast<uint32_t*>(&r); }
      }
      if (valid) {
        uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)img * OUT_CHUNKS * OP * OP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = cb * 4 + (xodd ? 2 : 0) + h;
          if (ch * 8 < NREAL) oimg[(ch * OP + py + 2) * OP + px + 2] = make_uint4(m1[4 * h], m1[4 * h + 1], m1[4 * h + 2], m1[4 * h + 3]);
        }
      }
    }
  }
}

// KIND_I8 with a 16-bit output writes fp16 (the int8 engine keeps its non-quantised tensors in fp16).
template <class Cfg, int KIND, int OUT>
__global__ void __launch_bounds__(kBlockThreads, 1)
acff_block_kernel(const uint16_t* __restrict__ in, const uint16_t* __restrict__ wimg,
                  const __grid_constant__ EpiParams<Cfg::N> par, uint16_t* __restrict__ out, int batch) {
  constexpr int N = Cfg::N, P = Cfg::P, G = Cfg::G, NBUF = Cfg::NBUF, T = Cfg::T, NG = Cfg::NG;
  constexpr int OP = Cfg::OP, OUT_H = Cfg::OUT_H;
  constexpr bool BF16 = KIND == KIND_BF16;                       // 16-bit output element type
  constexpr uint32_t IDESC = KIND == KIND_I8 ? instr_desc(2u, 1u, 128u, (uint32_t)N)      // s8 x s8 -> s32
                                             : instr_desc(1u, BF16 ? 1u : 0u, 128u, (uint32_t)N);
  constexpr int NREAL = Cfg::NREAL;
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? NREAL / 16 : NREAL / 8;     // 16-byte chunks per output pixel
  static_assert(Cfg::POOL || OUT != OUT_P16, "un-pooled int8 output is not needed");

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* s_in = smem;
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_in = bars;            // [1]
  uint64_t* w_full = bars + 1;        // [4]
  uint64_t* w_empty = bars + 5;       // [4]
  uint64_t* acc_full = bars + 9;      // [2]
  uint64_t* acc_empty = bars + 11;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  volatile uint32_t* abort_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img0 = blockIdx.x * Cfg::IMGS;
  const int nimg = min(Cfg::IMGS, batch - img0);

  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    mbar_init(bar_in, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      // weights are constants: their copies are issued before waiting for the previous kernel (PDL)
      mbar_expect_tx(bar_in, (uint32_t)(nimg * Cfg::IMG_BYTES + (Cfg::WRES ? Cfg::W_BYTES : 0)));
      if (Cfg::WRES) bulk_g2s(s_w, wimg, Cfg::W_BYTES, bar_in);
      if (!Cfg::WRES) {
        for (int it = 0; it < Cfg::WSTAGES; ++it) {
          mbar_expect_tx(&w_full[it], Cfg::TAP_BYTES);
          bulk_g2s(s_w + it * Cfg::TAP_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)it * Cfg::TAP_BYTES, Cfg::TAP_BYTES, &w_full[it]);
        }
      }
      pdl_wait();
      bulk_g2s(s_in, reinterpret_cast<const uint8_t*>(in) + (size_t)img0 * Cfg::IMG_BYTES, (uint32_t)(nimg * Cfg::IMG_BYTES), bar_in);
      if (!Cfg::WRES) {
        for (int it = Cfg::WSTAGES; it < NG * Cfg::TAPS; ++it) {
          const int s = it % Cfg::WSTAGES, use = it / Cfg::WSTAGES;
          if (use > 0 && !mbar_wait(&w_empty[s], (use - 1) & 1, abort_flag, 0x100u, it)) break;
          mbar_expect_tx(&w_full[s], Cfg::TAP_BYTES);
          bulk_g2s(s_w + s * Cfg::TAP_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)(it % Cfg::TAPS) * Cfg::TAP_BYTES,
                   Cfg::TAP_BYTES, &w_full[s]);
        }
      }
    }
    __syncwarp();
    pdl_wait();
    if (OUT != OUT_NHWC) {
      // zero halo of the output images (rows 0,1,OP-1 and cols 0,1,OP-1 of every chunk): the next block's
      // conv padding.  Done by the otherwise idle producer warp.
      constexpr int BORDER = 3 * OP + (OP - 3) * 3;
      for (int im = 0; im < nimg; ++im) {
        uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)(img0 + im) * OUT_CHUNKS * OP * OP;
        for (int i = lane; i < OUT_CHUNKS * BORDER; i += 32) {
          const int ch = i / BORDER, k = i - ch * BORDER;
          int r, c;
          if (k < 3 * OP) { r = k / OP; c = k - r * OP; if (r == 2) r = OP - 1; }
          else { const int k2 = k - 3 * OP; r = 2 + k2 / 3; c = k2 % 3; if (c == 2) c = OP - 1; }
          oimg[(ch * OP + r) * OP + c] = make_uint4(0, 0, 0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // ONE elected lane runs the whole issue loop.  Per MMA it spends an add on each descriptor's low word:
    // tile bases are computed once per group, the tap shift is a compile-time constant (block 1: loop fully
    // unrolled) or one multiply-add per tap, k-steps are immediates.
    if (elect_one()) {
      bool ok = mbar_wait(bar_in, 0, abort_flag, 0x200u);
      tc_fence_after();
      const uint32_t in_addr = smem_u32(s_in), w_addr = smem_u32(s_w);
      constexpr uint32_t A_HI = desc_hi(P * 16), B_HI = desc_hi(128);
      constexpr uint32_t A_KSTEP = (2 * Cfg::CHUNK_BYTES) >> 4, B_KSTEP = (2 * N * 16) >> 4;
      constexpr int TAP_UNROLL = Cfg::KSTEPS == 1 ? Cfg::TAPS : 1;
      const uint32_t w_lo0 = desc_lo(w_addr, N * 16);
      int ws = 0;                 // weight ring stage and its phase (streamed weights)
      uint32_t wphase = 0;
      for (int g = 0; g < NG && ok; ++g) {
        const int buf = g % NBUF, use = g / NBUF;
        if (use > 0) { ok = mbar_wait(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x201u, g); tc_fence_after(); }
        if (!ok) break;
        const int ntile = min(G, T - g * G);
        uint32_t a_lo[G];         // descriptor low word of each tile of the group at tap shift 0, k-step 0
#pragma unroll
        for (int tl = 0; tl < G; ++tl) {
          const int t = min(g * G + tl, T - 1);
          const int im = t / Cfg::TILES_PER_IMG, rem = t - im * Cfg::TILES_PER_IMG;
          const int ty = rem / Cfg::TCOLS, tx = rem - ty * Cfg::TCOLS;
          a_lo[tl] = desc_lo(in_addr + im * Cfg::IMG_BYTES + (uint32_t)(((ty * 16 + 2) * P + tx * 8 + 2) * 16), Cfg::CHUNK_BYTES);
        }
        const uint32_t d0 = tmem_base + (uint32_t)(buf * G * N);
#pragma unroll TAP_UNROLL
        for (int tap = 0; tap < Cfg::TAPS; ++tap) {
          uint32_t b_lo;
          if (Cfg::WRES) {
            b_lo = w_lo0 + (uint32_t)(tap * (Cfg::TAP_BYTES >> 4));
          } else {
            ok = mbar_wait(&w_full[ws], wphase, abort_flag, 0x202u, g * 32 + tap);
            if (!ok) break;
            tc_fence_after();
            b_lo = w_lo0 + (uint32_t)(ws * (Cfg::TAP_BYTES >> 4));
          }
          const uint32_t toff = Cfg::TAPS == 1 ? 0u : (uint32_t)(tap_dy(tap) * P + tap_dx(tap));   // tap shift in 16-byte units
#pragma unroll
          for (int tl = 0; tl < G; ++tl) {
            if (tl < ntile) {
#pragma unroll
              for (int ks = 0; ks < Cfg::KSTEPS; ++ks)
                if (KIND == KIND_I8)
                  mma_i8(d0 + tl * N, desc_make(a_lo[tl] + toff + ks * A_KSTEP, A_HI),
                         desc_make(b_lo + ks * B_KSTEP, B_HI), IDESC, (tap | ks) != 0 ? 1u : 0u);
                else
                  mma_f16(d0 + tl * N, desc_make(a_lo[tl] + toff + ks * A_KSTEP, A_HI),
                          desc_make(b_lo + ks * B_KSTEP, B_HI), IDESC, (tap | ks) != 0 ? 1u : 0u);
            }
          }
          if (!Cfg::WRES) {
            mma_commit(&w_empty[ws]);
            if (++ws == Cfg::WSTAGES) { ws = 0; wphase ^= 1; }
          }
        }
        if (ok) mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q4 = warp & 3;                                 // TMEM lane quarter this warp may read
    const int ehalf = (warp - 2) >> 2;                       // this warp takes tiles with (tl & 1) == ehalf
    const int rr = 4 * q4 + (lane >> 3), cc = lane & 7;
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int qsel = (xodd ? 2 : 0) + (yodd ? 1 : 0);        // which 8-channel chunk of a 32-channel block this lane keeps
    for (int g = 0; g < NG; ++g) {
      const int buf = g % NBUF, use = g / NBUF;
      if (!mbar_wait(&acc_full[buf], use & 1, abort_flag, 0x300u + warp, g)) break;
      tc_fence_after();
      const int ntile = min(G, T - g * G);
      for (int tl = ehalf; tl < ntile; tl += 2) {
        const int t = g * G + tl;
        const int im = t / Cfg::TILES_PER_IMG, rem = t - im * Cfg::TILES_PER_IMG;
        const int ty = rem / Cfg::TCOLS, tx = rem - ty * Cfg::TCOLS;
        const int y = ty * 16 + rr, x = tx * 8 + cc;
        const bool valid = (y < Cfg::HU) && (x < Cfg::HU) && (im < nimg);
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * G * N + tl * N);
        epilogue_tile<Cfg, KIND, OUT>(par, tbase, y, x, valid, xodd, yodd, qsel, out, img0 + im);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Block configurations (SURVEY.md section 7.5 sizes, Squeeze_ErNET).  NC counts 16-byte chunks per pixel:
// 8 channels at 16 bit, 16 channels at int8 (block 1 int8: 16 real channels + one zero chunk, K step = 32).
//   block 1: 16 -> 64,  69x69 in, 66x66 used, 45 tiles/img, weights resident, 2 x 4-tile TMEM buffers
//   block 2: 64 -> 96,  33x33 in, 30x30 used,  8 tiles/img, weights streamed through a ring
//   block 3: 96 -> 128, 15x15 in, 12x12 used,  2 tiles/img, 2 images per CTA, weights streamed
using CfgBlock1 = BlockCfg<2, 64, 69, 66, 1, 4, 2, true, 1>;
using CfgBlock2 = BlockCfg<8, 96, 33, 30, 1, 4, 1, false, 3>;
using CfgBlock3 = BlockCfg<12, 128, 15, 12, 2, 4, 1, false, 2>;
// Squeeze_RedConv (model/squeeze_ernet_redconv.py): block 1 sees 8 real + 8 zero channels; block 2 is split into an
// un-pooled ACFF2 instance and a 1-tap conv_red2 + pool instance (96 -> 48, N padded to 64); block 3 has 48 inputs.
using CfgBlock2R = BlockCfg<8, 96, 33, 30, 1, 4, 1, false, 3, /*POOL*/ false>;
using CfgRed2R = BlockCfg<12, 64, 30, 30, 1, 4, 2, true, 1, /*POOL*/ true, /*TAPS*/ 1, /*ACT*/ false, /*NREAL*/ 48>;
using CfgBlock3R = BlockCfg<6, 128, 15, 12, 2, 4, 1, false, 3>;
using CfgBlock1Q = BlockCfg<2, 64, 69, 66, 1, 4, 2, true, 1>;
using CfgBlock2Q = BlockCfg<4, 96, 33, 30, 1, 4, 1, false, 4>;
using CfgBlock3Q = BlockCfg<6, 128, 15, 12, 2, 4, 1, false, 3>;
using CfgBlock3RQ = BlockCfg<4, 128, 15, 12, 2, 4, 1, false, 3>;      // int8 Squeeze_RedConv: 48 real + 16 zero channels (weight-image size)

template <class Cfg, int KIND, int OUT>
inline int launch_acff_block(const void* in, const void* wimg, const EpiParams<Cfg::N>& par, void* out, int batch,
                             cudaStream_t stream) {
  const int grid = (batch + Cfg::IMGS - 1) / Cfg::IMGS;
  ERNET_CUDA(launch_pdl(acff_block_kernel<Cfg, KIND, OUT>, dim3(grid), dim3(kBlockThreads), Cfg::SMEM_BYTES, stream,
                        static_cast<const uint16_t*>(in), static_cast<const uint16_t*>(wimg), par, static_cast<uint16_t*>(out), batch));
  return ERNET_OK;
}

template <class Cfg, int KIND, int OUT>
inline int set_block_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(acff_block_kernel<Cfg, KIND, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  return ERNET_OK;
}

inline int set_all_block_attrs() {
  int rc;
  if ((rc = set_block_attr<CfgBlock1, KIND_BF16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgBlock1, KIND_F16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgBlock2, KIND_BF16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgBlock2, KIND_F16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgBlock3, KIND_BF16, OUT_NHWC>())) return rc;
  if ((rc = set_block_attr<CfgBlock3, KIND_F16, OUT_NHWC>())) return rc;
  if ((rc = set_block_attr<CfgBlock2R, KIND_BF16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgBlock2R, KIND_F16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgRed2R, KIND_BF16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgRed2R, KIND_F16, OUT_P8>())) return rc;
  if ((rc = set_block_attr<CfgBlock3R, KIND_BF16, OUT_NHWC>())) return rc;
  if ((rc = set_block_attr<CfgBlock3R, KIND_F16, OUT_NHWC>())) return rc;
  if ((rc = set_block_attr<CfgBlock1Q, KIND_I8, OUT_P16>())) return rc;
  if ((rc = set_block_attr<CfgBlock2Q, KIND_I8, OUT_P16>())) return rc;
  if ((rc = set_block_attr<CfgBlock3Q, KIND_I8, OUT_NHWC>())) return rc;
  return ERNET_OK;
}

// ---------------------------------------------------------------------------------- stem -> P8
// conv1 3x3/s2 (model/squeeze_ernet.py:11,25; RedConv: conv_red1 folded in) writing the P8 layout block 1
// stages: (B, 2 chunks, 72, 72, 8) with the zero halo included.  One thread per padded pixel.
// KIND_I8: channel c is quantised with out_inv.v[c] (= 1 / its int8 step) into chunk 0 of a P16 image, chunk 1 is zero.
struct StemInv { float v[16]; };
template <typename TI, int CS, int KIND, int HS = 69>
__global__ void __launch_bounds__(128)
stem_p8_kernel(const TI* __restrict__ x, long long sb, long long sc, long long sy, long long sx,
               const float* __restrict__ w /*[27][CS]*/, const float* __restrict__ bias, uint16_t* __restrict__ out, int total,
               const __grid_constant__ StemInv out_inv) {
  __shared__ float ws[27 * CS + CS];
  for (int i = threadIdx.x; i < 27 * CS + CS; i += blockDim.x) ws[i] = i < 27 * CS ? w[i] : bias[i - 27 * CS];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  constexpr int P = HS + 3;          // HS = conv1 output size: 69 (140x140 inputs), 119 (ErNET, 240x240)
  const int b = idx / (P * P);
  const int r = idx - b * P * P;
  const int pr = r / P, pc = r - pr * P;
  const int oy = pr - 2, ox = pc - 2;
  uint4* o = reinterpret_cast<uint4*>(out) + (size_t)b * 2 * P * P + pr * P + pc;
  if (oy < 0 || oy >= HS || ox < 0 || ox >= HS) {
    o[0] = make_uint4(0, 0, 0, 0);
    o[P * P] = make_uint4(0, 0, 0, 0);
    return;
  }
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = c < CS ? ws[27 * CS + c] : 0.f;
  const TI* p = x + b * sb + (2 * oy) * sy + (2 * ox) * sx;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = to_f32<TI>(p[ky * sy + kx * sx + c * sc]);
        const float* wr = ws + ((ky * 3 + kx) * 3 + c) * CS;
#pragma unroll
        for (int k = 0; k < CS; ++k) acc[k] = fmaf(v, wr[k], acc[k]);
      }
  if (KIND == KIND_I8) {
    uint32_t wq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t wv = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int q = __float2int_rn(acc[4 * j + e] * out_inv.v[4 * j + e]);
        q = max(-127, min(127, q));
        wv |= ((uint32_t)q & 0xffu) << (8 * e);
      }
      wq[j] = wv;
    }
    o[0] = make_uint4(wq[0], wq[1], wq[2], wq[3]);
    o[P * P] = make_uint4(0, 0, 0, 0);
    return;
  }
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    float t8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t8[k] = acc[ch * 8 + k];
    o[ch * P * P] = KIND == KIND_BF16 ? pack16<__nv_bfloat16>(t8) : pack16<__half>(t8);
  }
}

// debug: P16 (B, NC, H+3, W+3, 16) int8 -> dequantised fp32 NCHW (first C channels)
__global__ void tap_p16_to_nchw_f32(const int8_t* __restrict__ src, int NC, int C, int H, long long total,
                                    const float* __restrict__ scale /*[C]*/, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int P = H + 3;
  const int x = (int)(i % H);
  const int y = (int)((i / H) % H);
  const int c = (int)((i / ((long long)H * H)) % C);
  const long long b = i / ((long long)H * H * C);
  dst[i] = scale[c] * (float)src[(((b * NC + (c >> 4)) * P + y + 2) * P + x + 2) * 16 + (c & 15)];
}

// debug: P8 (B, NC, H+3, W+3, 8) 16-bit -> fp32 NCHW (B, NC*8 [first C], H, W)
template <bool BF16>
__global__ void tap_p8_to_nchw_f32(const uint16_t* __restrict__ src, int NC, int C, int H, long long total, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int P = H + 3;
  const int x = (int)(i % H);
  const int y = (int)((i / H) % H);
  const int c = (int)((i / ((long long)H * H)) % C);
  const long long b = i / ((long long)H * H * C);
  const uint16_t raw = src[(((b * NC + (c >> 3)) * P + y + 2) * P + x + 2) * 8 + (c & 7)];
  dst[i] = BF16 ? __uint_as_float((uint32_t)raw << 16) : __half2float(*reinterpret_cast<const __half*>(&raw));
}

}  // namespace tc
}  // namespace ernet
