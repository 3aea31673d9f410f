// ACFF4 + classifier head in one kernel (sm_100a).
//
//   a4     = BN(LeakyReLU(W_f . cat(dw_1(x), dw_2(x), dw_3(x)) + b_f))           model/acff.py:37-59
//   logits = W_eff . sum_{4x4}(a4) + b_fc ; probs = softmax(logits)              model/squeeze_ernet.py:33-41
//
// The map is only 6x6 -> 4x4 here, so unlike blocks 1-3 the 25-tap fold would waste the tensor cores on
// padding.  Instead ("design D"): a CTA takes IMGS images x 16 output pixels as rows of one M=128 MMA tile;
//   1. one bulk copy stages the input images (NHWC, contiguous in HBM),
//   2. all 8 warps compute the three dilated depthwise convs on CUDA cores (fp32 accumulate) and write
//      the results, rounded to 16 bit, directly as the A operand in the un-swizzled K-major UMMA layout
//      [k-chunk of 8][128 rows][8] - the concat of acff.py:46 is just the K order [branch][channel],
//   3. one lane issues K/16 tcgen05.mma (N=256) against weight slices streamed through a bulk-copy ring
//      that was started before step 2,
//   4. four epilogue warps read the accumulators (lane = pixel), apply bias/LeakyReLU/BN with
//      constant-bank operands, reduce the collapsed classifier per pixel, sum the 16 pixels of an image
//      with shuffles, softmax, write 5 probabilities (+ logits).
#pragma once
#include "tc_common.cuh"

namespace ernet {
namespace tc {

struct TailParams {          // kernel-parameter constant bank
  float bias[256];           // fused_conv.bias
  float scale[256];          // BN gamma / sqrt(var + eps)
  float shift[256];          // BN beta - mean * scale
  float weff[5][256];        // conv2 -> avgpool -> fc collapsed (pack.py)
  float bfc[5];
};

template <int C4_>
struct TailCfg {
  static constexpr int C4 = C4_, K = 3 * C4, N = 256;
  // images per CTA: 4 fill only rows 0..63 of the M=128 tile (the MMA is a negligible part of this kernel), which
  // doubles the number of CTAs sharing the CUDA-core depthwise stage
  static constexpr int IMGS = 4;
  static constexpr int KCHUNKS = K / 8, KSTEPS = K / 16;
  static constexpr int IN_BYTES = IMGS * 36 * C4 * 2;              // 8 images (6,6,C4) 16-bit
  static constexpr int A_BYTES = KCHUNKS * 128 * 16;               // [k-chunk][128 rows][16 B]
  static constexpr int DW_FLOATS = 30 * C4;                        // [3][9][C4] weights + [3][C4] bias
  static constexpr int KS_PER_STAGE = 2;
  static constexpr int STAGE_BYTES = KS_PER_STAGE * 2 * N * 16;    // 16 KB
  // weight ring: WPRE dedicated stages are filled while the depthwise stage runs; once that stage is done the
  // staged input is dead and its region provides WREUSE more stages (deep enough to cover L2 latency)
  static constexpr int WPRE = 2;
  static constexpr int WREUSE = (IMGS * 36 * C4_ * 2) / STAGE_BYTES;
  static constexpr int WSTAGES = WPRE + WREUSE;
  static constexpr int NSTAGE_LOADS = KSTEPS / KS_PER_STAGE;
  static constexpr int OFF_A = IN_BYTES;
  static constexpr int OFF_DW = OFF_A + A_BYTES;
  static constexpr int OFF_W = OFF_DW + DW_FLOATS * 4;
  static constexpr int OFF_BAR = OFF_W + WPRE * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static_assert(KSTEPS % KS_PER_STAGE == 0, "stage granularity");
  static_assert(WSTAGES <= 8, "barrier array size");
  static_assert(OFF_W % 128 == 0 && OFF_A % 128 == 0, "alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

constexpr int kTailThreads = 256;

template <class Cfg, bool BF16, bool WRITE_A4>
__global__ void __launch_bounds__(kTailThreads, 1)
acff4_head_kernel(const uint16_t* __restrict__ in /*(B,6,6,C4)*/, const float* __restrict__ dw_w /*[3][9][C4]*/,
                  const float* __restrict__ dw_b /*[3][C4]*/, const uint16_t* __restrict__ wimg /*[K/8][256][8]*/,
                  const __grid_constant__ TailParams par, float* __restrict__ probs, float* __restrict__ logits,
                  uint16_t* __restrict__ a4_out /*(B,4,4,256), debug*/, int batch) {
  constexpr int C4 = Cfg::C4, N = Cfg::N, CV = C4 / 8;
  constexpr uint32_t IDESC = instr_desc(1u, BF16 ? 1u : 0u, 128u, 256u);
  using T = typename std::conditional<BF16, __nv_bfloat16, __half>::type;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint4* s_in = reinterpret_cast<const uint4*>(smem);                       // [img][36][CV]
  uint4* s_a = reinterpret_cast<uint4*>(smem + Cfg::OFF_A);                       // [k-chunk][128]
  float* s_dw = reinterpret_cast<float*>(smem + Cfg::OFF_DW);
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_in = bars;          // [1]
  uint64_t* w_full = bars + 1;      // [8]
  uint64_t* w_empty = bars + 9;     // [8]
  uint64_t* acc_full = bars + 17;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  volatile uint32_t* abort_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img0 = blockIdx.x * Cfg::IMGS;
  const int nimg = min(Cfg::IMGS, batch - img0);

  ERNET_CHAIN_ENTRY(4);
  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    mbar_init(bar_in, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  for (int i = threadIdx.x; i < Cfg::DW_FLOATS / 4; i += kTailThreads)       // 27*C4 and 3*C4 are multiples of 4
    reinterpret_cast<float4*>(s_dw)[i] = i < 27 * C4 / 4 ? __ldg(reinterpret_cast<const float4*>(dw_w) + i)
                                                          : __ldg(reinterpret_cast<const float4*>(dw_b) + (i - 27 * C4 / 4));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  // ---- kick off the first weight stages (constants: before the PDL wait), then the input copy; both land while
  //      the depthwise stage runs
  if (threadIdx.x == 0) {
    for (int st = 0; st < Cfg::WPRE; ++st) {
      mbar_expect_tx(&w_full[st], Cfg::STAGE_BYTES);
      bulk_g2s(s_w + st * Cfg::STAGE_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)st * Cfg::STAGE_BYTES,
               Cfg::STAGE_BYTES, &w_full[st]);
    }
    pdl_wait();
    ERNET_CHAIN_WAITED(4);
    mbar_expect_tx(bar_in, (uint32_t)(nimg * 36 * C4 * 2));
    bulk_g2s(smem, reinterpret_cast<const uint8_t*>(in) + (size_t)img0 * 36 * C4 * 2, (uint32_t)(nimg * 36 * C4 * 2), bar_in);
  }

  // ---- depthwise trio -> A operand (all warps)
  bool ok = mbar_wait(bar_in, 0, abort_flag, 0x400u);
  if (ok) {
    for (int item = threadIdx.x; item < Cfg::IMGS * 16 * CV; item += kTailThreads) {
      const int v = item % CV, row = item / CV;
      const int im = row >> 4, oy = (row >> 2) & 3, ox = row & 3;
      float acc[3][8];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[d][k] = s_dw[27 * C4 + d * C4 + v * 8 + k];
      if (im < nimg) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int dil = d + 1;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int iy = oy + ky * dil - (dil - 1), ix = ox + kx * dil - (dil - 1);
              if (iy >= 0 && iy < 6 && ix >= 0 && ix < 6) {
                float xv[8];
                unpack16<T>(s_in[(im * 36 + iy * 6 + ix) * CV + v], xv);
                const float* wr = s_dw + (d * 9 + ky * 3 + kx) * C4 + v * 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[d][k] = fmaf(xv[k], wr[k], acc[d][k]);
              }
            }
        }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) s_a[(d * CV + v) * 128 + row] = pack16<T>(acc[d]);
    }
  }
  fence_proxy_async();          // generic-proxy writes of A must be visible to the tensor core (async proxy)
  __syncthreads();

  if (warp == 0) {
    // ---- weight ring producer: remaining loads; stages >= WPRE live in the (now dead) input region
    if (lane == 0) {
      for (int ld = Cfg::WPRE; ld < Cfg::NSTAGE_LOADS; ++ld) {
        const int st = ld % Cfg::WSTAGES, use = ld / Cfg::WSTAGES;
        if (use > 0 && !mbar_wait(&w_empty[st], (use - 1) & 1, abort_flag, 0x401u, ld)) break;
        uint8_t* dst = st < Cfg::WPRE ? s_w + st * Cfg::STAGE_BYTES : smem + (st - Cfg::WPRE) * Cfg::STAGE_BYTES;
        mbar_expect_tx(&w_full[st], Cfg::STAGE_BYTES);
        bulk_g2s(dst, reinterpret_cast<const uint8_t*>(wimg) + (size_t)ld * Cfg::STAGE_BYTES, Cfg::STAGE_BYTES, &w_full[st]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    if (elect_one()) {
      const uint32_t a_lo0 = desc_lo(smem_u32(s_a), 128 * 16);
      constexpr uint32_t AB_HI = desc_hi(128);
      bool okm = ok;
      for (int ld = 0; ld < Cfg::NSTAGE_LOADS && okm; ++ld) {
        const int st = ld % Cfg::WSTAGES, use = ld / Cfg::WSTAGES;
        okm = mbar_wait(&w_full[st], use & 1, abort_flag, 0x402u, ld);
        if (!okm) break;
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < Cfg::KS_PER_STAGE; ++j) {
          const int ks = ld * Cfg::KS_PER_STAGE + j;
          const uint8_t* wst = st < Cfg::WPRE ? s_w + st * Cfg::STAGE_BYTES : smem + (st - Cfg::WPRE) * Cfg::STAGE_BYTES;
          mma_f16(tmem_base, desc_make(a_lo0 + (uint32_t)(ks * ((2 * 128 * 16) >> 4)), AB_HI),
                  desc_make(desc_lo(smem_u32(wst) + (uint32_t)(j * 2 * N * 16), N * 16), AB_HI), IDESC, ks != 0 ? 1u : 0u);
        }
        mma_commit(&w_empty[st]);
      }
      if (okm) mma_commit(acc_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---- epilogue: lane = pixel (row of the tile), 16 consecutive lanes = one image
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane, im = row >> 4;
    if (q4 * 32 < Cfg::IMGS * 16 && mbar_wait(acc_full, 0, abort_flag, 0x403u, warp)) {
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16);
      float dot[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      uint32_t v[2][32];
      tmem_ld32(tbase, v[0]);
#pragma unroll
      for (int cb = 0; cb < N / 32; ++cb) {
        tmem_ld_wait();
        if (cb + 1 < N / 32) tmem_ld32(tbase + (cb + 1) * 32, v[(cb + 1) & 1]);
        float yv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = cb * 32 + j;
          float z = __uint_as_float(v[cb & 1][j]) + par.bias[n];
          z = fmaxf(z, 0.01f * z);
          const float y = fmaf(z, par.scale[n], par.shift[n]);
          yv[j] = y;
#pragma unroll
          for (int c = 0; c < 5; ++c) dot[c] = fmaf(y, par.weff[c][n], dot[c]);
        }
        if (WRITE_A4 && im < nimg) {
          uint4* o = reinterpret_cast<uint4*>(a4_out + ((size_t)(img0 + im) * 16 + (row & 15)) * 256 + cb * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float t8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t8[k] = yv[q * 8 + k];
            o[q] = pack16<T>(t8);
          }
        }
      }
      // sum over the 16 pixels of the image (squeeze_ernet.py:34-40 collapsed), then softmax (:41)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dot[c] += __shfl_xor_sync(0xffffffffu, dot[c], o);
      }
      if ((lane & 15) == 0 && im < nimg) {
        float z[5], m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 5; ++c) { z[c] = dot[c] + par.bfc[c]; m = fmaxf(m, z[c]); }
        float e[5], sum = 0.f;
#pragma unroll
        for (int c = 0; c < 5; ++c) { e[c] = expf(z[c] - m); sum += e[c]; }
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          probs[(size_t)(img0 + im) * 5 + c] = e[c] / sum;
          if (logits) logits[(size_t)(img0 + im) * 5 + c] = z[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  ERNET_CHAIN_EXIT(4);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

using TailCfg128 = TailCfg<128>;     // Squeeze_ErNET:  acff4 = ACFF(128, 256)
using TailCfg64 = TailCfg<64>;       // Squeeze_RedConv: acff4 = ACFF(64, 256)

template <class Cfg>
inline int launch_acff4_head(bool bf16, const void* in, const float* dw_w, const float* dw_b, const void* wimg,
                             const TailParams& par, float* probs, float* logits, void* a4_out, int batch, cudaStream_t stream) {
  const int grid = (batch + Cfg::IMGS - 1) / Cfg::IMGS;
  auto* i16 = static_cast<const uint16_t*>(in);
  auto* w16 = static_cast<const uint16_t*>(wimg);
  auto* a16 = static_cast<uint16_t*>(a4_out);
#define ERNET_TAIL(BF, WA) launch_pdl(acff4_head_kernel<Cfg, BF, WA>, dim3(grid), dim3(kTailThreads), Cfg::SMEM_BYTES, stream, i16, dw_w, dw_b, w16, par, probs, logits, a16, batch)
  cudaError_t ce;
  if (bf16) { ce = a4_out ? ERNET_TAIL(true, true) : ERNET_TAIL(true, false); }
  else      { ce = a4_out ? ERNET_TAIL(false, true) : ERNET_TAIL(false, false); }
#undef ERNET_TAIL
  if (ce != cudaSuccess) return fail(ERNET_ERR_CUDA, "launch of acff4_head_kernel failed: %s", cudaGetErrorString(ce));
  return ERNET_OK;
}

template <class Cfg>
inline int set_tail_attrs() {
  ERNET_CUDA(cudaFuncSetAttribute(acff4_head_kernel<Cfg, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  ERNET_CUDA(cudaFuncSetAttribute(acff4_head_kernel<Cfg, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  ERNET_CUDA(cudaFuncSetAttribute(acff4_head_kernel<Cfg, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  ERNET_CUDA(cudaFuncSetAttribute(acff4_head_kernel<Cfg, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  return ERNET_OK;
}

}  // namespace tc
}  // namespace ernet
