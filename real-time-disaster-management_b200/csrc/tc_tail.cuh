// ACFF4 + classifier head in one kernel (sm_100a).
//
//   a4     = BN(LeakyReLU(W_f . cat(dw_1(x), dw_2(x), dw_3(x)) + b_f))           model/acff.py:37-59
//   logits = W_eff . sum_{4x4}(a4) + b_fc ; probs = softmax(logits)              model/squeeze_ernet.py:33-41
//
// The map is only 6x6 -> 4x4 here, so unlike blocks 1-3 the 25-tap fold would waste the tensor cores on
// padding.  Instead ("design D"): a CTA takes 2 images x 16 output pixels as rows 0..31 of one M=128 MMA tile;
//   1. the weight slices (constants) are requested before the PDL wait - up to 9 x 16 KB in flight while the previous
//      kernel drains - then one bulk copy stages the input images (NHWC, contiguous in HBM),
//   2. all 16 warps compute the three dilated depthwise convs on CUDA cores (fp32 accumulate, one (pixel, 8 channels)
//      item per thread) and write the results, rounded to 16 bit, directly as the A operand in the un-swizzled
//      K-major UMMA layout [k-chunk of 8][32 rows][8] - the concat of acff.py:46 is just the K order [branch][channel],
//   3. one lane issues K/16 tcgen05.mma (N=256),
//   4. four epilogue warps (TMEM lane quarter 0, 64 columns each) apply bias/LeakyReLU/BN with constant-bank
//      operands and reduce the collapsed classifier per pixel and over the 16 pixels of an image; the four partial
//      sums are added in a fixed order, softmax, 5 probabilities (+ logits) out.
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
  // Two images per CTA: 32 of the 128 rows of the M = 128 tile.  The kernel is a latency chain (input -> depthwise ->
  // 12-24 MMAs -> epilogue), so what counts is many CTAs (128 for a batch of 256) and short stages, not tile occupancy.
  static constexpr int IMGS = 2, ROWS = IMGS * 16;
  static constexpr int KCHUNKS = K / 8, KSTEPS = K / 16;
  static constexpr int IN_BYTES = IMGS * 36 * C4 * 2;              // images (6,6,C4) 16-bit
  // A operand [k-chunk][ROWS][16 B]: rows 32..127 of the tile read on into the following chunks (finite data, results
  // of those rows are never used), so only ROWS rows per chunk are stored; + pad for the last chunk's overrun
  static constexpr int A_CHUNK = ROWS * 16;
  static constexpr int A_BYTES = KCHUNKS * A_CHUNK + 128 * 16;
  static constexpr int DW_FLOATS = 30 * C4;                        // [3][9][C4] weights + [3][C4] bias
  static constexpr int KS_PER_STAGE = 2;
  static constexpr int STAGE_BYTES = KS_PER_STAGE * 2 * N * 16;    // 16 KB
  static constexpr int NSTAGE_LOADS = KSTEPS / KS_PER_STAGE;
  // weight ring: everything that fits is requested BEFORE the PDL wait (weights are constants), i.e. while the
  // previous kernel is still running
  static constexpr int WSTAGES = NSTAGE_LOADS < 9 ? NSTAGE_LOADS : 9;
  static constexpr int OFF_A = (IN_BYTES + 127) / 128 * 128;
  static constexpr int OFF_DW = (OFF_A + A_BYTES + 127) / 128 * 128;
  static constexpr int OFF_RED = OFF_DW + DW_FLOATS * 4;            // [4 column groups][IMGS][5] partial classifier dots
  static constexpr int OFF_PAR = (OFF_RED + 4 * IMGS * 5 * 4 + 127) / 128 * 128;   // [256][8]: bias, scale, shift, weff[0..4] per channel
  static constexpr int OFF_W = OFF_PAR + 256 * 8 * 4;
  static constexpr int OFF_BAR = OFF_W + WSTAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static_assert(KSTEPS % KS_PER_STAGE == 0, "stage granularity");
  static_assert(WSTAGES <= 12, "barrier array size");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

constexpr int kTailThreads = 512;     // 16 warps: one depthwise item per thread, 4 epilogue warps on lane quarter 0

template <class Cfg, bool BF16, bool WRITE_A4>
__global__ void __launch_bounds__(kTailThreads, 1)
acff4_head_kernel(const uint16_t* __restrict__ in /*(B,6,6,C4)*/, const float* __restrict__ dw_w /*[3][9][C4]*/,
                  const float* __restrict__ dw_b /*[3][C4]*/, const uint16_t* __restrict__ wimg /*[K/8][256][8]*/,
                  const __grid_constant__ TailParams par, float* __restrict__ probs, float* __restrict__ logits,
                  uint16_t* __restrict__ a4_out /*(B,4,4,256), debug*/, int batch) {
  constexpr int C4 = Cfg::C4, N = Cfg::N, CV = C4 / 8, ROWS = Cfg::ROWS;
  constexpr uint32_t IDESC = instr_desc(1u, BF16 ? 1u : 0u, 128u, 256u);
  using T = typename std::conditional<BF16, __nv_bfloat16, __half>::type;

  extern __shared__ __align__(128) uint8_t smem[];
  const uint4* s_in = reinterpret_cast<const uint4*>(smem);                       // [img][36][CV]
  uint4* s_a = reinterpret_cast<uint4*>(smem + Cfg::OFF_A);                       // [k-chunk][ROWS]
  float* s_dw = reinterpret_cast<float*>(smem + Cfg::OFF_DW);
  float* s_red = reinterpret_cast<float*>(smem + Cfg::OFF_RED);
  float4* s_par = reinterpret_cast<float4*>(smem + Cfg::OFF_PAR);
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* bar_in = bars;          // [1]
  uint64_t* w_full = bars + 1;      // [12]
  uint64_t* w_empty = bars + 13;    // [12]
  uint64_t* acc_full = bars + 25;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
  volatile uint32_t* abort_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img0 = blockIdx.x * Cfg::IMGS;
  const int nimg = min(Cfg::IMGS, batch - img0);

  ERNET_CHAIN_ENTRY(4);
  constexpr int tl_kernel = 2;     // study builds: stamps go to unit slot 20 of the block-3 timeline
  (void)tl_kernel;
  if (threadIdx.x == 0) ERNET_TL(20, 0);
  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    mbar_init(bar_in, 1);
    for (int i = 0; i < 12; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    // constants first (before the PDL wait): as many weight stages as the ring holds
    for (int st = 0; st < Cfg::WSTAGES; ++st) {
      mbar_expect_tx(&w_full[st], Cfg::STAGE_BYTES);
      bulk_g2s(s_w + st * Cfg::STAGE_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)st * Cfg::STAGE_BYTES,
               Cfg::STAGE_BYTES, &w_full[st]);
    }
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  // Per-channel epilogue constants: kernel parameters -> shared memory, transposed to 8 floats per channel.  Read
  // straight from the constant bank in the epilogue every constant is a cold miss (each is used once per CTA): that
  // was measured at 33 k cycles for 64 channels; here the misses of 16 warps overlap and precede the PDL wait.
  if (threadIdx.x < 256) {
    const int n = threadIdx.x;
    s_par[2 * n] = make_float4(par.bias[n], par.scale[n], par.shift[n], par.weff[0][n]);
    s_par[2 * n + 1] = make_float4(par.weff[1][n], par.weff[2][n], par.weff[3][n], par.weff[4][n]);
  }
  for (int i = threadIdx.x; i < Cfg::DW_FLOATS / 4; i += kTailThreads)       // 27*C4 and 3*C4 are multiples of 4
    reinterpret_cast<float4*>(s_dw)[i] = i < 27 * C4 / 4 ? __ldg(reinterpret_cast<const float4*>(dw_w) + i)
                                                          : __ldg(reinterpret_cast<const float4*>(dw_b) + (i - 27 * C4 / 4));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    pdl_wait();
    ERNET_CHAIN_WAITED(4);
    ERNET_TL(20, 1);
    mbar_expect_tx(bar_in, (uint32_t)(nimg * 36 * C4 * 2));
    bulk_g2s(smem, reinterpret_cast<const uint8_t*>(in) + (size_t)img0 * 36 * C4 * 2, (uint32_t)(nimg * 36 * C4 * 2), bar_in);
  }

  // ---- depthwise trio -> A operand: one (pixel, 8-channel chunk) item per thread
  // 16 warps wait here from the moment the CTA is resident (8-20 k cycles before block 3 ends) until the input has landed:
  // bit 1 of g_epi_suspend selects the suspending form (study switch ERNET_EPI_SUSPEND=3)
  const bool susp = (g_epi_suspend & 2u) != 0;
  bool ok = mbar_wait_epi(bar_in, 0, abort_flag, 0x400u, 0, susp);
  if (threadIdx.x == 0) ERNET_TL(20, 2);
  if (ok) {
    for (int item = threadIdx.x; item < ROWS * CV; item += kTailThreads) {
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
      for (int d = 0; d < 3; ++d) s_a[(d * CV + v) * ROWS + row] = pack16<T>(acc[d]);
    }
  }
  fence_proxy_async();          // generic-proxy writes of A must be visible to the tensor core (async proxy)
  __syncthreads();
  if (threadIdx.x == 0) ERNET_TL(20, 3);

  if (warp == 2) {
    // ---- weight ring producer: loads beyond the ring depth (none for the 64-channel tail)
    if (lane == 0) {
      for (int ld = Cfg::WSTAGES; ld < Cfg::NSTAGE_LOADS; ++ld) {
        const int st = ld % Cfg::WSTAGES, use = ld / Cfg::WSTAGES;
        if (use > 0 && !mbar_wait(&w_empty[st], (use - 1) & 1, abort_flag, 0x401u, ld)) break;
        mbar_expect_tx(&w_full[st], Cfg::STAGE_BYTES);
        bulk_g2s(s_w + st * Cfg::STAGE_BYTES, reinterpret_cast<const uint8_t*>(wimg) + (size_t)ld * Cfg::STAGE_BYTES, Cfg::STAGE_BYTES, &w_full[st]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    if (elect_one()) {
      const uint32_t a_lo0 = desc_lo(smem_u32(s_a), Cfg::A_CHUNK);
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
          mma_f16(tmem_base, desc_make(a_lo0 + (uint32_t)(ks * ((2 * Cfg::A_CHUNK) >> 4)), AB_HI),
                  desc_make(desc_lo(smem_u32(s_w + st * Cfg::STAGE_BYTES) + (uint32_t)(j * 2 * N * 16), N * 16), AB_HI), IDESC, ks != 0 ? 1u : 0u);
        }
        mma_commit(&w_empty[st]);
      }
      if (okm) mma_commit(acc_full);
    }
    __syncwarp();
  }
  if ((warp & 3) == 0) {
    // ---- epilogue: the 32 rows sit in TMEM lane quarter 0, readable by warps 0, 4, 8, 12; each takes 64 of the 256
    //      columns, lane = pixel (16 consecutive lanes = one image)
    const int cg = warp >> 2;
    const int row = lane, im = row >> 4;
    if (mbar_wait_epi(acc_full, 0, abort_flag, 0x403u, warp, susp)) {
      if (threadIdx.x == 0) ERNET_TL(20, 4);
      tc_fence_after();
      float dot[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      uint32_t v[2][32];
      tmem_ld32(tmem_base + (uint32_t)(cg * 64), v[0]);
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        tmem_ld_wait();
        if (cb == 0) tmem_ld32(tmem_base + (uint32_t)(cg * 64 + 32), v[1]);
        float yv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 p0 = s_par[2 * (cg * 64 + cb * 32 + j)], p1 = s_par[2 * (cg * 64 + cb * 32 + j) + 1];   // broadcast loads
          float z = __uint_as_float(v[cb][j]);
          const float bias = p0.x, scale = p0.y, shift = p0.z, w0 = p0.w, w1 = p1.x, w2 = p1.y, w3 = p1.z, w4 = p1.w;
          z += bias;
          z = fmaxf(z, 0.01f * z);
          const float y = fmaf(z, scale, shift);
          yv[j] = y;
          dot[0] = fmaf(y, w0, dot[0]); dot[1] = fmaf(y, w1, dot[1]); dot[2] = fmaf(y, w2, dot[2]);
          dot[3] = fmaf(y, w3, dot[3]); dot[4] = fmaf(y, w4, dot[4]);
        }
        if (WRITE_A4 && im < nimg) {
          uint4* o = reinterpret_cast<uint4*>(a4_out + ((size_t)(img0 + im) * 16 + (row & 15)) * 256 + cg * 64 + cb * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float t8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t8[k] = yv[q * 8 + k];
            o[q] = pack16<T>(t8);
          }
        }
      }
      // sum over the 16 pixels of the image (squeeze_ernet.py:34-40 collapsed)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dot[c] += __shfl_xor_sync(0xffffffffu, dot[c], o);
      }
      if ((lane & 15) == 0) {
#pragma unroll
        for (int c = 0; c < 5; ++c) s_red[(cg * Cfg::IMGS + im) * 5 + c] = dot[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ERNET_TL(20, 5);
  if (threadIdx.x < Cfg::IMGS && (int)threadIdx.x < nimg && !*abort_flag) {
    // column groups added in a fixed order, then softmax (squeeze_ernet.py:41)
    const int im = threadIdx.x;
    float z[5], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      z[c] = ((s_red[(0 * Cfg::IMGS + im) * 5 + c] + s_red[(1 * Cfg::IMGS + im) * 5 + c]) +
              (s_red[(2 * Cfg::IMGS + im) * 5 + c] + s_red[(3 * Cfg::IMGS + im) * 5 + c])) + par.bfc[c];
      m = fmaxf(m, z[c]);
    }
    float e[5], sum = 0.f;
#pragma unroll
    for (int c = 0; c < 5; ++c) { e[c] = expf(z[c] - m); sum += e[c]; }
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      probs[(size_t)(img0 + im) * 5 + c] = e[c] / sum;
      if (logits) logits[(size_t)(img0 + im) * 5 + c] = z[c];
    }
  }
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
