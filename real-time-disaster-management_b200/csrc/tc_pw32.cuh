// fp32 engine: the 1x1 convolutions of ACFF blocks 1-3 (model/acff.py:31-34 + the 2x2 pool of squeeze_ernet.py:13) on the
// tcgen05 tensor cores in split-TF32 form.  An fp32 value v is written as hi + lo with hi = v with the low 13 mantissa
// bits cleared (exactly a TF32 number) and lo = v - hi (exact in fp32; the MMA reads its top 10 mantissa bits), and
//     a . b  ~=  a_lo . b_hi  +  a_hi . b_lo  +  a_hi . b_hi          (fp32 accumulation in TMEM)
// drops only the lo x lo term (2^-22 relative) - the three-MMA scheme known as 3xTF32.  The result agrees with the FFMA
// kernel (simt_layers.cuh: pointwise_kernel) to a few fp32 ulps of the accumulated sum, and the engine's logits stay
// within the same 1e-4 of the PyTorch fp32 reference (tests/test_gpu_parity.py).
//
// Schedule (persistent, one CTA per SM, units dealt round-robin).  A unit is a 16 x 16 pixel patch of one image = two
// M = 128 tiles (left / right 8 columns) that share every weight stage:
//   warp 0      TMA: per K stage of 16 channels one box (16 ch, 16 px, 16 rows) of the NHWC concat tensor -> raw ring
//               (rows outside the image are zero-filled by the TMA unit), plus that stage's weight block when the
//               weights are streamed;
//   warps 2-9   convert a raw stage into the two UMMA operand images (hi, lo) in the K-major core-matrix layout;
//   warp 1      MMA issuer: 2 tiles x 2 K steps x 3 MMAs (kind::tf32, M128 x N x K8) per stage, two TMEM accumulator
//               buffers so that the epilogue of unit k runs under the MMAs of unit k+1;
//   warps 10-17 epilogue, POOL FIRST: TMEM -> 2x2 max of the raw accumulators by lane shuffles -> bias, LeakyReLU, BN
//               affine on the pooled quarter -> fp32 NHWC.  max commutes with every monotone map, and bias + LeakyReLU +
//               BN is non-decreasing in the accumulator when the BN scale is >= 0; for channels with a negative scale
//               the packer negates the weight column (the accumulator becomes -acc, exactly) and the epilogue multiplies
//               the pooled value by -1 again: max(-acc) = -min(acc), which is what a decreasing map needs.
#pragma once
#include <cuda.h>

#include "tc_pblock.cuh"

namespace ernet {
namespace tc {

// FLAT (block 4: 4x4 maps, no pool): the concat tensor (batch, 16 pixels, K) is read as ONE 16-pixel-wide image whose rows
// are the images of the batch, so a unit is 16 images x 16 pixels; the N_ * NSPLIT_ output channels are dealt to NSPLIT_
// units per row group (more CTAs busy on a tiny GEMM, TMEM holds two accumulator buffers).
template <int K_, int N_, int H_, bool WRES_, int RAW_STAGES_, int OP_STAGES_, int ACC_BUFS_, bool FLAT_ = false, int NSPLIT_ = 1>
struct TfCfg {
  static constexpr int K = K_, N = N_, H = H_, RAW_STAGES = RAW_STAGES_, OP_STAGES = OP_STAGES_, ACC_BUFS = ACC_BUFS_;
  static constexpr bool WRES = WRES_, FLAT = FLAT_;
  static constexpr int NSPLIT = NSPLIT_, NTOT = N_ * NSPLIT_;
  static constexpr int KS = 16, CH = KS / 4, NKS = K / KS;           // channels / 16-byte chunks per stage, stages per unit
  static constexpr int RAW_BYTES = 256 * KS * 4;                      // 16 x 16 pixels x 16 channels fp32 (64-byte swizzled rows)
  static constexpr int B_HALF = CH * N * 16, B_STAGE = 2 * B_HALF;    // weights of one stage: hi then lo, [chunk][n][4 k]
  static constexpr int W_SMEM = WRES ? NKS * B_STAGE : OP_STAGES * B_STAGE;
  static constexpr int UX = (H + 15) / 16, UNITS_PER_IMG = UX * UX;
  static constexpr int W_SPLIT_BYTES = NKS * B_STAGE;                  // weight image of one N split
  static constexpr int OH = H / 2;
  // TMEM columns: accumulators [buf][tile][N], then the A operand ring [stage][tile][hi 16 | lo 16]
  static constexpr int A_BASE = ACC_BUFS * 2 * N, A_STAGE_COLS = 64;
  static constexpr int OFF_W = RAW_STAGES * RAW_BYTES;
  static constexpr int OFF_BAR = OFF_W + W_SMEM;
  static constexpr int OFF_PAR = OFF_BAR + 256;                        // bias | bn scale | bn shift | accumulator sign, NTOT floats each
  static constexpr int SMEM_BYTES = OFF_PAR + 4 * NTOT * 4;
  static constexpr int NCONV = 8, NEPI = 8;                            // converter / epilogue warps: (tile, TMEM lane quarter) each
  static constexpr int THREADS = (2 + NCONV + NEPI) * 32;
  static_assert(K % KS == 0 && N % 32 == 0 && N <= 128 && H % 2 == 0, "operand shape");
  static_assert(FLAT ? (H == 16 && !WRES) : NSPLIT == 1, "flat maps: 16 pixels per image, streamed weights; N splits only there");

  struct Unit { int img, uy, ux, nh; };
  static __device__ __forceinline__ int total_units(int batch) { return FLAT ? ((batch + 15) / 16) * NSPLIT : batch * UNITS_PER_IMG; }
  static __device__ __forceinline__ Unit unit(int u) {
    Unit r;
    if (FLAT) { r.nh = u % NSPLIT; r.uy = u / NSPLIT; r.ux = 0; r.img = 0; }
    else { r.img = u / UNITS_PER_IMG; const int q = u - r.img * UNITS_PER_IMG; r.uy = q / UX; r.ux = q - r.uy * UX; r.nh = 0; }
    return r;
  }
  static_assert(RAW_STAGES <= 6 && OP_STAGES <= 4 && ACC_BUFS >= 1 && ACC_BUFS <= 2, "barrier arrays");
  static_assert(A_BASE + OP_STAGES * A_STAGE_COLS <= 512, "TMEM columns");
  static_assert(OFF_W % 128 == 0 && OFF_BAR % 16 == 0, "alignment");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct Pw32Params {
  const float* bias;
  const float* bn_s;
  const float* bn_t;
};

// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 registers per thread -> 16 consecutive TMEM columns of the thread's lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// w [K][ldn] fp32, columns n0 .. n0 + N -> per stage of 16 k: [hi | lo][chunk c < 4][n][k = 16 s + 4 c + j, j < 4]
// neg_if (pooled blocks: the BN scale) != nullptr: columns whose entry is negative are stored negated (pool-first epilogue)
__global__ void pw32_pack_weights(const float* __restrict__ w, int K, int N, float* __restrict__ out, int ldn = 0, int n0 = 0,
                                  const float* __restrict__ neg_if = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i - k * N;
  float v = w[(size_t)k * (ldn ? ldn : N) + n0 + n];
  if (neg_if && neg_if[n0 + n] < 0.f) v = -v;
  const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  const int s = k >> 4, c = (k >> 2) & 3, j = k & 3;
  const size_t base = (size_t)s * (8 * N * 4) + ((size_t)c * N + n) * 4 + j;
  out[base] = hi;
  out[base + (size_t)4 * N * 4] = v - hi;
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
pw32_kernel(const __grid_constant__ CUtensorMap tmap_in, const float* __restrict__ wimg, const Pw32Params par,
            float* __restrict__ out, int batch) {
  constexpr int N = Cfg::N, H = Cfg::H, NKS = Cfg::NKS;
  constexpr uint32_t IDESC = instr_desc(1u, 2u, 128u, (uint32_t)N);   // D = f32, A/B = tf32, K-major
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem + Cfg::OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* raw_full = bars;           // [6]
  uint64_t* raw_empty = bars + 6;      // [6]
  uint64_t* op_full = bars + 12;       // [4]
  uint64_t* op_empty = bars + 16;      // [4]
  uint64_t* w_full = bars + 20;        // [4]  ([0] alone when the weights are resident)
  uint64_t* acc_full = bars + 24;      // [2]
  uint64_t* acc_empty = bars + 26;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  volatile uint32_t* abort_flag = tmem_slot + 1;
  float* s_par = reinterpret_cast<float*>(smem + Cfg::OFF_PAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = Cfg::total_units(batch);
  constexpr int NTOT = Cfg::NTOT;

  for (int i = threadIdx.x; i < NTOT; i += Cfg::THREADS) {   // constants: not ordered after the previous kernel
    s_par[i] = par.bias[i];
    s_par[NTOT + i] = par.bn_s[i];
    s_par[2 * NTOT + i] = par.bn_t[i];
    s_par[3 * NTOT + i] = (!Cfg::FLAT && par.bn_s[i] < 0.f) ? -1.f : 1.f;      // same rule as pw32_pack_weights
  }
  if (threadIdx.x == 0) {
    *abort_flag = 0u;
    for (int i = 0; i < 6; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], Cfg::NCONV); }
    for (int i = 0; i < 4; ++i) { mbar_init(&op_full[i], Cfg::NCONV); mbar_init(&op_empty[i], 1); mbar_init(&w_full[i], 1); }
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: raw ring (+ streamed weights)
    if (lane == 0) {
      if (Cfg::WRES) {
        mbar_expect_tx(&w_full[0], NKS * Cfg::B_STAGE);
        for (int s = 0; s < NKS; ++s)
          bulk_g2s(s_w + s * Cfg::B_STAGE, reinterpret_cast<const uint8_t*>(wimg) + (size_t)s * Cfg::B_STAGE, Cfg::B_STAGE, &w_full[0]);
      }
      pdl_wait();
      int it = 0;
      bool ok = true;
      for (int u = blockIdx.x; u < total_units && ok; u += gridDim.x) {
        const typename Cfg::Unit un = Cfg::unit(u);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(wimg) + (size_t)un.nh * Cfg::W_SPLIT_BYTES;
        for (int ks = 0; ks < NKS; ++ks, ++it) {
          const int rs = it % Cfg::RAW_STAGES, ruse = it / Cfg::RAW_STAGES;
          if (ruse > 0 && !(ok = mbar_wait(&raw_empty[rs], (ruse - 1) & 1, abort_flag, 0x700u, it))) break;
          mbar_expect_tx(&raw_full[rs], Cfg::RAW_BYTES);
          tma_load_4d(smem + rs * Cfg::RAW_BYTES, &tmap_in, ks * Cfg::KS, un.ux * 16, un.uy * 16, un.img, &raw_full[rs]);
          if (!Cfg::WRES) {
            // the weight block of this stage shares the ring index (and the release) of the A operand stage in TMEM
            const int os = it % Cfg::OP_STAGES, ouse = it / Cfg::OP_STAGES;
            if (ouse > 0 && !(ok = mbar_wait(&op_empty[os], (ouse - 1) & 1, abort_flag, 0x701u, it))) break;
            mbar_expect_tx(&w_full[os], Cfg::B_STAGE);
            bulk_g2s(s_w + os * Cfg::B_STAGE, wsrc + (size_t)ks * Cfg::B_STAGE, Cfg::B_STAGE, &w_full[os]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      bool ok = true;
      if (Cfg::WRES) ok = mbar_wait(&w_full[0], 0, abort_flag, 0x702u);
      const uint32_t w_addr = smem_u32(s_w);
      constexpr uint32_t HI = desc_hi(128);
      int it = 0, k = 0;
      for (int u = blockIdx.x; u < total_units && ok; u += gridDim.x, ++k) {
        const int ux = Cfg::unit(u).ux;
        const int ntile = (ux * 16 + 8 < H) ? 2 : 1;
        const int buf = k % Cfg::ACC_BUFS, use = k / Cfg::ACC_BUFS;
        if (use > 0 && !(ok = mbar_wait(&acc_empty[buf], (use - 1) & 1, abort_flag, 0x703u, k))) break;
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 2 * N);
        for (int ks = 0; ks < NKS; ++ks, ++it) {
          const int os = it % Cfg::OP_STAGES, ophase = (it / Cfg::OP_STAGES) & 1;
          if (!(ok = mbar_wait(&op_full[os], ophase, abort_flag, 0x704u, it))) break;
          if (!Cfg::WRES && !(ok = mbar_wait(&w_full[os], ophase, abort_flag, 0x705u, it))) break;
          tc_fence_after();
          const uint32_t a0 = tmem_base + (uint32_t)(Cfg::A_BASE + os * Cfg::A_STAGE_COLS);
          const uint32_t b0 = desc_lo(w_addr + (Cfg::WRES ? ks : os) * Cfg::B_STAGE, N * 16);
#pragma unroll
          for (int tl = 0; tl < 2; ++tl) {
            if (tl < ntile) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint32_t a_hi = a0 + (uint32_t)(tl * 32 + kk * 8), a_lo = a_hi + 16u;
                const uint32_t b_hi = b0 + (uint32_t)((kk * 2 * N * 16) >> 4);
                const uint32_t b_lo = b_hi + (uint32_t)(Cfg::B_HALF >> 4);
                const uint32_t d = d0 + (uint32_t)(tl * N);
                mma_tf32_ts(d, a_lo, desc_make(b_hi, HI), IDESC, (ks | kk) != 0 ? 1u : 0u);
                mma_tf32_ts(d, a_hi, desc_make(b_lo, HI), IDESC, 1u);
                mma_tf32_ts(d, a_hi, desc_make(b_hi, HI), IDESC, 1u);
              }
            }
          }
          mma_commit(&op_empty[os]);
        }
        if (ok) mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp < 2 + Cfg::NCONV) {
    // ------------------------------------------------------------------ converters: raw fp32 -> (hi, lo) A operand in TMEM.
    // Warp = (tile t, lane quarter q = warp % 4): thread = row 32 q + lane of the tile = pixel (4 q + lane / 8, 8 t + lane % 8)
    // of the unit; it reads its pixel's 16 channels (four 16-byte chunks of a 64-byte row, TMA 64 B swizzle: chunk ^
    // ((row >> 1) & 3), conflict-free) and stores them as 16 + 16 TMEM columns of its lane.
    const int q4 = warp & 3, t = (warp - 2) >> 2;
    const int px = (4 * q4 + (lane >> 3)) * 16 + 8 * t + (lane & 7);
    const int sw = (px >> 1) & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(Cfg::A_BASE + t * 32);
    int it = 0;
    bool ok = true;
    for (int u = blockIdx.x; u < total_units && ok; u += gridDim.x) {
      const int ux = Cfg::unit(u).ux;
      const bool live = t == 0 || (ux * 16 + 8 < H);            // the right tile of the last unit column is outside the image
      for (int ks = 0; ks < NKS; ++ks, ++it) {
        const int rs = it % Cfg::RAW_STAGES, os = it % Cfg::OP_STAGES, ouse = it / Cfg::OP_STAGES;
        if (!(ok = mbar_wait(&raw_full[rs], (it / Cfg::RAW_STAGES) & 1, abort_flag, 0x710u + warp, it))) break;
        if (ouse > 0 && !(ok = mbar_wait(&op_empty[os], (ouse - 1) & 1, abort_flag, 0x720u + warp, it))) break;
        if (live) {
          const uint4* row = reinterpret_cast<const uint4*>(smem + rs * Cfg::RAW_BYTES + px * 64);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 v = row[j ^ sw];
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hi[4 * j + e] = w4[e] & 0xFFFFE000u;
              lo[4 * j + e] = __float_as_uint(__uint_as_float(w4[e]) - __uint_as_float(hi[4 * j + e]));
            }
          }
          tc_fence_after();
          tmem_st16(lane_addr + (uint32_t)(os * Cfg::A_STAGE_COLS), hi);
          tmem_st16(lane_addr + (uint32_t)(os * Cfg::A_STAGE_COLS + 16), lo);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&op_full[os]); mbar_arrive(&raw_empty[rs]); }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp (quarter q4, tile e)
    const int q4 = warp & 3, e = (warp - 2 - Cfg::NCONV) >> 2;
    const int ry = 4 * q4 + (lane >> 3), rx = 8 * e + (lane & 7);
    const bool xodd = (lane & 1) != 0, yodd = ((lane >> 3) & 1) != 0;
    const int chsel = (xodd ? 16 : 0) + (yodd ? 8 : 0);
    int k = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++k) {
      const typename Cfg::Unit un = Cfg::unit(u);
      const int img = un.img, uy = un.uy, ux = un.ux;
      const int ntile = (ux * 16 + 8 < H) ? 2 : 1;
      const float* pp = s_par + un.nh * N;                       // this unit's slice of bias | bn scale | bn shift
      const int buf = k % Cfg::ACC_BUFS, use = k / Cfg::ACC_BUFS;
      if (!mbar_wait_suspend(&acc_full[buf], use & 1, abort_flag, 0x730u + warp, k)) break;
      tc_fence_after();
      if (e < ntile) {
        const int y = uy * 16 + ry, x = ux * 16 + rx;
        const bool valid = Cfg::FLAT ? (y < batch) : ((y < H) && (x < H));
        float* o = Cfg::FLAT ? out + ((size_t)y * 16 + x) * NTOT + un.nh * N        // un-pooled: (batch, 16 pixels, NTOT)
                             : out + (((size_t)img * Cfg::OH + (y >> 1)) * Cfg::OH + (x >> 1)) * N + chsel;
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * 2 * N + e * N);
#pragma unroll 1
        for (int cb = 0; cb < N / 32; ++cb) {
          uint32_t v[32];
          tmem_ld32(tbase + cb * 32, v);
          tmem_ld_wait();
          if (Cfg::FLAT) {                 // un-pooled block: every value goes through bias, LeakyReLU, BN
            if (valid) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 pb = *reinterpret_cast<const float4*>(pp + cb * 32 + j4 * 4);
                const float4 ps = *reinterpret_cast<const float4*>(pp + NTOT + cb * 32 + j4 * 4);
                const float4 pt = *reinterpret_cast<const float4*>(pp + 2 * NTOT + cb * 32 + j4 * 4);
                const float bb[4] = {pb.x, pb.y, pb.z, pb.w}, ss[4] = {ps.x, ps.y, ps.z, ps.w}, tt[4] = {pt.x, pt.y, pt.z, pt.w};
                float f[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  float z = __uint_as_float(v[j4 * 4 + jj]) + bb[jj];
                  z = z > 0.f ? z : 0.01f * z;
                  f[jj] = fmaf(z, ss[jj], tt[jj]);
                }
                *reinterpret_cast<float4*>(o + cb * 32 + j4 * 4) = make_float4(f[0], f[1], f[2], f[3]);
              }
            }
            continue;
          }
          // 2x2 max-pool of the raw accumulators: lanes l^1 (x neighbour) and l^8 (y neighbour) hold the other three pixels
          // of the window; each exchange halves the channels a lane keeps, so the four lanes end with 8 pooled channels each
          float g[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float keep = __uint_as_float(xodd ? v[j + 16] : v[j]);
            const float send = __uint_as_float(xodd ? v[j] : v[j + 16]);
            g[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
          }
          float m[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float keep = yodd ? g[j + 8] : g[j];
            const float send = yodd ? g[j] : g[j + 8];
            m[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
          }
          // bias, LeakyReLU, BN on the lane's 8 pooled channels (cb * 32 + chsel + 0..7)
#pragma unroll
          for (int j4 = 0; j4 < 2; ++j4) {
            const float* q = pp + cb * 32 + chsel + j4 * 4;
            const float4 pb = *reinterpret_cast<const float4*>(q);
            const float4 ps = *reinterpret_cast<const float4*>(q + NTOT);
            const float4 pt = *reinterpret_cast<const float4*>(q + 2 * NTOT);
            const float4 pg = *reinterpret_cast<const float4*>(q + 3 * NTOT);
            const float bb[4] = {pb.x, pb.y, pb.z, pb.w}, ss[4] = {ps.x, ps.y, ps.z, ps.w}, tt[4] = {pt.x, pt.y, pt.z, pt.w},
                        gg[4] = {pg.x, pg.y, pg.z, pg.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float z = fmaf(m[j4 * 4 + jj], gg[jj], bb[jj]);       // (+-1) * pooled accumulator + bias: exact sign restore
              z = fmaxf(z, 0.01f * z);                              // LeakyReLU(0.01)
              m[j4 * 4 + jj] = fmaf(z, ss[jj], tt[jj]);
            }
          }
          if (valid) {
            *reinterpret_cast<float4*>(o + cb * 32) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4*>(o + cb * 32 + 4) = make_float4(m[4], m[5], m[6], m[7]);
          }
        }
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

// ---- host side -----------------------------------------------------------------------------------------------
// Tensor map of the concat tensor (batch, H, H, K) fp32 NHWC; box = (16 channels, 16 px, 16 rows, 1 image).
template <class Cfg>
inline int make_pw32_map(CUtensorMap* map, const void* base, int batch) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)Cfg::K, (cuuint64_t)Cfg::H, Cfg::FLAT ? (cuuint64_t)batch : (cuuint64_t)Cfg::H,
                              Cfg::FLAT ? (cuuint64_t)1 : (cuuint64_t)batch};
  const cuuint64_t strides[3] = {(cuuint64_t)Cfg::K * 4, (cuuint64_t)Cfg::H * Cfg::K * 4,
                                 (Cfg::FLAT ? (cuuint64_t)batch : (cuuint64_t)Cfg::H) * Cfg::H * Cfg::K * 4};
  const cuuint32_t box[4] = {(cuuint32_t)Cfg::KS, 16, 16, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ERNET_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 concat) failed with CUresult %d", (int)r);
  return ERNET_OK;
}

template <class Cfg>
inline int launch_pw32(const float* a, const float* wimg, const float* bias, const float* bn_s, const float* bn_t, float* out,
                       int batch, int num_sms, cudaStream_t stream) {
  CUtensorMap map;
  int rc = make_pw32_map<Cfg>(&map, a, batch);
  if (rc) return rc;
  const int total = Cfg::FLAT ? ((batch + 15) / 16) * Cfg::NSPLIT : batch * Cfg::UNITS_PER_IMG;
  const int grid = total < num_sms ? total : num_sms;
  Pw32Params par{bias, bn_s, bn_t};
  ERNET_CUDA(launch_pdl(pw32_kernel<Cfg>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, map, wimg, par, out, batch));
  return ERNET_OK;
}

template <class Cfg>
inline int set_pw32_attr() {
  ERNET_CUDA(cudaFuncSetAttribute(pw32_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  return ERNET_OK;
}

inline size_t pw32_weight_floats(int K, int N) { return (size_t)2 * K * N; }

// Squeeze_ErNET fp32: block 1 (48 -> 64 on 66x66, weights resident), block 2 (192 -> 96 on 30x30), block 3 (288 -> 128 on 12x12);
using FPw1 = TfCfg<48, 64, 66, true, 6, 3, 2>;
using FPw2 = TfCfg<192, 96, 30, true, 4, 2, 2>;
using FPw3 = TfCfg<288, 128, 12, false, 4, 4, 1>;
// block 4 (384 -> 256 on 4x4, LeakyReLU + BN, no pool): 16 images x 16 pixels per unit, four 64-channel splits
using FPw4 = TfCfg<384, 64, 16, false, 4, 4, 2, true, 4>;

}  // namespace tc
}  // namespace ernet
