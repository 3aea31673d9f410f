// Pool-first epilogue of the pooled ACFF block kernels (model/acff.py:52-54 + squeeze_ernet.py:13,27): the 2x2 max-pool is
// taken on the raw fp32 / int32 accumulators and bias, LeakyReLU, BatchNorm, rounding run on the ONE surviving value of
// each 2x2 window - a quarter of the arithmetic of "activate every pixel, then pool".
//
// Exactness.  f(z) = s * lrelu(z + b) + t followed by the 16-bit (or int8) rounding is monotone in z: non-decreasing
// when s >= 0, non-increasing when s < 0, and every step (add, max, fma, convert) is itself monotone.  A monotone
// non-decreasing function commutes with max bit for bit, so max-then-f == f-then-max.  For channels with s < 0
// (negative BatchNorm gamma) the packer NEGATES the channel's folded weights, so the accumulator holds -z and the pool
// picks max(-z) = -min(z): the value f is largest at.  With u = max(-z) - b:  f = |s| * min(u', 0.01 u') + t where
// u' = -(z + b) - the same roundings as the reference order of operations (negation is exact).
//
// Data movement.  tcgen05.ld.16x256b hands each thread the accumulator rows r and r + 8 of a 16-lane half for two
// adjacent columns.  With the tile's pixels laid out as TMEM lane = 8 * (tile row) + (tile col), rows r and r + 8 are the
// vertical neighbours (y, x) / (y + 1, x): the y-pool is thread-local.  The x-neighbour (row r ^ 1) sits in lane ^ 4: the
// two threads swap one column each (one SHFL per two accumulator columns) and each keeps one pooled column per 8-column
// repetition.  The packer permutes the output channels (TMEM column 64 g + 8 i + m  <->  channel 64 g + 8 m + i) so that
// the 8 values a thread ends with are 8 CONSECUTIVE channels of one pooled pixel: one 16-byte store in the next block's
// P8 layout.  Per 1024 accumulators: 1 TMEM load, 24 max, 8 shuffles, 32-48 ALU for the activation of 256 survivors,
// 1 store (the activate-then-pool form: 128 + 16 converts + 24 for the pool of packed halves + 8 shuffles).
#pragma once
#include "tc_block.cuh"

namespace ernet {
namespace tc {

// Per-channel constants of the pool-first epilogue, staged in shared memory as two float4 arrays in TMEM-COLUMN order
// (entry j describes channel pf_channel_of_column(j)): the eight threads that differ in m then read eight consecutive
// float4 - no bank conflicts.
//   A[j] = { deq (int8: real value of one accumulator unit, else 1), sigma * b_eff, |s|, t }
//   B[j] = { out_inv (int8 output: 1 / step of the channel, else 1), sigma (+1 / -1: sign of the BatchNorm scale; the
//            channel's weights are stored multiplied by it), 0, 0 }
template <int N>
struct EpiPFTable { float4 A[N]; float4 B[N]; };

// TMEM column (= row of the weight image) -> output channel, and back
__host__ __device__ constexpr int pf_channel_of_column(int j) { return (j & ~63) + 8 * (j & 7) + ((j >> 3) & 7); }

__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// One 16x8 output tile, one warp = one TMEM lane quarter (4 tile rows x 8 columns of pixels -> 2 x 4 pooled pixels).
//   tbase  TMEM address of the tile's first column at this warp's lane quarter
//   yq     first tile row of the quarter (global output row, multiple of 4); xt = first column of the tile (multiple of 8)
//   s_tab  shared-memory copy of the channel table; any_neg: some channel has sigma < 0
template <class Cfg, int KIND, int OUT>
__device__ __forceinline__ void epilogue_tile_pf(const float4* __restrict__ s_tab /*A[N] then B[N]*/, bool any_neg, uint32_t tbase, int yq, int xt,
                                                 bool dup, uint16_t* __restrict__ out, int img, int lane) {
  constexpr int N = Cfg::N, NREAL = Cfg::NREAL, OP = Cfg::OP, OUT_H = Cfg::OUT_H;
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr int OUT_CHUNKS = OUT == OUT_P16 ? NREAL / 16 : NREAL / 8;
  static_assert(Cfg::POOL && N % 64 == 0 && Cfg::HU % 2 == 0, "pool-first epilogue: pooled blocks, 64-column groups");
  const int q = lane & 3, r = lane >> 2, e = r & 1, m = 2 * q + e;
  const int px = (xt + r) >> 1;
  const bool xok = xt + (r & ~1) < Cfg::HU;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int yy = yq + 2 * half, py = yy >> 1;
    const bool valid = xok && yy < Cfg::HU && !dup;
#pragma unroll
    for (int g = 0; g < N / 64; ++g) {
      uint32_t v[32];
      tmem_ld_16x256b_x8(tbase + ((uint32_t)(16 * half) << 16) + (uint32_t)(64 * g), v);
      tmem_ld_wait();
      float pooled[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // v[4i], v[4i+1]: row r, columns 8i + 2q, 8i + 2q + 1;  v[4i+2], v[4i+3]: row r + 8 (the pixel below)
        if (KIND == KIND_I8) {
          const int c0 = max((int)v[4 * i], (int)v[4 * i + 2]), c1 = max((int)v[4 * i + 1], (int)v[4 * i + 3]);
          const int keep = e ? c1 : c0, send = e ? c0 : c1;
          pooled[i] = __int2float_rn(max(keep, __shfl_xor_sync(0xffffffffu, send, 4)));
        } else {
          const float c0 = fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 2]));
          const float c1 = fmaxf(__uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 3]));
          const float keep = e ? c1 : c0, send = e ? c0 : c1;
          pooled[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4));
        }
      }
      // the thread now owns channels 64 g + 8 m + (0..7) of pooled pixel (py, px)
      const float4* tA = s_tab + 64 * g + m;               // entry of column 64 g + 8 i + m: tA[8 i]
      const float4* tB = tA + N;
      float yv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 A = tA[8 * i];                        // deq, bias, scale, shift
        float u = KIND == KIND_I8 ? fmaf(pooled[i], A.x, A.y) : pooled[i] + A.y;
        if (Cfg::ACT) {
          const float c = 0.01f * u;
          float gsel = fmaxf(u, c);                        // LeakyReLU(0.01), acff.py:33
          if (any_neg) { if (tB[8 * i].y < 0.f) gsel = fminf(u, c); }
          u = fmaf(gsel, A.z, A.w);                        // eval BatchNorm, acff.py:34
        }
        yv[i] = u;
      }
      const int chunk = 8 * g + m;
      if (OUT == OUT_P16) {
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int qv = __float2int_rn(yv[i] * tB[8 * i].x);
          qv = max(-127, min(127, qv));
          w[i >> 2] |= ((uint32_t)qv & 0xffu) << (8 * (i & 3));
        }
        if (valid && chunk * 8 < NREAL) {
          uint2* oimg = reinterpret_cast<uint2*>(out) + ((size_t)img * OUT_CHUNKS * OP * OP) * 2;
          oimg[((size_t)((chunk >> 1) * OP + py + 2) * OP + px + 2) * 2 + (chunk & 1)] = make_uint2(w[0], w[1]);
        }
      } else {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(yv[2 * j], yv[2 * j + 1]); pk[j] = *reinterpret_cast<uint32_t*>(&h); }
          else      { __half2 h = __floats2half2_rn(yv[2 * j], yv[2 * j + 1]);            pk[j] = *reinterpret_cast<uint32_t*>(&h); }
        }
        if (valid && chunk * 8 < NREAL) {
          const uint4 o4 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if (OUT == OUT_P8) {
            uint4* oimg = reinterpret_cast<uint4*>(out) + (size_t)img * OUT_CHUNKS * OP * OP;
            oimg[(chunk * OP + py + 2) * OP + px + 2] = o4;
          } else {
            uint16_t* o = out + ((size_t)(img * OUT_H + py) * OUT_H + px) * NREAL + chunk * 8;
            *reinterpret_cast<uint4*>(o) = o4;
          }
        }
      }
    }
  }
}

// ---- host side: channel table + permuted / sign-folded weight image ------------------------------------------------
// par: the constant-bank epilogue parameters of the activate-then-pool kernels (same numbers, channel order).
template <int N>
inline bool build_pf_table(const EpiParams<N>& par, EpiPFTable<N>* tab) {
  bool any_neg = false;
  for (int j = 0; j < N; ++j) {
    const int c = pf_channel_of_column(j);
    const float sg = par.scale[c] < 0.f ? -1.f : 1.f;
    any_neg |= sg < 0.f;
    tab->A[j] = make_float4(par.deq[c], sg * par.bias[c], sg * par.scale[c], par.shift[c]);
    tab->B[j] = make_float4(par.out_inv[c], sg, 0.f, 0.f);
  }
  return any_neg;
}

// Weight image rows: [outer][N][16 bytes] (outer = tap x chunk, or pair x 2) -> same shape with row j holding channel
// pf_channel_of_column(j), multiplied by sigma (16-bit kinds: flip the sign bit of the eight elements; int8: negate).
template <int N>
inline void permute_weight_rows_pf(const uint8_t* src, uint8_t* dst, size_t outer, const EpiPFTable<N>& tab, bool int8) {
  for (size_t o = 0; o < outer; ++o)
    for (int j = 0; j < N; ++j) {
      const int c = pf_channel_of_column(j);
      const uint8_t* s = src + (o * N + c) * 16;
      uint8_t* d = dst + (o * N + j) * 16;
      const bool neg = tab.B[j].y < 0.f;
      for (int b = 0; b < 16; ++b) {
        if (!neg) d[b] = s[b];
        else if (int8) d[b] = (uint8_t)(-(int8_t)s[b]);
        else d[b] = (b & 1) ? (uint8_t)(s[b] ^ 0x80u) : s[b];       // little-endian 16-bit elements: sign bit in the odd byte
      }
    }
}

}  // namespace tc
}  // namespace ernet
