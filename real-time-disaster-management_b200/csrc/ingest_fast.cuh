// Fused eval transform + conv1 for the 16-bit tensor-core engines, 5-tap frames (e.g. 240x240 -> 159): word-wide version
// of ingest_stem_kernel<.., KS = 5> (ingest.cuh).  Same integer resample (bit-exact with Pillow: two passes, 22-bit
// coefficients, rounding between the passes - dataloaders/aider.py:421-423), different data movement and a different
// split of the float work:
//   phase 0  raw rows of the band: ONE bulk copy (cp.async.bulk, TMA engine) of the contiguous bytes into shared memory
//   phase 1  horizontal pass: 5 aligned word loads + funnel shifts per output pixel, bytes extracted with PRMT, 15 IMAD
//            on 4x coefficients (the 8-bit result is the top byte: no shift, no clamp), one RGBX word out
//   phase 2  vertical pass: four pixels per thread (5 x LDS.128), 60 IMAD, packed RGB bytes out [rows][420]
//   phase 3  conv1 (model/squeeze_ernet.py:11,25) on mma.sync m16n8k16 fp16 straight from those BYTES: ToTensor and
//            Normalize (aider.py:424-425) are affine per channel, so they are folded into the conv weights and bias
//            (w' = w / (255 std_c), b' = b - sum w mean_c / std_c); the A fragments are the uint8 pixels themselves,
//            turned into exact fp16 by PRMT (0x6400 | v = 1024 + v) and one HSUB2.  This removes the 256x3 table
//            lookup and the rounding of the normalised tensor to 16 bit that the two-kernel path has; the result is
//            closer to the fp32 reference, not bit-identical to the two-kernel path.
// K order of the implicit GEMM: k = ky*10 + (kx*3 + c), 9 real + 1 zero column per ky, 30 -> 32.
#pragma once
#include "ingest.cuh"
#include "tc_common.cuh"

namespace ernet {

constexpr int kFastThreads = 448;            // 14 warps, 2 CTAs per SM; phase 1 uses 3 row groups x 140 columns

struct FastGeom {                             // shared-memory carve-up (bytes), computed on the host
  int off_hbuf, off_vbuf, off_bar, total;
};

// Folded conv1 constants in mma.sync fragment order, built once per handle (build_stem_fragments below):
//   frag[order][lane][s][j][hh]  fp16 pairs w'[k][n], w'[k+1][n], k = 16s + 2t + 8hh, n = 8j + g (lane = 4g + t)
//   then CS floats: folded bias
struct StemFrag { uint32_t frag[2][32][8]; float bias[16]; };
// fp32 engine: the folded weights as TWO fp16 images, w' = hi + lo / kStemLoScale with hi = fp16(w') and
// lo = fp16((w' - hi) * kStemLoScale) (22 significant bits; the uint8 pixels are exact in fp16 and the products are exact in
// the fp32 accumulators, one accumulator per image).  `base` first: the band function reads it through a StemFrag pointer.
struct StemFrag32 { StemFrag base; uint32_t lo[2][32][8]; };
constexpr float kStemLoScale = 65536.f;

// Thread-group synchronisation of one band: the whole CTA (stand-alone kernel) or a named barrier over the NT helper
// threads of the fused transform + block-1 kernel (tc_fblock.cuh).
struct BandSyncCta { static __device__ __forceinline__ void sync() { __syncthreads(); } };
template <int NT> struct BandSyncNamed {
  static __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
};

// Where the raw rows of one band live and how they are staged.
struct BandCopy {
  const uint8_t* a0;      // 16-byte aligned start of the copy
  uint32_t bytes;         // multiple of 16
  int off;                // byte offset of the band's first row inside the staged bytes
  int r0, in_rows;        // first raw row, number of raw rows
  bool bulk;              // one bulk copy (TMA engine); else the threads copy with guards (first / last band of an unaligned buffer)
};
// TAB: 0 = resize tables and conv1 fragments are read from global memory (__ldg), 1 = from generic pointers into shared
// memory (the fused kernel stages them once per CTA: its few helper warps cannot hide global-load latency)
template <int TAB> __device__ __forceinline__ int tab_ld(const int* p) { return TAB ? *p : __ldg(p); }

template <int TAB>
__device__ __forceinline__ BandCopy band_geometry(int b, int band_idx, const uint8_t* frames, const uint8_t* frames_end, int H, int W,
                                                  const int* ymin, const int* ylen) {
  const int y0 = band_idx * kStemBand, y1 = min(y0 + kStemBand, 69);
  const int n0 = 2 * y0, n1 = 2 * (y1 - 1) + 2;
  BandCopy c;
  c.r0 = tab_ld<TAB>(ymin + n0);
  c.in_rows = tab_ld<TAB>(ymin + n1) + tab_ld<TAB>(ylen + n1) - c.r0;
  const int rowb = W * 3;
  // The raw rows of the band are contiguous.  The copy starts at the 16-byte boundary below the band; `off` is carried
  // into the byte offsets of phase 1.  A band whose rounded range leaves [frames, frames_end) is copied by the threads.
  const uint8_t* band = frames + ((size_t)b * H + c.r0) * rowb;
  c.a0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<size_t>(band) & ~(size_t)15);
  c.off = (int)(band - c.a0);
  c.bytes = (uint32_t)((c.off + c.in_rows * rowb + 15) & ~15);
  c.bulk = c.a0 >= frames && c.a0 + c.bytes <= frames_end;
  return c;
}
// one thread: ONE bulk copy (TMA engine) of the band's raw rows, no instructions per byte
__device__ __forceinline__ void band_issue_bulk(uint8_t* fsm, uint64_t* bar, const BandCopy& c) {
  tc::mbar_expect_tx(bar, c.bytes);
  tc::bulk_g2s(fsm, c.a0, c.bytes, bar);
}
struct BandNoHook { __device__ __forceinline__ void operator()() const {} };

// One band (kStemBand conv1 rows) of one frame by a group of NT threads (tid = 0 .. NT-1).  `bar` is an initialised
// mbarrier (count 1) on which the band's bulk copy (already issued when `issued`, else issued here) completes; it has
// completed `bar_uses` phases before.  `after_h()` is called by every thread once the horizontal pass no longer needs the
// raw rows (the fused kernel starts the NEXT band's bulk copy there, under phases 2 and 3 of this one).
// OUT = FS_P8: 16-bit stem tensor (B,2,72,72,8); FS_P16: int8 stem tensor (B,2,72,72,16) quantised with q.inv[c] (the
// int8 engine); zero_chunk1: also write the all-zero second chunk (not needed when block 1 pairs taps and never reads it).
template <typename T, int CS, int OUT, int NT, class Sync, int TAB = 0, class Hook = BandNoHook>
__device__ __forceinline__ void ingest_stem5_band(uint8_t* __restrict__ fsm, const FastGeom& geo, uint64_t* bar, uint32_t bar_uses,
                                                  const BandCopy& cp, bool issued, int tid, int b, int band_idx,
                                                  const uint8_t* __restrict__ frames, const uint8_t* __restrict__ frames_end, int W, int bgr,
                                                  const int* xmin, const int* kx, const int* ymin, const int* ky, const StemFrag* sf,
                                                  const StemQ& q, int zero_chunk1, void* __restrict__ out, Hook after_h = Hook()) {
  const uint32_t* raww = reinterpret_cast<const uint32_t*>(fsm);           // [in_rows][W*3] packed RGB bytes (bulk copy)
  uint32_t* hbuf4 = reinterpret_cast<uint32_t*>(fsm + geo.off_hbuf);       // [in_rows + 5][140] RGBX words
  uint8_t* vbuf = fsm + geo.off_vbuf;                                      // [2*band + 2][420] RGB bytes
  constexpr int NW = NT / 32;

  const int y0 = band_idx * kStemBand;
  const int y1 = min(y0 + kStemBand, 69);
  const int n0 = 2 * y0, n1 = 2 * (y1 - 1) + 2;
  const int r0 = cp.r0, in_rows = cp.in_rows, off = cp.off;
  const int rowb = W * 3;                                                  // bytes per raw row

  // ---- phase 0: stage the raw rows
  if (cp.bulk) {
    if (!issued && tid == 0) band_issue_bulk(fsm, bar, cp);
  } else {
    for (int i = tid; i < (int)(cp.bytes >> 4); i += NT) {
      const uint8_t* pv = cp.a0 + (size_t)i * 16;
      if (pv >= frames && pv + 16 <= frames_end) {
        reinterpret_cast<uint4*>(fsm)[i] = __ldg(reinterpret_cast<const uint4*>(pv));
      } else {
        for (int e = 0; e < 16; ++e) fsm[i * 16 + e] = (pv + e >= frames && pv + e < frames_end) ? __ldg(pv + e) : (uint8_t)0;
      }
    }
  }
  // per-thread constants of phase 1 while the copy is in flight
  constexpr int RG = NT / kCrop;                                          // row groups of phase 1
  const int ox = tid % kCrop, rg = tid / kCrop;
  uint32_t kc[5];
#pragma unroll
  for (int t = 0; t < 5; ++t) kc[t] = (uint32_t)tab_ld<TAB>(kx + ox * 5 + t) << 2;      // 4k: the result byte is the top byte
  const int bo = off + 3 * tab_ld<TAB>(xmin + ox);
  const int sh = (bo & 3) * 8;
  // suspending wait: the kernel is issue-bound, and 14 warps polling for their band's bytes would take issue slots from the
  // co-resident CTA that is computing
  if (cp.bulk) { while (!tc::mbar_try_wait(bar, bar_uses & 1u)) { } }
  else Sync::sync();

  // ---- phase 1: horizontal pass.  Thread = output column, TWO rows per iteration (independent chains: the few helper
  // warps of the fused kernel live on instruction-level parallelism); 5 aligned words cover the 15 bytes of the 5 taps, a
  // funnel shift aligns them to the pixel, bytes come out with constant PRMT selectors.  acc = 4 * (2^21 + sum k p) < 2^32.
  if (tid < RG * kCrop) {
    const int rstride = RG * (rowb >> 2);
    const uint32_t* p = raww + rg * (rowb >> 2) + (bo >> 2);
    uint32_t* h = hbuf4 + rg * kCrop + ox;
    auto hrow = [&](const uint32_t (&w)[5]) -> uint32_t {
      const uint32_t v[4] = {__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh)};
      uint32_t a[3] = {1u << 23, 1u << 23, 1u << 23};
#pragma unroll
      for (int t = 0; t < 5; ++t)
#pragma unroll
        for (int c = 0; c < 3; ++c) a[c] += kc[t] * __byte_perm(v[(3 * t + c) >> 2], 0u, 0x4440u + (uint32_t)((3 * t + c) & 3));
      return __byte_perm(__byte_perm(a[0], a[1], 0x0073), a[2], 0x7710);    // R | G << 8 | B << 16
    };
    int r = rg;
    for (; r + RG < in_rows; r += 2 * RG, p += 2 * rstride, h += 2 * RG * kCrop) {
      uint32_t wa[5], wb[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) { wa[i] = p[i]; wb[i] = p[rstride + i]; }
      const uint32_t ra = hrow(wa), rb = hrow(wb);
      h[0] = ra; h[RG * kCrop] = rb;
    }
    if (r < in_rows) {
      uint32_t wa[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) wa[i] = p[i];
      h[0] = hrow(wa);
    }
  }
  Sync::sync();
  after_h();

  // ---- phase 2: vertical pass, four pixels per task -> packed RGB bytes
  const int nrows = n1 - n0 + 1;
  if (tid < (NT / 35) * 35)
  for (int rn = tid / 35, g = tid - (tid / 35) * 35; rn < nrows; rn += NT / 35) {
    const int oy = n0 + rn;
    const uint4* hp = reinterpret_cast<const uint4*>(hbuf4 + (tab_ld<TAB>(ymin + oy) - r0) * kCrop) + g;
    uint32_t cy[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) cy[t] = (uint32_t)tab_ld<TAB>(ky + oy * 5 + t) << 2;
    uint32_t acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 1u << 23;
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const uint32_t c = cy[t];
      const uint4 qd = hp[t * (kCrop / 4)];
      const uint32_t px[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[3 * j + 0] += c * __byte_perm(px[j], 0u, 0x4440);
        acc[3 * j + 1] += c * __byte_perm(px[j], 0u, 0x4441);
        acc[3 * j + 2] += c * __byte_perm(px[j], 0u, 0x4442);
      }
    }
    uint32_t* vp = reinterpret_cast<uint32_t*>(vbuf + rn * (kCrop * 3)) + 3 * g;
#pragma unroll
    for (int wd = 0; wd < 3; ++wd)
      vp[wd] = __byte_perm(__byte_perm(acc[4 * wd], acc[4 * wd + 1], 0x0073), __byte_perm(acc[4 * wd + 2], acc[4 * wd + 3], 0x0073), 0x5410);
  }
  Sync::sync();

  // ---- phase 3: conv1 3x3 / stride 2 as an implicit GEMM on mma.sync (16 pixels x 8 channels per instruction); one task =
  // 16 output pixels of one row, dealt round-robin to the warps
  {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int NTL = CS / 8;
    int koff[2][2];                                        // byte offset of this lane's k pairs inside a pixel's window
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = 16 * s + 2 * t + 8 * hh;
        koff[s][hh] = (k / 10) * (kCrop * 3) + (k % 10);
      }
    uint32_t bfrag[2][2][2];
    {
      const uint4* fp = reinterpret_cast<const uint4*>(sf->frag[bgr ? 1 : 0][lane]);
      const uint4 f0 = TAB ? fp[0] : __ldg(fp), f1 = TAB ? fp[1] : __ldg(fp + 1);
      bfrag[0][0][0] = f0.x; bfrag[0][0][1] = f0.y; bfrag[0][1][0] = f0.z; bfrag[0][1][1] = f0.w;
      bfrag[1][0][0] = f1.x; bfrag[1][0][1] = f1.y; bfrag[1][1][0] = f1.z; bfrag[1][1][1] = f1.w;
    }
    constexpr bool F32 = OUT == FS_NHWC;                    // fp32 engine: second (lo) weight image, fp32 NHWC stem tensor
    uint32_t bfrag_lo[F32 ? 2 : 1][2][2];
    if (F32) {
      const uint4* fp = reinterpret_cast<const uint4*>(reinterpret_cast<const StemFrag32*>(sf)->lo[bgr ? 1 : 0][lane]);
      const uint4 f0 = TAB ? fp[0] : __ldg(fp), f1 = TAB ? fp[1] : __ldg(fp + 1);
      bfrag_lo[0][0][0] = f0.x; bfrag_lo[0][0][1] = f0.y; bfrag_lo[0][1][0] = f0.z; bfrag_lo[0][1][1] = f0.w;
      bfrag_lo[F32 ? 1 : 0][0][0] = f1.x; bfrag_lo[F32 ? 1 : 0][0][1] = f1.y; bfrag_lo[F32 ? 1 : 0][1][0] = f1.z; bfrag_lo[F32 ? 1 : 0][1][1] = f1.w;
    }
    float bia[NTL][2];
#pragma unroll
    for (int j = 0; j < NTL; ++j) {
      bia[j][0] = TAB ? sf->bias[8 * j + 2 * t] : __ldg(sf->bias + 8 * j + 2 * t);
      bia[j][1] = TAB ? sf->bias[8 * j + 2 * t + 1] : __ldg(sf->bias + 8 * j + 2 * t + 1);
    }
    const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
    const int brow = y1 - y0;
    uint4* img = reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72;
    uint32_t* orow = reinterpret_cast<uint32_t*>(img + (y0 + 2) * 72 + 2) + t;     // lane's 4-byte slot of chunk 0, band row 0, pixel 0
    float* orow32 = reinterpret_cast<float*>(out) + ((size_t)b * 69 + y0) * 69 * CS + 2 * t;   // FS_NHWC: (B,69,69,CS) fp32
    for (int task = warp; task < brow * 5; task += NW) {
      const int ly = task / 5, xg = task - ly * 5;
      const uint8_t* vrow = vbuf + (2 * ly) * (kCrop * 3);
      uint32_t* orow_l = orow + ly * (72 * 4);
      {
        const int ox0 = xg * 16 + g, ox1 = ox0 + 8;
        const uint8_t* base0 = vrow + 6 * min(ox0, 68);
        const uint8_t* base1 = vrow + 6 * min(ox1, 68);
        uint32_t afrag[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t u0 = *reinterpret_cast<const uint16_t*>(base0 + koff[s][hh]);
            const uint32_t u1 = *reinterpret_cast<const uint16_t*>(base1 + koff[s][hh]);
            uint32_t f0 = __byte_perm(u0, 0x64646464u, 0x4140), f1 = __byte_perm(u1, 0x64646464u, 0x4140);   // (1024 + v) fp16 pairs
            __half2 h0 = __hsub2(*reinterpret_cast<const __half2*>(&f0), k1024), h1 = __hsub2(*reinterpret_cast<const __half2*>(&f1), k1024);
            afrag[s][2 * hh + 0] = *reinterpret_cast<const uint32_t*>(&h0);      // row g
            afrag[s][2 * hh + 1] = *reinterpret_cast<const uint32_t*>(&h1);      // row g + 8
          }
        float acc[NTL][4];
#pragma unroll
        for (int j = 0; j < NTL; ++j) { acc[j][0] = acc[j][2] = bia[j][0]; acc[j][1] = acc[j][3] = bia[j][1]; }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int j = 0; j < NTL; ++j) StemMma<__half>::mma(acc[j], afrag[s], bfrag[s][j]);
        if (F32) {
          float acl[NTL][4];
#pragma unroll
          for (int j = 0; j < NTL; ++j) acl[j][0] = acl[j][1] = acl[j][2] = acl[j][3] = 0.f;
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int j = 0; j < NTL; ++j) StemMma<__half>::mma(acl[j], afrag[s], bfrag_lo[F32 ? s : 0][j]);
#pragma unroll
          for (int j = 0; j < NTL; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = fmaf(acl[j][e], 1.f / kStemLoScale, acc[j][e]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int oxx = half ? ox1 : ox0;
          if (oxx >= 69) continue;
          if (F32) {
#pragma unroll
            for (int j = 0; j < NTL; ++j)
              *reinterpret_cast<float2*>(orow32 + ((size_t)ly * 69 + oxx) * CS + 8 * j) = make_float2(acc[j][2 * half], acc[j][2 * half + 1]);
          } else if (OUT == FS_P16) {
            uint16_t* o16 = reinterpret_cast<uint16_t*>(orow_l + oxx * 4 - t) + t;      // pixel's 16 bytes, this lane's channel pairs
#pragma unroll
            for (int j = 0; j < NTL; ++j) {
              int q0 = __float2int_rn(acc[j][2 * half] * q.inv[8 * j + 2 * t]), q1 = __float2int_rn(acc[j][2 * half + 1] * q.inv[8 * j + 2 * t + 1]);
              q0 = max(-127, min(127, q0)); q1 = max(-127, min(127, q1));
              o16[4 * j] = (uint16_t)(((uint32_t)q0 & 0xffu) | (((uint32_t)q1 & 0xffu) << 8));
            }
          } else {
#pragma unroll
            for (int j = 0; j < NTL; ++j) orow_l[j * (72 * 72 * 4) + oxx * 4] = StemMma<T>::pack(acc[j][2 * half], acc[j][2 * half + 1]);
          }
        }
      }
    }
    // zero halo columns of the band's rows (and the all-zero second chunk of an 8-channel stem)
    if (!F32)
    for (int i = tid; i < brow * 72; i += NT) {
      const int ly = i / 72, pc = i - ly * 72;
      const bool halo = pc < 2 || pc >= 71;
      if (halo) img[(y0 + ly + 2) * 72 + pc] = make_uint4(0, 0, 0, 0);
      if (zero_chunk1 ? (halo || CS == 8 || OUT == FS_P16) : (halo && CS == 16 && OUT == FS_P8)) img[72 * 72 + (y0 + ly + 2) * 72 + pc] = make_uint4(0, 0, 0, 0);
    }
    const int nch = (zero_chunk1 || (CS == 16 && OUT == FS_P8)) ? 2 : 1;         // chunks whose halo rows are written
    if (!F32 && band_idx == 0)
      for (int i = tid; i < nch * 2 * 72; i += NT) img[(i / 144) * 72 * 72 + (i % 144)] = make_uint4(0, 0, 0, 0);
    if (!F32 && y1 == 69)
      for (int i = tid; i < nch * 72; i += NT) img[(i / 72) * 72 * 72 + 71 * 72 + (i % 72)] = make_uint4(0, 0, 0, 0);
  }
}

template <typename T, int CS, int OUT = FS_P8>
__global__ void __launch_bounds__(kFastThreads, 2)
ingest_stem5_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ frames_end, int H, int W, int bgr,
                    const int* __restrict__ xmin, const int* __restrict__ kx, const int* __restrict__ ymin, const int* __restrict__ ylen,
                    const int* __restrict__ ky, const StemFrag* __restrict__ sf, const __grid_constant__ FastGeom geo,
                    const __grid_constant__ StemQ q, int zero_chunk1, void* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t fsm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(fsm + geo.off_bar);
  ERNET_CHAIN_ENTRY(0);
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_mbar_init(); }
  __syncthreads();
  pdl_wait();                 // frames may come from the previous kernel of the stream; the stem tensor is read by block 1
  pdl_launch_dependents();
  ERNET_CHAIN_WAITED(0);
  const BandCopy cp = band_geometry<0>(blockIdx.y, blockIdx.x, frames, frames_end, H, W, ymin, ylen);
  ingest_stem5_band<T, CS, OUT, kFastThreads, BandSyncCta>(fsm, geo, bar, 0u, cp, false, threadIdx.x, blockIdx.y, blockIdx.x, frames, frames_end, W, bgr,
                                                           xmin, kx, ymin, ky, sf, q, zero_chunk1, out);
  ERNET_CHAIN_EXIT(0);
}

// Host: fold ToTensor + Normalize (aider.py:424-425) into conv1 and lay the result out as mma.sync B fragments.
// w = [27][CS] fp32 (k = (ky*3+kx)*3 + c), bias = [CS].
inline void build_stem_fragments(const float* w, const float* bias, int cs, StemFrag* out) {
  const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
  memset(out, 0, sizeof(*out));
  for (int order = 0; order < 2; ++order) {
    auto wk = [&](int k, int n) -> float {                 // folded weight of GEMM row k (ky*10 + kx*3 + byte), column n
      const int kyy = k / 10, off = k % 10;
      if (kyy >= 3 || off >= 9 || n >= cs) return 0.f;
      const int kxx = off / 3, cb = off % 3;
      const int c = order ? 2 - cb : cb;                    // model channel of this byte (order 1 = BGR frames)
      return (float)((double)w[((kyy * 3 + kxx) * 3 + c) * cs + n] / (255.0 * stdv[c]));
    };
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int s = 0; s < 2; ++s)
        for (int j = 0; j < 2; ++j)
          for (int hh = 0; hh < 2; ++hh) {
            const int k0 = 16 * s + 2 * t + 8 * hh, n = 8 * j + g;
            const __half lo = __float2half_rn(wk(k0, n)), hi = __float2half_rn(wk(k0 + 1, n));
            uint16_t lob, hib;
            memcpy(&lob, &lo, 2); memcpy(&hib, &hi, 2);
            out->frag[order][lane][(s * 2 + j) * 2 + hh] = (uint32_t)lob | ((uint32_t)hib << 16);
          }
    }
  }
  for (int n = 0; n < cs; ++n) {
    double s = bias[n];
    for (int k = 0; k < 27; ++k) s -= (double)w[k * cs + n] * mean[k % 3] / stdv[k % 3];
    out->bias[n] = (float)s;
  }
}

// fp32 engine: the same fold, every weight as the (hi, lo) fp16 pair of StemFrag32
inline void build_stem_fragments32(const float* w, const float* bias, int cs, StemFrag32* out) {
  build_stem_fragments(w, bias, cs, &out->base);            // bias (fp32) and the hi image = fp16(w')
  const double stdv[3] = {0.229, 0.224, 0.225};
  memset(out->lo, 0, sizeof(out->lo));
  for (int order = 0; order < 2; ++order) {
    auto lo_of = [&](int k, int n) -> __half {
      const int kyy = k / 10, off = k % 10;
      if (kyy >= 3 || off >= 9 || n >= cs) return __float2half_rn(0.f);
      const int kxx = off / 3, cb = off % 3;
      const int c = order ? 2 - cb : cb;
      const double wd = (double)w[((kyy * 3 + kxx) * 3 + c) * cs + n] / (255.0 * stdv[c]);
      const double hi = (double)__half2float(__float2half_rn((float)wd));
      return __float2half_rn((float)((wd - hi) * (double)kStemLoScale));
    };
    for (int lane = 0; lane < 32; ++lane) {
      const int g = lane >> 2, t = lane & 3;
      for (int s = 0; s < 2; ++s)
        for (int j = 0; j < 2; ++j)
          for (int hh = 0; hh < 2; ++hh) {
            const int k0 = 16 * s + 2 * t + 8 * hh, n = 8 * j + g;
            const __half lo = lo_of(k0, n), hi = lo_of(k0 + 1, n);
            uint16_t lob, hib;
            memcpy(&lob, &lo, 2); memcpy(&hib, &hi, 2);
            out->lo[order][lane][(s * 2 + j) * 2 + hh] = (uint32_t)lob | ((uint32_t)hib << 16);
          }
    }
  }
}

// Geometry + eligibility of the fast path: exactly 5 non-negative taps per axis whose sums leave the 8-bit result in
// the top byte of 4*acc, rows that are whole words, two CTAs per SM.
inline bool fast5_geometry(const IngestTables& t, const void* frames, FastGeom& g) {
  (void)frames;
  if (!t.fast5_ok || t.fs_max_in_rows <= 0 || (t.W % 4) != 0) return false;
  const int raw = t.fs_max_in_rows * t.W * 3 + 64;
  g.off_hbuf = (raw + 127) / 128 * 128;
  g.off_vbuf = g.off_hbuf + (t.fs_max_in_rows + 5) * kCrop * 4;
  g.off_bar = (g.off_vbuf + (kStemXnRows + 1) * kCrop * 3 + 16 + 15) / 16 * 16;
  g.total = g.off_bar + 16;
  return g.total <= 110 * 1024;
}

template <typename T, int CS, int OUT = FS_P8>
inline int launch_ingest_stem5(const IngestTables& t, const FastGeom& g, const uint8_t* frames, int batch, int bgr, const StemFrag* sf,
                               const StemQ& q, bool zero_chunk1, void* out, cudaStream_t stream) {
  dim3 grid((69 + kStemBand - 1) / kStemBand, batch);
  ERNET_CUDA(launch_pdl(ingest_stem5_kernel<T, CS, OUT>, grid, dim3(kFastThreads), (size_t)g.total, stream, frames, frames + (size_t)batch * t.H * t.W * 3, t.H, t.W, bgr,
                        t.d_xmin, t.d_kx, t.d_ymin, t.d_ylen, t.d_ky, sf, g, q, zero_chunk1 ? 1 : 0, out));
  return ERNET_OK;
}

}  // namespace ernet
