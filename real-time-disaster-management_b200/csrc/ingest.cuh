// Ingest: uint8 HWC frames -> Resize(159) -> CenterCrop(140) -> ToTensor -> Normalize.
//
// Replaces `squeeze_transforms` (code/disaster_detection/dataloaders/aider.py:412-426,431).
// Integer arithmetic is bit-exact with Pillow's 8-bit resampler (two passes, 22-bit fixed-point
// coefficients, round-half-up and clip to uint8 between passes); only the 140x140 crop window of
// the resized image is ever computed.  The float stage is a 256x3 table built on the host with
// the same fp32 operations torch applies, so it is bit-exact as well.
#pragma once
#include <math.h>

#include <vector>

#include "common.cuh"

namespace ernet {

constexpr int kCrop = 140;
constexpr int kResizeShort = 159;
constexpr int kPrecisionBits = 22;
constexpr int kMaxTaps = 64;  // frames up to ~31x down-scaling
constexpr int kStemBand = 14;                 // conv1 output rows per CTA of the fused transform+stem kernel
constexpr int kStemXnRows = 2 * kStemBand + 1;
constexpr size_t kRawStageBytes = 48 * 1024;

struct IngestTables {
  int H = 0, W = 0, new_h = 0, new_w = 0, top = 0, left = 0;
  int crop = 140, resize = 159;  // CenterCrop / Resize of this handle's transform: 140 / 159 (squeeze_transforms), 240 / 273 (aider_transforms, ErNET)
  int ksx = 0, ksy = 0;          // taps per output column / row (padded table width)
  int band_rows = 0;             // output rows per CTA
  int max_in_rows = 0;           // input rows a band needs at most
  // fused transform+stem kernel: bands of kStemBand conv1 rows
  int fs_max_in_rows = 0;        // input rows such a band needs at most
  int fs_stage_raw = 0;          // 1: the raw rows of a band fit in shared memory and are staged with 16-byte loads
  int col_lo = 0, col_hi = 0;    // same for columns
  int row_lo = 0, row_hi = 0;    // input rows [row_lo, row_hi) are the only ones the crop window reads
  int fast5_ok = 0;              // 1: 5 non-negative taps per axis, sums within the no-clamp bound of ingest_fast.cuh
  // device arrays (one allocation): per crop column / row
  int* d_base = nullptr;
  int *d_xmin = nullptr, *d_xlen = nullptr, *d_kx = nullptr;
  int *d_ymin = nullptr, *d_ylen = nullptr, *d_ky = nullptr;
  float* d_lut = nullptr;        // [256][3]
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter,
// restricted to the `count` outputs starting at `first`.  in == out -> identity pass.
inline void host_coeffs(int in_size, int out_size, int first, int count, int& ksize,
                        std::vector<int>& xmin, std::vector<int>& xlen, std::vector<int>& kk) {
  xmin.assign(count, 0);
  xlen.assign(count, 0);
  if (in_size == out_size) {  // Pillow skips the pass entirely
    ksize = 1;
    kk.assign(count, 1 << kPrecisionBits);
    for (int i = 0; i < count; ++i) { xmin[i] = first + i; xlen[i] = 1; }
    return;
  }
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  kk.assign((size_t)count * ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> k(ksize);
  for (int i = 0; i < count; ++i) {
    const int xx = first + i;
    const double center = (xx + 0.5) * scale;
    int x0 = (int)(center - support + 0.5);
    if (x0 < 0) x0 = 0;
    int x1 = (int)(center + support + 0.5);
    if (x1 > in_size) x1 = in_size;
    const int n = x1 - x0;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      double a = (x + x0 - center + 0.5) * ss;
      if (a < 0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < n; ++x) {
      double v = (ww != 0.0) ? k[x] / ww : k[x];
      kk[(size_t)i * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    xmin[i] = x0;
    xlen[i] = n;
  }
}

inline int py_round_half_even(double v) {  // Python round(): CenterCrop uses int(round(x / 2.0))
  double f = floor(v);
  double d = v - f;
  if (d > 0.5) return (int)f + 1;
  if (d < 0.5) return (int)f;
  return ((long long)f % 2 == 0) ? (int)f : (int)f + 1;
}

// ------------------------------------------------------------------------------------------
// One CTA = one band of `band_rows` output rows of one frame.
//   phase 1: horizontal pass over the input rows the band needs -> smem uint8 [rows][140][3]
//   phase 2: vertical pass + table lookup -> output tensor
template <typename TO>
__global__ void __launch_bounds__(256)
ingest_kernel(const uint8_t* __restrict__ frames, int H, int W, int bgr,
              const int* __restrict__ xmin, const int* __restrict__ xlen, const int* __restrict__ kx, int ksx,
              const int* __restrict__ ymin, const int* __restrict__ ylen, const int* __restrict__ ky, int ksy,
              const float* __restrict__ lut, int band_rows, int crop,
              TO* __restrict__ out, long long out_sb, long long out_sc, long long out_sy, long long out_sx) {
  extern __shared__ uint8_t hbuf[];  // [in_rows][crop*3]
  const int b = blockIdx.y;
  const int oy0 = blockIdx.x * band_rows;
  const int oy1 = min(oy0 + band_rows, crop);
  const int r0 = ymin[oy0];
  const int r1 = ymin[oy1 - 1] + ylen[oy1 - 1];
  const int in_rows = r1 - r0;
  const uint8_t* src = frames + (size_t)b * H * W * 3;

  for (int idx = threadIdx.x; idx < in_rows * crop; idx += blockDim.x) {
    const int r = idx / crop, ox = idx - r * crop;
    const int x0 = xmin[ox], n = xlen[ox];
    const uint8_t* p = src + ((size_t)(r0 + r) * W + x0) * 3;
    const int* k = kx + ox * ksx;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int c = __ldg(k + t);
      a0 += c * (int)__ldg(p + 3 * t);
      a1 += c * (int)__ldg(p + 3 * t + 1);
      a2 += c * (int)__ldg(p + 3 * t + 2);
    }
    uint8_t* h = hbuf + (size_t)idx * 3;
    h[0] = (uint8_t)min(max(a0 >> kPrecisionBits, 0), 255);
    h[1] = (uint8_t)min(max(a1 >> kPrecisionBits, 0), 255);
    h[2] = (uint8_t)min(max(a2 >> kPrecisionBits, 0), 255);
  }
  __syncthreads();

  const int rows = oy1 - oy0;
  for (int idx = threadIdx.x; idx < rows * crop; idx += blockDim.x) {
    const int ry = idx / crop, ox = idx - ry * crop;
    const int oy = oy0 + ry;
    const int y0 = ymin[oy] - r0, n = ylen[oy];
    const int* k = ky + oy * ksy;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int c = __ldg(k + t);
      const uint8_t* h = hbuf + ((size_t)(y0 + t) * crop + ox) * 3;
      a0 += c * (int)h[0];
      a1 += c * (int)h[1];
      a2 += c * (int)h[2];
    }
    int v[3] = {min(max(a0 >> kPrecisionBits, 0), 255), min(max(a1 >> kPrecisionBits, 0), 255),
                min(max(a2 >> kPrecisionBits, 0), 255)};
    if (bgr) { int t = v[0]; v[0] = v[2]; v[2] = t; }
    TO* o = out + b * out_sb + oy * out_sy + ox * out_sx;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * out_sc] = from_f32<TO>(__ldg(lut + v[c] * 3 + c));
  }
}


// ------------------------------------------------------------------------------------------
// Fused eval transform + conv1 (model/squeeze_ernet.py:11,25; RedConv: conv_red1 folded in) for the frames
// path.  One CTA = one band of kStemBand conv1 output rows of one frame:
//   phase 0  the raw uint8 rows the band needs are contiguous in HBM: staged with 16-byte loads
//   phase 1  horizontal resample pass  -> smem uint8 [rows][140][3]
//   phase 2  vertical pass + ToTensor/Normalize table -> smem T [<=29][140][3]  (rounded to the engine's
//            element type exactly like the standalone ingest kernel, so both paths agree bit for bit)
//   phase 3  3x3/s2 conv, fp32 accumulate in the same order as stem_kernel -> stem tensor
// Output formats: FS_NHWC (B,69,69,CS) T | FS_P8 (B,2,72,72,8) 16-bit + zero halo | FS_P16 (B,2,72,72,16) int8.
enum : int { FS_NHWC = 0, FS_P8 = 1, FS_P16 = 2 };
struct StemQ { float inv[16]; };   // FS_P16: 1 / int8 step of each stem channel

constexpr int kFusedThreads = 512;   // 2 CTAs per SM (<= 64 registers); the 5-tap fast path uses 3 row groups x 140 columns
constexpr int kFusedRowGroups = 3;


// ---- warp-level tensor-core helper for the stem conv (legacy mma.sync path; the stem is 4.5 % of the MACs and
// lives inside the CUDA-core transform kernel, so a TMEM pipeline would not pay) ------------------------------
template <typename T> struct StemMma;
template <> struct StemMma<__half> {
  static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
};
template <> struct StemMma<__nv_bfloat16> {
  static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
};
template <> struct StemMma<float> {   // never used (fp32 engine keeps FFMA); present so the template instantiates
  static __device__ __forceinline__ void mma(float (&)[4], const uint32_t (&)[4], const uint32_t (&)[2]) {}
  static __device__ __forceinline__ uint32_t pack(float, float) { return 0u; }
};

// KS = 5: specialisation for frames whose resample needs exactly 5 taps per axis (e.g. 240x240 -> 159) with the raw
// rows staged in shared memory: every thread owns one output column, keeps its 5 horizontal coefficients in
// registers and walks down the rows with fully unrolled taps (tables are zero-padded to 5 taps, so border columns
// just multiply neighbouring bytes by 0).  KS = 0: generic run-time tap counts.
template <typename T, int CS, int OUT, int KS>
__global__ void __launch_bounds__(kFusedThreads, 2)
ingest_stem_kernel(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ frames_end, int H, int W, int bgr, int stage_raw,
                   const int* __restrict__ xmin, const int* __restrict__ xlen, const int* __restrict__ kx, int ksx,
                   const int* __restrict__ ymin, const int* __restrict__ ylen, const int* __restrict__ ky, int ksy,
                   const float* __restrict__ lut, const float* __restrict__ w /*[27][CS]*/, const float* __restrict__ bias,
                   const __grid_constant__ StemQ q, int hbuf_bytes, void* __restrict__ out) {
  pdl_wait();                 // writes the stem tensor the previous step's block 1 read: wait before anything else
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t fs_smem[];
  // layout: [xn: kStemXnRows*140*3 T][weights: 28*CS float][hbuf: hbuf_bytes][raw: rest]
  T* xn = reinterpret_cast<T*>(fs_smem);
  constexpr int XN_BYTES = (kStemXnRows * kCrop * 3 * (int)sizeof(T) + 15) / 16 * 16;
  float* ws = reinterpret_cast<float*>(fs_smem + XN_BYTES);
  uint8_t* hbuf = fs_smem + XN_BYTES + 28 * CS * 4;
  uint8_t* raw = hbuf + hbuf_bytes;

  const int b = blockIdx.y;
  const int y0 = blockIdx.x * kStemBand;
  const int y1 = min(y0 + kStemBand, 69);
  const int n0 = 2 * y0, n1 = 2 * (y1 - 1) + 2;          // transform rows n0 .. n1 (inclusive)
  const int r0 = ymin[n0];
  const int r1 = ymin[n1] + ylen[n1];
  const int in_rows = r1 - r0;
  const uint8_t* src = frames + (size_t)b * H * W * 3;

  for (int i = threadIdx.x; i < 28 * CS; i += blockDim.x) ws[i] = i < 27 * CS ? w[i] : bias[i - 27 * CS];

  // ---- phase 0: stage raw rows
  const uint8_t* band = src + (size_t)r0 * W * 3;
  int raw_off = 0;
  if (stage_raw) {
    const size_t a0 = reinterpret_cast<size_t>(band) & ~(size_t)15;
    raw_off = (int)(reinterpret_cast<size_t>(band) - a0);
    const int nvec = (raw_off + in_rows * W * 3 + 15) / 16;
    const uint8_t* g = reinterpret_cast<const uint8_t*>(a0);
    // vectors that would touch bytes outside [frames, frames_end) are assembled byte-wise
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const uint8_t* pv = g + (size_t)i * 16;
      if (pv >= frames && pv + 16 <= frames_end) {
        reinterpret_cast<uint4*>(raw)[i] = __ldg(reinterpret_cast<const uint4*>(pv));
      } else {
        for (int e = 0; e < 16; ++e) raw[i * 16 + e] = (pv + e >= frames && pv + e < frames_end) ? __ldg(pv + e) : (uint8_t)0;
      }
    }
    __syncthreads();
  }

  // ---- phase 1: horizontal pass
  if (KS == 5) {
    if (threadIdx.x < kFusedRowGroups * kCrop) {
      const int ox = threadIdx.x % kCrop, rg = threadIdx.x / kCrop;
      int kc[5];
#pragma unroll
      for (int t = 0; t < 5; ++t) kc[t] = __ldg(kx + ox * 5 + t);
      const uint8_t* p = raw + raw_off + (rg * W + xmin[ox]) * 3;
      uint8_t* h = hbuf + (rg * kCrop + ox) * 3;
      for (int r = rg; r < in_rows; r += kFusedRowGroups, p += kFusedRowGroups * W * 3, h += kFusedRowGroups * kCrop * 3) {
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          a0 += kc[t] * (int)p[3 * t]; a1 += kc[t] * (int)p[3 * t + 1]; a2 += kc[t] * (int)p[3 * t + 2];
        }
        h[0] = (uint8_t)min(max(a0 >> kPrecisionBits, 0), 255);
        h[1] = (uint8_t)min(max(a1 >> kPrecisionBits, 0), 255);
        h[2] = (uint8_t)min(max(a2 >> kPrecisionBits, 0), 255);
      }
    }
  } else
  for (int idx = threadIdx.x; idx < in_rows * kCrop; idx += blockDim.x) {
    const int r = idx / kCrop, ox = idx - r * kCrop;
    const int x0 = xmin[ox], n = xlen[ox];
    const int* k = kx + ox * ksx;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    if (stage_raw) {
      const uint8_t* p = raw + raw_off + (r * W + x0) * 3;
      for (int t = 0; t < n; ++t) {
        const int c = __ldg(k + t);
        a0 += c * (int)p[3 * t]; a1 += c * (int)p[3 * t + 1]; a2 += c * (int)p[3 * t + 2];
      }
    } else {
      const uint8_t* p = band + ((size_t)r * W + x0) * 3;
      for (int t = 0; t < n; ++t) {
        const int c = __ldg(k + t);
        a0 += c * (int)__ldg(p + 3 * t); a1 += c * (int)__ldg(p + 3 * t + 1); a2 += c * (int)__ldg(p + 3 * t + 2);
      }
    }
    uint8_t* h = hbuf + (size_t)idx * 3;
    h[0] = (uint8_t)min(max(a0 >> kPrecisionBits, 0), 255);
    h[1] = (uint8_t)min(max(a1 >> kPrecisionBits, 0), 255);
    h[2] = (uint8_t)min(max(a2 >> kPrecisionBits, 0), 255);
  }
  __syncthreads();

  // ---- phase 2: vertical pass + normalise -> xn
  const int nrows = n1 - n0 + 1;
  if (KS == 5) {
    if (threadIdx.x < kFusedRowGroups * kCrop) {
      const int ox = threadIdx.x % kCrop, rg = threadIdx.x / kCrop;
      for (int rn = rg; rn < nrows; rn += kFusedRowGroups) {
        const int oy = n0 + rn;
        const uint8_t* h = hbuf + ((ymin[oy] - r0) * kCrop + ox) * 3;
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll
        for (int t = 0; t < 5; ++t) {
          const int c = __ldg(ky + oy * 5 + t);
          a0 += c * (int)h[t * kCrop * 3]; a1 += c * (int)h[t * kCrop * 3 + 1]; a2 += c * (int)h[t * kCrop * 3 + 2];
        }
        int v[3] = {min(max(a0 >> kPrecisionBits, 0), 255), min(max(a1 >> kPrecisionBits, 0), 255),
                    min(max(a2 >> kPrecisionBits, 0), 255)};
        if (bgr) { int t = v[0]; v[0] = v[2]; v[2] = t; }
#pragma unroll
        for (int c = 0; c < 3; ++c) xn[(rn * kCrop + ox) * 3 + c] = from_f32<T>(__ldg(lut + v[c] * 3 + c));
      }
    }
  } else
  for (int idx = threadIdx.x; idx < nrows * kCrop; idx += blockDim.x) {
    const int rn = idx / kCrop, ox = idx - rn * kCrop;
    const int oy = n0 + rn;
    const int yy = ymin[oy] - r0, n = ylen[oy];
    const int* k = ky + oy * ksy;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int c = __ldg(k + t);
      const uint8_t* h = hbuf + ((size_t)(yy + t) * kCrop + ox) * 3;
      a0 += c * (int)h[0]; a1 += c * (int)h[1]; a2 += c * (int)h[2];
    }
    int v[3] = {min(max(a0 >> kPrecisionBits, 0), 255), min(max(a1 >> kPrecisionBits, 0), 255),
                min(max(a2 >> kPrecisionBits, 0), 255)};
    if (bgr) { int t = v[0]; v[0] = v[2]; v[2] = t; }
#pragma unroll
    for (int c = 0; c < 3; ++c) xn[idx * 3 + c] = from_f32<T>(__ldg(lut + v[c] * 3 + c));
  }
  __syncthreads();

  // ---- phase 3: conv1 (3x3, stride 2) out of shared memory
  const int brow = y1 - y0;
  constexpr int PW = OUT == FS_NHWC ? 69 : 72;           // pixels per output row handled here (incl. halo cols)
  if constexpr (sizeof(T) == 2) {
    // 16-bit engines: im2col fragments gathered from the xn tile feed mma.sync m16n8k16 (16 pixels x 8 channels,
    // K = 27 padded to 32, fp32 accumulate; weights rounded to T).  k = (ky*3+kx)*3+c, and inside one ky the 9
    // values of a pixel are contiguous in xn, so element (k, ox) sits at  (2*ly + k/9)*420 + 6*ox + k%9.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int NT = CS / 8;                             // n-tiles of 8 output channels
    int koff[2][4];                                        // smem offsets of this lane's k indices; -1 = zero padding
#pragma unroll
    for (int ss = 0; ss < 2; ++ss)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 16 * ss + 2 * t + (e & 1) + (e >> 1) * 8;
        koff[ss][e] = k < 27 ? (k / 9) * (kCrop * 3) + (k % 9) : -1;
      }
    uint32_t bfrag[2][NT][2];                              // B fragments: w[k][n], k = 2t(+1)(+8), n = 8j + g
#pragma unroll
    for (int ss = 0; ss < 2; ++ss)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int k0 = 16 * ss + 2 * t + 8 * h2;
          const float w0 = k0 < 27 ? ws[k0 * CS + 8 * j + g] : 0.f;
          const float w1 = k0 + 1 < 27 ? ws[(k0 + 1) * CS + 8 * j + g] : 0.f;
          bfrag[ss][j][h2] = StemMma<T>::pack(w0, w1);
        }
    const uint16_t* xs = reinterpret_cast<const uint16_t*>(xn);
    const int ngroups = brow * 5;                          // 5 groups of 16 pixels cover the 69 columns
    for (int grp = warp; grp < ngroups; grp += kFusedThreads / 32) {
      const int ly = grp / 5, xg = grp - ly * 5;
      const int oy = y0 + ly;
      const int ox0 = xg * 16 + g, ox1 = ox0 + 8;
      const int b0 = (2 * ly) * (kCrop * 3) + 6 * min(ox0, 68), b1 = (2 * ly) * (kCrop * 3) + 6 * min(ox1, 68);
      uint32_t afrag[2][4];
#pragma unroll
      for (int ss = 0; ss < 2; ++ss) {
        uint32_t e0[4], e1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          e0[e] = koff[ss][e] >= 0 ? (uint32_t)xs[b0 + koff[ss][e]] : 0u;
          e1[e] = koff[ss][e] >= 0 ? (uint32_t)xs[b1 + koff[ss][e]] : 0u;
        }
        afrag[ss][0] = e0[0] | (e0[1] << 16);              // row g,   k = 2t, 2t+1
        afrag[ss][1] = e1[0] | (e1[1] << 16);              // row g+8
        afrag[ss][2] = e0[2] | (e0[3] << 16);              // row g,   k = 2t+8, 2t+9
        afrag[ss][3] = e1[2] | (e1[3] << 16);              // row g+8
      }
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        acc[j][0] = acc[j][2] = ws[27 * CS + 8 * j + 2 * t];
        acc[j][1] = acc[j][3] = ws[27 * CS + 8 * j + 2 * t + 1];
      }
#pragma unroll
      for (int ss = 0; ss < 2; ++ss)
#pragma unroll
        for (int j = 0; j < NT; ++j) StemMma<T>::mma(acc[j], afrag[ss], bfrag[ss][j]);
      // lane holds channels 8j+2t, 8j+2t+1 of pixels ox0 (acc[j][0..1]) and ox1 (acc[j][2..3])
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int ox = half ? ox1 : ox0;
        if (ox >= 69) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float v0 = acc[j][2 * half], v1 = acc[j][2 * half + 1];
          if (OUT == FS_NHWC) {
            T* o = static_cast<T*>(out) + ((size_t)(b * 69 + oy) * 69 + ox) * CS + 8 * j + 2 * t;
            *reinterpret_cast<uint32_t*>(o) = StemMma<T>::pack(v0, v1);
          } else if (OUT == FS_P8) {
            uint32_t* o = reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72 + (size_t)j * 72 * 72 + (oy + 2) * 72 + ox + 2);
            o[t] = StemMma<T>::pack(v0, v1);
          } else {
            int q0 = __float2int_rn(v0 * q.inv[8 * j + 2 * t]), q1 = __float2int_rn(v1 * q.inv[8 * j + 2 * t + 1]);
            q0 = max(-127, min(127, q0)); q1 = max(-127, min(127, q1));
            uint16_t* o = reinterpret_cast<uint16_t*>(reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72 + (oy + 2) * 72 + ox + 2);
            o[4 * j + t] = (uint16_t)(((uint32_t)q0 & 0xffu) | (((uint32_t)q1 & 0xffu) << 8));
          }
        }
      }
    }
    if (OUT != FS_NHWC) {
      // P8 with CS = 8: chunk 1 is all zero; P16: chunk 1 is all zero; plus the zero halo columns of the band's rows
      uint4* img = reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72;
      const bool zero_chunk1 = (OUT == FS_P16) || (CS == 8);
      for (int i = threadIdx.x; i < brow * 72; i += blockDim.x) {
        const int ly = i / 72, pc = i - ly * 72;
        const bool halo = pc < 2 || pc >= 71;
        if (halo) img[(y0 + ly + 2) * 72 + pc] = make_uint4(0, 0, 0, 0);
        if (halo || zero_chunk1) img[72 * 72 + (y0 + ly + 2) * 72 + pc] = make_uint4(0, 0, 0, 0);
      }
    }
  } else
  for (int idx = threadIdx.x; idx < brow * PW; idx += blockDim.x) {
    const int ly = idx / PW, pc = idx - ly * PW;
    const int oy = y0 + ly;
    const int ox = OUT == FS_NHWC ? pc : pc - 2;
    if (OUT != FS_NHWC && (ox < 0 || ox >= 69)) {          // zero halo columns of this row
      uint4* o = reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72 + (oy + 2) * 72 + pc;
      o[0] = make_uint4(0, 0, 0, 0);
      o[72 * 72] = make_uint4(0, 0, 0, 0);
      continue;
    }
    // keep the 27*CS tap weights in shared memory: without this barrier the compiler hoists all of them out of the pixel
    // loop and, under the 64-register cap of this kernel, parks them in local memory (fp32 instance: 1768 bytes of stack,
    // 0.67 ms per 256 frames instead of the 16-bit instances' 0.05)
    asm volatile("" ::: "memory");
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = c < CS ? ws[27 * CS + c] : 0.f;
    const T* p = xn + ((2 * ly) * kCrop + 2 * ox) * 3;
#pragma unroll
    for (int kyy = 0; kyy < 3; ++kyy)
#pragma unroll
      for (int kxx = 0; kxx < 3; ++kxx)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v = to_f32<T>(p[(kyy * kCrop + kxx) * 3 + c]);
          const float* wr = ws + ((kyy * 3 + kxx) * 3 + c) * CS;
#pragma unroll
          for (int k = 0; k < CS; ++k) acc[k] = fmaf(v, wr[k], acc[k]);
        }
    if (OUT == FS_NHWC) {
      constexpr int NV = Vec16<T>::NV;
      uint4* o4 = reinterpret_cast<uint4*>(static_cast<T*>(out) + ((size_t)(b * 69 + oy) * 69 + ox) * CS);
#pragma unroll
      for (int v = 0; v < CS / NV; ++v) {
        float t[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) t[i] = acc[v * NV + i];
        o4[v] = pack16<T>(t);
      }
    } else {
      uint4* o = reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72 + (oy + 2) * 72 + pc;
      if (OUT == FS_P16) {
        uint32_t wq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t wv = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int qq = __float2int_rn(acc[4 * j + e] * q.inv[4 * j + e]);
            qq = max(-127, min(127, qq));
            wv |= ((uint32_t)qq & 0xffu) << (8 * e);
          }
          wq[j] = wv;
        }
        o[0] = make_uint4(wq[0], wq[1], wq[2], wq[3]);
        o[72 * 72] = make_uint4(0, 0, 0, 0);
      } else {
        if constexpr (sizeof(T) == 2) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            float t8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t8[k] = acc[ch * 8 + k];
            o[ch * 72 * 72] = pack16<T>(t8);
          }
        }
      }
    }
  }
  if (OUT != FS_NHWC) {                                     // zero halo rows 0,1 (first band) and 71 (last band)
    uint4* img = reinterpret_cast<uint4*>(out) + (size_t)b * 2 * 72 * 72;
    if (blockIdx.x == 0)
      for (int i = threadIdx.x; i < 2 * 2 * 72; i += blockDim.x) img[(i / 144) * 72 * 72 + (i % 144)] = make_uint4(0, 0, 0, 0);
    if (y1 == 69)
      for (int i = threadIdx.x; i < 2 * 72; i += blockDim.x) img[(i / 72) * 72 * 72 + 71 * 72 + (i % 72)] = make_uint4(0, 0, 0, 0);
  }
}

template <typename T, int CS, int OUT>
inline size_t ingest_stem_smem(const IngestTables& t) {
  const size_t xn = ((size_t)kStemXnRows * kCrop * 3 * sizeof(T) + 15) / 16 * 16;
  const size_t hb = ((size_t)(t.fs_max_in_rows + 5) * kCrop * 3 + 15) / 16 * 16;   // +5 rows: zero-weight taps past the band
  const size_t raw = t.fs_stage_raw ? ((size_t)t.fs_max_in_rows * t.W * 3 + 64 + 15) / 16 * 16 : 0;
  return xn + 28 * CS * 4 + hb + raw;
}

template <typename T, int CS, int OUT>
inline int launch_ingest_stem(const IngestTables& t, const uint8_t* frames, int batch, int bgr, const float* w,
                              const float* bias, const StemQ& q, void* out, cudaStream_t stream) {
  const size_t hb = ((size_t)(t.fs_max_in_rows + 5) * kCrop * 3 + 15) / 16 * 16;
  dim3 grid((69 + kStemBand - 1) / kStemBand, batch);
  const bool fast5 = t.fs_stage_raw && t.ksx == 5 && t.ksy == 5;
  auto kern = fast5 ? ingest_stem_kernel<T, CS, OUT, 5> : ingest_stem_kernel<T, CS, OUT, 0>;
  ERNET_CUDA(launch_pdl(kern, grid, dim3(kFusedThreads), ingest_stem_smem<T, CS, OUT>(t), stream,
                        frames, frames + (size_t)batch * t.H * t.W * 3, t.H, t.W, bgr, t.fs_stage_raw, t.d_xmin, t.d_xlen, t.d_kx, t.ksx,
                        t.d_ymin, t.d_ylen, t.d_ky, t.ksy, t.d_lut, w, bias, q, (int)hb, out));
  return ERNET_OK;
}

// ToTensor + Normalize as a 256x3 table, each step rounded to fp32 like torch (aider.py:424-425).
inline void host_lut(float* lut) {
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int v = 0; v < 256; ++v)
    for (int c = 0; c < 3; ++c) {
      volatile float x = (float)v / 255.0f;   // ToTensor: .div(255)
      volatile float y = x - mean[c];         // Normalize: .sub_(mean)
      volatile float z = y / stdv[c];         //            .div_(std)
      lut[v * 3 + c] = z;
    }
}

inline int build_ingest_tables(IngestTables& t, int H, int W, int crop = kCrop, int resize = kResizeShort) {
  if (H < 1 || W < 1) return fail(ERNET_ERR_INVALID_ARG, "bad frame size %dx%d", H, W);
  t.H = H; t.W = W; t.crop = crop; t.resize = resize;
  if (W <= H) { t.new_w = resize; t.new_h = (int)((double)resize * H / W); }
  else        { t.new_h = resize; t.new_w = (int)((double)resize * W / H); }
  if (t.new_h < crop || t.new_w < crop) return fail(ERNET_ERR_BAD_SHAPE, "resized frame smaller than crop");
  t.top = py_round_half_even((t.new_h - crop) / 2.0);
  t.left = py_round_half_even((t.new_w - crop) / 2.0);
  std::vector<int> xmin, xlen, kx, ymin, ylen, ky;
  host_coeffs(W, t.new_w, t.left, crop, t.ksx, xmin, xlen, kx);
  host_coeffs(H, t.new_h, t.top, crop, t.ksy, ymin, ylen, ky);
  if (t.ksx > kMaxTaps || t.ksy > kMaxTaps) return fail(ERNET_ERR_UNSUPPORTED, "frame %dx%d needs too many taps", H, W);
  t.col_lo = xmin[0];
  t.col_hi = xmin[crop - 1] + xlen[crop - 1];
  for (int i = 0; i < crop; ++i) { if (xmin[i] < t.col_lo) t.col_lo = xmin[i]; if (xmin[i] + xlen[i] > t.col_hi) t.col_hi = xmin[i] + xlen[i]; }
  t.row_lo = ymin[0];
  t.row_hi = ymin[crop - 1] + ylen[crop - 1];
  for (int i = 0; i < crop; ++i) { if (ymin[i] < t.row_lo) t.row_lo = ymin[i]; if (ymin[i] + ylen[i] > t.row_hi) t.row_hi = ymin[i] + ylen[i]; }
  // band size: largest that keeps the uint8 row buffer within 64 KB
  static const int kBands[] = {28, 20, 14, 10, 7, 5, 4, 2, 1};
  t.band_rows = 1;
  for (int br : kBands) {
    int worst = 0;
    for (int oy0 = 0; oy0 < crop; oy0 += br) {
      int oy1 = oy0 + br < crop ? oy0 + br : crop;
      int rows = ymin[oy1 - 1] + ylen[oy1 - 1] - ymin[oy0];
      if (rows > worst) worst = rows;
    }
    if ((size_t)worst * crop * 3 <= 64 * 1024) { t.band_rows = br; t.max_in_rows = worst; break; }
  }
  if (t.max_in_rows == 0) return fail(ERNET_ERR_UNSUPPORTED, "frame %dx%d too large for the ingest row buffer", H, W);
  if (crop == kCrop) {  // geometry of the fused transform+stem bands (kStemBand conv1 rows -> 2*band+1 transform rows)
    int worst = 0;
    for (int y0 = 0; y0 < 69; y0 += kStemBand) {
      const int y1 = y0 + kStemBand < 69 ? y0 + kStemBand : 69;
      const int n0 = 2 * y0, n1 = 2 * (y1 - 1) + 2;
      const int rows = ymin[n1] + ylen[n1] - ymin[n0];
      if (rows > worst) worst = rows;
    }
    t.fs_max_in_rows = worst;
    t.fs_stage_raw = ((size_t)worst * W * 3 + 64 <= kRawStageBytes) ? 1 : 0;
    if ((size_t)worst * crop * 3 > 72 * 1024) t.fs_max_in_rows = 0;      // fused kernel unavailable: fall back to two kernels
  }

  {  // ingest_fast.cuh keeps 4 * (2^21 + sum k p) in 32 bits and takes the top byte: needs k >= 0 and 1020 * sum k + 2^23 < 2^32
    bool ok = t.ksx == 5 && t.ksy == 5 && crop == kCrop;     // the fast kernel is written for the 140-crop
    auto check = [&](const std::vector<int>& kk, int ks) {
      for (int i = 0; i < crop && ok; ++i) {
        long long sum = 0;
        for (int j = 0; j < ks; ++j) { if (kk[(size_t)i * ks + j] < 0) ok = false; sum += kk[(size_t)i * ks + j]; }
        if (1020LL * sum + (1LL << 23) >= (1LL << 32) || sum < (1LL << 22) - 4096) ok = false;
      }
    };
    if (ok) { check(kx, t.ksx); check(ky, t.ksy); }
    t.fast5_ok = ok ? 1 : 0;
  }

  float lut[256 * 3];
  host_lut(lut);

  const size_t n_int = (size_t)crop * (2 + t.ksx) + (size_t)crop * (2 + t.ksy);
  const size_t bytes = n_int * sizeof(int) + sizeof(lut);
  ERNET_CUDA(cudaMalloc(&t.d_base, bytes));
  std::vector<int> host(n_int);
  size_t o = 0;
  auto put = [&](const std::vector<int>& v, int*& dptr) {
    memcpy(host.data() + o, v.data(), v.size() * sizeof(int));
    dptr = t.d_base + o;
    o += v.size();
  };
  put(xmin, t.d_xmin); put(xlen, t.d_xlen); put(kx, t.d_kx);
  put(ymin, t.d_ymin); put(ylen, t.d_ylen); put(ky, t.d_ky);
  t.d_lut = reinterpret_cast<float*>(t.d_base + n_int);
  ERNET_CUDA(cudaMemcpy(t.d_base, host.data(), n_int * sizeof(int), cudaMemcpyHostToDevice));
  ERNET_CUDA(cudaMemcpy(t.d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
  return ERNET_OK;
}

template <typename TO>
inline int launch_ingest(const IngestTables& t, const uint8_t* frames, int batch, int bgr, TO* out,
                         long long sb, long long sc, long long sy, long long sx, cudaStream_t stream) {
  const size_t smem = (size_t)t.max_in_rows * t.crop * 3;
  dim3 grid((t.crop + t.band_rows - 1) / t.band_rows, batch);
  ingest_kernel<TO><<<grid, 256, smem, stream>>>(frames, t.H, t.W, bgr, t.d_xmin, t.d_xlen, t.d_kx, t.ksx,
                                                 t.d_ymin, t.d_ylen, t.d_ky, t.ksy, t.d_lut, t.band_rows, t.crop,
                                                 out, sb, sc, sy, sx);
  ERNET_LAUNCH_CHECK("ingest_kernel");
  return ERNET_OK;
}

}  // namespace ernet
