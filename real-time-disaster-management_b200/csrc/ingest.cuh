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

struct IngestTables {
  int H = 0, W = 0, new_h = 0, new_w = 0, top = 0, left = 0;
  int ksx = 0, ksy = 0;          // taps per output column / row (padded table width)
  int band_rows = 0;             // output rows per CTA
  int max_in_rows = 0;           // input rows a band needs at most
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
              const float* __restrict__ lut, int band_rows,
              TO* __restrict__ out, long long out_sb, long long out_sc, long long out_sy, long long out_sx) {
  extern __shared__ uint8_t hbuf[];  // [in_rows][140*3]
  const int b = blockIdx.y;
  const int oy0 = blockIdx.x * band_rows;
  const int oy1 = min(oy0 + band_rows, kCrop);
  const int r0 = ymin[oy0];
  const int r1 = ymin[oy1 - 1] + ylen[oy1 - 1];
  const int in_rows = r1 - r0;
  const uint8_t* src = frames + (size_t)b * H * W * 3;

  for (int idx = threadIdx.x; idx < in_rows * kCrop; idx += blockDim.x) {
    const int r = idx / kCrop, ox = idx - r * kCrop;
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
  for (int idx = threadIdx.x; idx < rows * kCrop; idx += blockDim.x) {
    const int ry = idx / kCrop, ox = idx - ry * kCrop;
    const int oy = oy0 + ry;
    const int y0 = ymin[oy] - r0, n = ylen[oy];
    const int* k = ky + oy * ksy;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int t = 0; t < n; ++t) {
      const int c = __ldg(k + t);
      const uint8_t* h = hbuf + ((size_t)(y0 + t) * kCrop + ox) * 3;
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

inline int build_ingest_tables(IngestTables& t, int H, int W) {
  if (H < 1 || W < 1) return fail(ERNET_ERR_INVALID_ARG, "bad frame size %dx%d", H, W);
  t.H = H; t.W = W;
  if (W <= H) { t.new_w = kResizeShort; t.new_h = (int)((double)kResizeShort * H / W); }
  else        { t.new_h = kResizeShort; t.new_w = (int)((double)kResizeShort * W / H); }
  if (t.new_h < kCrop || t.new_w < kCrop) return fail(ERNET_ERR_BAD_SHAPE, "resized frame smaller than crop");
  t.top = py_round_half_even((t.new_h - kCrop) / 2.0);
  t.left = py_round_half_even((t.new_w - kCrop) / 2.0);
  std::vector<int> xmin, xlen, kx, ymin, ylen, ky;
  host_coeffs(W, t.new_w, t.left, kCrop, t.ksx, xmin, xlen, kx);
  host_coeffs(H, t.new_h, t.top, kCrop, t.ksy, ymin, ylen, ky);
  if (t.ksx > kMaxTaps || t.ksy > kMaxTaps) return fail(ERNET_ERR_UNSUPPORTED, "frame %dx%d needs too many taps", H, W);
  // band size: largest that keeps the uint8 row buffer within 64 KB
  static const int kBands[] = {28, 20, 14, 10, 7, 5, 4, 2, 1};
  t.band_rows = 1;
  for (int br : kBands) {
    int worst = 0;
    for (int oy0 = 0; oy0 < kCrop; oy0 += br) {
      int oy1 = oy0 + br < kCrop ? oy0 + br : kCrop;
      int rows = ymin[oy1 - 1] + ylen[oy1 - 1] - ymin[oy0];
      if (rows > worst) worst = rows;
    }
    if ((size_t)worst * kCrop * 3 <= 64 * 1024) { t.band_rows = br; t.max_in_rows = worst; break; }
  }
  if (t.max_in_rows == 0) return fail(ERNET_ERR_UNSUPPORTED, "frame %dx%d too large for the ingest row buffer", H, W);

  float lut[256 * 3];
  host_lut(lut);

  const size_t n_int = (size_t)kCrop * (2 + t.ksx) + (size_t)kCrop * (2 + t.ksy);
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
  const size_t smem = (size_t)t.max_in_rows * kCrop * 3;
  dim3 grid((kCrop + t.band_rows - 1) / t.band_rows, batch);
  ingest_kernel<TO><<<grid, 256, smem, stream>>>(frames, t.H, t.W, bgr, t.d_xmin, t.d_xlen, t.d_kx, t.ksx,
                                                 t.d_ymin, t.d_ylen, t.d_ky, t.ksy, t.d_lut, t.band_rows,
                                                 out, sb, sc, sy, sx);
  ERNET_LAUNCH_CHECK("ingest_kernel");
  return ERNET_OK;
}

}  // namespace ernet
