// CUDA-core ("simt") kernels of the Squeeze-ErNet forward pass: fp32 arithmetic, activations stored
// NHWC in T (fp32 / fp16 / bf16).  This is the product path for precision=fp32 (tensor cores top out
// at TF32, which cannot hold the 1e-4 logit bound) and the layer-wise reference implementation that
// the fused tensor-core kernels are validated against on the device.
//
//   stem_kernel          conv1 3x3/s2 3->16            model/squeeze_ernet.py:11,25
//                        (+conv_red1 folded, 3->8)     model/squeeze_ernet_redconv.py:12,28-29
//   acff_dw_kernel       three dilated depthwise 3x3 + concat      model/acff.py:25-30,46
//   pointwise_kernel     1x1 conv + bias [+LeakyReLU] [+BN affine] [+2x2 max-pool]
//                                                      model/acff.py:31-34,51-53, squeeze_ernet.py:13
//   head_kernel          conv2 -> AvgPool(5,1,1) -> view -> fc -> softmax, collapsed
//                                                      model/squeeze_ernet.py:19-22,33-41
#pragma once
#include "common.cuh"

namespace ernet {

// ---------------------------------------------------------------------------------- stem
// One thread = one output pixel, all CS output channels.  Input addressed through element strides so
// NCHW (the reference's layout) and NHWC (the ingest kernel's) are both read in place.
template <typename TI, typename TO, int CS>
__global__ void __launch_bounds__(128)
stem_kernel(const TI* __restrict__ x, long long sb, long long sc, long long sy, long long sx,
            const float* __restrict__ w /*[27][CS]*/, const float* __restrict__ bias /*[CS]*/,
            TO* __restrict__ out /*(B,HO,HO,CS)*/, int total, int HO = 69) {
  __shared__ float ws[27 * CS + CS];
  for (int i = threadIdx.x; i < 27 * CS + CS; i += blockDim.x) ws[i] = i < 27 * CS ? w[i] : bias[i - 27 * CS];
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = idx / (HO * HO);
  const int r = idx - b * HO * HO;
  const int oy = r / HO, ox = r - oy * HO;
  float acc[CS];
#pragma unroll
  for (int o = 0; o < CS; ++o) acc[o] = ws[27 * CS + o];
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
        for (int o = 0; o < CS; ++o) acc[o] = fmaf(v, wr[o], acc[o]);
      }
  constexpr int NV = Vec16<TO>::NV;
  uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)idx * CS);
#pragma unroll
  for (int v = 0; v < CS / NV; ++v) {
    float t[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) t[i] = acc[v * NV + i];
    o4[v] = pack16<TO>(t);
  }
}

// ---------------------------------------------------------------------------------- depthwise trio
// CTA = one (th x tw) tile of output pixels of one image, all C channels.  The (th+6)x(tw+6) input halo
// (rows/cols y-2 .. y+4: the union of the three dilations) is staged in shared memory with 16-byte
// NHWC vectors, out-of-image taps read zeros (= the conv padding 0/1/2 of acff.py:25-30).  Each thread
// then produces one 16-byte channel vector of one pixel for all three branches and stores them at
// channel offsets 0, C, 2C (the concat of acff.py:46).
template <typename T>
__global__ void __launch_bounds__(256)
acff_dw_kernel(const T* __restrict__ x, int H, int W, int C, int out_h, int out_w, int th, int tw,
               const float* __restrict__ w /*[3][9][C]*/, const float* __restrict__ bias /*[3][C]*/,
               T* __restrict__ out /*(B,out_h,out_w,3C)*/) {
  constexpr int NV = Vec16<T>::NV;
  extern __shared__ uint4 dw_smem[];
  const int cv = C / NV;                       // vectors per pixel
  const int hh = th + 6, hw = tw + 6;
  uint4* tile = dw_smem;                       // [hh][hw][cv]
  float* wsm = reinterpret_cast<float*>(dw_smem + (size_t)hh * hw * cv);  // [27][C] then bias [3][C]
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * th, x0 = blockIdx.x * tw;
  const T* xb = x + (size_t)b * H * W * C;

  for (int i = threadIdx.x; i < 30 * C; i += blockDim.x) wsm[i] = i < 27 * C ? w[i] : bias[i - 27 * C];
  // halo staging, 8 independent 16-byte loads in flight per thread (one load per iteration left the kernel waiting on
  // a chain of ~15 dependent HBM latencies per CTA)
  for (int i0 = threadIdx.x; i0 < hh * hw * cv; i0 += 8 * blockDim.x) {
    uint4 val[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      const int v = i % cv;
      const int p = i / cv;
      const int px = p % hw, py = p / hw;
      const int gy = y0 - 2 + py, gx = x0 - 2 + px;
      val[u] = make_uint4(0, 0, 0, 0);
      if (i < hh * hw * cv && gy >= 0 && gy < H && gx >= 0 && gx < W)
        val[u] = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)gy * W + gx) * C) + v);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < hh * hw * cv) tile[i] = val[u];
    }
  }
  __syncthreads();

  // Each thread produces a strip of PX consecutive pixels for one channel vector: the PX + 6 columns of a tile row
  // are loaded once and feed every branch that has taps in that row (dy = -2..4: d3, d2, d1, d1+d2+d3, d1, d2, d3), the
  // tap's weight vector is loaded once per strip.  24 (fp32) shared-memory vectors per pixel instead of 54: the
  // kernel was bound by shared-memory bandwidth, not by HBM.
  constexpr int PX = NV == 4 ? 4 : 2;
  const int sxn = tw / PX;                     // tw is a multiple of PX (dw_pick_tile)
  const int items = th * sxn * cv;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int v = i % cv;
    const int p = i / cv;
    const int lx0 = (p % sxn) * PX, ly = p / sxn;
    const int oy = y0 + ly, ox0 = x0 + lx0;
    if (oy >= out_h || ox0 >= out_w) continue;
    float acc[3][PX][NV];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int px = 0; px < PX; ++px)
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[d][px][k] = wsm[27 * C + d * C + v * NV + k];
#pragma unroll
    for (int ry = 0; ry < 7; ++ry) {
      const int dy = ry - 2;
      const uint4* rowp = tile + ((size_t)(ly + ry) * hw + lx0) * cv + v;
      float xr[PX + 6][NV];
#pragma unroll
      for (int c = 0; c < PX + 6; ++c) unpack16<T>(rowp[(size_t)c * cv], xr[c]);     // unused columns are dropped by the compiler
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int dil = d + 1;
        // branch d has taps in this row when dy = ky*dil - (dil-1) for ky in 0..2
        const int t = dy + (dil - 1);
        if (t < 0 || t % dil != 0 || t / dil > 2) continue;
        const int ky = t / dil;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wr = wsm + (d * 9 + ky * 3 + kx) * C + v * NV;
          float wv[NV];
#pragma unroll
          for (int k = 0; k < NV; ++k) wv[k] = wr[k];
          const int col = 2 + kx * dil - (dil - 1);          // column of pixel 0's tap in xr
#pragma unroll
          for (int px = 0; px < PX; ++px)
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[d][px][k] = fmaf(xr[px + col][k], wv[k], acc[d][px][k]);
        }
      }
    }
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      if (ox0 + px >= out_w) break;
      T* o = out + (((size_t)b * out_h + oy) * out_w + ox0 + px) * (3 * C) + v * NV;
#pragma unroll
      for (int d = 0; d < 3; ++d) *reinterpret_cast<uint4*>(o + d * C) = pack16<T>(acc[d][px]);
    }
  }
}

// fp32 register-tile form (the product path of precision=fp32).  The shared-memory kernel above moves 24 words
// per output through shared memory (tile data + the tap weights re-read per strip) and that, not HBM, is what
// bounded it (2.2-2.3 TB/s, l1tex data-pipe wavefronts 76 % busy in ncu).  Here one thread owns ONE channel and a
// PY x 4 patch of output pixels for all three branches: its 27 tap weights live in registers for the whole patch,
// every input word of the (PY+6) x 10 footprint is loaded once per thread straight from global memory (lanes =
// consecutive channels, so a warp load is whole 32-byte sectors; neighbouring patches' overlap is served by L1),
// i.e. 6.25 words per output at PY = 4 and no weight traffic.  The channel count is a template parameter so that
// every column / branch offset is an immediate in the load / store instruction (with a run-time C the first
// version spent 4.7x the useful instruction count on address arithmetic and bounds selects and was issue-bound),
// and patches whose footprint lies inside the image take a path without bounds checks.  Per accumulator the FMA
// order is bias, then ky-major / kx-minor - the same as the shared-memory kernel, so the two are bit-identical.
// ADD = the add-fusion flavour of the block (victim_localization/yolov3/models.py:302: conv1(x) + conv2(x) + conv3(x)):
// the three branch sums are added, (b1+S1) + (b2+S2) + (b3+S3) in that order, and C channels are stored instead of 3C.
// C = 0 selects the run-time channel count `c_rt` (channel counts outside the instantiated list).
template <int C, int PY, bool EDGE, bool ADD>
__device__ __forceinline__ void dw_tile_patch(const float* __restrict__ xb /* image base + channel */, int H, int W, int c_rt,
                                              int out_h, int out_w, int oy0, int ox0, const float (&wr)[27],
                                              const float (&bv)[3], float* __restrict__ ob /* image base + channel */) {
  constexpr int PX = 4;
  const int Cc = C > 0 ? C : c_rt;
  float acc[3][PY][PX];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int py = 0; py < PY; ++py)
#pragma unroll
      for (int px = 0; px < PX; ++px) acc[d][py][px] = bv[d];
  const size_t rs = (size_t)W * Cc;                      // row stride in elements
  const float* p0 = xb + ((ptrdiff_t)(oy0 - 2) * W + (ox0 - 2)) * Cc;
#pragma unroll
  for (int ry = 0; ry < PY + 6; ++ry) {
    const float* rp = p0 + ry * rs;
    float xr[PX + 6];
    if constexpr (EDGE) {
      const int iy = oy0 + ry - 2;
      const bool rowok = iy >= 0 && iy < H;
#pragma unroll
      for (int cx = 0; cx < PX + 6; ++cx) {
        const int ix = ox0 + cx - 2;
        xr[cx] = 0.f;
        if (rowok && ix >= 0 && ix < W) xr[cx] = __ldg(rp + cx * Cc);
      }
    } else {
#pragma unroll
      for (int cx = 0; cx < PX + 6; ++cx) xr[cx] = __ldg(rp + cx * Cc);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int dil = d + 1;
#pragma unroll
      for (int py = 0; py < PY; ++py) {
        // input row ry is tap row ky of branch d for output row py when ry - 2 - py = ky*dil - (dil-1)
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
  const int OC = ADD ? Cc : 3 * Cc;
  const size_t os = (size_t)out_w * OC;
  float* o0 = ob + ((size_t)oy0 * out_w + ox0) * OC;
#pragma unroll
  for (int py = 0; py < PY; ++py) {
    if (EDGE && oy0 + py >= out_h) break;
    float* orow = o0 + py * os;
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      if (EDGE && ox0 + px >= out_w) break;
      if constexpr (ADD) {
        orow[px * OC] = (acc[0][py][px] + acc[1][py][px]) + acc[2][py][px];
      } else {
#pragma unroll
        for (int d = 0; d < 3; ++d) orow[px * OC + d * Cc] = acc[d][py][px];
      }
    }
  }
}

template <int C, int PY, int MINB, bool ADD>
__global__ void __launch_bounds__(256, MINB)
acff_dw_tile_kernel(const float* __restrict__ x, int H, int W, int c_rt, int out_h, int out_w, int TX, int TY,
                    long long total, const float* __restrict__ w /*[3][9][C]*/, const float* __restrict__ bias /*[3][C]*/,
                    float* __restrict__ out /*(B,out_h,out_w,3C) or (B,out_h,out_w,C) with ADD*/) {
  constexpr int PX = 4;
  const int Cc = C > 0 ? C : c_rt;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int c = (int)(gid % Cc);
  unsigned t = (unsigned)(gid / Cc);
  const int tx = t % TX; t /= TX;
  const int ty = t % TY;
  const int b = t / TY;
  const int ox0 = tx * PX, oy0 = ty * PY;
  float wr[27], bv[3];
#pragma unroll
  for (int i = 0; i < 27; ++i) wr[i] = __ldg(w + i * Cc + c);
#pragma unroll
  for (int d = 0; d < 3; ++d) bv[d] = __ldg(bias + d * Cc + c);
  const float* xb = x + (size_t)b * H * W * Cc + c;
  float* ob = out + (size_t)b * out_h * out_w * (ADD ? Cc : 3 * Cc) + c;
  const bool interior = oy0 >= 2 && oy0 + PY + 3 < H && ox0 >= 2 && ox0 + PX + 3 < W && oy0 + PY <= out_h && ox0 + PX <= out_w;
  if (interior) dw_tile_patch<C, PY, false, ADD>(xb, H, W, c_rt, out_h, out_w, oy0, ox0, wr, bv, ob);
  else          dw_tile_patch<C, PY, true, ADD>(xb, H, W, c_rt, out_h, out_w, oy0, ox0, wr, bv, ob);
}

inline int g_dw_fp32_form = 2;    // 2 = TMA-staged kernel where it serves the shape, else register tile (default); 1 = register tile;
                                  // 0 = shared-memory kernel (ernet_set_depthwise_form)
// dw_tma.cuh (included after this file by every translation unit that launches the depthwise stage)
inline int launch_acff_dw_tma(const float* x, int batch, int H, int W, int C, int out_h, int out_w, const float* w,
                              const float* bias, float* out, cudaStream_t stream);

template <int C, int PY = 4, int MINB = 2, bool ADD = false>
inline int launch_acff_dw_tile_c(const float* x, int batch, int H, int W, int c_rt, int out_h, int out_w,
                                 const float* w, const float* bias, float* out, cudaStream_t stream) {
  const int Cc = C > 0 ? C : c_rt;
  const int TX = (out_w + 3) / 4, TY = (out_h + PY - 1) / PY;
  const long long total = (long long)batch * TY * TX * Cc;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL || total / Cc > 0xffffffffLL) return fail(ERNET_ERR_INVALID_ARG, "depthwise: batch too large for one launch");
  acff_dw_tile_kernel<C, PY, MINB, ADD><<<(unsigned)blocks, 256, 0, stream>>>(x, H, W, c_rt, out_h, out_w, TX, TY, total, w, bias, out);
  ERNET_LAUNCH_CHECK("acff_dw_tile_kernel");
  return ERNET_OK;
}

// Channel counts of the three classifier architectures (Squeeze_ErNET 16/64/96/128, Squeeze_RedConv 8/64/48/64, ErNET
// 16..128) and of the detector's add-fusion blocks (128/256) are compiled in; with `generic` any other C runs the
// run-time-C instantiation, without it -1 is returned (the concat caller falls back to the shared-memory kernel).
template <int PY = 4, int MINB = 2, bool ADD = false>
inline int launch_acff_dw_tile(const float* x, int batch, int H, int W, int C, int out_h, int out_w,
                               const float* w, const float* bias, float* out, cudaStream_t stream, bool generic = false) {
  switch (C) {
    case 8:   return launch_acff_dw_tile_c<8, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 16:  return launch_acff_dw_tile_c<16, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 32:  return launch_acff_dw_tile_c<32, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 48:  return launch_acff_dw_tile_c<48, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 64:  return launch_acff_dw_tile_c<64, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 96:  return launch_acff_dw_tile_c<96, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 128: return launch_acff_dw_tile_c<128, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    case 256: return launch_acff_dw_tile_c<256, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
    default:
      if (generic) return launch_acff_dw_tile_c<0, PY, MINB, ADD>(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
      return -1;
  }
}

// ---------------------------------------------------------------------------------- evaluation bookkeeping
// argmax over the class scores of each image + confusion-matrix update, cm[target][prediction] += 1 (what
// evaluate-classification-metrics.py:81-87 does with output.argmax(dim=1) and torchmetrics' ConfusionMatrix).
// Ties resolve to the lowest class index; a NaN score wins (first one), exactly like torch.argmax, so a broken forward
// shows up in the confusion matrix instead of being masked.  Targets outside [0, nc) are counted in *bad.
__global__ void confusion_update_kernel(const float* __restrict__ scores, const long long* __restrict__ targets, int batch,
                                        int nc, long long* __restrict__ cm, long long* __restrict__ pred_out,
                                        unsigned long long* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const float* s = scores + (size_t)i * nc;
  int best = 0;
  float bv = s[0];
  for (int k = 1; k < nc; ++k) {
    const float v = s[k];
    if (v > bv || (v != v && bv == bv)) { bv = v; best = k; }      // a NaN beats every number, the first NaN stays: torch.argmax
  }
  if (pred_out) pred_out[i] = best;
  if (targets && cm) {
    const long long t = targets[i];
    if (t < 0 || t >= nc) { if (bad) atomicAdd(bad, 1ull); return; }
    atomicAdd(reinterpret_cast<unsigned long long*>(cm) + t * nc + best, 1ull);
  }
}

inline void dw_pick_tile(int C, int esize, int out_h, int out_w, int& th, int& tw, size_t& smem) {
  // largest square-ish tile whose halo fits in ~96 KB (two CTAs per SM); the width is a whole number of the strips
  // a thread computes (4 pixels at fp32, 2 at 16 bit), columns past the image are skipped in the kernel
  const int px = esize == 4 ? 4 : 2;
  static const int cand[] = {24, 22, 16, 15, 12, 11, 10, 8, 6, 5, 4, 3, 2, 1};
  for (int t : cand) {
    int tth = t < out_h ? t : out_h, ttw = t < out_w ? t : out_w;
    ttw = (ttw + px - 1) / px * px;
    size_t s = (size_t)(tth + 6) * (ttw + 6) * C * esize + 30 * (size_t)C * sizeof(float);
    if (s <= 96 * 1024) { th = tth; tw = ttw; smem = s; return; }
  }
  th = 1; tw = px;
  smem = (size_t)7 * (px + 6) * C * esize + 30 * (size_t)C * sizeof(float);
}

template <typename T>
inline int launch_acff_dw(const T* x, int batch, int H, int W, int C, int out_h, int out_w,
                          const float* w, const float* bias, T* out, cudaStream_t stream) {
  if (C % Vec16<T>::NV) return fail(ERNET_ERR_INVALID_ARG, "depthwise: C=%d not a multiple of %d", C, Vec16<T>::NV);
  if (out_h > H - 2 || out_w > W - 2 || out_h < 1 || out_w < 1)
    return fail(ERNET_ERR_INVALID_ARG, "depthwise: bad output size %dx%d for input %dx%d", out_h, out_w, H, W);
  if constexpr (std::is_same<T, float>::value) {
    if (g_dw_fp32_form == 2) {
      const int rc = launch_acff_dw_tma(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
      if (rc != -1) return rc;
    }
    if (g_dw_fp32_form) {
      const int rc = launch_acff_dw_tile(x, batch, H, W, C, out_h, out_w, w, bias, out, stream);
      if (rc != -1) return rc;
    }
  }
  int th, tw;
  size_t smem;
  dw_pick_tile(C, (int)sizeof(T), out_h, out_w, th, tw, smem);
  dim3 grid((out_w + tw - 1) / tw, (out_h + th - 1) / th, batch);
  acff_dw_kernel<T><<<grid, 256, smem, stream>>>(x, H, W, C, out_h, out_w, th, tw, w, bias, out);
  ERNET_LAUNCH_CHECK("acff_dw_kernel");
  return ERNET_OK;
}

// ---------------------------------------------------------------------------------- pointwise (1x1)
// Register-tiled FFMA GEMM: rows = pixels, K = input channels, N = output channels.  CTA tile 128 x BN,
// thread tile 8 x 4.  With POOL the 128 rows are 32 pooled pixels x their 2x2 quad, so the max-pool is an
// in-thread max over 4 consecutive rows and the un-pooled map never reaches memory.
struct PwGeom {
  int H, W;        // spatial size of the input map (rows of A = B*H*W, compact NHWC)
  int Hp, Wp;      // pooled size (floor(H/2), floor(W/2)) when pooling
  int m_total;     // number of output rows (pooled pixels when pooling)
};

template <typename T, int BN, bool POOL>
__global__ void __launch_bounds__(16 * BN / 4)
pointwise_kernel(const T* __restrict__ A, int K, int N, const float* __restrict__ Wt /*[K][N]*/,
                 const float* __restrict__ bias, const float* __restrict__ bn_s, const float* __restrict__ bn_t,
                 int leaky, PwGeom g, T* __restrict__ out) {
  constexpr int BM = 128, BK = 8, NT = 16 * BN / 4;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  __shared__ long long rowoff[BM];   // element offset of each A row, -1 = out of range

  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * BN;
  const int out_rows_per_tile = POOL ? BM / 4 : BM;
  const int o0 = blockIdx.x * out_rows_per_tile;

  for (int r = tid; r < BM; r += NT) {
    long long off = -1;
    if (POOL) {
      const int p = o0 + (r >> 2), q = r & 3;
      if (p < g.m_total) {
        const int b = p / (g.Hp * g.Wp);
        const int rem = p - b * g.Hp * g.Wp;
        const int py = rem / g.Wp, px = rem - py * g.Wp;
        off = (((long long)b * g.H + (2 * py + (q >> 1))) * g.W + (2 * px + (q & 1))) * K;
      }
    } else {
      const int m = o0 + r;
      if (m < g.m_total) off = (long long)m * K;
    }
    rowoff[r] = off;
  }
  __syncthreads();

  const int tr = tid / (BN / 4), tc = tid % (BN / 4);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // Software pipeline over K: the global loads of tile k+1 are issued into registers before the FMAs of tile k (the
  // kernel used to load, sync, compute, sync - every 8-deep K step paid a full global-memory latency with nothing else
  // in flight).  Per thread: at most 2 A vectors (BM*2 = 256 of them per tile) and 1 B vector (BK*BN/4 <= NT).
  constexpr int A_ITEMS = (BM * 2 + NT - 1) / NT;
  static_assert(BK * BN / 4 <= NT && A_ITEMS <= 2, "prefetch registers sized for these tiles");
  float pa[A_ITEMS][4];
  float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int k0) {
#pragma unroll
    for (int u = 0; u < A_ITEMS; ++u) {
      const int i = tid + u * NT;
      pa[u][0] = pa[u][1] = pa[u][2] = pa[u][3] = 0.f;
      if (i < BM * 2) {
        const int r = i >> 1, kh = (i & 1) * 4;
        const long long off = rowoff[r];
        if (off >= 0) {
          const T* p = A + off + k0 + kh;
          if constexpr (sizeof(T) == 4) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(p));
            pa[u][0] = f.x; pa[u][1] = f.y; pa[u][2] = f.z; pa[u][3] = f.w;
          } else {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
            const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) pa[u][j] = to_f32<T>(e[j]);
          }
        }
      }
    }
    if (tid < BK * BN / 4) {
      const int kk = tid / (BN / 4), c4 = tid % (BN / 4);
      pb = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)(k0 + kk) * N + n0) + c4);
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int u = 0; u < A_ITEMS; ++u) {
      const int i = tid + u * NT;
      if (i < BM * 2) {
        const int r = i >> 1, kh = (i & 1) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) As[kh + j][r] = pa[u][j];
      }
    }
    if (tid < BK * BN / 4) {
      const int kk = tid / (BN / 4), c4 = tid % (BN / 4);
      *reinterpret_cast<float4*>(&Bs[kk][c4 * 4]) = pb;
    }
  };
  gload(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    sstore();
    __syncthreads();
    if (k0 + BK < K) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tr * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][tr * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tc * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  float bsv[4], ssv[4], tsv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tc * 4 + j;
    bsv[j] = bias ? bias[n] : 0.f;
    ssv[j] = bn_s ? bn_s[n] : 1.f;
    tsv[j] = bn_t ? bn_t[n] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bsv[j];
      if (leaky) v = leaky_relu(v);
      acc[i][j] = fmaf(v, ssv[j], tsv[j]);
    }

  auto store4 = [&](int orow, const float (&v)[4]) {
    T* o = out + (size_t)orow * N + n0 + tc * 4;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      T e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = from_f32<T>(v[j]);
      *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(e);
    }
  };
  if (POOL) {
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      const int p = o0 + tr * 2 + qd;
      if (p < g.m_total) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = fmaxf(fmaxf(acc[qd * 4][j], acc[qd * 4 + 1][j]), fmaxf(acc[qd * 4 + 2][j], acc[qd * 4 + 3][j]));
        store4(p, v);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = o0 + tr * 8 + i;
      if (m < g.m_total) store4(m, acc[i]);
    }
  }
}

template <typename T>
inline int launch_pointwise(const T* a, int batch, int H, int W, int K, int N, const float* w, const float* bias,
                            const float* bn_s, const float* bn_t, int leaky, int pool, T* out, cudaStream_t stream) {
  if (K % 8) return fail(ERNET_ERR_INVALID_ARG, "pointwise: K=%d not a multiple of 8", K);
  PwGeom g;
  g.H = H; g.W = W; g.Hp = H / 2; g.Wp = W / 2;
  g.m_total = pool ? batch * g.Hp * g.Wp : batch * H * W;
  if (g.m_total <= 0) return fail(ERNET_ERR_INVALID_ARG, "pointwise: empty output");
  const int rows_per_tile = pool ? 32 : 128;
  const int bn = (N % 64 == 0) ? 64 : ((N % 48 == 0) ? 48 : ((N % 32 == 0) ? 32 : 0));
  if (!bn) return fail(ERNET_ERR_INVALID_ARG, "pointwise: N=%d must be a multiple of 32 or 48", N);
  dim3 grid((g.m_total + rows_per_tile - 1) / rows_per_tile, N / bn);
#define ERNET_PW(BN_)                                                                                        \
  do {                                                                                                       \
    if (pool) pointwise_kernel<T, BN_, true><<<grid, 16 * BN_ / 4, 0, stream>>>(a, K, N, w, bias, bn_s, bn_t, leaky, g, out); \
    else      pointwise_kernel<T, BN_, false><<<grid, 16 * BN_ / 4, 0, stream>>>(a, K, N, w, bias, bn_s, bn_t, leaky, g, out); \
  } while (0)
  if (bn == 64) ERNET_PW(64); else if (bn == 48) ERNET_PW(48); else ERNET_PW(32);
#undef ERNET_PW
  ERNET_LAUNCH_CHECK("pointwise_kernel");
  return ERNET_OK;
}

// ---------------------------------------------------------------------------------- head
// At 140x140 input acff4 yields a 4x4 map; AvgPool2d(5,1,1) on a 4x4 map gives a 2x2 map whose four
// entries are all sum(4x4)/25, so conv2 -> pool -> view -> fc is logits = W_eff . sum_pixels(a4) + b_fc
// with W_eff[5][256] precomputed by the packer.  One CTA of 256 threads per image.
template <typename T>
__global__ void __launch_bounds__(256)
head_kernel(const T* __restrict__ a4 /*(B,4,4,256)*/, const float* __restrict__ w_eff /*[5][256]*/,
            const float* __restrict__ b_fc, float* __restrict__ probs, float* __restrict__ logits) {
  __shared__ float red[8][5];
  const int b = blockIdx.x, c = threadIdx.x;
  const T* p = a4 + (size_t)b * 16 * 256 + c;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += to_f32<T>(p[i * 256]);
  float part[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) part[j] = s * w_eff[j * 256 + c];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
  if ((c & 31) == 0)
#pragma unroll
    for (int j = 0; j < 5; ++j) red[c >> 5][j] = part[j];
  __syncthreads();
  if (c == 0) {
    float z[5], m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float t = b_fc[j];
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += red[wv][j];
      z[j] = t;
      m = fmaxf(m, t);
    }
    float e[5], sum = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { e[j] = expf(z[j] - m); sum += e[j]; }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      probs[b * 5 + j] = e[j] / sum;
      if (logits) logits[b * 5 + j] = z[j];
    }
  }
}

// ---------------------------------------------------------------------------------- debug tap
template <typename T>
__global__ void tap_nhwc_to_nchw_f32(const T* __restrict__ src, int C, int HW, long long total, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW index
  if (i >= total) return;
  const int p = (int)(i % HW);
  const long long bc = i / HW;
  const int c = (int)(bc % C);
  const long long b = bc / C;
  dst[i] = to_f32<T>(src[(b * HW + p) * C + c]);
}

// ---------------------------------------------------------------------------------- ErNET head
// conv2 (1x1, 256 -> 5) -> AvgPool2d(5, stride 1) on the 7x7 map -> view(-1, 45) -> fc -> softmax (model/ernet.py:20-23,35-44):
// everything before the softmax is linear, so the packer collapses it to a position-dependent W_eff[5][49][256]
// (pack.py) and logits = b + sum_{pixel, channel} W_eff * acff6.  One CTA of 256 threads (= channels) per image.
template <typename T>
__global__ void __launch_bounds__(256)
ernet_head_kernel(const T* __restrict__ a6 /*(B,7,7,256)*/, const float* __restrict__ w_eff /*[5][49][256]*/,
                  const float* __restrict__ b_fc, float* __restrict__ probs, float* __restrict__ logits) {
  __shared__ float red[8][5];
  const int b = blockIdx.x, c = threadIdx.x;
  const T* p = a6 + (size_t)b * 49 * 256 + c;
  float part[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < 49; ++i) {
    const float v = to_f32<T>(p[i * 256]);
#pragma unroll
    for (int j = 0; j < 5; ++j) part[j] = fmaf(v, __ldg(w_eff + (j * 49 + i) * 256 + c), part[j]);
  }
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
  if ((c & 31) == 0)
#pragma unroll
    for (int j = 0; j < 5; ++j) red[c >> 5][j] = part[j];
  __syncthreads();
  if (c == 0) {
    float z[5], m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      float t = b_fc[j];
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += red[wv][j];
      z[j] = t;
      m = fmaxf(m, t);
    }
    float e[5], sum = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) { e[j] = expf(z[j] - m); sum += e[j]; }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      probs[(size_t)b * 5 + j] = e[j] / sum;
      if (logits) logits[(size_t)b * 5 + j] = z[j];
    }
  }
}

}  // namespace ernet
