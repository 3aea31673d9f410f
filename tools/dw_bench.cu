// Micro-benchmark of the ACFF depthwise trio kernels (fp32): register-tile variants against the shared-memory
// kernel, on the four Squeeze-ErNet block shapes.  Prints algorithmic GB/s (input once + 3x output once) and
// checks every variant bitwise against the shared-memory kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I include -o tools/dw_bench tools/dw_bench.cu
#include <stdarg.h>
#include <type_traits>
#include <vector>
#include <cstdlib>
#include "../real-time-disaster-management_b200/csrc/dw_tma.cuh"
namespace ernet { thread_local char g_err[512]; }
using namespace ernet;

#define CK(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { printf("%s: %s\n", #e, cudaGetErrorString(_e)); return 1; } } while (0)

template <int PY, int MINB> static int tile(const float* x, int b, int H, int W, int C, int oh, int ow, const float* w, const float* bi, float* o, cudaStream_t s) { return launch_acff_dw_tile<PY, MINB, false>(x, b, H, W, C, oh, ow, w, bi, o, s); }
template <int MINB, int NT, int PXW> static int tma16(const float* x, int b, int H, int W, int C, int oh, int ow, const float* w, const float* bi, float* o, cudaStream_t s) {
  if (C != 16 || oh < 24) return -1;
  return launch_acff_dw_tma_c<16, false, MINB, NT, 24, 16, PXW>(x, b, H, W, oh, ow, w, bi, o, s);
}
typedef int (*launch_fn)(const float*, int, int, int, int, int, int, const float*, const float*, float*, cudaStream_t);
static int launch_smem(const float* x, int b, int H, int W, int C, int oh, int ow, const float* w, const float* bi, float* o, cudaStream_t s) {
  g_dw_fp32_form = 0; int rc = launch_acff_dw<float>(x, b, H, W, C, oh, ow, w, bi, o, s); g_dw_fp32_form = 1; return rc;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256;
  CK(cudaFuncSetAttribute(acff_dw_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  struct Shape { int H, C, oh; } shapes[] = {{69, 16, 67}, {33, 64, 31}, {69, 8, 67}, {119, 16, 117}, {69, 16, 66}, {15, 96, 13}, {6, 128, 4}};
  struct Var { const char* name; launch_fn fn; } vars[] = {
      {"smem", launch_smem},
      {"tile py4 occ2", tile<4, 2>}, {"tma (default)", launch_acff_dw_tma}, {"tma16 px4 nt192 o3", tma16<3, 192, 4>}};
  float* flush; const size_t flush_n = 160u << 20;   // 640 MB > L2
  CK(cudaMalloc(&flush, flush_n * 4));
  {   // what a pure write stream and a copy reach on this GPU (context for the 74 %-write depthwise traffic)
    float* src; CK(cudaMalloc(&src, flush_n * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k = 0; k < 2; ++k) {
      float ms = 0;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0, 0);
        if (k == 0) cudaMemsetAsync(flush, it, flush_n * 4, 0); else cudaMemcpyAsync(flush, src, flush_n * 4, cudaMemcpyDeviceToDevice, 0);
        cudaEventRecord(e1, 0); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("%s of %.0f MB: %.1f us = %.1f GB/s\n", k ? "copy (read+write bytes)" : "memset", flush_n * 4 / 1e6, ms * 1e3, (k ? 2.0 : 1.0) * flush_n * 4 / (ms * 1e-3) / 1e9);
    }
    cudaFree(src);
  }
  const int nshapes = getenv("DW_SHAPES") ? atoi(getenv("DW_SHAPES")) : 7;
  const int iters = getenv("DW_ITERS") ? atoi(getenv("DW_ITERS")) : 10;
  const char* only = getenv("DW_VARS");     // e.g. "01" = variants 0 and 1
  int si = 0;
  for (auto sh : shapes) {
    if (si++ >= nshapes) break;
    const int H = sh.H, C = sh.C, oh = sh.oh;
    const size_t nin = (size_t)B * H * H * C, nout = (size_t)B * oh * oh * 3 * C;
    std::vector<float> hx(nin), hw(27 * C), hb(3 * C);
    srand(H * 7 + C);
    for (auto& v : hx) v = (rand() % 2001 - 1000) / 500.f;
    for (auto& v : hw) v = (rand() % 2001 - 1000) / 2500.f;
    for (auto& v : hb) v = (rand() % 2001 - 1000) / 10000.f;
    float *x, *w, *bi, *o, *oref;
    CK(cudaMalloc(&x, nin * 4)); CK(cudaMalloc(&w, hw.size() * 4)); CK(cudaMalloc(&bi, hb.size() * 4));
    CK(cudaMalloc(&o, nout * 4)); CK(cudaMalloc(&oref, nout * 4));
    CK(cudaMemcpy(x, hx.data(), nin * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(bi, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> ref(nout), got(nout);
    const double bytes = (double)(nin + nout) * 4;
    printf("shape B=%d H=%d C=%d out=%d  algorithmic %.1f MB\n", B, H, C, oh, bytes / 1e6);
    for (auto& v : vars) {
      if (only && &v != &vars[0] && !strchr(only, '0' + (int)(&v - vars))) continue;
      float* dst = (&v == &vars[0]) ? oref : o;
      CK(cudaMemset(dst, 0xff, nout * 4));
      if (int rc_ = v.fn(x, B, H, H, C, oh, oh, w, bi, dst, 0)) { if (rc_ == -1) { printf("  %s: not served\n", v.name); continue; } printf("  %s: launch failed: %s\n", v.name, g_err); continue; }
      CK(cudaDeviceSynchronize());
      bool same = true;
      if (&v == &vars[0]) CK(cudaMemcpy(ref.data(), dst, nout * 4, cudaMemcpyDeviceToHost));
      else { CK(cudaMemcpy(got.data(), dst, nout * 4, cudaMemcpyDeviceToHost)); same = memcmp(got.data(), ref.data(), nout * 4) == 0; }
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e30f, sum = 0;
      for (int it = 0; it < iters + 2; ++it) {
        CK(cudaMemsetAsync(flush, it, flush_n * 4, 0));      // L2 flush: inputs come from HBM
        cudaEventRecord(e0, 0);
        v.fn(x, B, H, H, C, oh, oh, w, bi, dst, 0);
        cudaEventRecord(e1, 0);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
      }
      printf("  %-17s %8.1f us avg %8.1f us best  %7.1f GB/s avg  %s\n", v.name, sum / iters * 1e3, best * 1e3,
             bytes / (sum / iters * 1e-3) / 1e9, same ? "bit-identical" : "MISMATCH");
    }
    cudaFree(x); cudaFree(w); cudaFree(bi); cudaFree(o); cudaFree(oref);
  }
  return 0;
}
