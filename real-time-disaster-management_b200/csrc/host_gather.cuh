// Host path (ernet_classify_frames_host*): the frame bytes the eval transform reads, pulled over PCIe by a kernel.
//
// Resize(160) + CenterCrop(140) (dataloaders/aider.py:421-423) only reads the footprint of the crop window - for a 240x240
// frame rows 14..226 and columns 14..226, 79 % of the bytes.  The copy engine moves the ROW range at full PCIe rate (one
// strided copy, rows of a frame are contiguous) but a 3-D copy that also skips the columns runs at ~10 GB/s (639-byte
// rows, one descriptor each).  This kernel reads the footprint straight from the caller's PINNED host buffer (mapped into
// the device address space by unified addressing) with 16-byte loads - a few CTAs keep hundreds of KB in flight, enough to
// cover the PCIe round trip - and writes it to the same offsets of the device frame buffer, so the transform kernels see
// the layout they always see.  Bytes outside the footprint are never read by them (zero-weight taps multiply stale bytes
// of an integer buffer).
#pragma once
#include "common.cuh"

namespace ernet {

struct GatherGeom {
  unsigned long long f_img;   // bytes per frame
  int rowb;                   // bytes per row
  int row_lo, nrows;          // first footprint row, number of rows
  int xb0, xb1;               // byte range of the footprint columns within a row
  int vpr;                    // upper bound of 16-byte vectors per row segment
};

// 128 threads at <= 64 registers: 8 K registers per CTA, so a gather CTA fits NEXT TO the persistent compute CTA of the
// previous sub-chunk on the same SM (block kernels leave 10-13 K registers and no shared memory is needed here)
constexpr int kGatherThreads = 128;
constexpr int kGatherUnroll = 8;

// src and dst have the same 16-byte phase (both are 16-byte aligned); `total` = bytes of the whole source buffer.
__global__ void __launch_bounds__(kGatherThreads, 8)
host_gather_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_frames, unsigned long long total, const GatherGeom g) {
  const long long items = (long long)n_frames * g.nrows * g.vpr;
  const long long stride = (long long)gridDim.x * kGatherThreads;
  const unsigned long long vend = total & ~15ull;            // vectors may not run past the end of the caller's buffer
  for (long long i0 = (long long)blockIdx.x * kGatherThreads + threadIdx.x; i0 < items; i0 += stride * kGatherUnroll) {
    uint4 v[kGatherUnroll];
    uint32_t off[kGatherUnroll];                               // in 16-byte vectors (buffers up to 64 GB)
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const long long i = i0 + u * stride;
      off[u] = ~0u;
      if (i < items) {
        const long long seg = i / g.vpr;
        const int j = (int)(i - seg * g.vpr);
        const int f = (int)(seg / g.nrows), r = (int)(seg - (long long)f * g.nrows);
        const unsigned long long a = (unsigned long long)f * g.f_img + (unsigned long long)(g.row_lo + r) * g.rowb;
        const unsigned long long o = ((a + g.xb0) & ~15ull) + 16ull * j;
        if (o < a + g.xb1 && o + 16 <= vend) off[u] = (uint32_t)(o >> 4);
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u)
      if (off[u] != ~0u) v[u] = __ldcs(reinterpret_cast<const uint4*>(src) + off[u]);
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u)
      if (off[u] != ~0u) reinterpret_cast<uint4*>(dst)[off[u]] = v[u];
  }
  // the last (< 16) bytes of the buffer, if the footprint of the last frame reaches into them
  if (blockIdx.x == 0 && threadIdx.x < (total & 15ull)) {
    const unsigned long long o = vend + threadIdx.x;
    dst[o] = src[o];
  }
}

// bytes this kernel moves per frame (16-byte vectors that overlap the footprint; exact for 16-byte aligned rows)
inline size_t gather_bytes_per_frame(const GatherGeom& g) {
  size_t n = 0;
  for (int r = 0; r < g.nrows; ++r) {
    const unsigned long long a = (unsigned long long)(g.row_lo + r) * g.rowb;
    n += (size_t)((((a + g.xb1) + 15) & ~15ull) - ((a + g.xb0) & ~15ull));
  }
  return n;
}

}  // namespace ernet
