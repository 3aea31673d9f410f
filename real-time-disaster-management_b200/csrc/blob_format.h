// Packed-weights blob shared by the host-side packer (pack.py) and the CUDA library.
//
//   header   : ernet_blob_header (32 bytes)
//   table    : n_entries x ernet_blob_entry (24 bytes each)
//   payload  : tensors, each 256-byte aligned relative to the start of the blob
//
// The packer derives every tensor in fp64 from the reference's state_dict
//  and stores it in the layout the kernels read.
#pragma once
#include <stdint.h>

#define ERNET_BLOB_MAGIC 0x424E5245u /* 'ERNB' */
#define ERNET_BLOB_VERSION 2u

struct ernet_blob_header {
  uint32_t magic, version, arch, precision, n_entries, reserved[3];
};
struct ernet_blob_entry {
  uint32_t id, dtype;  // dtype: ernet_dtype, or 16 = raw bytes
  uint64_t offset, nbytes;
};

// ---- tensor ids -----------------------------------------------------------------------------
// CUDA-core ("simt") path, all fp32:
#define ERNET_T_STEM_W 0        // [3][3][3][CS]  (ky,kx,cin,cout); RedConv: conv_red1 folded in, CS=8
#define ERNET_T_STEM_B 1        // [CS]           (zeros for squeeze-ernet)
#define ERNET_T_BLOCK_BASE 8    // + 8*k, k = 0..3 (acff1..acff4)
#define ERNET_T_DW_W 0          //   [3][9][C]    branch, tap (ky*3+kx), channel
#define ERNET_T_DW_B 1          //   [3][C]
#define ERNET_T_PW_W 2          //   [3C][N]      k = branch*C + c  (concat order, acff.py:46)
#define ERNET_T_PW_B 3          //   [N]
#define ERNET_T_BN_S 4          //   [N]  gamma / sqrt(var + eps)
#define ERNET_T_BN_T 5          //   [N]  beta - mean * scale
#define ERNET_T_RED2_W 40       // [96][48]   conv_red2, k-major
#define ERNET_T_RED2_B 41
#define ERNET_T_RED3_W 42       // [128][64]  conv_red3
#define ERNET_T_RED3_B 43
#define ERNET_T_HEAD_W 44       // [5][256]   conv2 o avgpool o fc collapsed
#define ERNET_T_HEAD_B 45       // [5]
// tensor-core path (16-bit / int8 operand images), see tc_*.cuh:
#define ERNET_T_TC_BASE 64      // + 4*k, k = 0..2 (acff1..acff3)
#define ERNET_T_TC_WIMG 0       //   [25 taps][C/8][N][8] 16-bit: folded depthwise x 1x1 weights (pack_tc.py)
#define ERNET_T_TC_BIAS 1       //   [N] fp32: b_f + W_f . cat(b_d)
#define ERNET_T_TC_DEQ 2        //   [N] fp32, int8 only: per-output-channel weight scale s_w[n] = max|W'_eff[n]|/127
#define ERNET_T_TC4_WIMG 76     // [3*C4/8][256][8] 16-bit: acff4.fused_conv.weight as the K-major UMMA B operand (tc_tail.cuh)
#define ERNET_T_TC_RED2_WIMG 77 // [1][12][64][8] 16-bit: conv_red2 (96 -> 48, N padded to 64) as a 1-tap block-kernel image
#define ERNET_T_TC_RED2_BIAS 78 // [64] fp32 (48 real)
#define ERNET_T_Q_SCALES 80     // [16+64+96] fp32, int8 only: real value of one int8 step of every channel of the stem /
                                // pool1 / pool2 tensors (per-channel equalisation of a per-tensor int8 scale; folded into
                                // the producer's epilogue and the consumer's weights, so the runtime tensor scale is 1)
#define ERNET_T_EBLOCK_BASE 88  // ErNET (arch 2) blocks 5 and 6: + 8*(k-4), same six tensors as ERNET_T_BLOCK_BASE
#define ERNET_T_EHEAD_W 104     // ErNET: [5][49][256] conv2 o AvgPool(5,1,0) o view o fc collapsed per pixel (pack.py)
#define ERNET_T_ETC_BASE 106    // ErNET tensor-core images: + 2*k, k = 0..5: [25][C/8][N][8] 16-bit, then [N] fp32 folded bias
#define ERNET_T_MAX 128
