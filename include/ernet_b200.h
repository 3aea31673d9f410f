/*
 * ernet_b200.h — C ABI of the B200-native Squeeze-ErNet / Squeeze-ErNet-RedConv engine.
 *
 * This is the drop-in boundary for ONE hot path of qazi0/real-time-disaster-management:
 * the `model(x)` call of the AIDER classifier and the eval transform in front of it.
 * Every entry point cites the reference interface it replaces; paths are relative to
 * code/disaster_detection/ in the reference repository.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all device pointers are raw CUDA device addresses;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *     enqueued asynchronously on it (the reference times `model(data)` followed by
 *     `torch.cuda.synchronize()`, evaluate-classification-metrics.py:76-79);
 *   - the caller owns inputs, outputs and the workspace; the library owns only the packed
 *     weights it copied at ernet_load_packed() and small per-frame-size ingest tables;
 *   - every function returning int returns ERNET_OK (0) or a negative ernet_status; the
 *     message of the last failure on the calling thread is ernet_last_error();
 *   - a handle is bound to one device and is not safe for concurrent calls; multi-GPU use
 *     is one handle (and one process or thread) per device, batches sharded by the caller.
 *
 * Environment variables read by ernet_create() (initial values of a new handle; study / cross-check switches, the
 * defaults are the fast paths): ERNET_TAIL_TILES=0 block 1 on 16x8 tiles only (no tail unit), ERNET_FP32_TC=0 fp32
 * engine with the FFMA 1x1 kernels instead of the split-TF32 tcgen05 GEMMs, ERNET_FUSE_INGEST=1, ERNET_HOST_GATHER=1
 * (+ ERNET_GATHER_CTAS=n), ERNET_PAIR_TAPS=0, ERNET_PAIR_BLOCK1=1, ERNET_DUAL_COPY=1, ERNET_TRIM_COLUMNS=1, ERNET_EPI_SUSPEND=0
 * (polling instead of suspending epilogue waits), ERNET_SMALL_BATCH_UNITS=0 (two-tile pair-kernel units at every batch size), ERNET_NVTX=1 (an NVTX range per stage launch).  Results are bit-identical across all of them except ERNET_FP32_TC (fp32 rounding
 * level) - tests/test_gpu_parity.py cross-checks each.
 */
#ifndef ERNET_B200_H_
#define ERNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERNET_ABI_VERSION 1

typedef struct ernet_handle ernet_handle;

typedef enum ernet_status {
  ERNET_OK = 0,
  ERNET_ERR_INVALID_ARG = -1,   /* bad enum, null pointer, non-positive batch ... (-> ValueError)   */
  ERNET_ERR_BAD_SHAPE = -2,     /* anything that is not 3x140x140 for model(x) (-> ValueError)       */
  ERNET_ERR_NOT_LOADED = -3,    /* forward before ernet_load_packed                                   */
  ERNET_ERR_BAD_BLOB = -4,      /* packed weights: magic/version/arch/precision/table mismatch        */
  ERNET_ERR_WORKSPACE = -5,     /* workspace null or smaller than ernet_workspace_bytes()             */
  ERNET_ERR_CUDA = -6,          /* a CUDA runtime call failed; message has cudaGetErrorString         */
  ERNET_ERR_UNSUPPORTED = -7    /* valid request this build does not implement                        */
} ernet_status;

/* `--model {squeeze-ernet,squeeze-redconv}` of aider-predict.py:124-138 / load_model :22-32.        */
typedef enum ernet_arch { ERNET_ARCH_SQUEEZE = 0, ERNET_ARCH_REDCONV = 1,
                          ERNET_ARCH_ERNET = 2 /* baseline ErNET, (B,3,240,240) inputs, layer-wise CUDA-core kernels */ } ernet_arch;

/* `--quant {fp32,fp16,int8}` of build_tensorrt_model.py:320-330 (+ bf16, new).                      */
typedef enum ernet_precision {
  ERNET_PREC_FP32 = 0, ERNET_PREC_FP16 = 1, ERNET_PREC_BF16 = 2, ERNET_PREC_INT8 = 3
} ernet_precision;

typedef enum ernet_dtype { ERNET_F32 = 0, ERNET_F16 = 1, ERNET_BF16 = 2, ERNET_U8 = 3 } ernet_dtype;
typedef enum ernet_layout { ERNET_NCHW = 0, ERNET_NHWC = 1 } ernet_layout;
/* Frames from OpenCV are BGR and the reference converts them (aider-predict.py:62).                 */
typedef enum ernet_channel_order { ERNET_RGB = 0, ERNET_BGR = 1 } ernet_channel_order;

/* Intermediate tensors that ernet_debug_tap() can read back (test-only introspection).              */
typedef enum ernet_tap {
  ERNET_TAP_INGEST = 0,  /* (B,3,140,140) output of the eval transform                               */
  ERNET_TAP_STEM = 1,    /* conv1 [+conv_red1]            (B,16|8,69,69)                             */
  ERNET_TAP_POOL1 = 2,   /* pool1(acff1(.))               (B,64,33,33)                               */
  ERNET_TAP_POOL2 = 3,   /* pool2([conv_red2](acff2(.)))  (B,96|48,15,15)                            */
  ERNET_TAP_POOL3 = 4,   /* [conv_red3](pool3(acff3(.)))  (B,128|64,6,6)                             */
  ERNET_TAP_ACFF4 = 5,   /* acff4(.)                      (B,256,4,4)                                */
  ERNET_TAP_COUNT = 6
} ernet_tap;

/* ---- lifetime ----------------------------------------------------------------------------------
 * Replaces the constructors `Squeeze_ErNET()` / `Squeeze_RedConv()` (model/squeeze_ernet.py:8-22,
 * model/squeeze_ernet_redconv.py:8-25) plus `.to(device)` (aider-predict.py:43).                    */
int ernet_create(ernet_handle** out, int arch, int precision, int device);
void ernet_destroy(ernet_handle* h);

/* Replaces `model.load_state_dict(...)` (aider-predict.py:35-41).  `blob` is the byte string the
 * host-side packer builds from the reference's 56/62-key state_dict; it is copied to the device,
 * the caller keeps ownership.  May be called again to swap weights.                                  */
int ernet_load_packed(ernet_handle* h, const void* blob, size_t bytes);

/* Scratch the caller must provide to the forward calls for `batch` images (no hidden cudaMalloc in the
 * forward path).  Large batches are processed in chunks of ernet_get_chunk() images, so the
 * size saturates.                                                                                    */
size_t ernet_workspace_bytes(const ernet_handle* h, int batch);
int ernet_set_chunk(ernet_handle* h, int images_per_chunk);
int ernet_get_chunk(const ernet_handle* h);

/* Kernel family.  AUTO (default): tcgen05 tensor-core block kernels where they exist (16-bit
 * Squeeze_ErNET), CUDA-core kernels otherwise (fp32: tensor cores top out at TF32, which cannot hold
 * the 1e-4 logit bound).  SIMT forces the CUDA-core kernels (used to cross-check the two on device).   */
typedef enum ernet_engine { ERNET_ENGINE_AUTO = 0, ERNET_ENGINE_SIMT = 1, ERNET_ENGINE_TC = 2 } ernet_engine;
int ernet_set_engine(ernet_handle* h, int engine);
int ernet_get_engine(const ernet_handle* h);   /* the family the next forward will use (SIMT or TC) */
/* Schedule of the tensor-core block kernels: 2 (default) = persistent CTAs fed by TMA box loads, with blocks 2 and 3
 * on CTA pairs (tcgen05 cta_group::2, tc_cblock.cuh); 1 = persistent single CTAs (tc_pblock.cuh); 0 = one image per
 * CTA (tc_block.cuh).  Same folded weights; 1 and 0 are bit-identical, 2 accumulates the K steps in another order. */
int ernet_set_persistent(ernet_handle* h, int on);
/* Frames path of the 16-bit tensor-core engines, 5-tap frames (240x240): 1 (default) = word-wide fused transform +
 * conv1 with ToTensor/Normalize folded into the conv weights (ingest_fast.cuh); 0 = the table-lookup kernel that is
 * bit-identical to ernet_ingest_u8 followed by ernet_forward.                                              */
int ernet_set_fast_ingest(ernet_handle* h, int on);
/* Frames path, tensor-core engines, 5-tap frames: 1 = the eval transform + conv1 run INSIDE block 1's persistent kernel
 * on dedicated helper warps, under its tcgen05 MMAs (tc_fblock.cuh; one launch less, bit-identical results); 0
 * (default) = a kernel of their own in front of block 1.  Measured at parity on B200, hence off
 * by default.  The environment variable ERNET_FUSE_INGEST=1 sets the initial value of new handles.            */
int ernet_set_fuse_ingest(ernet_handle* h, int on);
/* Host path (ernet_classify_frames_host*): 1 = when the caller's frame buffer is pinned host memory, a small kernel
 * pulls only the FOOTPRINT of the crop window (rows and columns the eval transform reads, 240x240: 79 % of the frame)
 * over PCIe with 16-byte loads instead of a copy-engine transfer of the row range (89 %); pageable or unaligned
 * buffers silently take the copy-engine path.  0 = always the copy engine.  `ctas` = CTAs of that kernel (0 keeps
 * the current value).  The environment variable ERNET_HOST_GATHER sets the initial value of new handles.        */
int ernet_set_host_gather(ernet_handle* h, int on, int ctas);
/* Fused kernels keep some intermediates on chip (acff4 inside the ACFF4+head kernel).  With debug taps
 * on they are also written to the workspace so that ernet_debug_tap() can read them (test use).      */
int ernet_set_debug_taps(ernet_handle* h, int on);

/* ---- the hot path ------------------------------------------------------------------------------
 * Replaces `output = model(data)` (aider-predict.py:76, evaluate-classification-metrics.py:77):
 * Squeeze_ErNET.forward (model/squeeze_ernet.py:24-46) / Squeeze_RedConv.forward
 * (model/squeeze_ernet_redconv.py:27-52).  `x` is (batch,3,140,140) in `x_layout` order with
 * element type `x_dtype` (fp32 / fp16 / bf16) on the handle's device.  Writes softmax
 * probabilities (batch,5) fp32 to `probs_out` and, if non-null, the pre-softmax `fc` output to
 * `logits_out` (the reference exposes no logits API; parity is judged on them).                     */
int ernet_forward(ernet_handle* h, const void* x, int x_dtype, int x_layout, int batch,
                  float* probs_out, float* logits_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Replaces `squeeze_transforms` = Resize(159) -> CenterCrop(140) -> ToTensor -> Normalize
 * (dataloaders/aider.py:412-426,431; applied at aider-predict.py:66) for a batch of equally sized
 * uint8 HWC frames already on the device.  Bit-exact with Pillow's 8-bit antialiased bilinear
 * resampler; output (batch,3,140,140) in `out_layout` order, element type `out_dtype`.               */
int ernet_ingest_u8(ernet_handle* h, const uint8_t* frames_hwc, int batch, int height, int width,
                    int channel_order, void* x_out, int out_dtype, int out_layout, void* stream);

/* Builds (and caches in the handle) the resampling tables for one frame size.  ernet_ingest_u8 /
 * ernet_forward_frames call it on first use of a size; call it up front to keep allocation out of
 * a timed or graph-captured region.                                                                  */
int ernet_prepare_ingest(ernet_handle* h, int height, int width);

/* Transform + model in one call on device-resident frames:
 * `model(transform(frame))` of aider-predict.py:57-76 for a whole batch.                             */
int ernet_forward_frames(ernet_handle* h, const uint8_t* frames_hwc, int batch, int height, int width,
                         int channel_order, float* probs_out, float* logits_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Bytes of one height x width frame that ernet_classify_frames_host really sends to the device: only the rows the
 * crop window of the eval transform reads (240x240: 213 of 240 rows) - with ernet_set_host_gather on, only the
 * 16-byte vectors that overlap the window's rows AND columns.  0 on error.                                 */
size_t ernet_host_copy_bytes_per_frame(ernet_handle* h, int height, int width);
/* Same with HOST buffers: `predict()` of aider-predict.py:47-86 / the loop body of
 * evaluate-classification-metrics.py:69-82 for a batch.  Host->device copies of the frames and the
 * device->host copy of the results happen inside, double-buffered against the kernels on internal
 * streams; returns after the results are in `probs_host` (and `logits_host`, nullable).  Staging
 * buffers are allocated on first use and kept in the handle.                                         */
int ernet_classify_frames_host(ernet_handle* h, const uint8_t* frames_hwc_host, int batch,
                               int height, int width, int channel_order,
                               float* probs_host, float* logits_host);
/* The same in two halves, so that a caller can keep ONE call in flight while it submits the next (the copy of batch
 * i+1 then runs under the kernels of batch i): submit enqueues copies, kernels and the read-back and returns a ticket
 * (0 / 1, alternating); wait blocks until that call's outputs are in probs_host / logits_host.  Host buffers must
 * stay valid (and should be pinned) until the wait.  At most two calls may be outstanding.                  */
int ernet_classify_frames_host_submit(ernet_handle* h, const uint8_t* frames_host, int batch, int height, int width,
                                      int channel_order, float* probs_host, float* logits_host, int* ticket);
int ernet_classify_frames_host_wait(ernet_handle* h, int ticket);

/* ---- building blocks exposed for unit tests and micro-benchmarks --------------------------------
 * ACFF depthwise trio (model/acff.py:25-30,46): x (batch,H,W,C) NHWC -> (batch,out_h,out_w,3C)
 * with out_h<=H-2, out_w<=W-2 (the floor-mode max-pool that follows never reads an odd last
 * row/column, model/squeeze_ernet.py:13).  `w` is [3][9][C] fp32, `b` is [3][C] fp32.               */
int ernet_acff_depthwise(const void* x, int dtype, int batch, int H, int W, int C, int out_h, int out_w,
                         const float* w, const float* b, void* out, void* stream);

/* fp32 form of the trio: 2 (default) = TMA-staged tiles (one box copy per 24x24 / 8x8 output tile, zero fill = the conv
 * padding) for C = 8 / 16 / 64, register-tile kernel otherwise; 1 = register-tile kernel (one channel x 4x4 output patch
 * per thread, tap weights in registers, no shared memory); 0 = the shared-memory halo kernel (what fp16/bf16 always use).
 * Process-wide; the three forms are bit-identical.  Returns the previous value.                        */
int ernet_set_depthwise_form(int form);

/* Depthwise stage of the ADD-fusion ACFF block of the detector half (victim_localization/yolov3/models.py:277-282,302:
 * out = conv1(x) + conv2(x) + conv3(x), the same three dilated depthwise 3x3 convolutions, summed instead of
 * concatenated): x (batch,H,W,C) NHWC -> (batch,out_h,out_w,C), out_h<=H-2, out_w<=W-2.  fp32 only (dtype must be
 * ERNET_F32), any C >= 1.  Followed by ernet_pointwise (fused_conv + LeakyReLU + BN, models.py:307-309) it is the whole
 * block.  `w` is [3][9][C] fp32, `b` is [3][C] fp32.                                                                */
int ernet_acff_add_depthwise(const void* x, int dtype, int batch, int H, int W, int C, int out_h, int out_w,
                             const float* w, const float* b, void* out, void* stream);

/* Evaluation bookkeeping on the device (evaluate-classification-metrics.py:81-87): prediction = argmax over the
 * `num_classes` scores of each image (lowest index on ties), cm[target][prediction] += 1.  `scores` (batch,num_classes)
 * fp32, `targets` (batch) int64, `cm` [num_classes][num_classes] int64 accumulated in place; `pred_out` (batch) int64
 * and `bad` (count of targets outside [0,num_classes)) are optional.  `targets` and `cm` may both be NULL when only
 * predictions are wanted.  All pointers are device pointers.                                                         */
int ernet_confusion_update(const float* scores, const long long* targets, int batch, int num_classes,
                           long long* cm, long long* pred_out, unsigned long long* bad, void* stream);

/* 1x1 convolution (model/acff.py:31) + bias [+ LeakyReLU(0.01)] [+ per-channel affine = eval BN,
 * acff.py:33-34] [+ 2x2/2 max-pool, squeeze_ernet.py:13].  a: (batch,H,W,K) NHWC; w: [K][N] fp32.   */
int ernet_pointwise(const void* a, int dtype, int batch, int H, int W, int K, int N,
                    const float* w, const float* bias, const float* bn_scale, const float* bn_shift,
                    int leaky, int pool, void* out, void* stream);

/* Copies intermediate tensor `tap` of the most recent forward chunk out of `workspace` as fp32
 * NCHW.  `out_elems` must equal batch*C*H*W of that tap.                                            */
int ernet_debug_tap(ernet_handle* h, int tap, const void* workspace, int batch,
                    float* out_nchw, size_t out_elems, void* stream);

/* Host-only (no GPU needed): the resampling tables ernet_ingest_u8 uses for one frame size, so their
 * bit-exactness against Pillow can be tested without a device.  meta = {new_h,new_w,top,left,ksy,ksx};
 * xmin/xlen/ymin/ylen hold 140 ints, kx/ky up to 140*64 ints, lut 256*3 floats (nullable).            */
int ernet_ingest_tables_host(int height, int width, int* meta, int* xmin, int* xlen, int* kx,
                             int* ymin, int* ylen, int* ky, float* lut);

/* Device-side watchdog record of the tensor-core kernels on the current device: out8[0] = number of
 * pipeline waits that timed out since the last reset (0 in a healthy run), out8[1..3] = tag / block /
 * aux of the first one.  A timed-out kernel terminates normally but its results are invalid.          */
int ernet_debug_device_status(unsigned int* out8, int reset);
/* Call after synchronising with work enqueued by ernet_forward / ernet_forward_frames (the reference's own
 * sync point is evaluate-classification-metrics.py:79, torch.cuda.synchronize()): ERNET_ERR_CUDA (and the record
 * is cleared) when a kernel's watchdog fired since the last check, else ERNET_OK.  Reads one word of mapped
 * host memory: no device round trip.  ernet_classify_frames_host_wait and every forward call check it too.   */
int ernet_check_watchdog(void);
/* sha256 of the sources this library was compiled from (set by build.py; "unknown" for a hand build).           */
const char* ernet_source_hash(void);
/* Study builds (-DERNET_TIMELINE) only: per-CTA clock64 stamps of the persistent block kernels, [3 kernels][148][32][8].
 * Returns ERNET_ERR_UNSUPPORTED in the normal build.                                                    */
int ernet_debug_timeline(unsigned long long* out, size_t count);
/* Study builds only: global-timer view of the forward chain, [8 kernels][first entry, first PDL-wait return, last exit,
 * last entry] in ns; reset != 0 re-arms the minima / maxima.                                               */
int ernet_debug_chain(unsigned long long* out32, int reset);

/* Per-stage device timing with CUDA events recorded on the launch stream around every kernel of the
 * forward path (bench.py's live roofline measurement).  Off by default; when on, each forward adds
 * two event records per stage.  ernet_profile_read() synchronises on the recorded events, sums the
 * elapsed milliseconds and launch counts per ernet_stage since the previous read and clears them.    */
typedef enum ernet_stage {
  ERNET_STAGE_INGEST = 0, ERNET_STAGE_STEM = 1,
  ERNET_STAGE_DW1 = 2, ERNET_STAGE_PW1 = 3, ERNET_STAGE_DW2 = 4, ERNET_STAGE_PW2 = 5, ERNET_STAGE_RED2 = 6,
  ERNET_STAGE_DW3 = 7, ERNET_STAGE_PW3 = 8, ERNET_STAGE_RED3 = 9, ERNET_STAGE_DW4 = 10, ERNET_STAGE_PW4 = 11,
  ERNET_STAGE_HEAD = 12,
  ERNET_STAGE_TC_BLOCK1 = 13, ERNET_STAGE_TC_BLOCK2 = 14, ERNET_STAGE_TC_BLOCK3 = 15, ERNET_STAGE_TC_BLOCK4 = 16,
  ERNET_STAGE_COUNT = 17
} ernet_stage;
int ernet_profile_enable(ernet_handle* h, int on);
int ernet_profile_read(ernet_handle* h, double* ms_by_stage, int* launches_by_stage, int n_stages);

/* Number of kernels ernet_forward_frames() launches for `batch` frames (bench.py's gpu_launches).   */
int ernet_launches_per_forward(const ernet_handle* h, int batch, int with_ingest);

const char* ernet_last_error(void);
int ernet_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif  /* ERNET_B200_H_ */
