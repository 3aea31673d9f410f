// C ABI of the B200-native Squeeze-ErNet engine (see include/ernet_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include <stdarg.h>

#include <algorithm>
#include <map>
#include <type_traits>
#include <new>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler injects the NVTX library

#include "common.cuh"
#include "ingest.cuh"
#include "simt_layers.cuh"
#include "ingest_fast.cuh"
#include "tc_block.cuh"
#include "tc_tail.cuh"
#include "tc_pblock.cuh"
#include "tc_cblock.cuh"
#include "tc_fblock.cuh"
#include "tc_pw32.cuh"
#include "dw_tma.cuh"
#include "host_gather.cuh"

namespace ernet {

thread_local char g_err[512] = "";

struct Tensor {
  const void* dev = nullptr;
  size_t nbytes = 0;
  int dtype = 0;
};

struct Plan {                 // byte offsets into the caller's workspace for one chunk
  size_t ingest, stem, cat1, p1, cat2, a2, p2, cat3, p3, r3, cat4, a4, total;
  size_t ecat, ea, eb;        // ErNET: concat buffer and two ping-pong activation buffers
  size_t e4, e5, e6;          // ErNET tensor-core path: outputs of blocks 4-6 (stem / p1 / p2 / p3 hold the first three)
  bool tc;                    // stem/p1/p2 are in the padded P8 layout of tc_block.cuh
};

}  // namespace ernet

using namespace ernet;

constexpr int kSyncImages = 65536;     // >= the largest chunk (ernet_set_chunk)

struct ernet_handle {
  int arch = 0, precision = 0, device = 0;
  int chunk = 1024;
  int engine = ERNET_ENGINE_AUTO;
  bool loaded = false;
  bool has_tc = false;          // blob carries the tensor-core operand images
  tc::EpiParams<64> epi1;       // host copies of the per-channel epilogue constants (kernel parameters)
  tc::EpiParams<96> epi2;
  tc::EpiParams<128> epi3;
  tc::EpiParams<64> epi_r2;     // RedConv conv_red2 (bias only)
  tc::EpiParams<128> ee4, ee5;  // ErNET blocks 4, 5 (blocks 1-3 reuse epi1-3)
  tc::EpiParams<256> ee6;       // ErNET block 6
  bool has_tail = false;        // blob carries the ACFF4+head tensor-core image
  bool debug_taps = false;      // keep intermediates the fused kernels would not write (acff4)
  int persistent = 2;           // block-kernel schedule: 0 one image per CTA, 1 persistent CTAs, 2 persistent + CTA pairs for blocks 2, 3
  int num_sms = 148;
  // fused transform + conv1 + block 1 (tc_fblock.cuh): per-image band counters, re-armed by the kernel itself
  uint32_t* d_sync = nullptr;        // [2][kSyncImages]: ready, taken
  bool fuse_ingest = false;          // ERNET_FUSE_INGEST=1: transform + conv1 under block 1 in one kernel (tc_fblock.cuh; off until it beats the two kernels)
  bool last_fused = false;           // the most recent frames chunk took the fused kernel (ernet_launches_per_forward)
  void* d_w1_pair = nullptr;         // block-1 weights regrouped for tap pairing ([13][2][64][16 B], tc_pblock.cuh)
  float* d_pw32[4] = {nullptr, nullptr, nullptr, nullptr};   // fp32 Squeeze_ErNET: (hi, lo) TF32 images of the four 1x1 weights (tc_pw32.cuh)
  bool fp32_tc = true;               // fp32 engine: blocks 1-4 1x1 convolutions as split-TF32 tcgen05 GEMMs (ERNET_FP32_TC=0: FFMA kernel)
  StemFrag* d_stem_frag = nullptr;   // folded conv1 in mma.sync fragment order (16-bit engines)
  StemFrag32* d_stem_frag32 = nullptr;   // fp32 engine: the same as (hi, lo) fp16 images (ingest_fast.cuh)
  bool dual_copy = false;       // host path: alternate two copy streams (ERNET_DUAL_COPY=1)
  bool trim_columns = false;    // host path: also skip the columns outside the crop footprint (ERNET_TRIM_COLUMNS=1).  Off:
                                // measured 78 K img/s against 281 K - a 3-D copy of 639-byte rows runs at ~10 GB/s
  bool small_batch_units = true;   // pair kernels: one-tile units for small batches (CBlock2S / CBlock3S; ERNET_SMALL_BATCH_UNITS=0: off)
  bool nvtx = false;            // ERNET_NVTX=1: an NVTX range per stage launch
  bool tail_tiles = true;       // block 1: output rows 64, 65 as one tail unit per image (PCfg::TAIL, tc_pblock.cuh; ERNET_TAIL_TILES=0: 16x8 tiles only)
  bool host_gather = false;     // host path: pinned frames are pulled by host_gather_kernel (footprint rows AND columns, host_gather.cuh)
  int gather_ctas = 32;
  bool pair_taps = true;        // two taps per MMA in block 1 when its input has one real chunk (ERNET_PAIR_TAPS=0 switches it off)
  bool pair_block1 = false;     // block 1 on the CTA-pair kernel as well (experiment switch: ERNET_PAIR_BLOCK1=0)
  bool fast_ingest = true;      // word-wide fused transform+conv1 with Normalize folded into conv1 (ingest_fast.cuh)
  tc::TailParams tail;
  void* d_blob = nullptr;
  size_t blob_bytes = 0;
  Tensor t[ERNET_T_MAX];
  std::map<std::pair<int, int>, IngestTables> ingest;
  // host-buffer path (ernet_classify_frames_host)
  cudaStream_t s_copy = nullptr, s_copy2 = nullptr, s_compute = nullptr;   // two copy streams: the copies of consecutive sub-chunks overlap
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_call[2] = {nullptr, nullptr};
  unsigned long long sub_it = 0, calls = 0;   // sub-chunks / calls submitted so far (buffer and ticket alternation)
  uint8_t* d_frames[2] = {nullptr, nullptr};
  size_t d_frames_bytes = 0;
  void* d_ws = nullptr;
  size_t d_ws_bytes = 0;
  float* d_res = nullptr;       // probs then logits
  size_t d_res_elems = 0;
  // optional per-stage CUDA-event timing (ernet_profile_*)
  bool profiling = false;
  struct ProfRec { int stage; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_pool;

  bool red() const { return arch == ERNET_ARCH_REDCONV; }
  bool ernet() const { return arch == ERNET_ARCH_ERNET; }
  int in_hw() const { return ernet() ? 240 : 140; }    // model input size (ErNET is the 240x240-native model)
  int cs() const { return red() ? 8 : 16; }            // stem output channels
  int c3() const { return red() ? 48 : 96; }           // acff3 input channels
  int c4() const { return red() ? 64 : 128; }          // acff4 input channels
  size_t esize() const { return precision == ERNET_PREC_FP32 ? 4 : 2; }
  // tensor-core path: 16-bit handles of both architectures, int8 Squeeze_ErNET
  bool use_tc() const {
    if (precision == ERNET_PREC_INT8) return has_tc;                // int8 exists only as tensor-core kernels
    if (engine == ERNET_ENGINE_SIMT) return false;
    return has_tc && (precision == ERNET_PREC_BF16 || precision == ERNET_PREC_FP16);
  }
  float q_scales[16 + 64 + 96];          // int8: per-channel int8 step of the stem / pool1 / pool2 tensors
  tc::StemInv stem_inv;
  const float* f(int id) const { return static_cast<const float*>(t[id].dev); }
  const float* blk(int k, int what) const { return f(ERNET_T_BLOCK_BASE + 8 * k + what); }
};

namespace ernet {

static Plan make_plan(const ernet_handle* h, int n) {
  const size_t e = h->esize(), N = (size_t)n;
  Plan p{};
  size_t o = 0;
  auto take = [&](size_t elems) { size_t r = o; o += align_up(elems * e, 256); return r; };
  p.tc = h->use_tc();
  if (h->ernet() && p.tc) {                      // P8 images between the tensor-core kernels (+ slack for the last pair-unit)
    p.ingest = take(N * 240 * 240 * 3);          // frames path: the transformed tensor, NHWC
    p.stem = take(N * 2 * 122 * 122 * 8);        // (B,2,122,122,8)
    p.p1 = take(N * 8 * 61 * 61 * 8);            // (B,8,61,61,8)
    p.p2 = take(N * 12 * 31 * 31 * 8);           // (B,12,31,31,8)
    p.p3 = take(N * 16 * 16 * 16 * 8);           // (B,16,16,16,8)
    p.e4 = take(N * 16 * 14 * 14 * 8);           // (B,16,14,14,8)
    p.e5 = take(N * 16 * 12 * 12 * 8);           // (B,16,12,12,8)
    p.e6 = take(N * 7 * 7 * 256);                // NHWC
    p.total = o;
    return p;
  }
  if (h->ernet()) {                              // layer-wise path, NHWC: sizes of the largest user of each buffer
    p.ingest = take(N * 240 * 240 * 3);
    p.stem = take(N * 119 * 119 * 16);
    p.ecat = take(N * 117 * 117 * 48);           // ACFF1 concat (ACFF2: 56*56*192 is smaller)
    p.ea = take(N * 58 * 58 * 64);               // pool1 (also pool3 13*13*128, acff5 9*9*128)
    p.eb = take(N * 28 * 28 * 96);               // pool2 (also acff4 11*11*128, acff6 7*7*256)
    p.total = o;
    return p;
  }
  p.ingest = take(N * 140 * 140 * 3);
  if (p.tc && h->precision == ERNET_PREC_INT8 && h->red()) {   // int8 Squeeze_RedConv (sizes in 2-byte units)
    p.stem = take(N * 2 * 72 * 72 * 8);        // P8 fp16 (B,2,72,72,8): 8 real channels in chunk 0 (ACFF1 runs in fp16)
    p.p1 = take(N * 4 * 36 * 36 * 8);          // P16 (B,4,36,36,16)
    p.a2 = take(N * 12 * 33 * 33 * 8);         // P8 fp16 (B,12,33,33,8): un-pooled acff2 output (conv_red2 input)
    p.p2 = take(N * 4 * 18 * 18 * 8 + 4 * 18 * 18 * 8);     // P16 (B,4,18,18,16): 48 real + 16 zero channels, + slack
    p.p3 = take(N * 6 * 6 * 128);              // NHWC fp16
    p.r3 = take(N * 6 * 6 * 64);               // NHWC fp16
    p.cat4 = take(N * 4 * 4 * 3 * h->c4());
    p.a4 = take(N * 4 * 4 * 256);
    p.total = o;
    return p;
  }
  if (p.tc && h->precision == ERNET_PREC_INT8) {   // P16 images: 16 int8 channels per 16-byte chunk (sizes in 2-byte units)
    p.stem = take(N * 2 * 72 * 72 * 8);        // (B,2,72,72,16) int8
    p.p1 = take(N * 4 * 36 * 36 * 8);          // (B,4,36,36,16)
    p.p2 = take(N * 6 * 18 * 18 * 8 + 6 * 18 * 18 * 8);     // (B,6,18,18,16) + slack
    p.p3 = take(N * 6 * 6 * 128);              // NHWC fp16
    p.cat4 = take(N * 4 * 4 * 3 * h->c4());
    p.a4 = take(N * 4 * 4 * 256);
    p.total = o;
    return p;
  }
  if (p.tc && h->red()) {
    p.stem = take(N * 2 * 72 * 72 * 8);        // P8 (B,2,72,72,8): 8 real + 8 zero channels
    p.p1 = take(N * 8 * 36 * 36 * 8);          // P8 (B,8,36,36,8)
    p.a2 = take(N * 12 * 33 * 33 * 8);         // P8 (B,12,33,33,8): un-pooled acff2 output (conv_red2 input)
    p.p2 = take(N * 6 * 18 * 18 * 8 + 6 * 18 * 18 * 8);     // P8 (B,6,18,18,8) + one image of slack
    p.p3 = take(N * 6 * 6 * 128);              // NHWC
    p.r3 = take(N * 6 * 6 * 64);               // NHWC
    p.cat4 = take(N * 4 * 4 * 3 * h->c4());
    p.a4 = take(N * 4 * 4 * 256);
    p.total = o;
    return p;
  }
  if (p.tc) {
    p.stem = take(N * 2 * 72 * 72 * 8);        // P8 (B,2,72,72,8)
    p.p1 = take(N * 8 * 36 * 36 * 8);          // P8 (B,8,36,36,8)
    p.p2 = take(N * 12 * 18 * 18 * 8 + 12 * 18 * 18 * 8);   // P8 (B,12,18,18,8) + one image of slack (2 images per CTA)
    p.p3 = take(N * 6 * 6 * 128);              // NHWC
    p.cat4 = take(N * 4 * 4 * 3 * h->c4());
    p.a4 = take(N * 4 * 4 * 256);
    p.total = o;
    return p;
  }
  p.stem = take(N * 69 * 69 * h->cs());
  p.cat1 = take(N * 66 * 66 * 3 * h->cs());
  p.p1 = take(N * 33 * 33 * 64);
  p.cat2 = take(N * 30 * 30 * 192);
  p.a2 = h->red() ? take(N * 30 * 30 * 96) : 0;
  p.p2 = take(N * 15 * 15 * h->c3());
  p.cat3 = take(N * 12 * 12 * 3 * h->c3());
  p.p3 = take(N * 6 * 6 * 128);
  p.r3 = h->red() ? take(N * 6 * 6 * 64) : 0;
  p.cat4 = take(N * 4 * 4 * 3 * h->c4());
  p.a4 = take(N * 4 * 4 * 256);
  p.total = o;
  return p;
}

static int check_tensor(const ernet_handle* h, int id, size_t elems) {
  if (!h->t[id].dev) return fail(ERNET_ERR_BAD_BLOB, "packed weights: tensor id %d missing", id);
  if (h->t[id].nbytes != elems * sizeof(float))
    return fail(ERNET_ERR_BAD_BLOB, "packed weights: tensor id %d has %zu bytes, expected %zu", id, h->t[id].nbytes,
                elems * sizeof(float));
  return ERNET_OK;
}

static const int kErnetCin[6] = {16, 64, 96, 128, 128, 128}, kErnetCout[6] = {64, 96, 128, 128, 128, 256};
static inline int ernet_block_base(int k) { return k < 4 ? ERNET_T_BLOCK_BASE + 8 * k : ERNET_T_EBLOCK_BASE + 8 * (k - 4); }

static int validate_simt_tensors(const ernet_handle* h) {
  int rc;
  if (h->ernet()) {
    if ((rc = check_tensor(h, ERNET_T_STEM_W, 27 * 16))) return rc;
    if ((rc = check_tensor(h, ERNET_T_STEM_B, 16))) return rc;
    for (int k = 0; k < 6; ++k) {
      const int base = ernet_block_base(k), ci = kErnetCin[k], co = kErnetCout[k];
      if ((rc = check_tensor(h, base + ERNET_T_DW_W, 27 * ci))) return rc;
      if ((rc = check_tensor(h, base + ERNET_T_DW_B, 3 * ci))) return rc;
      if ((rc = check_tensor(h, base + ERNET_T_PW_W, (size_t)3 * ci * co))) return rc;
      if ((rc = check_tensor(h, base + ERNET_T_PW_B, co))) return rc;
      if ((rc = check_tensor(h, base + ERNET_T_BN_S, co))) return rc;
      if ((rc = check_tensor(h, base + ERNET_T_BN_T, co))) return rc;
    }
    if ((rc = check_tensor(h, ERNET_T_EHEAD_W, 5 * 49 * 256))) return rc;
    return check_tensor(h, ERNET_T_HEAD_B, 5);
  }
  const int cs = h->cs();
  if ((rc = check_tensor(h, ERNET_T_STEM_W, 27 * cs))) return rc;
  if ((rc = check_tensor(h, ERNET_T_STEM_B, cs))) return rc;
  const int cin[4] = {cs, 64, h->c3(), h->c4()}, cout[4] = {64, 96, 128, 256};
  for (int k = 0; k < 4; ++k) {
    const int base = ERNET_T_BLOCK_BASE + 8 * k;
    if ((rc = check_tensor(h, base + ERNET_T_DW_W, 27 * cin[k]))) return rc;
    if ((rc = check_tensor(h, base + ERNET_T_DW_B, 3 * cin[k]))) return rc;
    if ((rc = check_tensor(h, base + ERNET_T_PW_W, (size_t)3 * cin[k] * cout[k]))) return rc;
    if ((rc = check_tensor(h, base + ERNET_T_PW_B, cout[k]))) return rc;
    if ((rc = check_tensor(h, base + ERNET_T_BN_S, cout[k]))) return rc;
    if ((rc = check_tensor(h, base + ERNET_T_BN_T, cout[k]))) return rc;
  }
  if (h->red()) {
    if ((rc = check_tensor(h, ERNET_T_RED2_W, 96 * 48))) return rc;
    if ((rc = check_tensor(h, ERNET_T_RED2_B, 48))) return rc;
    if ((rc = check_tensor(h, ERNET_T_RED3_W, 128 * 64))) return rc;
    if ((rc = check_tensor(h, ERNET_T_RED3_B, 64))) return rc;
  }
  if ((rc = check_tensor(h, ERNET_T_HEAD_W, 5 * 256))) return rc;
  if ((rc = check_tensor(h, ERNET_T_HEAD_B, 5))) return rc;
  return ERNET_OK;
}

template <typename TI, typename T>
static int launch_stem(const ernet_handle* h, const TI* x, long long sb, long long sc, long long sy, long long sx,
                       T* out, int n, cudaStream_t s) {
  const int total = n * 69 * 69;
  const int grid = (total + 127) / 128;
  if (h->red())
    stem_kernel<TI, T, 8><<<grid, 128, 0, s>>>(x, sb, sc, sy, sx, h->f(ERNET_T_STEM_W), h->f(ERNET_T_STEM_B), out, total);
  else
    stem_kernel<TI, T, 16><<<grid, 128, 0, s>>>(x, sb, sc, sy, sx, h->f(ERNET_T_STEM_W), h->f(ERNET_T_STEM_B), out, total);
  ERNET_LAUNCH_CHECK("stem_kernel");
  return ERNET_OK;
}

// Brackets one stage with CUDA events on the launch stream when profiling is on, and with an NVTX range named after the
// stage when ERNET_NVTX=1 (profiler timelines: SURVEY section 5).
static const char* const kStageNames[ERNET_STAGE_COUNT] = {"ingest", "stem", "dw1", "pw1", "dw2", "pw2", "red2", "dw3", "pw3", "red3",
                                                           "dw4", "pw4", "head", "tc_block1", "tc_block2", "tc_block3", "tc_block4"};
struct StageTimer {
  ernet_handle* h; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr; int stage; bool range = false;
  StageTimer(ernet_handle* h_, int stage_, cudaStream_t s_) : h(h_), s(s_), stage(stage_) {
    if (h->nvtx && stage >= 0 && stage < ERNET_STAGE_COUNT) { nvtxRangePushA(kStageNames[stage]); range = true; }
    if (!h->profiling) return;
    auto get = [&]() { cudaEvent_t e = nullptr;
      if (!h->prof_pool.empty()) { e = h->prof_pool.back(); h->prof_pool.pop_back(); } else cudaEventCreate(&e);
      return e; };
    a = get(); b = get();
    cudaEventRecord(a, s);
  }
  ~StageTimer() {
    if (range) nvtxRangePop();
    if (!a) return;
    cudaEventRecord(b, s);
    h->prof.push_back({stage, a, b});
  }
};
#define ERNET_STAGE(id, call)                                   \
  do { StageTimer _t(h, id, s); if ((rc = (call))) return rc; } while (0)

// ACFF4 + head: one fused tensor-core kernel when the blob has its weight image (16-bit / int8 handles),
// else depthwise + pointwise + head CUDA-core kernels.
template <typename T>
static int run_tail(ernet_handle* h, const T* in4, T* cat4, T* a4, int n, float* probs, float* logits, cudaStream_t s) {
  int rc;
  const int c4 = h->c4();
  if (h->has_tail && h->engine != ERNET_ENGINE_SIMT && sizeof(T) == 2) {
    constexpr bool BF16 = std::is_same<T, __nv_bfloat16>::value;
    StageTimer _t(h, ERNET_STAGE_TC_BLOCK4, s);
    void* dbg = h->debug_taps ? a4 : nullptr;
    if (c4 == 128) rc = tc::launch_acff4_head<tc::TailCfg128>(BF16, in4, h->blk(3, ERNET_T_DW_W), h->blk(3, ERNET_T_DW_B), h->t[ERNET_T_TC4_WIMG].dev, h->tail, probs, logits, dbg, n, s);
    else           rc = tc::launch_acff4_head<tc::TailCfg64>(BF16, in4, h->blk(3, ERNET_T_DW_W), h->blk(3, ERNET_T_DW_B), h->t[ERNET_T_TC4_WIMG].dev, h->tail, probs, logits, dbg, n, s);
    return rc;
  }
  ERNET_STAGE(ERNET_STAGE_DW4, launch_acff_dw<T>(in4, n, 6, 6, c4, 4, 4, h->blk(3, ERNET_T_DW_W), h->blk(3, ERNET_T_DW_B), cat4, s));
  bool tf = false;
  if constexpr (std::is_same<T, float>::value) tf = h->fp32_tc && !h->red() && h->d_pw32[3];
  if (tf) {
    if constexpr (std::is_same<T, float>::value)
      ERNET_STAGE(ERNET_STAGE_PW4, (tc::launch_pw32<tc::FPw4>(cat4, h->d_pw32[3], h->blk(3, ERNET_T_PW_B), h->blk(3, ERNET_T_BN_S),
                                    h->blk(3, ERNET_T_BN_T), a4, n, h->num_sms, s)));
  } else
  ERNET_STAGE(ERNET_STAGE_PW4, launch_pointwise<T>(cat4, n, 4, 4, 3 * c4, 256, h->blk(3, ERNET_T_PW_W), h->blk(3, ERNET_T_PW_B),
                                h->blk(3, ERNET_T_BN_S), h->blk(3, ERNET_T_BN_T), 1, 0, a4, s));
  {
    StageTimer _t(h, ERNET_STAGE_HEAD, s);
    head_kernel<T><<<n, 256, 0, s>>>(a4, h->f(ERNET_T_HEAD_W), h->f(ERNET_T_HEAD_B), probs, logits);
    ERNET_LAUNCH_CHECK("head_kernel");
  }
  return ERNET_OK;
}

// One chunk of n images through the layer-wise CUDA-core pipeline.
template <typename T>
static int run_chunk_simt(ernet_handle* h, const void* x, int x_dtype, int x_layout, const uint8_t* frames,
                          const IngestTables* tab, int order, int n, float* probs, float* logits, char* ws,
                          cudaStream_t s) {
  const Plan p = make_plan(h, n);
  auto buf = [&](size_t off) { return reinterpret_cast<T*>(ws + off); };
  int rc;
  if (frames && tab->fs_max_in_rows > 0 && !h->debug_taps) {
    // transform + conv1 fused (the transformed tensor never reaches HBM)
    StemQ q{};
    FastGeom fg;
    bool fast = false;
    if constexpr (std::is_same<T, float>::value) fast = h->fast_ingest && h->d_stem_frag32 && fast5_geometry(*tab, frames, fg);
    if (fast) {
      // fp32 engine, 5-tap frames: word-wide transform, conv1 on mma.sync with (hi, lo) fp16 weight images (ingest_fast.cuh)
      if constexpr (std::is_same<T, float>::value) {
        if (h->red()) ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem5<float, 8, FS_NHWC>(*tab, fg, frames, n, order == ERNET_BGR, &h->d_stem_frag32->base, q, false, buf(p.stem), s)));
        else          ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem5<float, 16, FS_NHWC>(*tab, fg, frames, n, order == ERNET_BGR, &h->d_stem_frag32->base, q, false, buf(p.stem), s)));
      }
    } else
    if (h->red()) ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem<T, 8, FS_NHWC>(*tab, frames, n, order == ERNET_BGR, h->f(ERNET_T_STEM_W), h->f(ERNET_T_STEM_B), q, buf(p.stem), s)));
    else          ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem<T, 16, FS_NHWC>(*tab, frames, n, order == ERNET_BGR, h->f(ERNET_T_STEM_W), h->f(ERNET_T_STEM_B), q, buf(p.stem), s)));
  } else if (frames) {
    ERNET_STAGE(ERNET_STAGE_INGEST, launch_ingest<T>(*tab, frames, n, order == ERNET_BGR, buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, s));
    ERNET_STAGE(ERNET_STAGE_STEM, (launch_stem<T, T>(h, buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, buf(p.stem), n, s)));
  } else {
    long long sb = 3LL * 140 * 140, sc, sy, sx;
    if (x_layout == ERNET_NCHW) { sc = 140 * 140; sy = 140; sx = 1; }
    else                        { sc = 1; sy = 140 * 3; sx = 3; }
    StageTimer _t(h, ERNET_STAGE_STEM, s);
    if (x_dtype == ERNET_F32) rc = launch_stem<float, T>(h, static_cast<const float*>(x), sb, sc, sy, sx, buf(p.stem), n, s);
    else if (x_dtype == ERNET_F16) rc = launch_stem<__half, T>(h, static_cast<const __half*>(x), sb, sc, sy, sx, buf(p.stem), n, s);
    else rc = launch_stem<__nv_bfloat16, T>(h, static_cast<const __nv_bfloat16*>(x), sb, sc, sy, sx, buf(p.stem), n, s);
    if (rc) return rc;
  }
  const int cs = h->cs(), c3 = h->c3(), c4 = h->c4();
  // acff1 + pool1
  ERNET_STAGE(ERNET_STAGE_DW1, launch_acff_dw<T>(buf(p.stem), n, 69, 69, cs, 66, 66, h->blk(0, ERNET_T_DW_W), h->blk(0, ERNET_T_DW_B), buf(p.cat1), s));
  // fp32 Squeeze_ErNET: the three big 1x1 convolutions run on the tensor cores in split-TF32 form (tc_pw32.cuh)
  bool tf = false;
  if constexpr (std::is_same<T, float>::value) tf = h->fp32_tc && !h->red() && h->d_pw32[0] && h->d_pw32[1] && h->d_pw32[2];
  if (tf) {
    if constexpr (std::is_same<T, float>::value)
      ERNET_STAGE(ERNET_STAGE_PW1, (tc::launch_pw32<tc::FPw1>(buf(p.cat1), h->d_pw32[0], h->blk(0, ERNET_T_PW_B), h->blk(0, ERNET_T_BN_S),
                                    h->blk(0, ERNET_T_BN_T), buf(p.p1), n, h->num_sms, s)));
  } else
  ERNET_STAGE(ERNET_STAGE_PW1, launch_pointwise<T>(buf(p.cat1), n, 66, 66, 3 * cs, 64, h->blk(0, ERNET_T_PW_W), h->blk(0, ERNET_T_PW_B),
                                h->blk(0, ERNET_T_BN_S), h->blk(0, ERNET_T_BN_T), 1, 1, buf(p.p1), s));
  // acff2 [+conv_red2] + pool2
  ERNET_STAGE(ERNET_STAGE_DW2, launch_acff_dw<T>(buf(p.p1), n, 33, 33, 64, 30, 30, h->blk(1, ERNET_T_DW_W), h->blk(1, ERNET_T_DW_B), buf(p.cat2), s));
  if (h->red()) {
    ERNET_STAGE(ERNET_STAGE_PW2, launch_pointwise<T>(buf(p.cat2), n, 30, 30, 192, 96, h->blk(1, ERNET_T_PW_W), h->blk(1, ERNET_T_PW_B),
                                  h->blk(1, ERNET_T_BN_S), h->blk(1, ERNET_T_BN_T), 1, 0, buf(p.a2), s));
    ERNET_STAGE(ERNET_STAGE_RED2, launch_pointwise<T>(buf(p.a2), n, 30, 30, 96, 48, h->f(ERNET_T_RED2_W), h->f(ERNET_T_RED2_B), nullptr, nullptr,
                                  0, 1, buf(p.p2), s));
  } else if (tf) {
    if constexpr (std::is_same<T, float>::value)
      ERNET_STAGE(ERNET_STAGE_PW2, (tc::launch_pw32<tc::FPw2>(buf(p.cat2), h->d_pw32[1], h->blk(1, ERNET_T_PW_B), h->blk(1, ERNET_T_BN_S),
                                    h->blk(1, ERNET_T_BN_T), buf(p.p2), n, h->num_sms, s)));
  } else {
    ERNET_STAGE(ERNET_STAGE_PW2, launch_pointwise<T>(buf(p.cat2), n, 30, 30, 192, 96, h->blk(1, ERNET_T_PW_W), h->blk(1, ERNET_T_PW_B),
                                  h->blk(1, ERNET_T_BN_S), h->blk(1, ERNET_T_BN_T), 1, 1, buf(p.p2), s));
  }
  // acff3 + pool3 [+conv_red3]
  ERNET_STAGE(ERNET_STAGE_DW3, launch_acff_dw<T>(buf(p.p2), n, 15, 15, c3, 12, 12, h->blk(2, ERNET_T_DW_W), h->blk(2, ERNET_T_DW_B), buf(p.cat3), s));
  if (tf) {
    if constexpr (std::is_same<T, float>::value)
      ERNET_STAGE(ERNET_STAGE_PW3, (tc::launch_pw32<tc::FPw3>(buf(p.cat3), h->d_pw32[2], h->blk(2, ERNET_T_PW_B), h->blk(2, ERNET_T_BN_S),
                                    h->blk(2, ERNET_T_BN_T), buf(p.p3), n, h->num_sms, s)));
  } else
  ERNET_STAGE(ERNET_STAGE_PW3, launch_pointwise<T>(buf(p.cat3), n, 12, 12, 3 * c3, 128, h->blk(2, ERNET_T_PW_W), h->blk(2, ERNET_T_PW_B),
                                h->blk(2, ERNET_T_BN_S), h->blk(2, ERNET_T_BN_T), 1, 1, buf(p.p3), s));
  const T* in4 = buf(p.p3);
  if (h->red()) {
    ERNET_STAGE(ERNET_STAGE_RED3, launch_pointwise<T>(buf(p.p3), n, 6, 6, 128, 64, h->f(ERNET_T_RED3_W), h->f(ERNET_T_RED3_B), nullptr, nullptr,
                                  0, 0, buf(p.r3), s));
    in4 = buf(p.r3);
  }
  // acff4 + head
  (void)c4;
  return run_tail<T>(h, in4, buf(p.cat4), buf(p.a4), n, probs, logits, s);
}

// One chunk through the tensor-core pipeline: ingest -> stem (P8/P16) -> blocks 1-3 on tcgen05 -> block 4 + head.
// T is the 16-bit element type of the non-quantised tensors (int8 engine: fp16).
template <typename T, int KIND>
static int run_chunk_tc(ernet_handle* h, const void* x, int x_dtype, int x_layout, const uint8_t* frames,
                        const IngestTables* tab, int order, int n, float* probs, float* logits, char* ws,
                        cudaStream_t s) {
  const Plan p = make_plan(h, n);
  auto buf = [&](size_t off) { return reinterpret_cast<T*>(ws + off); };
  auto u16 = [&](size_t off) { return reinterpret_cast<uint16_t*>(ws + off); };
  int rc;
  const int total = n * 72 * 72, grid = (total + 127) / 128;
  const float* sw = h->f(ERNET_T_STEM_W);
  const float* sb_ = h->f(ERNET_T_STEM_B);
  const tc::StemInv& s_inv = h->stem_inv;
  if (h->red()) {
    if constexpr (KIND != tc::KIND_I8) {
      // ---- Squeeze_RedConv: conv1+conv_red1 -> P8 (8 real channels), blocks 1-3 with conv_red2 as a 1-tap instance
      bool fused1 = false;
      if (frames && tab->fs_max_in_rows > 0 && !h->debug_taps) {
        StemQ q{};
        FastGeom fg;
        if (h->fast_ingest && h->d_stem_frag && fast5_geometry(*tab, frames, fg)) {
          if (h->fuse_ingest && h->persistent == 2 && h->d_w1_pair && h->pair_taps && tc::fused_fits<tc::PBlock1P>(fg) && n <= kSyncImages) {
            fused1 = h->last_fused = true;     // transform + conv1 run under block 1 in ONE kernel (tc_fblock.cuh)
            ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_ingest_block1<tc::PBlock1P, KIND, tc::OUT_P8, T, 8, FS_P8>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, false,
                        u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), h->num_sms, h->d_sync, h->d_sync + kSyncImages, s)));
          } else {
            ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem5<T, 8>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, !(h->persistent == 2 && h->d_w1_pair && h->pair_taps), u16(p.stem), s)));
          }
        } else {
          ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem<T, 8, FS_P8>(*tab, frames, n, order == ERNET_BGR, sw, sb_, q, u16(p.stem), s)));
        }
      } else if (frames) {
        ERNET_STAGE(ERNET_STAGE_INGEST, launch_ingest<T>(*tab, frames, n, order == ERNET_BGR, buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, s));
        StageTimer _t(h, ERNET_STAGE_STEM, s);
        tc::stem_p8_kernel<T, 8, KIND><<<grid, 128, 0, s>>>(buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, sw, sb_, u16(p.stem), total, s_inv);
        ERNET_LAUNCH_CHECK("stem_p8_kernel");
      } else {
        long long sb = 3LL * 140 * 140, sc, sy, sx;
        if (x_layout == ERNET_NCHW) { sc = 140 * 140; sy = 140; sx = 1; }
        else                        { sc = 1; sy = 140 * 3; sx = 3; }
        StageTimer _t(h, ERNET_STAGE_STEM, s);
        if (x_dtype == ERNET_F32) tc::stem_p8_kernel<float, 8, KIND><<<grid, 128, 0, s>>>(static_cast<const float*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
        else if (x_dtype == ERNET_F16) tc::stem_p8_kernel<__half, 8, KIND><<<grid, 128, 0, s>>>(static_cast<const __half*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
        else tc::stem_p8_kernel<__nv_bfloat16, 8, KIND><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
        ERNET_LAUNCH_CHECK("stem_p8_kernel");
      }
      auto wimgr = [&](int k) { return h->t[ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_WIMG].dev; };
      if (h->persistent == 2) {
        if (fused1) {
        } else if (h->d_w1_pair && h->pair_taps) {
          {
            if (h->tail_tiles) ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1PT, KIND, tc::OUT_P8>(u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), n, h->num_sms, s)));
            else ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1P, KIND, tc::OUT_P8>(u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), n, h->num_sms, s)));
          }
        } else {
          ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1, KIND, tc::OUT_P8>(u16(p.stem), wimgr(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
        }
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_cblock<tc::CBlock2R, KIND, tc::OUT_P8>(u16(p.p1), wimgr(1), h->epi2, u16(p.a2), n, h->num_sms, s)));
        ERNET_STAGE(ERNET_STAGE_RED2, (tc::launch_acff_pblock<tc::PRed2R, KIND, tc::OUT_P8>(u16(p.a2), h->t[ERNET_T_TC_RED2_WIMG].dev, h->epi_r2, u16(p.p2), n, h->num_sms, s)));
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_cblock<tc::CBlock3R, KIND, tc::OUT_NHWC>(u16(p.p2), wimgr(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
      } else {
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_block<tc::CfgBlock1, KIND, tc::OUT_P8>(u16(p.stem), wimgr(0), h->epi1, u16(p.p1), n, s)));
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_block<tc::CfgBlock2R, KIND, tc::OUT_P8>(u16(p.p1), wimgr(1), h->epi2, u16(p.a2), n, s)));
        ERNET_STAGE(ERNET_STAGE_RED2, (tc::launch_acff_block<tc::CfgRed2R, KIND, tc::OUT_P8>(u16(p.a2), h->t[ERNET_T_TC_RED2_WIMG].dev, h->epi_r2, u16(p.p2), n, s)));
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_block<tc::CfgBlock3R, KIND, tc::OUT_NHWC>(u16(p.p2), wimgr(2), h->epi3, u16(p.p3), n, s)));
      }
      ERNET_STAGE(ERNET_STAGE_RED3, launch_pointwise<T>(buf(p.p3), n, 6, 6, 128, 64, h->f(ERNET_T_RED3_W), h->f(ERNET_T_RED3_B), nullptr, nullptr, 0, 0, buf(p.r3), s));
      return run_tail<T>(h, buf(p.r3), buf(p.cat4), buf(p.a4), n, probs, logits, s);
    } else {
      // ---- int8 Squeeze_RedConv (persistent kernels only): conv1+conv_red1 -> fp16 P8 (8 real channels), ACFF1 in fp16 with
      // tap pairing (13 MMAs per tile: what int8 would issue too) writing the int8 pool1 tensor, ACFF2 int8 -> fp16
      // un-pooled, conv_red2 as a 16-bit 1-tap instance writing the int8 pool2 tensor, ACFF3 int8 -> fp16 NHWC, conv_red3 +
      // ACFF4 + head in fp16 (first conv and head in >= fp16)
      bool fused1 = false;
      StemQ q{};
      const bool paired = h->d_w1_pair && h->pair_taps;
      auto wimgq = [&](int k) { return h->t[ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_WIMG].dev; };
      if (frames && tab->fs_max_in_rows > 0 && !h->debug_taps) {
        FastGeom fg;
        if (h->fast_ingest && h->d_stem_frag && fast5_geometry(*tab, frames, fg)) {
          if (h->fuse_ingest && paired && tc::fused_fits<tc::PBlock1P>(fg) && n <= kSyncImages) {
            fused1 = h->last_fused = true;
            ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_ingest_block1<tc::PBlock1P, tc::KIND_F16, tc::OUT_P16, T, 8, FS_P8>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, false,
                        u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), h->num_sms, h->d_sync, h->d_sync + kSyncImages, s)));
          } else {
            ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem5<T, 8>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, !paired, u16(p.stem), s)));
          }
        } else {
          ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem<T, 8, FS_P8>(*tab, frames, n, order == ERNET_BGR, sw, sb_, q, u16(p.stem), s)));
        }
      } else if (frames) {
        ERNET_STAGE(ERNET_STAGE_INGEST, launch_ingest<T>(*tab, frames, n, order == ERNET_BGR, buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, s));
        StageTimer _t(h, ERNET_STAGE_STEM, s);
        tc::stem_p8_kernel<T, 8, tc::KIND_F16><<<grid, 128, 0, s>>>(buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, sw, sb_, u16(p.stem), total, s_inv);
        ERNET_LAUNCH_CHECK("stem_p8_kernel");
      } else {
        long long sb = 3LL * 140 * 140, sc, sy, sx;
        if (x_layout == ERNET_NCHW) { sc = 140 * 140; sy = 140; sx = 1; }
        else                        { sc = 1; sy = 140 * 3; sx = 3; }
        StageTimer _t(h, ERNET_STAGE_STEM, s);
        if (x_dtype == ERNET_F32) tc::stem_p8_kernel<float, 8, tc::KIND_F16><<<grid, 128, 0, s>>>(static_cast<const float*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
        else if (x_dtype == ERNET_F16) tc::stem_p8_kernel<__half, 8, tc::KIND_F16><<<grid, 128, 0, s>>>(static_cast<const __half*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
        else tc::stem_p8_kernel<__nv_bfloat16, 8, tc::KIND_F16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
        ERNET_LAUNCH_CHECK("stem_p8_kernel");
      }
      if (fused1) {
      } else if (paired) {
        {
          if (h->tail_tiles) ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1PT, tc::KIND_F16, tc::OUT_P16>(u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), n, h->num_sms, s)));
          else ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1P, tc::KIND_F16, tc::OUT_P16>(u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), n, h->num_sms, s)));
        }
      } else {
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1, tc::KIND_F16, tc::OUT_P16>(u16(p.stem), wimgq(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
      }
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_cblock<tc::CBlock2RQ, tc::KIND_I8, tc::OUT_P8>(u16(p.p1), wimgq(1), h->epi2, u16(p.a2), n, h->num_sms, s)));
      ERNET_STAGE(ERNET_STAGE_RED2, (tc::launch_acff_pblock<tc::PRed2RQ, tc::KIND_F16, tc::OUT_P16>(u16(p.a2), h->t[ERNET_T_TC_RED2_WIMG].dev, h->epi_r2, u16(p.p2), n, h->num_sms, s)));
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_cblock<tc::CBlock3RQ, tc::KIND_I8, tc::OUT_NHWC>(u16(p.p2), wimgq(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
      ERNET_STAGE(ERNET_STAGE_RED3, launch_pointwise<T>(buf(p.p3), n, 6, 6, 128, 64, h->f(ERNET_T_RED3_W), h->f(ERNET_T_RED3_B), nullptr, nullptr, 0, 0, buf(p.r3), s));
      return run_tail<T>(h, buf(p.r3), buf(p.cat4), buf(p.a4), n, probs, logits, s);
    }
  }
  // SK = element kind of the stem tensor, ST = its type
  bool fused1 = false;
  auto run_stem = [&](auto st_tag, auto sk_tag) -> int {
    using ST = decltype(st_tag);
    constexpr int SK = decltype(sk_tag)::value;
    if (frames && tab->fs_max_in_rows > 0 && !h->debug_taps) {
      StemQ q{};
      for (int i = 0; i < 16; ++i) q.inv[i] = s_inv.v[i];
      constexpr int FSOUT = SK == tc::KIND_I8 ? FS_P16 : FS_P8;
      FastGeom fg;
      if (h->fast_ingest && h->d_stem_frag && fast5_geometry(*tab, frames, fg)) {
        const bool zc1 = !(SK == tc::KIND_I8 && h->persistent == 2 && h->d_w1_pair && h->pair_taps);
        const bool can_fuse = h->fuse_ingest && h->persistent >= 1 && !(h->persistent == 2 && h->pair_block1) && n <= kSyncImages;
        if (can_fuse && SK == tc::KIND_I8 && h->persistent == 2 && !zc1 && tc::fused_fits<tc::PBlock1P>(fg)) {
          fused1 = h->last_fused = true;
          if constexpr (SK == tc::KIND_I8)
            ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_ingest_block1<tc::PBlock1P, tc::KIND_I8, tc::OUT_P16, ST, 16, FS_P16>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, false,
                        u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), h->num_sms, h->d_sync, h->d_sync + kSyncImages, s)));
        } else if (can_fuse && SK != tc::KIND_I8 && tc::fused_fits<tc::PBlock1>(fg)) {
          fused1 = h->last_fused = true;     // transform + conv1 run under block 1 in ONE kernel (tc_fblock.cuh)
          if constexpr (SK != tc::KIND_I8)
            ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_ingest_block1<tc::PBlock1, SK, tc::OUT_P8, ST, 16, FS_P8>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, true,
                        u16(p.stem), h->t[ERNET_T_TC_BASE + ERNET_T_TC_WIMG].dev, h->epi1, u16(p.p1), h->num_sms, h->d_sync, h->d_sync + kSyncImages, s)));
        } else {
          ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem5<ST, 16, FSOUT>(*tab, fg, frames, n, order == ERNET_BGR, h->d_stem_frag, q, zc1, u16(p.stem), s)));
        }
      } else {
        ERNET_STAGE(ERNET_STAGE_INGEST, (launch_ingest_stem<ST, 16, FSOUT>(*tab, frames, n, order == ERNET_BGR, sw, sb_, q, u16(p.stem), s)));
      }
    } else if (frames) {
      ERNET_STAGE(ERNET_STAGE_INGEST, launch_ingest<T>(*tab, frames, n, order == ERNET_BGR, buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, s));
      StageTimer _t(h, ERNET_STAGE_STEM, s);
      tc::stem_p8_kernel<T, 16, SK><<<grid, 128, 0, s>>>(buf(p.ingest), 140LL * 140 * 3, 1, 140 * 3, 3, sw, sb_, u16(p.stem), total, s_inv);
      ERNET_LAUNCH_CHECK("stem_p8_kernel");
    } else {
      long long sb = 3LL * 140 * 140, sc, sy, sx;
      if (x_layout == ERNET_NCHW) { sc = 140 * 140; sy = 140; sx = 1; }
      else                        { sc = 1; sy = 140 * 3; sx = 3; }
      StageTimer _t(h, ERNET_STAGE_STEM, s);
      if (x_dtype == ERNET_F32) tc::stem_p8_kernel<float, 16, SK><<<grid, 128, 0, s>>>(static_cast<const float*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
      else if (x_dtype == ERNET_F16) tc::stem_p8_kernel<__half, 16, SK><<<grid, 128, 0, s>>>(static_cast<const __half*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
      else tc::stem_p8_kernel<__nv_bfloat16, 16, SK><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), sb, sc, sy, sx, sw, sb_, u16(p.stem), total, s_inv);
      ERNET_LAUNCH_CHECK("stem_p8_kernel");
    }
    return ERNET_OK;
  };
  rc = run_stem(T{}, std::integral_constant<int, KIND>{});
  if (rc) return rc;
  auto wimg = [&](int k) { return h->t[ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_WIMG].dev; };
  if (KIND == tc::KIND_I8 && h->persistent == 2) {
    if (fused1) {
    } else if (h->d_w1_pair && h->pair_taps) {
      {
        if (h->tail_tiles) ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1PT, tc::KIND_I8, tc::OUT_P16>(u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), n, h->num_sms, s)));
        else ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1P, tc::KIND_I8, tc::OUT_P16>(u16(p.stem), h->d_w1_pair, h->epi1, u16(p.p1), n, h->num_sms, s)));
      }
    } else {
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1, tc::KIND_I8, tc::OUT_P16>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
    }
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_cblock<tc::CBlock2Q, tc::KIND_I8, tc::OUT_P16>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, h->num_sms, s)));
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_cblock<tc::CBlock3Q, tc::KIND_I8, tc::OUT_NHWC>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
  } else if (KIND == tc::KIND_I8) {
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_block<tc::CfgBlock1Q, tc::KIND_I8, tc::OUT_P16>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, s)));
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_block<tc::CfgBlock2Q, tc::KIND_I8, tc::OUT_P16>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, s)));
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_block<tc::CfgBlock3Q, tc::KIND_I8, tc::OUT_NHWC>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, s)));
  } else if (h->persistent) {
    constexpr int K16 = KIND == tc::KIND_I8 ? tc::KIND_F16 : KIND;
    if (fused1) {
    } else if (h->persistent == 2 && h->pair_block1) {
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_cblock<tc::CBlock1, K16, tc::OUT_P8>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
    } else if (h->tail_tiles) {
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1T, K16, tc::OUT_P8>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
    } else {
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::PBlock1, K16, tc::OUT_P8>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
    }
    if (h->persistent == 2) {
      if (h->small_batch_units && n <= tc::kSmallBatch2)
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_cblock<tc::CBlock2S, K16, tc::OUT_P8>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, h->num_sms, s)));
      else
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_cblock<tc::CBlock2, K16, tc::OUT_P8>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, h->num_sms, s)));
      if (h->small_batch_units && n <= tc::kSmallBatch3)
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_cblock<tc::CBlock3S, K16, tc::OUT_NHWC>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
      else
        ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_cblock<tc::CBlock3, K16, tc::OUT_NHWC>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
    } else {
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_pblock<tc::PBlock2, K16, tc::OUT_P8>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, h->num_sms, s)));
      ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_pblock<tc::PBlock3, K16, tc::OUT_NHWC>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
    }
  } else {
    constexpr int K16 = KIND == tc::KIND_I8 ? tc::KIND_F16 : KIND;
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_block<tc::CfgBlock1, K16, tc::OUT_P8>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, s)));
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_block<tc::CfgBlock2, K16, tc::OUT_P8>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, s)));
    ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_block<tc::CfgBlock3, K16, tc::OUT_NHWC>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, s)));
  }
  return run_tail<T>(h, buf(p.p3), buf(p.cat4), buf(p.a4), n, probs, logits, s);
}

// Baseline ErNET (model/ernet.py:24-48), one chunk through the layer-wise CUDA-core kernels: conv1, then per ACFF block
// the depthwise trio + concat and the 1x1 conv with LeakyReLU, BN (and the 2x2 max-pool of blocks 1-3) in its epilogue.
template <typename T>
static int run_chunk_ernet(ernet_handle* h, const void* x, int x_dtype, int x_layout, int n, float* probs, float* logits,
                           char* ws, cudaStream_t s) {
  const Plan p = make_plan(h, n);
  auto buf = [&](size_t off) { return reinterpret_cast<T*>(ws + off); };
  int rc;
  {
    long long sb = 3LL * 240 * 240, sc, sy, sx;
    if (x_layout == ERNET_NCHW) { sc = 240 * 240; sy = 240; sx = 1; }
    else                        { sc = 1; sy = 240 * 3; sx = 3; }
    const int total = n * 119 * 119, grid = (total + 127) / 128;
    const float* w = h->f(ERNET_T_STEM_W); const float* b = h->f(ERNET_T_STEM_B);
    StageTimer _t(h, ERNET_STAGE_STEM, s);
    if (x_dtype == ERNET_F32) stem_kernel<float, T, 16><<<grid, 128, 0, s>>>(static_cast<const float*>(x), sb, sc, sy, sx, w, b, buf(p.stem), total, 119);
    else if (x_dtype == ERNET_F16) stem_kernel<__half, T, 16><<<grid, 128, 0, s>>>(static_cast<const __half*>(x), sb, sc, sy, sx, w, b, buf(p.stem), total, 119);
    else stem_kernel<__nv_bfloat16, T, 16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), sb, sc, sy, sx, w, b, buf(p.stem), total, 119);
    ERNET_LAUNCH_CHECK("stem_kernel");
  }
  // (input size, pooled?) per block: 119 -> 117 -> 58 -> 56 -> 28 -> 26 -> 13 -> 11 -> 9 -> 7
  const int hin[6] = {119, 58, 28, 13, 11, 9};
  const int pool[6] = {1, 1, 1, 0, 0, 0};
  const int st_dw[6] = {ERNET_STAGE_DW1, ERNET_STAGE_DW2, ERNET_STAGE_DW3, ERNET_STAGE_DW4, ERNET_STAGE_DW4, ERNET_STAGE_DW4};   // profiling labels: blocks 4-6 share one
  const int st_pw[6] = {ERNET_STAGE_PW1, ERNET_STAGE_PW2, ERNET_STAGE_PW3, ERNET_STAGE_PW4, ERNET_STAGE_PW4, ERNET_STAGE_PW4};
  const T* cur = buf(p.stem);
  T* pp[2] = {buf(p.ea), buf(p.eb)};
  for (int k = 0; k < 6; ++k) {
    const int base = ernet_block_base(k), ci = kErnetCin[k], co = kErnetCout[k];
    const int H = hin[k], Ho = H - 2;
    // with a pool only the even-sized region it keeps is computed (58 = floor(117 / 2) needs rows/cols 0..115)
    const int Hu = pool[k] ? (Ho / 2) * 2 : Ho;
    ERNET_STAGE(st_dw[k], launch_acff_dw<T>(cur, n, H, H, ci, Hu, Hu, h->f(base + ERNET_T_DW_W), h->f(base + ERNET_T_DW_B), buf(p.ecat), s));
    T* dst = pp[k & 1];
    ERNET_STAGE(st_pw[k], launch_pointwise<T>(buf(p.ecat), n, Hu, Hu, 3 * ci, co, h->f(base + ERNET_T_PW_W), h->f(base + ERNET_T_PW_B),
                                  h->f(base + ERNET_T_BN_S), h->f(base + ERNET_T_BN_T), 1, pool[k], dst, s));
    cur = dst;
  }
  {
    StageTimer _t(h, ERNET_STAGE_HEAD, s);
    ernet_head_kernel<T><<<n, 256, 0, s>>>(cur, h->f(ERNET_T_EHEAD_W), h->f(ERNET_T_HEAD_B), probs, logits);
    ERNET_LAUNCH_CHECK("ernet_head_kernel");
  }
  return ERNET_OK;
}

// ErNET on the tensor-core block kernels: conv1 -> P8, block 1 on the persistent kernel, blocks 2-6 on CTA pairs
// (blocks 4-6 without a pool; block 6 with N = 256), head on the CUDA cores.
template <typename T, int KIND>
static int run_chunk_ernet_tc(ernet_handle* h, const void* x, int x_dtype, int x_layout, int n, float* probs, float* logits,
                              char* ws, cudaStream_t s) {
  const Plan p = make_plan(h, n);
  auto u16 = [&](size_t off) { return reinterpret_cast<uint16_t*>(ws + off); };
  auto wimg = [&](int k) { return h->t[ERNET_T_ETC_BASE + 2 * k].dev; };
  int rc;
  {
    long long sb = 3LL * 240 * 240, sc, sy, sx;
    if (x_layout == ERNET_NCHW) { sc = 240 * 240; sy = 240; sx = 1; }
    else                        { sc = 1; sy = 240 * 3; sx = 3; }
    const int total = n * 122 * 122, grid = (total + 127) / 128;
    const float* w = h->f(ERNET_T_STEM_W); const float* b = h->f(ERNET_T_STEM_B);
    StageTimer _t(h, ERNET_STAGE_STEM, s);
    if (x_dtype == ERNET_F32) tc::stem_p8_kernel<float, 16, KIND, 119><<<grid, 128, 0, s>>>(static_cast<const float*>(x), sb, sc, sy, sx, w, b, u16(p.stem), total, h->stem_inv);
    else if (x_dtype == ERNET_F16) tc::stem_p8_kernel<__half, 16, KIND, 119><<<grid, 128, 0, s>>>(static_cast<const __half*>(x), sb, sc, sy, sx, w, b, u16(p.stem), total, h->stem_inv);
    else tc::stem_p8_kernel<__nv_bfloat16, 16, KIND, 119><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), sb, sc, sy, sx, w, b, u16(p.stem), total, h->stem_inv);
    ERNET_LAUNCH_CHECK("stem_p8_kernel");
  }
  ERNET_STAGE(ERNET_STAGE_TC_BLOCK1, (tc::launch_acff_pblock<tc::EBlock1, KIND, tc::OUT_P8>(u16(p.stem), wimg(0), h->epi1, u16(p.p1), n, h->num_sms, s)));
  ERNET_STAGE(ERNET_STAGE_TC_BLOCK2, (tc::launch_acff_cblock<tc::EBlock2, KIND, tc::OUT_P8>(u16(p.p1), wimg(1), h->epi2, u16(p.p2), n, h->num_sms, s)));
  ERNET_STAGE(ERNET_STAGE_TC_BLOCK3, (tc::launch_acff_cblock<tc::EBlock3, KIND, tc::OUT_P8>(u16(p.p2), wimg(2), h->epi3, u16(p.p3), n, h->num_sms, s)));
  ERNET_STAGE(ERNET_STAGE_TC_BLOCK4, (tc::launch_acff_cblock<tc::EBlock4, KIND, tc::OUT_P8>(u16(p.p3), wimg(3), h->ee4, u16(p.e4), n, h->num_sms, s)));
  ERNET_STAGE(ERNET_STAGE_TC_BLOCK4, (tc::launch_acff_cblock<tc::EBlock5, KIND, tc::OUT_P8>(u16(p.e4), wimg(4), h->ee5, u16(p.e5), n, h->num_sms, s)));
  ERNET_STAGE(ERNET_STAGE_TC_BLOCK4, (tc::launch_acff_cblock<tc::EBlock6, KIND, tc::OUT_NHWC>(u16(p.e5), wimg(5), h->ee6, u16(p.e6), n, h->num_sms, s)));
  {
    StageTimer _t(h, ERNET_STAGE_HEAD, s);
    ernet_head_kernel<T><<<n, 256, 0, s>>>(reinterpret_cast<const T*>(ws + p.e6), h->f(ERNET_T_EHEAD_W), h->f(ERNET_T_HEAD_B), probs, logits);
    ERNET_LAUNCH_CHECK("ernet_head_kernel");
  }
  return ERNET_OK;
}

static int run_chunk(ernet_handle* h, const void* x, int x_dtype, int x_layout, const uint8_t* frames,
                     const IngestTables* tab, int order, int n, float* probs, float* logits, char* ws, cudaStream_t s) {
  if (frames) h->last_fused = false;
  if (h->ernet()) {
    if (frames) {
      // aider_transforms (aider.py:430: Resize(273) -> CenterCrop(240) -> ToTensor -> Normalize) on the table-driven kernel
      // into the workspace (NHWC, the engine's type), then the same chain as for tensors
      const Plan p = make_plan(h, n);
      int rc;
      const long long sb = 3LL * 240 * 240;
      switch (h->precision) {
        case ERNET_PREC_FP32: rc = launch_ingest<float>(*tab, frames, n, order == ERNET_BGR, reinterpret_cast<float*>(ws + p.ingest), sb, 1, 240 * 3, 3, s); x_dtype = ERNET_F32; break;
        case ERNET_PREC_FP16: rc = launch_ingest<__half>(*tab, frames, n, order == ERNET_BGR, reinterpret_cast<__half*>(ws + p.ingest), sb, 1, 240 * 3, 3, s); x_dtype = ERNET_F16; break;
        case ERNET_PREC_BF16: rc = launch_ingest<__nv_bfloat16>(*tab, frames, n, order == ERNET_BGR, reinterpret_cast<__nv_bfloat16*>(ws + p.ingest), sb, 1, 240 * 3, 3, s); x_dtype = ERNET_BF16; break;
        default: return fail(ERNET_ERR_UNSUPPORTED, "int8 is implemented for Squeeze_ErNET only");
      }
      if (rc) return rc;
      x = ws + p.ingest;
      x_layout = ERNET_NHWC;
    }
    if (h->use_tc()) {
      if (h->precision == ERNET_PREC_BF16) return run_chunk_ernet_tc<__nv_bfloat16, tc::KIND_BF16>(h, x, x_dtype, x_layout, n, probs, logits, ws, s);
      return run_chunk_ernet_tc<__half, tc::KIND_F16>(h, x, x_dtype, x_layout, n, probs, logits, ws, s);
    }
    switch (h->precision) {
      case ERNET_PREC_FP32: return run_chunk_ernet<float>(h, x, x_dtype, x_layout, n, probs, logits, ws, s);
      case ERNET_PREC_FP16: return run_chunk_ernet<__half>(h, x, x_dtype, x_layout, n, probs, logits, ws, s);
      case ERNET_PREC_BF16: return run_chunk_ernet<__nv_bfloat16>(h, x, x_dtype, x_layout, n, probs, logits, ws, s);
      default: return fail(ERNET_ERR_UNSUPPORTED, "int8 is implemented for Squeeze_ErNET only");
    }
  }
  if (h->use_tc()) {
    if (h->precision == ERNET_PREC_BF16) return run_chunk_tc<__nv_bfloat16, tc::KIND_BF16>(h, x, x_dtype, x_layout, frames, tab, order, n, probs, logits, ws, s);
    if (h->precision == ERNET_PREC_INT8) return run_chunk_tc<__half, tc::KIND_I8>(h, x, x_dtype, x_layout, frames, tab, order, n, probs, logits, ws, s);
    return run_chunk_tc<__half, tc::KIND_F16>(h, x, x_dtype, x_layout, frames, tab, order, n, probs, logits, ws, s);
  }
  switch (h->precision) {
    case ERNET_PREC_FP32: return run_chunk_simt<float>(h, x, x_dtype, x_layout, frames, tab, order, n, probs, logits, ws, s);
    case ERNET_PREC_FP16: return run_chunk_simt<__half>(h, x, x_dtype, x_layout, frames, tab, order, n, probs, logits, ws, s);
    case ERNET_PREC_BF16: return run_chunk_simt<__nv_bfloat16>(h, x, x_dtype, x_layout, frames, tab, order, n, probs, logits, ws, s);
    default: return fail(ERNET_ERR_UNSUPPORTED, "precision %d is not implemented in this build", h->precision);
  }
}

template <typename T>
static int set_fused_ingest_attrs() {
  const int lim = 200 * 1024;
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<T, 16, FS_NHWC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<T, 8, FS_NHWC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<T, 16, FS_NHWC, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<T, 8, FS_NHWC, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  return ERNET_OK;
}

template <typename T>
static int set_smem_attrs() {
  ERNET_CUDA(cudaFuncSetAttribute(acff_dw_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  return ERNET_OK;
}

static int init_device_attrs() {
  int rc;
  if ((rc = set_smem_attrs<float>())) return rc;
  if ((rc = set_smem_attrs<__half>())) return rc;
  if ((rc = set_smem_attrs<__nv_bfloat16>())) return rc;
  if ((rc = set_fused_ingest_attrs<float>())) return rc;
  if ((rc = set_fused_ingest_attrs<__half>())) return rc;
  if ((rc = set_fused_ingest_attrs<__nv_bfloat16>())) return rc;
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__half, 16, FS_P8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__nv_bfloat16, 16, FS_P8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__half, 16, FS_P16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__half, 8, FS_P8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__nv_bfloat16, 8, FS_P8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__half, 8, FS_P8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__nv_bfloat16, 8, FS_P8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__half, 16, FS_P8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__nv_bfloat16, 16, FS_P8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem_kernel<__half, 16, FS_P16, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<__half, 16, FS_P16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<__nv_bfloat16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<__half, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<__nv_bfloat16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<__half, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<float, 16, FS_NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  ERNET_CUDA(cudaFuncSetAttribute(ingest_stem5_kernel<float, 8, FS_NHWC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  if ((rc = tc::set_pw32_attr<tc::FPw1>())) return rc;
  if ((rc = tc::set_pw32_attr<tc::FPw2>())) return rc;
  if ((rc = tc::set_pw32_attr<tc::FPw3>())) return rc;
  if ((rc = tc::set_pw32_attr<tc::FPw4>())) return rc;
  if ((rc = tc::set_all_block_attrs())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1T, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1T, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1PT, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1PT, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1PT, tc::KIND_F16, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1PT, tc::KIND_I8, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock2, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock2, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock3, tc::KIND_BF16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock3, tc::KIND_F16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock1, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock1, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2S, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2S, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3S, tc::KIND_BF16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3S, tc::KIND_F16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3, tc::KIND_BF16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3, tc::KIND_F16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2R, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2R, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3R, tc::KIND_BF16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3R, tc::KIND_F16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PRed2R, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PRed2R, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1, tc::KIND_I8, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1P, tc::KIND_I8, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1P, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1P, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2Q, tc::KIND_I8, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3Q, tc::KIND_I8, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::EBlock1, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::EBlock1, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock2, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock2, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock3, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock3, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock4, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock4, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock5, tc::KIND_BF16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock5, tc::KIND_F16, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock6, tc::KIND_BF16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::EBlock6, tc::KIND_F16, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_fblock_attr<tc::PBlock1, tc::KIND_BF16, tc::OUT_P8, __nv_bfloat16, 16, FS_P8>())) return rc;
  if ((rc = tc::set_fblock_attr<tc::PBlock1, tc::KIND_F16, tc::OUT_P8, __half, 16, FS_P8>())) return rc;
  if ((rc = tc::set_fblock_attr<tc::PBlock1P, tc::KIND_I8, tc::OUT_P16, __half, 16, FS_P16>())) return rc;
  if ((rc = tc::set_fblock_attr<tc::PBlock1P, tc::KIND_BF16, tc::OUT_P8, __nv_bfloat16, 8, FS_P8>())) return rc;
  if ((rc = tc::set_fblock_attr<tc::PBlock1P, tc::KIND_F16, tc::OUT_P8, __half, 8, FS_P8>())) return rc;
  if ((rc = tc::set_fblock_attr<tc::PBlock1P, tc::KIND_F16, tc::OUT_P16, __half, 8, FS_P8>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1P, tc::KIND_F16, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PBlock1, tc::KIND_F16, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock2RQ, tc::KIND_I8, tc::OUT_P8>())) return rc;
  if ((rc = tc::set_cblock_attr<tc::CBlock3RQ, tc::KIND_I8, tc::OUT_NHWC>())) return rc;
  if ((rc = tc::set_pblock_attr<tc::PRed2RQ, tc::KIND_F16, tc::OUT_P16>())) return rc;
  if ((rc = tc::set_tail_attrs<tc::TailCfg128>())) return rc;
  if ((rc = tc::set_tail_attrs<tc::TailCfg64>())) return rc;
  return ERNET_OK;
}

static int get_tables(ernet_handle* h, int H, int W, const IngestTables** out) {
  auto key = std::make_pair(H, W);
  auto it = h->ingest.find(key);
  if (it == h->ingest.end()) {
    if (h->ingest.size() >= 64) {                          // bounded cache: a long evaluation over many frame sizes must not grow
      cudaDeviceSynchronize();                             // device memory without limit; nothing may still be reading a table
      for (auto& kv : h->ingest) if (kv.second.d_base) cudaFree(kv.second.d_base);
      h->ingest.clear();
    }
    IngestTables t;
    int rc = build_ingest_tables(t, H, W, h->in_hw(), h->ernet() ? 273 : 159);   // int(image_size * 1.14), aider.py:422
    if (rc) return rc;
    it = h->ingest.emplace(key, t).first;
  }
  *out = &it->second;
  return ERNET_OK;
}

// The mbarrier watchdog of the tensor-core kernels (tc_common.cuh) ends a stuck CTA with partial output instead of
// hanging the device; it also raises a word in mapped pinned host memory, which the synchronising entry points read
// (no device round trip) and turn into an error instead of returning wrong probabilities.
static volatile unsigned int* g_watchdog_host = nullptr;
static int init_watchdog_flag() {
  if (g_watchdog_host) return ERNET_OK;
  unsigned int* hp = nullptr;
  ERNET_CUDA(cudaHostAlloc(&hp, sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable));
  *hp = 0u;
  g_watchdog_host = hp;
  return ERNET_OK;
}
static int publish_watchdog_flag() {                       // per device: the device-side pointer to the host word
  unsigned int* dp = nullptr;
  ERNET_CUDA(cudaHostGetDevicePointer(&dp, const_cast<unsigned int*>(g_watchdog_host), 0));
  ERNET_CUDA(cudaMemcpyToSymbol(tc::g_tc_host_flag, &dp, sizeof(dp)));
  return ERNET_OK;
}
static int check_watchdog(const char* where) {
  if (g_watchdog_host && *g_watchdog_host) {
    unsigned int st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyFromSymbol(st, tc::g_tc_status, sizeof(st));
    *g_watchdog_host = 0u;
    unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(tc::g_tc_status, z, sizeof(z));
    return fail(ERNET_ERR_CUDA, "%s: a tensor-core kernel timed out in an mbarrier wait (tag 0x%x, block %u, aux %u, %u waits): results are invalid",
                where, st[1], st[2], st[3], st[0]);
  }
  return ERNET_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace ernet

// ================================================================================================ C ABI
extern "C" {

#ifndef ERNET_SOURCE_HASH
#define ERNET_SOURCE_HASH "unknown"
#endif
// "ERNET_SOURCE_HASH=<sha256 of the sources>": found by build.py in the file itself to decide whether to rebuild
extern const char ernet_source_hash_str[] = "ERNET_SOURCE_HASH=" ERNET_SOURCE_HASH;
const char* ernet_source_hash(void) { return ernet_source_hash_str + 18; }

const char* ernet_last_error(void) { return g_err; }
int ernet_abi_version(void) { return ERNET_ABI_VERSION; }

int ernet_create(ernet_handle** out, int arch, int precision, int device) {
  if (!out) return fail(ERNET_ERR_INVALID_ARG, "ernet_create: out is null");
  *out = nullptr;
  if (arch != ERNET_ARCH_SQUEEZE && arch != ERNET_ARCH_REDCONV && arch != ERNET_ARCH_ERNET)
    return fail(ERNET_ERR_INVALID_ARG, "Unsupported model: arch=%d", arch);   // aider-predict.py:32
  if (arch == ERNET_ARCH_ERNET && precision == ERNET_PREC_INT8)
    return fail(ERNET_ERR_UNSUPPORTED, "int8 is implemented for Squeeze_ErNET only");
  if (precision < ERNET_PREC_FP32 || precision > ERNET_PREC_INT8)
    return fail(ERNET_ERR_INVALID_ARG, "unknown precision %d", precision);
  int ndev = 0;
  ERNET_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(ERNET_ERR_INVALID_ARG, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  ERNET_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(ERNET_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  DeviceGuard g(device);
  if (!g.ok) return fail(ERNET_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  int rc = init_device_attrs();
  if (rc) return rc;
  if ((rc = init_watchdog_flag()) || (rc = publish_watchdog_flag())) return rc;
  ernet_handle* h = new (std::nothrow) ernet_handle();
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "out of host memory");
  h->arch = arch; h->precision = precision; h->device = device;
  h->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("ERNET_PAIR_BLOCK1")) h->pair_block1 = atoi(e) != 0;
  if (const char* e = getenv("ERNET_PAIR_TAPS")) h->pair_taps = atoi(e) != 0;
  if (const char* e = getenv("ERNET_FUSE_INGEST")) h->fuse_ingest = atoi(e) != 0;
  if (cudaMalloc(&h->d_sync, 2 * kSyncImages * sizeof(uint32_t)) != cudaSuccess || cudaMemset(h->d_sync, 0, 2 * kSyncImages * sizeof(uint32_t)) != cudaSuccess) {
    delete h;
    return fail(ERNET_ERR_CUDA, "allocating the band counters failed");
  }
  if (const char* e = getenv("ERNET_FP32_TC")) h->fp32_tc = atoi(e) != 0;
  if (const char* e = getenv("ERNET_TRIM_COLUMNS")) h->trim_columns = atoi(e) != 0;
  if (const char* e = getenv("ERNET_HOST_GATHER")) h->host_gather = atoi(e) != 0;
  if (const char* e = getenv("ERNET_TAIL_TILES")) h->tail_tiles = atoi(e) != 0;
  if (const char* e = getenv("ERNET_NVTX")) h->nvtx = atoi(e) != 0;
  if (const char* e = getenv("ERNET_SMALL_BATCH_UNITS")) h->small_batch_units = atoi(e) != 0;
  if (const char* e = getenv("ERNET_EPI_SUSPEND")) {       // study switch, device-wide (tc_common.cuh)
    const unsigned int v = (unsigned int)atoi(e) & 3u;       // bit 0: block kernels' epilogue warps, bit 1: ACFF4 + head kernel
    cudaMemcpyToSymbol(tc::g_epi_suspend, &v, sizeof(v));
  }
  if (const char* e = getenv("ERNET_GATHER_CTAS")) { const int c = atoi(e); if (c >= 1 && c <= 1024) h->gather_ctas = c; }
  if (const char* e = getenv("ERNET_DUAL_COPY")) h->dual_copy = atoi(e) != 0;
  *out = h;
  return ERNET_OK;
}

void ernet_destroy(ernet_handle* h) {
  if (!h) return;
  DeviceGuard g(h->device);
  if (h->d_blob) cudaFree(h->d_blob);
  if (h->d_stem_frag) cudaFree(h->d_stem_frag);
  if (h->d_stem_frag32) cudaFree(h->d_stem_frag32);
  if (h->d_w1_pair) cudaFree(h->d_w1_pair);
  for (int k = 0; k < 4; ++k) if (h->d_pw32[k]) cudaFree(h->d_pw32[k]);
  if (h->d_sync) cudaFree(h->d_sync);
  for (auto& kv : h->ingest) if (kv.second.d_base) cudaFree(kv.second.d_base);
  for (int i = 0; i < 2; ++i) {
    if (h->d_frames[i]) cudaFree(h->d_frames[i]);
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    if (h->ev_call[i]) cudaEventDestroy(h->ev_call[i]);
  }
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->prof_pool) cudaEventDestroy(e);
  if (h->d_ws) cudaFree(h->d_ws);
  if (h->d_res) cudaFree(h->d_res);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->s_copy2) cudaStreamDestroy(h->s_copy2);
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  delete h;
}

int ernet_load_packed(ernet_handle* h, const void* blob, size_t bytes) {
  if (!h || !blob) return fail(ERNET_ERR_INVALID_ARG, "ernet_load_packed: null argument");
  if (bytes < sizeof(ernet_blob_header)) return fail(ERNET_ERR_BAD_BLOB, "blob too small (%zu bytes)", bytes);
  ernet_blob_header hd;
  memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != ERNET_BLOB_MAGIC) return fail(ERNET_ERR_BAD_BLOB, "bad magic 0x%08x", hd.magic);
  if (hd.version != ERNET_BLOB_VERSION) return fail(ERNET_ERR_BAD_BLOB, "blob version %u, library expects %u", hd.version, ERNET_BLOB_VERSION);
  if ((int)hd.arch != h->arch || (int)hd.precision != h->precision)
    return fail(ERNET_ERR_BAD_BLOB, "blob is arch=%u precision=%u, handle is arch=%d precision=%d", hd.arch, hd.precision, h->arch, h->precision);
  const size_t table_end = sizeof(hd) + (size_t)hd.n_entries * sizeof(ernet_blob_entry);
  if (hd.n_entries > 4096 || table_end > bytes) return fail(ERNET_ERR_BAD_BLOB, "entry table truncated");
  std::vector<ernet_blob_entry> ent(hd.n_entries);
  memcpy(ent.data(), static_cast<const char*>(blob) + sizeof(hd), hd.n_entries * sizeof(ernet_blob_entry));
  for (auto& e : ent) {
    if (e.id >= ERNET_T_MAX) return fail(ERNET_ERR_BAD_BLOB, "tensor id %u out of range", e.id);
    if (e.offset % 256 || e.offset < table_end || e.offset + e.nbytes > bytes)
      return fail(ERNET_ERR_BAD_BLOB, "tensor id %u: bad extent [%llu,+%llu) in a %zu-byte blob", e.id,
                  (unsigned long long)e.offset, (unsigned long long)e.nbytes, bytes);
  }
  DeviceGuard g(h->device);
  void* d = nullptr;
  ERNET_CUDA(cudaMalloc(&d, bytes));
  cudaError_t ce = cudaMemcpy(d, blob, bytes, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { cudaFree(d); return fail(ERNET_ERR_CUDA, "copying weights failed: %s", cudaGetErrorString(ce)); }
  Tensor old[ERNET_T_MAX];
  memcpy(old, h->t, sizeof(old));
  for (auto& t : h->t) t = Tensor();
  for (auto& e : ent) {
    h->t[e.id].dev = static_cast<char*>(d) + e.offset;
    h->t[e.id].nbytes = e.nbytes;
    h->t[e.id].dtype = (int)e.dtype;
  }
  int rc = validate_simt_tensors(h);
  if (rc) { memcpy(h->t, old, sizeof(old)); cudaFree(d); return rc; }
  if (h->ernet()) {
    // tensor-core images of the six blocks (16-bit blobs): all present with the sizes the kernel configurations expect?
    const size_t wb[6] = {(size_t)25 * 2 * 64 * 16, (size_t)25 * 8 * 96 * 16, (size_t)25 * 12 * 128 * 16, (size_t)25 * 16 * 128 * 16,
                          (size_t)25 * 16 * 128 * 16, (size_t)25 * 16 * 256 * 16};
    bool all = h->precision == ERNET_PREC_BF16 || h->precision == ERNET_PREC_FP16;
    for (int k = 0; k < 6 && all; ++k) {
      const Tensor& w = h->t[ERNET_T_ETC_BASE + 2 * k];
      const Tensor& b = h->t[ERNET_T_ETC_BASE + 2 * k + 1];
      all = w.dev && b.dev && w.nbytes == wb[k] && b.nbytes == (size_t)kErnetCout[k] * sizeof(float);
    }
    h->has_tc = all;
    h->has_tail = false;
    if (all) {
      auto host_f32 = [&](int id) { return reinterpret_cast<const float*>(static_cast<const char*>(blob) + (static_cast<const char*>(h->t[id].dev) - static_cast<const char*>(d))); };
      auto fill = [&](int k, float* bias, float* scale, float* shift, float* deq, float* out_inv, int n) {
        memcpy(bias, host_f32(ERNET_T_ETC_BASE + 2 * k + 1), n * sizeof(float));
        memcpy(scale, host_f32(ernet_block_base(k) + ERNET_T_BN_S), n * sizeof(float));
        memcpy(shift, host_f32(ernet_block_base(k) + ERNET_T_BN_T), n * sizeof(float));
        for (int i = 0; i < n; ++i) { deq[i] = 1.f; out_inv[i] = 1.f; }
      };
      fill(0, h->epi1.bias, h->epi1.scale, h->epi1.shift, h->epi1.deq, h->epi1.out_inv, 64);
      fill(1, h->epi2.bias, h->epi2.scale, h->epi2.shift, h->epi2.deq, h->epi2.out_inv, 96);
      fill(2, h->epi3.bias, h->epi3.scale, h->epi3.shift, h->epi3.deq, h->epi3.out_inv, 128);
      fill(3, h->ee4.bias, h->ee4.scale, h->ee4.shift, h->ee4.deq, h->ee4.out_inv, 128);
      fill(4, h->ee5.bias, h->ee5.scale, h->ee5.shift, h->ee5.deq, h->ee5.out_inv, 128);
      fill(5, h->ee6.bias, h->ee6.scale, h->ee6.shift, h->ee6.deq, h->ee6.out_inv, 256);
      for (int i = 0; i < 16; ++i) h->stem_inv.v[i] = 1.f;
    }
  } else {
    const bool q = h->precision == ERNET_PREC_INT8;
    const size_t wimg_bytes[3] = {q ? (size_t)tc::CfgBlock1Q::W_BYTES : (size_t)tc::CfgBlock1::W_BYTES,
                                  q ? (size_t)tc::CfgBlock2Q::W_BYTES : (size_t)tc::CfgBlock2::W_BYTES,
                                  q ? (h->red() ? (size_t)tc::CfgBlock3RQ::W_BYTES : (size_t)tc::CfgBlock3Q::W_BYTES)
                                    : (h->red() ? (size_t)tc::CfgBlock3R::W_BYTES : (size_t)tc::CfgBlock3::W_BYTES)};
    const size_t nout[3] = {64, 96, 128};
    bool all = h->precision != ERNET_PREC_FP32;
    if (h->red()) {
      const Tensor& wr = h->t[ERNET_T_TC_RED2_WIMG];
      const Tensor& br = h->t[ERNET_T_TC_RED2_BIAS];
      all = all && wr.dev && br.dev && wr.nbytes == (size_t)tc::CfgRed2R::W_BYTES && br.nbytes == 64 * sizeof(float);
    }
    for (int k = 0; k < 3 && all; ++k) {
      const Tensor& w = h->t[ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_WIMG];
      const Tensor& b = h->t[ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_BIAS];
      all = w.dev && b.dev && w.nbytes == wimg_bytes[k] && b.nbytes == nout[k] * sizeof(float);
      if (q) {
        const Tensor& dq = h->t[ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_DEQ];
        all = all && dq.dev && dq.nbytes == nout[k] * sizeof(float);
      }
    }
    if (q) all = all && h->t[ERNET_T_Q_SCALES].dev && h->t[ERNET_T_Q_SCALES].nbytes == sizeof(h->q_scales);
    h->has_tc = all;
    if (all) {
      auto host_f32 = [&](int id) { return reinterpret_cast<const float*>(static_cast<const char*>(blob) + (static_cast<const char*>(h->t[id].dev) - static_cast<const char*>(d))); };
      if (q || h->red()) {
        // Block 1 sees one real 16-byte chunk per pixel (16 int8 channels / RedConv's 8 16-bit channels; chunk 1 of the
        // packed image multiplies zeros): regroup [25 taps][2 chunks][64][16 B] into 13 tap pairs for the PAIR kernel
        const uint8_t* src = reinterpret_cast<const uint8_t*>(host_f32(ERNET_T_TC_BASE + ERNET_T_TC_WIMG));
        std::vector<uint8_t> pr(13 * 2 * 64 * 16, 0);
        for (int p = 0; p < 13; ++p) {
          const int tapA = p == 0 ? 0 : 2 * p - 1, tapB = p == 0 ? -1 : 2 * p;
          memcpy(pr.data() + (size_t)(p * 2 + 0) * 1024, src + (size_t)(tapA * 2) * 1024, 1024);
          if (tapB >= 0) memcpy(pr.data() + (size_t)(p * 2 + 1) * 1024, src + (size_t)(tapB * 2) * 1024, 1024);
        }
        if (!h->d_w1_pair) ERNET_CUDA(cudaMalloc(&h->d_w1_pair, pr.size()));
        ERNET_CUDA(cudaMemcpy(h->d_w1_pair, pr.data(), pr.size(), cudaMemcpyHostToDevice));
      }
      {  // conv1 with ToTensor/Normalize folded in, in mma.sync fragment order (ingest_fast.cuh)
        StemFrag sfh;
        build_stem_fragments(host_f32(ERNET_T_STEM_W), host_f32(ERNET_T_STEM_B), h->cs(), &sfh);
        if (!h->d_stem_frag) ERNET_CUDA(cudaMalloc(&h->d_stem_frag, sizeof(StemFrag)));
        ERNET_CUDA(cudaMemcpy(h->d_stem_frag, &sfh, sizeof(StemFrag), cudaMemcpyHostToDevice));
      }
      if (q) {
        memcpy(h->q_scales, host_f32(ERNET_T_Q_SCALES), sizeof(h->q_scales));
        for (int i = 0; i < 16; ++i) h->stem_inv.v[i] = 1.f / h->q_scales[i];
      } else {
        for (int i = 0; i < 16; ++i) h->stem_inv.v[i] = 1.f;
      }
      auto fill = [&](int k, float* bias, float* scale, float* shift, float* deq, float* out_inv, int n) {
        memcpy(bias, host_f32(ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_BIAS), n * sizeof(float));
        memcpy(scale, host_f32(ERNET_T_BLOCK_BASE + 8 * k + ERNET_T_BN_S), n * sizeof(float));
        memcpy(shift, host_f32(ERNET_T_BLOCK_BASE + 8 * k + ERNET_T_BN_T), n * sizeof(float));
        if (q) memcpy(deq, host_f32(ERNET_T_TC_BASE + 4 * k + ERNET_T_TC_DEQ), n * sizeof(float));
        else for (int i = 0; i < n; ++i) deq[i] = 1.f;
        const float* qs = h->q_scales + (k == 0 ? 16 : 16 + 64);      // blocks 1, 2 requantise for the next int8 block
        for (int i = 0; i < n; ++i) out_inv[i] = (q && k < 2) ? 1.f / qs[i] : 1.f;
      };
      fill(0, h->epi1.bias, h->epi1.scale, h->epi1.shift, h->epi1.deq, h->epi1.out_inv, 64);
      fill(1, h->epi2.bias, h->epi2.scale, h->epi2.shift, h->epi2.deq, h->epi2.out_inv, 96);
      fill(2, h->epi3.bias, h->epi3.scale, h->epi3.shift, h->epi3.deq, h->epi3.out_inv, 128);
      if (h->red()) {
        memcpy(h->epi_r2.bias, host_f32(ERNET_T_TC_RED2_BIAS), 64 * sizeof(float));
        for (int i = 0; i < 64; ++i) { h->epi_r2.scale[i] = 1.f; h->epi_r2.shift[i] = 0.f; h->epi_r2.deq[i] = 1.f; h->epi_r2.out_inv[i] = 1.f; }
        if (q) {   // int8: conv_red2 writes the int8 pool2 tensor; ACFF2's own (un-pooled, fp16) output is not re-quantised
          for (int i = 0; i < 48; ++i) h->epi_r2.out_inv[i] = 1.f / h->q_scales[16 + 64 + i];
          for (int i = 0; i < 96; ++i) h->epi2.out_inv[i] = 1.f;
        }
      }
    }
    if (q && !all) {
      memcpy(h->t, old, sizeof(old)); cudaFree(d);
      return fail(ERNET_ERR_BAD_BLOB, "int8 blob lacks the calibrated tensor-core images (pack with act_scales)");
    }
  }
  {
    const Tensor& w4 = h->t[ERNET_T_TC4_WIMG];
    h->has_tail = !h->ernet() && h->precision != ERNET_PREC_FP32 && w4.dev && w4.nbytes == (size_t)3 * h->c4() * 256 * 2;
    if (h->has_tail) {
      auto host_f32 = [&](int id) { return reinterpret_cast<const float*>(static_cast<const char*>(blob) + (static_cast<const char*>(h->t[id].dev) - static_cast<const char*>(d))); };
      memcpy(h->tail.bias, host_f32(ERNET_T_BLOCK_BASE + 24 + ERNET_T_PW_B), 256 * sizeof(float));
      memcpy(h->tail.scale, host_f32(ERNET_T_BLOCK_BASE + 24 + ERNET_T_BN_S), 256 * sizeof(float));
      memcpy(h->tail.shift, host_f32(ERNET_T_BLOCK_BASE + 24 + ERNET_T_BN_T), 256 * sizeof(float));
      memcpy(h->tail.weff, host_f32(ERNET_T_HEAD_W), 5 * 256 * sizeof(float));
      memcpy(h->tail.bfc, host_f32(ERNET_T_HEAD_B), 5 * sizeof(float));
    }
  }
  if (h->d_blob) { cudaDeviceSynchronize(); cudaFree(h->d_blob); }
  h->d_blob = d; h->blob_bytes = bytes; h->loaded = true;
  if (h->precision == ERNET_PREC_FP32 && h->arch != ERNET_ARCH_ERNET) {
    // conv1 (+conv_red1) with ToTensor/Normalize folded in as (hi, lo) fp16 images in mma.sync fragment order (ingest_fast.cuh)
    const Tensor& sw_ = h->t[ERNET_T_STEM_W];
    const Tensor& sb_ = h->t[ERNET_T_STEM_B];
    if (sw_.dev && sb_.dev && sw_.nbytes == (size_t)27 * h->cs() * sizeof(float)) {
      auto host_f32 = [&](const Tensor& t) { return reinterpret_cast<const float*>(static_cast<const char*>(blob) + (static_cast<const char*>(t.dev) - static_cast<const char*>(d))); };
      StemFrag32 sfh;
      build_stem_fragments32(host_f32(sw_), host_f32(sb_), h->cs(), &sfh);
      if (!h->d_stem_frag32) ERNET_CUDA(cudaMalloc(&h->d_stem_frag32, sizeof(StemFrag32)));
      ERNET_CUDA(cudaMemcpy(h->d_stem_frag32, &sfh, sizeof(StemFrag32), cudaMemcpyHostToDevice));
    }
  }
  if (h->precision == ERNET_PREC_FP32 && h->arch == ERNET_ARCH_SQUEEZE) {
    // (hi, lo) TF32 images of the 1x1 weights of blocks 1-3, in the stage order tc_pw32.cuh streams them
    const int kk[4] = {48, 192, 288, 384}, nn[4] = {64, 96, 128, 256};
    for (int k = 0; k < 4; ++k) {
      const Tensor& w = h->t[ERNET_T_BLOCK_BASE + 8 * k + ERNET_T_PW_W];
      if (!w.dev || w.nbytes != (size_t)kk[k] * nn[k] * sizeof(float)) continue;
      if (!h->d_pw32[k]) ERNET_CUDA(cudaMalloc(&h->d_pw32[k], tc::pw32_weight_floats(kk[k], nn[k]) * sizeof(float)));
      const int splits = k == 3 ? tc::FPw4::NSPLIT : 1, nsub = nn[k] / splits;      // block 4: one image per 64-channel split
      for (int sp = 0; sp < splits; ++sp) {
        tc::pw32_pack_weights<<<(kk[k] * nsub + 255) / 256, 256>>>(h->blk(k, ERNET_T_PW_W), kk[k], nsub,
                                                                  h->d_pw32[k] + (size_t)sp * tc::pw32_weight_floats(kk[k], nsub), nn[k], sp * nsub,
                                                                  k < 3 ? h->blk(k, ERNET_T_BN_S) : nullptr);   // pooled blocks: pool-first epilogue
        ERNET_LAUNCH_CHECK("pw32_pack_weights");
      }
    }
    ERNET_CUDA(cudaDeviceSynchronize());
  }
  return ERNET_OK;
}

int ernet_set_chunk(ernet_handle* h, int n) {
  if (!h || n < 1 || n > 65535) return fail(ERNET_ERR_INVALID_ARG, "chunk must be in [1,65535]");
  h->chunk = n;
  return ERNET_OK;
}
int ernet_get_chunk(const ernet_handle* h) { return h ? h->chunk : 0; }

int ernet_set_engine(ernet_handle* h, int engine) {
  if (!h || engine < ERNET_ENGINE_AUTO || engine > ERNET_ENGINE_TC) return fail(ERNET_ERR_INVALID_ARG, "bad engine %d", engine);
  if (engine == ERNET_ENGINE_TC && h->loaded && !(h->has_tc && h->precision != ERNET_PREC_FP32))
    return fail(ERNET_ERR_UNSUPPORTED, "tensor-core engine needs a 16-bit or int8 handle");
  h->engine = engine;
  return ERNET_OK;
}
int ernet_set_persistent(ernet_handle* h, int on) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  if (on < 0 || on > 2) return fail(ERNET_ERR_INVALID_ARG, "schedule must be 0, 1 or 2");
  h->persistent = on;
  return ERNET_OK;
}
int ernet_set_fast_ingest(ernet_handle* h, int on) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  h->fast_ingest = on != 0;
  return ERNET_OK;
}
int ernet_set_fuse_ingest(ernet_handle* h, int on) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  h->fuse_ingest = on != 0;
  return ERNET_OK;
}
int ernet_set_host_gather(ernet_handle* h, int on, int ctas) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  if (ctas < 0 || ctas > 1024) return fail(ERNET_ERR_INVALID_ARG, "host gather: ctas must be 0 (keep) .. 1024, got %d", ctas);
  h->host_gather = on != 0;
  if (ctas) h->gather_ctas = ctas;
  return ERNET_OK;
}
int ernet_set_debug_taps(ernet_handle* h, int on) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  h->debug_taps = on != 0;
  return ERNET_OK;
}
int ernet_get_engine(const ernet_handle* h) { return h ? (h->use_tc() ? ERNET_ENGINE_TC : ERNET_ENGINE_SIMT) : -1; }

size_t ernet_workspace_bytes(const ernet_handle* h, int batch) {
  if (!h || batch < 1) return 0;
  return make_plan(h, batch < h->chunk ? batch : h->chunk).total;
}

static int forward_common(ernet_handle* h, const void* x, int x_dtype, int x_layout, const uint8_t* frames, int H, int W,
                          int order, int batch, float* probs, float* logits, void* ws, size_t ws_bytes, void* stream) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  if (!h->loaded) return fail(ERNET_ERR_NOT_LOADED, "no weights loaded (call ernet_load_packed first)");
  if (batch < 1) return fail(ERNET_ERR_INVALID_ARG, "batch must be >= 1, got %d", batch);
  if (!probs) return fail(ERNET_ERR_INVALID_ARG, "probs_out is null");
  if (!frames) {
    if (!x) return fail(ERNET_ERR_INVALID_ARG, "x is null");
    if (x_dtype != ERNET_F32 && x_dtype != ERNET_F16 && x_dtype != ERNET_BF16)
      return fail(ERNET_ERR_INVALID_ARG, "x_dtype %d must be fp32, fp16 or bf16", x_dtype);
    if (x_layout != ERNET_NCHW && x_layout != ERNET_NHWC) return fail(ERNET_ERR_INVALID_ARG, "bad x_layout %d", x_layout);
  } else if (order != ERNET_RGB && order != ERNET_BGR) {
    return fail(ERNET_ERR_INVALID_ARG, "bad channel_order %d", order);
  }
  { int wrc = check_watchdog("an earlier forward call"); if (wrc) { cudaMemset(h->d_sync, 0, 2 * kSyncImages * sizeof(uint32_t)); return wrc; } }
  const size_t need = ernet_workspace_bytes(h, batch);
  if (!ws || ws_bytes < need) return fail(ERNET_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", need, ws ? ws_bytes : 0);
  DeviceGuard g(h->device);
  const IngestTables* tab = nullptr;
  if (frames) { int rc = get_tables(h, H, W, &tab); if (rc) return rc; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t x_img = 3ull * h->in_hw() * h->in_hw() * dtype_size(x_dtype);
  const size_t f_img = (size_t)H * W * 3;
  for (int b0 = 0; b0 < batch; b0 += h->chunk) {
    const int n = batch - b0 < h->chunk ? batch - b0 : h->chunk;
    int rc = run_chunk(h, x ? static_cast<const char*>(x) + (size_t)b0 * x_img : nullptr, x_dtype, x_layout,
                       frames ? frames + (size_t)b0 * f_img : nullptr, tab, order, n, probs + (size_t)b0 * 5,
                       logits ? logits + (size_t)b0 * 5 : nullptr, static_cast<char*>(ws), s);
    if (rc) return rc;
  }
  return ERNET_OK;
}

int ernet_forward(ernet_handle* h, const void* x, int x_dtype, int x_layout, int batch, float* probs_out,
                  float* logits_out, void* workspace, size_t workspace_bytes, void* stream) {
  return forward_common(h, x, x_dtype, x_layout, nullptr, 0, 0, 0, batch, probs_out, logits_out, workspace,
                        workspace_bytes, stream);
}

int ernet_forward_frames(ernet_handle* h, const uint8_t* frames, int batch, int height, int width, int channel_order,
                         float* probs_out, float* logits_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!frames) return fail(ERNET_ERR_INVALID_ARG, "frames is null");
  return forward_common(h, nullptr, 0, 0, frames, height, width, channel_order, batch, probs_out, logits_out, workspace,
                        workspace_bytes, stream);
}

int ernet_prepare_ingest(ernet_handle* h, int height, int width) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  DeviceGuard g(h->device);
  const IngestTables* tab;
  return get_tables(h, height, width, &tab);
}

int ernet_ingest_u8(ernet_handle* h, const uint8_t* frames, int batch, int height, int width, int channel_order,
                    void* x_out, int out_dtype, int out_layout, void* stream) {
  if (!h || !frames || !x_out) return fail(ERNET_ERR_INVALID_ARG, "ernet_ingest_u8: null argument");
  if (batch < 1) return fail(ERNET_ERR_INVALID_ARG, "batch must be >= 1, got %d", batch);
  if (channel_order != ERNET_RGB && channel_order != ERNET_BGR) return fail(ERNET_ERR_INVALID_ARG, "bad channel_order");
  DeviceGuard g(h->device);
  const IngestTables* tab;
  int rc = get_tables(h, height, width, &tab);
  if (rc) return rc;
  const long long S = h->in_hw();
  long long sb = 3LL * S * S, sc, sy, sx;
  if (out_layout == ERNET_NCHW) { sc = S * S; sy = S; sx = 1; }
  else if (out_layout == ERNET_NHWC) { sc = 1; sy = S * 3; sx = 3; }
  else return fail(ERNET_ERR_INVALID_ARG, "bad out_layout %d", out_layout);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool bgr = channel_order == ERNET_BGR;
  for (int b0 = 0; b0 < batch; b0 += 32768) {     // gridDim.y limit
    const int n = batch - b0 < 32768 ? batch - b0 : 32768;
    const uint8_t* f = frames + (size_t)b0 * height * width * 3;
    const size_t o = (size_t)b0 * sb;
    if (out_dtype == ERNET_F32) rc = launch_ingest<float>(*tab, f, n, bgr, static_cast<float*>(x_out) + o, sb, sc, sy, sx, s);
    else if (out_dtype == ERNET_F16) rc = launch_ingest<__half>(*tab, f, n, bgr, static_cast<__half*>(x_out) + o, sb, sc, sy, sx, s);
    else if (out_dtype == ERNET_BF16) rc = launch_ingest<__nv_bfloat16>(*tab, f, n, bgr, static_cast<__nv_bfloat16*>(x_out) + o, sb, sc, sy, sx, s);
    else return fail(ERNET_ERR_INVALID_ARG, "bad out_dtype %d", out_dtype);
    if (rc) return rc;
  }
  return ERNET_OK;
}

static GatherGeom gather_geometry(const IngestTables* tab, int width) {
  GatherGeom g;
  g.f_img = (unsigned long long)tab->H * width * 3;
  g.rowb = width * 3;
  g.row_lo = tab->row_lo; g.nrows = tab->row_hi - tab->row_lo;
  g.xb0 = tab->col_lo * 3; g.xb1 = tab->col_hi * 3;
  g.vpr = (g.xb1 - g.xb0 + 15) / 16 + 1;
  return g;
}

size_t ernet_host_copy_bytes_per_frame(ernet_handle* h, int height, int width) {
  if (!h) return 0;
  DeviceGuard g(h->device);
  const IngestTables* tab;
  if (get_tables(h, height, width, &tab)) return 0;
  if (h->host_gather && ((size_t)height * width * 3) % 16 == 0) return gather_bytes_per_frame(gather_geometry(tab, width));
  return (size_t)(tab->row_hi - tab->row_lo) * (h->trim_columns ? (size_t)(tab->col_hi - tab->col_lo) : (size_t)width) * 3;
}

int ernet_classify_frames_host_submit(ernet_handle* h, const uint8_t* frames_host, int batch, int height, int width,
                                      int channel_order, float* probs_host, float* logits_host, int* ticket) {
  if (!h || !frames_host || !probs_host || !ticket) return fail(ERNET_ERR_INVALID_ARG, "ernet_classify_frames_host: null argument");
  if (!h->loaded) return fail(ERNET_ERR_NOT_LOADED, "no weights loaded (call ernet_load_packed first)");
  if (batch < 1) return fail(ERNET_ERR_INVALID_ARG, "batch must be >= 1, got %d", batch);
  DeviceGuard g(h->device);
  const IngestTables* tab;
  int rc = get_tables(h, height, width, &tab);
  if (rc) return rc;
  if (!h->s_copy) {
    int prio_lo = 0, prio_hi = 0;                            // copy streams first: the gather kernel's few CTAs must not queue behind compute CTAs
    ERNET_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    ERNET_CUDA(cudaStreamCreateWithPriority(&h->s_copy, cudaStreamNonBlocking, prio_hi));
    ERNET_CUDA(cudaStreamCreateWithPriority(&h->s_copy2, cudaStreamNonBlocking, prio_hi));
    ERNET_CUDA(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      ERNET_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      ERNET_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
      ERNET_CUDA(cudaEventCreateWithFlags(&h->ev_call[i], cudaEventDisableTiming));
    }
  }
  // sub-chunks so that the H2D copy of one overlaps the kernels of the previous one even inside a single call
  int chunk = batch < h->chunk ? batch : h->chunk;
  if (batch >= 128) {                                      // quarter batches, but never below 32 frames nor above the caller's chunk
    int q = (batch + 3) / 4;
    if (q < 32) q = 32;
    if (q < chunk) chunk = q;
  }
  const size_t f_img = (size_t)height * width * 3;
  const size_t fbytes = (size_t)chunk * f_img;
  // pinned (mapped) frames can be pulled by a kernel that skips the columns outside the crop footprint as well
  const uint8_t* gather_src = nullptr;
  if (h->host_gather && f_img % 16 == 0) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, frames_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
        (reinterpret_cast<size_t>(pa.devicePointer) & 15) == 0)
      gather_src = static_cast<const uint8_t*>(pa.devicePointer);
    else
      cudaGetLastError();
  }
  if (h->d_frames_bytes < fbytes || h->d_ws_bytes < ernet_workspace_bytes(h, chunk) || h->d_res_elems < (size_t)batch * 10)
    ERNET_CUDA(cudaStreamSynchronize(h->s_compute));        // growing a buffer: nothing may still be using the old one
  if (h->d_frames_bytes < fbytes) {
    for (int i = 0; i < 2; ++i) {
      if (h->d_frames[i]) { cudaFree(h->d_frames[i]); h->d_frames[i] = nullptr; }
      ERNET_CUDA(cudaMalloc(&h->d_frames[i], fbytes));
    }
    h->d_frames_bytes = fbytes;
  }
  const size_t wbytes = ernet_workspace_bytes(h, chunk);
  if (h->d_ws_bytes < wbytes) {
    if (h->d_ws) { cudaFree(h->d_ws); h->d_ws = nullptr; }
    ERNET_CUDA(cudaMalloc(&h->d_ws, wbytes));
    h->d_ws_bytes = wbytes;
  }
  if (h->d_res_elems < (size_t)batch * 10) {
    if (h->d_res) { cudaFree(h->d_res); h->d_res = nullptr; }
    ERNET_CUDA(cudaMalloc(&h->d_res, (size_t)batch * 10 * sizeof(float)));
    h->d_res_elems = (size_t)batch * 10;
  }
  float* d_probs = h->d_res;
  float* d_logits = h->d_res + (size_t)batch * 5;
  // the two frame buffers alternate across sub-chunks AND across calls (a submitted call may still be running)
  for (int b0 = 0; b0 < batch; b0 += chunk, ++h->sub_it) {
    const int n = batch - b0 < chunk ? batch - b0 : chunk;
    const unsigned long long it = h->sub_it;
    const int s = (int)(it & 1);
    cudaStream_t cs = (h->dual_copy && (it & 1)) ? h->s_copy2 : h->s_copy;
    if (it >= 2) ERNET_CUDA(cudaStreamWaitEvent(cs, h->ev_done[s], 0));   // frame buffer s is free again
    // only the rows the crop window of the eval transform reads travel over PCIe (240x240: rows 14..226, 89 % of
    // the frame); one strided copy, each "row" of it is the contiguous row range of one frame
    const size_t rowb = (size_t)width * 3, lo = (size_t)tab->row_lo * rowb, span = (size_t)(tab->row_hi - tab->row_lo) * rowb;
    if (gather_src) {
      host_gather_kernel<<<h->gather_ctas, kGatherThreads, 0, cs>>>(gather_src + (size_t)b0 * f_img, h->d_frames[s], n, (unsigned long long)n * f_img,
                                                                    gather_geometry(tab, width));
      ERNET_LAUNCH_CHECK("host_gather_kernel");
    } else if (h->trim_columns) {
      // rows AND columns of the crop window's footprint: a 3-D strided copy (x = bytes of the column range, y = rows, z = frames)
      cudaMemcpy3DParms cp = {};
      const size_t xoff = (size_t)tab->col_lo * 3, xbytes = (size_t)(tab->col_hi - tab->col_lo) * 3;
      cp.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(frames_host) + (size_t)b0 * f_img + lo + xoff, rowb, rowb, (size_t)height);
      cp.dstPtr = make_cudaPitchedPtr(h->d_frames[s] + lo + xoff, rowb, rowb, (size_t)height);
      cp.extent = make_cudaExtent(xbytes, (size_t)(tab->row_hi - tab->row_lo), (size_t)n);
      cp.kind = cudaMemcpyHostToDevice;
      ERNET_CUDA(cudaMemcpy3DAsync(&cp, cs));
    } else {
      ERNET_CUDA(cudaMemcpy2DAsync(h->d_frames[s] + lo, f_img, frames_host + (size_t)b0 * f_img + lo, f_img, span, (size_t)n,
                                   cudaMemcpyHostToDevice, cs));
    }
    ERNET_CUDA(cudaEventRecord(h->ev_copied[s], cs));
    ERNET_CUDA(cudaStreamWaitEvent(h->s_compute, h->ev_copied[s], 0));
    rc = run_chunk(h, nullptr, 0, 0, h->d_frames[s], tab, channel_order, n, d_probs + (size_t)b0 * 5,
                   d_logits + (size_t)b0 * 5, static_cast<char*>(h->d_ws), h->s_compute);
    if (rc) return rc;
    ERNET_CUDA(cudaEventRecord(h->ev_done[s], h->s_compute));
  }
  ERNET_CUDA(cudaMemcpyAsync(probs_host, d_probs, (size_t)batch * 5 * sizeof(float), cudaMemcpyDeviceToHost, h->s_compute));
  if (logits_host)
    ERNET_CUDA(cudaMemcpyAsync(logits_host, d_logits, (size_t)batch * 5 * sizeof(float), cudaMemcpyDeviceToHost, h->s_compute));
  const int tk = (int)(h->calls++ & 1);
  ERNET_CUDA(cudaEventRecord(h->ev_call[tk], h->s_compute));
  *ticket = tk;
  return ERNET_OK;
}

int ernet_classify_frames_host_wait(ernet_handle* h, int ticket) {
  if (!h || ticket < 0 || ticket > 1 || !h->ev_call[ticket]) return fail(ERNET_ERR_INVALID_ARG, "ernet_classify_frames_host_wait: bad ticket");
  DeviceGuard g(h->device);
  ERNET_CUDA(cudaEventSynchronize(h->ev_call[ticket]));
  const int wrc = check_watchdog("ernet_classify_frames_host_wait");
  if (wrc) cudaMemset(h->d_sync, 0, 2 * kSyncImages * sizeof(uint32_t));     // an aborted fused kernel leaves its band counters dirty
  return wrc;
}

int ernet_classify_frames_host(ernet_handle* h, const uint8_t* frames_host, int batch, int height, int width,
                               int channel_order, float* probs_host, float* logits_host) {
  int ticket = 0;
  int rc = ernet_classify_frames_host_submit(h, frames_host, batch, height, width, channel_order, probs_host, logits_host, &ticket);
  if (rc) return rc;
  return ernet_classify_frames_host_wait(h, ticket);
}

int ernet_acff_depthwise(const void* x, int dtype, int batch, int H, int W, int C, int out_h, int out_w,
                         const float* w, const float* b, void* out, void* stream) {
  if (!x || !w || !b || !out || batch < 1) return fail(ERNET_ERR_INVALID_ARG, "ernet_acff_depthwise: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static thread_local bool attrs = false;
  if (!attrs) { int rc = init_device_attrs(); if (rc) return rc; attrs = true; }
  for (int b0 = 0; b0 < batch; b0 += 32768) {
    const int n = batch - b0 < 32768 ? batch - b0 : 32768;
    const size_t xi = (size_t)b0 * H * W * C, oi = (size_t)b0 * out_h * out_w * 3 * C;
    int rc;
    if (dtype == ERNET_F32) rc = launch_acff_dw<float>(static_cast<const float*>(x) + xi, n, H, W, C, out_h, out_w, w, b, static_cast<float*>(out) + oi, s);
    else if (dtype == ERNET_F16) rc = launch_acff_dw<__half>(static_cast<const __half*>(x) + xi, n, H, W, C, out_h, out_w, w, b, static_cast<__half*>(out) + oi, s);
    else if (dtype == ERNET_BF16) rc = launch_acff_dw<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x) + xi, n, H, W, C, out_h, out_w, w, b, static_cast<__nv_bfloat16*>(out) + oi, s);
    else return fail(ERNET_ERR_INVALID_ARG, "bad dtype %d", dtype);
    if (rc) return rc;
  }
  return ERNET_OK;
}

int ernet_acff_add_depthwise(const void* x, int dtype, int batch, int H, int W, int C, int out_h, int out_w,
                             const float* w, const float* b, void* out, void* stream) {
  if (!x || !w || !b || !out || batch < 1 || C < 1) return fail(ERNET_ERR_INVALID_ARG, "ernet_acff_add_depthwise: bad argument");
  if (dtype != ERNET_F32) return fail(ERNET_ERR_INVALID_ARG, "ernet_acff_add_depthwise: fp32 only (dtype %d)", dtype);
  if (out_h > H - 2 || out_w > W - 2 || out_h < 1 || out_w < 1)
    return fail(ERNET_ERR_INVALID_ARG, "add-fusion depthwise: bad output size %dx%d for input %dx%d", out_h, out_w, H, W);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long per_img = (long long)((out_h + 3) / 4) * ((out_w + 3) / 4) * C;
  const int step = (int)std::max<long long>(1, std::min<long long>(batch, 0x60000000LL / per_img));   // one launch = < 2^31 threads
  for (int b0 = 0; b0 < batch; b0 += step) {
    const int n = batch - b0 < step ? batch - b0 : step;
    const size_t xi = (size_t)b0 * H * W * C, oi = (size_t)b0 * out_h * out_w * C;
    int rc = g_dw_fp32_form == 2 ? launch_acff_add_dw_tma(static_cast<const float*>(x) + xi, n, H, W, C, out_h, out_w, w, b,
                                                          static_cast<float*>(out) + oi, s) : -1;
    if (rc == -1)
      rc = launch_acff_dw_tile<4, 2, true>(static_cast<const float*>(x) + xi, n, H, W, C, out_h, out_w, w, b,
                                           static_cast<float*>(out) + oi, s, true);
    if (rc) return rc;
  }
  return ERNET_OK;
}

int ernet_confusion_update(const float* scores, const long long* targets, int batch, int num_classes,
                           long long* cm, long long* pred_out, unsigned long long* bad, void* stream) {
  if (!scores || batch < 1 || num_classes < 1 || (!cm != !targets) || (!cm && !pred_out))
    return fail(ERNET_ERR_INVALID_ARG, "ernet_confusion_update: bad argument");
  confusion_update_kernel<<<(batch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, targets, batch, num_classes, cm, pred_out, bad);
  ERNET_LAUNCH_CHECK("confusion_update_kernel");
  return ERNET_OK;
}

int ernet_set_depthwise_form(int form) {
  const int prev = g_dw_fp32_form;
  g_dw_fp32_form = form < 0 ? 0 : (form > 2 ? 2 : form);
  return prev;
}

int ernet_pointwise(const void* a, int dtype, int batch, int H, int W, int K, int N, const float* w, const float* bias,
                    const float* bn_scale, const float* bn_shift, int leaky, int pool, void* out, void* stream) {
  if (!a || !w || !out || batch < 1) return fail(ERNET_ERR_INVALID_ARG, "ernet_pointwise: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == ERNET_F32) return launch_pointwise<float>(static_cast<const float*>(a), batch, H, W, K, N, w, bias, bn_scale, bn_shift, leaky, pool, static_cast<float*>(out), s);
  if (dtype == ERNET_F16) return launch_pointwise<__half>(static_cast<const __half*>(a), batch, H, W, K, N, w, bias, bn_scale, bn_shift, leaky, pool, static_cast<__half*>(out), s);
  if (dtype == ERNET_BF16) return launch_pointwise<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(a), batch, H, W, K, N, w, bias, bn_scale, bn_shift, leaky, pool, static_cast<__nv_bfloat16*>(out), s);
  return fail(ERNET_ERR_INVALID_ARG, "bad dtype %d", dtype);
}

int ernet_debug_tap(ernet_handle* h, int tap, const void* workspace, int batch, float* out, size_t out_elems, void* stream) {
  if (!h || !workspace || !out) return fail(ERNET_ERR_INVALID_ARG, "ernet_debug_tap: null argument");
  if (batch < 1 || batch > h->chunk) return fail(ERNET_ERR_INVALID_ARG, "tap batch must be within one chunk");
  const Plan p = make_plan(h, batch);
  size_t off; int C, HW;
  switch (tap) {
    case ERNET_TAP_INGEST:
      if (!h->debug_taps) return fail(ERNET_ERR_UNSUPPORTED, "the fused transform+conv1 kernel keeps the transformed tensor on chip; call ernet_set_debug_taps(h, 1) before the forward");
      off = p.ingest; C = 3; HW = 140 * 140; break;
    case ERNET_TAP_STEM: off = p.stem; C = h->cs(); HW = 69 * 69; break;
    case ERNET_TAP_POOL1: off = p.p1; C = 64; HW = 33 * 33; break;
    case ERNET_TAP_POOL2: off = p.p2; C = h->c3(); HW = 15 * 15; break;
    case ERNET_TAP_POOL3: off = h->red() ? p.r3 : p.p3; C = h->c4(); HW = 36; break;
    case ERNET_TAP_ACFF4:
      if (h->has_tail && h->engine != ERNET_ENGINE_SIMT && !h->debug_taps)
        return fail(ERNET_ERR_UNSUPPORTED, "the fused ACFF4+head kernel keeps acff4 on chip; call ernet_set_debug_taps(h, 1) before the forward");
      off = p.a4; C = 256; HW = 16; break;
    default: return fail(ERNET_ERR_INVALID_ARG, "unknown tap %d", tap);
  }
  const long long total = (long long)batch * C * HW;
  if ((size_t)total != out_elems) return fail(ERNET_ERR_BAD_SHAPE, "tap %d has %lld elements, caller expects %zu", tap, total, out_elems);
  DeviceGuard g(h->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const char* src = static_cast<const char*>(workspace) + off;
  const int grid = (int)((total + 255) / 256);
  if (p.tc && h->precision == ERNET_PREC_INT8 && ((tap == ERNET_TAP_STEM && !h->red()) || tap == ERNET_TAP_POOL1 || tap == ERNET_TAP_POOL2)) {
    const int Hh = tap == ERNET_TAP_STEM ? 69 : (tap == ERNET_TAP_POOL1 ? 33 : 15);
    const int NCc = tap == ERNET_TAP_STEM ? 2 : (tap == ERNET_TAP_POOL1 ? 4 : (h->red() ? 4 : 6));
    const float* sc = h->f(ERNET_T_Q_SCALES) + (tap == ERNET_TAP_STEM ? 0 : (tap == ERNET_TAP_POOL1 ? 16 : 16 + 64));
    tc::tap_p16_to_nchw_f32<<<grid, 256, 0, s>>>(reinterpret_cast<const int8_t*>(src), NCc, C, Hh, total, sc, out);
    ERNET_LAUNCH_CHECK("tap_p16_to_nchw_f32");
    return ERNET_OK;
  }
  if (p.tc && (tap == ERNET_TAP_STEM || tap == ERNET_TAP_POOL1 || tap == ERNET_TAP_POOL2)) {
    const int Hh = tap == ERNET_TAP_STEM ? 69 : (tap == ERNET_TAP_POOL1 ? 33 : 15);
    const int NCc = tap == ERNET_TAP_STEM ? 2 : (tap == ERNET_TAP_POOL1 ? 8 : (h->red() ? 6 : 12));
    if (h->precision == ERNET_PREC_BF16) tc::tap_p8_to_nchw_f32<true><<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(src), NCc, C, Hh, total, out);
    else tc::tap_p8_to_nchw_f32<false><<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(src), NCc, C, Hh, total, out);
    ERNET_LAUNCH_CHECK("tap_p8_to_nchw_f32");
    return ERNET_OK;
  }
  if (h->precision == ERNET_PREC_FP32) tap_nhwc_to_nchw_f32<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(src), C, HW, total, out);
  else if (h->precision == ERNET_PREC_FP16 || h->precision == ERNET_PREC_INT8) tap_nhwc_to_nchw_f32<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(src), C, HW, total, out);
  else tap_nhwc_to_nchw_f32<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), C, HW, total, out);
  ERNET_LAUNCH_CHECK("tap_nhwc_to_nchw_f32");
  return ERNET_OK;
}

int ernet_ingest_tables_host(int height, int width, int* meta, int* xmin, int* xlen, int* kx, int* ymin, int* ylen,
                              int* ky, float* lut) {
  if (!meta || !xmin || !xlen || !kx || !ymin || !ylen || !ky) return fail(ERNET_ERR_INVALID_ARG, "null argument");
  IngestTables t;
  if (height < 1 || width < 1) return fail(ERNET_ERR_INVALID_ARG, "bad frame size %dx%d", height, width);
  if (width <= height) { t.new_w = kResizeShort; t.new_h = (int)((double)kResizeShort * height / width); }
  else                 { t.new_h = kResizeShort; t.new_w = (int)((double)kResizeShort * width / height); }
  if (t.new_h < kCrop || t.new_w < kCrop) return fail(ERNET_ERR_BAD_SHAPE, "resized frame smaller than crop");
  t.top = py_round_half_even((t.new_h - kCrop) / 2.0);
  t.left = py_round_half_even((t.new_w - kCrop) / 2.0);
  std::vector<int> vxmin, vxlen, vkx, vymin, vylen, vky;
  host_coeffs(width, t.new_w, t.left, kCrop, t.ksx, vxmin, vxlen, vkx);
  host_coeffs(height, t.new_h, t.top, kCrop, t.ksy, vymin, vylen, vky);
  if (t.ksx > kMaxTaps || t.ksy > kMaxTaps) return fail(ERNET_ERR_UNSUPPORTED, "too many taps");
  meta[0] = t.new_h; meta[1] = t.new_w; meta[2] = t.top; meta[3] = t.left; meta[4] = t.ksy; meta[5] = t.ksx;
  memcpy(xmin, vxmin.data(), kCrop * sizeof(int)); memcpy(xlen, vxlen.data(), kCrop * sizeof(int));
  memcpy(ymin, vymin.data(), kCrop * sizeof(int)); memcpy(ylen, vylen.data(), kCrop * sizeof(int));
  memcpy(kx, vkx.data(), vkx.size() * sizeof(int)); memcpy(ky, vky.data(), vky.size() * sizeof(int));
  if (lut) host_lut(lut);
  return ERNET_OK;
}

int ernet_debug_device_status(unsigned int* out8, int reset) {
  if (!out8) return fail(ERNET_ERR_INVALID_ARG, "null argument");
  ERNET_CUDA(cudaMemcpyFromSymbol(out8, tc::g_tc_status, 8 * sizeof(unsigned int)));
  if (reset) {
    unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ERNET_CUDA(cudaMemcpyToSymbol(tc::g_tc_status, z, sizeof(z)));
  }
  return ERNET_OK;
}

int ernet_check_watchdog(void) { return check_watchdog("ernet_check_watchdog"); }

int ernet_debug_timeline(unsigned long long* out, size_t count) {
#ifdef ERNET_TIMELINE
  if (!out || count > 3 * 148 * 32 * 8) return fail(ERNET_ERR_INVALID_ARG, "bad timeline request");
  ERNET_CUDA(cudaMemcpyFromSymbol(out, tc::g_timeline, count * sizeof(unsigned long long)));
  return ERNET_OK;
#else
  (void)out; (void)count;
  return fail(ERNET_ERR_UNSUPPORTED, "library was built without -DERNET_TIMELINE");
#endif
}

int ernet_debug_chain(unsigned long long* out32, int reset) {
#ifdef ERNET_TIMELINE
  if (out32) ERNET_CUDA(cudaMemcpyFromSymbol(out32, tc::g_chain, 32 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[32];
    for (int k = 0; k < 8; ++k) { z[4 * k] = z[4 * k + 1] = ~0ULL; z[4 * k + 2] = z[4 * k + 3] = 0ULL; }
    ERNET_CUDA(cudaMemcpyToSymbol(tc::g_chain, z, sizeof(z)));
  }
  return ERNET_OK;
#else
  (void)out32; (void)reset;
  return fail(ERNET_ERR_UNSUPPORTED, "library was built without -DERNET_TIMELINE");
#endif
}

int ernet_profile_enable(ernet_handle* h, int on) {
  if (!h) return fail(ERNET_ERR_INVALID_ARG, "null handle");
  h->profiling = on != 0;
  return ERNET_OK;
}

int ernet_profile_read(ernet_handle* h, double* ms_by_stage, int* launches_by_stage, int n_stages) {
  if (!h || !ms_by_stage || !launches_by_stage || n_stages < ERNET_STAGE_COUNT)
    return fail(ERNET_ERR_INVALID_ARG, "ernet_profile_read: need arrays of ERNET_STAGE_COUNT entries");
  DeviceGuard g(h->device);
  for (int i = 0; i < n_stages; ++i) { ms_by_stage[i] = 0.0; launches_by_stage[i] = 0; }
  for (auto& r : h->prof) {
    ERNET_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    ERNET_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_stage[r.stage] += ms;
    launches_by_stage[r.stage] += 1;
    h->prof_pool.push_back(r.a);
    h->prof_pool.push_back(r.b);
  }
  h->prof.clear();
  return ERNET_OK;
}

int ernet_launches_per_forward(const ernet_handle* h, int batch, int with_ingest) {
  if (!h || batch < 1) return 0;
  const int chunks = (batch + h->chunk - 1) / h->chunk;
  if (h->ernet()) return chunks * (h->use_tc() ? 8 : 14);   // conv1, 6 fused blocks, head | conv1, 6 x (depthwise, 1x1), head
  const bool tail = h->has_tail && h->engine != ERNET_ENGINE_SIMT;
  const int tail_launches = tail ? 1 : 3;
  // frames path: transform + conv1 are one kernel (two with debug taps on); tensor path: conv1 only
  const int front = with_ingest ? (h->debug_taps ? 2 : (h->last_fused ? 0 : 1)) : 1;   // fused: transform + conv1 ride in block 1's kernel
  const int per = h->use_tc() ? front + 3 + (h->red() ? 2 : 0) + tail_launches : front + 6 + (h->red() ? 2 : 0) + tail_launches;
  return chunks * per;
}

}  // extern "C"
