"""Operand images for the tensor-core (tcgen05) ACFF block kernels (csrc/tc_block.cuh).

Each ACFF block (model/acff.py:25-31,46,51) is linear up to the LeakyReLU, so its three dilated
depthwise convs, the concat and the 1x1 conv fold into one dense conv over the 25 distinct taps:

    W_eff[n, tap, c] = sum_{d in {1,2,3} : tap in stencil d} W_f[n, (d-1)*C + c] * w_d[c, ky, kx]
    b_eff[n]         = b_f[n] + sum_{d,c} W_f[n, (d-1)*C + c] * b_d[c]

computed in fp64 and rounded once to bf16 / fp16.  The image is laid out exactly as the kernel's
shared-memory B operand: [tap][channel chunk of 8][n][8], so one 1-D bulk copy per tap stages it.
"""
from __future__ import annotations

import numpy as np

T_TC_BASE = 64
T_ETC_BASE = 106     # ErNET: + 2*k, k = 0..5 -> (weight image, bias) of acff1..acff6
T_TC_WIMG, T_TC_BIAS, T_TC_DEQ = 0, 1, 2   # + 4*k for block k
T_Q_SCALES = 80
T_TC4_WIMG = 76
T_TC_RED2_WIMG, T_TC_RED2_BIAS = 77, 78
DT_F32, DT_RAW = 0, 16

# offsets from the output coordinate, sorted by (dy, dx); mirrored by kTapDy/kTapDx in csrc/tc_common.cuh
TAPS = sorted({(ky * d - (d - 1), kx * d - (d - 1)) for d in (1, 2, 3) for ky in range(3) for kx in range(3)})
assert len(TAPS) == 25


def to_bits16(a, precision):
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    if precision == "fp16":
        return a32.astype(np.float16).view(np.uint16)
    u = a32.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)      # round-to-nearest-even bf16
    return r


def from_bits16(bits, precision):
    bits = np.asarray(bits, dtype=np.uint16)
    if precision == "fp16":
        return bits.view(np.float16).astype(np.float64)
    return (bits.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def fold_block(sd, prefix, c_real, c_pad):
    """-> (W_eff [N][25][c_pad] fp64, b_eff [N] fp64)."""
    from .pack import _np64
    wf = _np64(sd, f"{prefix}.fused_conv.weight")[:, :, 0, 0]               # (N, 3C)
    n_out = wf.shape[0]
    weff = np.zeros((n_out, 25, c_pad))
    beff = _np64(sd, f"{prefix}.fused_conv.bias").copy()
    for d in (1, 2, 3):
        wd = _np64(sd, f"{prefix}.conv{d}.weight")[:, 0]                     # (C,3,3)
        bd = _np64(sd, f"{prefix}.conv{d}.bias")
        wfd = wf[:, (d - 1) * c_real:d * c_real]                             # (N, C)
        beff += wfd @ bd
        for ky in range(3):
            for kx in range(3):
                tap = TAPS.index((ky * d - (d - 1), kx * d - (d - 1)))
                weff[:, tap, :c_real] += wfd * wd[:, ky, kx][None, :]
    return weff, beff


def weight_image(weff, precision):
    """[N][taps][C] -> uint16 image [taps][C/8][N][8]."""
    n_out, taps, c = weff.shape
    img = weff.transpose(1, 2, 0).reshape(taps, c // 8, 8, n_out).transpose(0, 1, 3, 2)   # [tap][chunk][n][8]
    return to_bits16(img, precision)


def quantize_weights(weff):
    """TensorRT-style symmetric per-output-channel int8: s_w[n] = max|W_eff[n]| / 127, q = round(W / s_w).
    -> (int8 [N][25][C], s_w [N])."""
    amax = np.abs(weff).reshape(weff.shape[0], -1).max(axis=1)
    s_w = np.where(amax > 0, amax / 127.0, 1.0)
    q = np.clip(np.rint(weff / s_w[:, None, None]), -127, 127).astype(np.int8)
    return q, s_w


def weight_image_i8(wq):
    """int8 [N][25][C] (C multiple of 32) -> image [25][C/16][N][16]: the K-major operand with 16 int8
    channels per 16-byte chunk."""
    n_out, _, c = wq.shape
    return np.ascontiguousarray(wq.transpose(1, 2, 0).reshape(25, c // 16, 16, n_out).transpose(0, 1, 3, 2))


ACT_CHANNELS = (16, 64, 96)      # channels of the three quantised tensors: stem, pool1, pool2 outputs (Squeeze_ErNET)
ACT_CHANNELS_RED = (8, 64, 48)   # Squeeze_RedConv: conv_red1 / pool1 / pool2 (after conv_red2) outputs


def act_channels(arch):
    return ACT_CHANNELS_RED if arch == "squeeze-redconv" else ACT_CHANNELS


def normalize_act_scales(act_scales, arch="squeeze-ernet"):
    """-> three fp64 arrays of per-channel int8 steps.  Accepts three scalars (plain per-tensor scales) or
    three arrays (per-channel equalised)."""
    if act_scales is None or len(act_scales) != 3:
        raise ValueError("act_scales must have three entries (stem, pool1, pool2)")
    out = []
    for v, c in zip(act_scales, act_channels(arch)):
        a = np.asarray(v, dtype=np.float64)
        a = np.full(c, float(a)) if a.ndim == 0 else a.reshape(-1)
        if a.shape != (c,) or not np.all(np.isfinite(a)) or np.any(a <= 0):
            raise ValueError(f"activation scales for a {c}-channel tensor must be {c} positive numbers")
        out.append(a)
    return out


def derive_tc_int8(sd, arch, act_scales):
    """int8 operand images.  ``act_scales[k][c]`` = real value of one int8 step of channel c of the k-th
    quantised tensor (stem, pool1, pool2 outputs; from calibration).  The producer divides channel c by its
    step in its epilogue (for pool1/pool2 that is a change of the BN affine constants), the consumer's
    folded weights are multiplied by it, so at run time the int8 tensor has ONE scale (1.0) - the
    per-channel part is cross-layer equalisation done at pack time."""
    from .pack import widths, _np64
    if arch not in ("squeeze-ernet", "squeeze-redconv"):
        raise ValueError("int8 is implemented for squeeze-ernet and squeeze-redconv")
    s_act = normalize_act_scales(act_scales, arch)
    out = {}
    for k, (c, _co) in enumerate(widths(arch)[:3]):
        if arch == "squeeze-redconv" and k == 0:
            # Squeeze_RedConv's ACFF1 stays in fp16: its input has 8 channels = ONE 16-byte chunk per pixel at 16 bit, so
            # tap pairing already issues 13 MMAs per tile - exactly what an int8 block 1 would issue.  int8 would buy
            # nothing and would push the whole image through an 8-channel int8 bottleneck (measured: 13/15 on the real
            # frames with it).  The block's OUTPUT (pool1) is still quantised for ACFF2.
            weff, beff = fold_block(sd, "acff1", c, 16)
            base = T_TC_BASE
            out[base + T_TC_WIMG] = (weight_image(weff, "fp16"), DT_RAW)
            out[base + T_TC_BIAS] = (beff.astype(np.float32), DT_F32)
            out[base + T_TC_DEQ] = (np.ones(weff.shape[0], dtype=np.float32), DT_F32)
            s_act[0] = np.ones_like(s_act[0])       # the stem tensor is fp16: no int8 step
            continue
        c_pad = (c + 31) // 32 * 32             # K step of kind::i8 is 32 (RedConv: 48 -> 64)
        weff, beff = fold_block(sd, f"acff{k + 1}", c, c_pad)
        weff[:, :, :c] *= s_act[k].reshape(1, 1, -1)
        wq, s_w = quantize_weights(weff)
        base = T_TC_BASE + 4 * k
        out[base + T_TC_WIMG] = (weight_image_i8(wq), DT_RAW)
        out[base + T_TC_BIAS] = (beff.astype(np.float32), DT_F32)
        out[base + T_TC_DEQ] = (s_w.astype(np.float32), DT_F32)
    # the handle keeps the steps at fixed offsets [16][64][96]; unused tails (RedConv: 8 of 16, 48 of 96) are 1.0
    qs = np.ones(16 + 64 + 96, dtype=np.float64)
    for off, a in zip((0, 16, 80), s_act):
        qs[off:off + len(a)] = a
    out[T_Q_SCALES] = (qs.astype(np.float32), DT_F32)
    if arch == "squeeze-redconv":
        # conv_red2 (squeeze_ernet_redconv.py:16,33) stays a 16-bit 1-tap instance (fp16 weights, fp16 un-pooled ACFF2
        # output in, int8 pool2 tensor out); N padded to 64: the 16 extra outputs are exact zeros = the zero half of the
        # fourth 16-channel chunk that block 3's K step of 32 wants
        wr = np.zeros((64, 1, 96))
        wr[:48, 0, :] = _np64(sd, "conv_red2.weight")[:, :, 0, 0]
        br = np.zeros(64)
        br[:48] = _np64(sd, "conv_red2.bias")
        out[T_TC_RED2_WIMG] = (weight_image(wr, "fp16"), DT_RAW)
        out[T_TC_RED2_BIAS] = (br.astype(np.float32), DT_F32)
    return out


def tail_weight_image(sd, precision):
    """acff4.fused_conv.weight (256, 3*C4, 1, 1) -> uint16 image [3*C4/8][256][8] (K-major B operand of the
    ACFF4+head kernel; K order = concat order [branch][channel], acff.py:46)."""
    from .pack import _np64
    w = _np64(sd, "acff4.fused_conv.weight")[:, :, 0, 0]                  # (256, K)
    n, k = w.shape
    return to_bits16(w.reshape(n, k // 8, 8).transpose(1, 0, 2), precision)


def derive_tc(sd, arch, precision, act_scales=None):
    from .pack import widths
    if precision == "int8":
        if act_scales is None:
            raise ValueError("int8 needs calibrated activation scales (model.calibrate(frames))")
        out = derive_tc_int8(sd, arch, act_scales)
        out[T_TC4_WIMG] = (tail_weight_image(sd, "fp16"), DT_RAW)        # the int8 engine's tail runs in fp16
        return out
    if precision not in ("fp16", "bf16"):
        return {}
    out = {}
    if arch == "ernet":
        # all six ACFF blocks of model/ernet.py:12-19 in the 25-tap dense form; the head stays on the CUDA cores
        for k, (c, _co) in enumerate(widths(arch)):
            weff, beff = fold_block(sd, f"acff{k + 1}", c, max(16, c))
            out[T_ETC_BASE + 2 * k] = (weight_image(weff, precision), DT_RAW)
            out[T_ETC_BASE + 2 * k + 1] = (beff.astype(np.float32), DT_F32)
        return out
    for k, (c, _co) in enumerate(widths(arch)[:3]):
        c_pad = max(16, c)
        weff, beff = fold_block(sd, f"acff{k + 1}", c, c_pad)
        base = T_TC_BASE + 4 * k
        out[base + T_TC_WIMG] = (weight_image(weff, precision), DT_RAW)
        out[base + T_TC_BIAS] = (beff.astype(np.float32), DT_F32)
    out[T_TC4_WIMG] = (tail_weight_image(sd, precision), DT_RAW)
    if arch == "squeeze-redconv":
        # conv_red2 (squeeze_ernet_redconv.py:16,33) as a 1-tap instance of the block kernel: 96 -> 48, N padded to 64
        from .pack import _np64
        wr = np.zeros((64, 1, 96))
        wr[:48, 0, :] = _np64(sd, "conv_red2.weight")[:, :, 0, 0]
        br = np.zeros(64)
        br[:48] = _np64(sd, "conv_red2.bias")
        out[T_TC_RED2_WIMG] = (weight_image(wr, precision), DT_RAW)
        out[T_TC_RED2_BIAS] = (br.astype(np.float32), DT_F32)
    return out
