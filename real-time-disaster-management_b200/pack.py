"""state_dict -> packed device blob (csrc/blob_format.h).

Consumes exactly the reference's checkpoint format: the 56-key (Squeeze_ErNET) / 62-key
(Squeeze_RedConv) state_dict defined by model/squeeze_ernet.py:8-22,
model/squeeze_ernet_redconv.py:8-25 and model/acff.py:25-35.  All derived tensors are computed in
fp64 and rounded once.

Exact algebraic rewrites applied here:
  * BatchNorm (eval) -> per-channel scale/shift applied after LeakyReLU (acff.py:52-53);
  * Squeeze_RedConv: conv_red1(conv1(x)) has no nonlinearity in between
    (squeeze_ernet_redconv.py:28-29) -> one 3x3/s2 conv 3->8 with bias;
  * conv2 -> AvgPool2d(5,1,1) -> view -> fc collapses, at 140x140 input, to
    logits = W_eff . sum_{4x4}(acff4) + b_fc   (squeeze_ernet.py:33-40).
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = 0x424E5245
VERSION = 2
BN_EPS = 1e-5

# tensor ids, mirrored from csrc/blob_format.h (tests/test_pack.py checks they agree)
T_STEM_W, T_STEM_B = 0, 1
T_BLOCK_BASE = 8
T_DW_W, T_DW_B, T_PW_W, T_PW_B, T_BN_S, T_BN_T = 0, 1, 2, 3, 4, 5
T_RED2_W, T_RED2_B, T_RED3_W, T_RED3_B = 40, 41, 42, 43
T_HEAD_W, T_HEAD_B = 44, 45
T_TC_BASE = 64

T_EBLOCK_BASE = 88          # ErNET blocks 5, 6 (k = 4, 5): 88 + 8*(k-4)
T_EHEAD_W = 104             # ErNET head: conv2 o AvgPool(5,1,0) o view o fc collapsed to [5][49][256]

ARCH_ID = {"squeeze-ernet": 0, "squeeze-redconv": 1, "ernet": 2}
PREC_ID = {"fp32": 0, "fp16": 1, "bf16": 2, "int8": 3}
DT_F32, DT_F16, DT_BF16, DT_U8, DT_RAW = 0, 1, 2, 3, 16


def widths(arch):
    if arch == "ernet":                       # model/ernet.py:11-19
        return [(16, 64), (64, 96), (96, 128), (128, 128), (128, 128), (128, 256)]
    return [(8, 64), (64, 96), (48, 128), (64, 256)] if arch == "squeeze-redconv" else \
           [(16, 64), (64, 96), (96, 128), (128, 256)]


def block_base(k):
    """blob id base of ACFF block k (0-based)."""
    return T_BLOCK_BASE + 8 * k if k < 4 else T_EBLOCK_BASE + 8 * (k - 4)


def expected_shapes(arch):
    red = arch == "squeeze-redconv"
    ks = {"conv1.weight": (16, 3, 3, 3)}
    if red:
        ks.update({"conv_red1.weight": (8, 16, 1, 1), "conv_red1.bias": (8,),
                   "conv_red2.weight": (48, 96, 1, 1), "conv_red2.bias": (48,),
                   "conv_red3.weight": (64, 128, 1, 1), "conv_red3.bias": (64,)})
    for k, (c, co) in enumerate(widths(arch), start=1):
        p = f"acff{k}"
        for j in (1, 2, 3):
            ks[f"{p}.conv{j}.weight"] = (c, 1, 3, 3)
            ks[f"{p}.conv{j}.bias"] = (c,)
        ks[f"{p}.fused_conv.weight"] = (co, 3 * c, 1, 1)
        ks[f"{p}.fused_conv.bias"] = (co,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            ks[f"{p}.batch_norm.{n}"] = (co,)
        ks[f"{p}.batch_norm.num_batches_tracked"] = ()
    ks.update({"conv2.weight": (5, 256, 1, 1), "fc.weight": (5, 45 if arch == "ernet" else 20), "fc.bias": (5,)})
    return ks


def _np64(sd, key):
    v = sd[key]
    if hasattr(v, "detach"):
        v = v.detach().to("cpu").double().numpy()
    return np.asarray(v, dtype=np.float64)


def validate_state_dict(sd, arch):
    want = expected_shapes(arch)
    missing = [k for k in want if k not in sd]
    unexpected = [k for k in sd if k not in want]
    if missing or unexpected:
        raise ValueError(f"state_dict does not match {arch}: missing={missing[:4]} unexpected={unexpected[:4]}")
    for k, shp in want.items():
        got = tuple(sd[k].shape)
        if got != tuple(shp):
            raise ValueError(f"size mismatch for {k}: expected {tuple(shp)}, got {got}")


def derive_simt(sd, arch):
    """fp64 tensors of the CUDA-core path, keyed by blob id."""
    red = arch == "squeeze-redconv"
    out = {}
    w1 = _np64(sd, "conv1.weight")                                   # (16,3,3,3) [o][c][ky][kx]
    if red:
        wr = _np64(sd, "conv_red1.weight")[:, :, 0, 0]               # (8,16)
        w1 = np.einsum("po,ocyx->pcyx", wr, w1)                      # (8,3,3,3)
        b1 = _np64(sd, "conv_red1.bias")
    else:
        b1 = np.zeros(16)
    out[T_STEM_W] = np.transpose(w1, (2, 3, 1, 0)).copy()            # [ky][kx][c][o]
    out[T_STEM_B] = b1
    for k, (c, co) in enumerate(widths(arch)):
        p = f"acff{k + 1}"
        base = block_base(k)
        dw = np.stack([_np64(sd, f"{p}.conv{j}.weight")[:, 0].reshape(c, 9).T for j in (1, 2, 3)], 0)  # [3][9][C]
        db = np.stack([_np64(sd, f"{p}.conv{j}.bias") for j in (1, 2, 3)], 0)                          # [3][C]
        pw = _np64(sd, f"{p}.fused_conv.weight")[:, :, 0, 0].T.copy()                                   # [3C][N]
        g, b = _np64(sd, f"{p}.batch_norm.weight"), _np64(sd, f"{p}.batch_norm.bias")
        mu, var = _np64(sd, f"{p}.batch_norm.running_mean"), _np64(sd, f"{p}.batch_norm.running_var")
        s = g / np.sqrt(var + BN_EPS)
        out[base + T_DW_W], out[base + T_DW_B] = dw, db
        out[base + T_PW_W], out[base + T_PW_B] = pw, _np64(sd, f"{p}.fused_conv.bias")
        out[base + T_BN_S], out[base + T_BN_T] = s, b - mu * s
    if red:
        out[T_RED2_W] = _np64(sd, "conv_red2.weight")[:, :, 0, 0].T.copy()   # [96][48]
        out[T_RED2_B] = _np64(sd, "conv_red2.bias")
        out[T_RED3_W] = _np64(sd, "conv_red3.weight")[:, :, 0, 0].T.copy()   # [128][64]
        out[T_RED3_B] = _np64(sd, "conv_red3.bias")
    wc2 = _np64(sd, "conv2.weight")[:, :, 0, 0]                              # (5,256)
    if arch == "ernet":
        # conv2 (1x1, no bias) -> AvgPool2d(5, stride 1, no padding) on the 7x7 map -> view(-1, 45) -> fc (ernet.py:35-43):
        # all linear, so logits = b + sum_{y,x,k} W_eff[o][y][x][k] * acff6[k][y][x] with
        # W_eff[o][y][x][k] = sum_c conv2[c][k] * (1/25) * sum_{i,j : 0 <= y-i < 5, 0 <= x-j < 5} fc[o][c*9 + i*3 + j]
        wfc = _np64(sd, "fc.weight").reshape(5, 5, 3, 3)                     # [o][c][i][j]
        cover = np.zeros((5, 5, 7, 7))
        for i in range(3):
            for j in range(3):
                cover[:, :, i:i + 5, j:j + 5] += wfc[:, :, i, j][:, :, None, None]
        out[T_EHEAD_W] = np.einsum("ocyx,ck->oyxk", cover, wc2).reshape(5, 49, 256) / 25.0
        out[T_HEAD_B] = _np64(sd, "fc.bias")
        return out
    wfc = _np64(sd, "fc.weight").reshape(5, 5, 4).sum(axis=2)                # (5 out, 5 conv2-ch): sum of the 2x2 taps
    out[T_HEAD_W] = (wfc @ wc2) / 25.0                                       # (5,256)
    out[T_HEAD_B] = _np64(sd, "fc.bias")
    return out


def assemble(arch, precision, tensors):
    """tensors: id -> (np.ndarray with final dtype, dtype code)."""
    ids = sorted(tensors)
    header_bytes = 32 + 24 * len(ids)
    off = (header_bytes + 255) // 256 * 256
    table, payload = [], []
    for i in ids:
        arr, code = tensors[i]
        raw = np.ascontiguousarray(arr).tobytes()
        table.append((i, code, off, len(raw)))
        pad = (-len(raw)) % 256
        payload.append(raw + b"\0" * pad)
        off += len(raw) + pad
    blob = struct.pack("<8I", MAGIC, VERSION, ARCH_ID[arch], PREC_ID[precision], len(ids), 0, 0, 0)
    for i, code, o, n in table:
        blob += struct.pack("<IIQQ", i, code, o, n)
    blob += b"\0" * ((-len(blob)) % 256)
    blob += b"".join(payload)
    return blob


def pack_state_dict(sd, arch, precision, act_scales=None):
    """Build the blob for ernet_load_packed().  ``act_scales`` (int8 only): the three calibrated
    per-tensor activation scales, see pack_tc.derive_tc_int8."""
    if arch not in ARCH_ID:
        raise ValueError(f"Unsupported model: {arch}")                 # aider-predict.py:32
    if precision not in PREC_ID:
        raise ValueError(f"unknown precision {precision}")
    validate_state_dict(sd, arch)
    tensors = {i: (v.astype(np.float32), DT_F32) for i, v in derive_simt(sd, arch).items()}
    if arch == "ernet" and precision == "int8":
        raise ValueError("int8 is implemented for squeeze-ernet and squeeze-redconv")
    if precision in ("fp16", "bf16", "int8"):
        from . import pack_tc
        tensors.update(pack_tc.derive_tc(sd, arch, precision, act_scales))
    return assemble(arch, precision, tensors)
