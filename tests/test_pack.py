"""Packer: blob structure, id agreement with the C header, and the algebraic folds (BN affine,
conv_red1 merge, head collapse) validated against the oracle on CPU."""
import os
import re

import numpy as np
import pytest

import fixtures
import packed_eval
import rtdm_b200.pack as P
from oracle import ernet_numpy as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ids_match_header():
    hdr = open(os.path.join(ROOT, "real-time-disaster-management_b200", "csrc", "blob_format.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+ERNET_(\w+)\s+(0x[0-9A-Fa-f]+|\d+)u?", hdr)}
    assert defs["BLOB_MAGIC"] == P.MAGIC and defs["BLOB_VERSION"] == P.VERSION
    for name in ("T_STEM_W", "T_STEM_B", "T_BLOCK_BASE", "T_DW_W", "T_DW_B", "T_PW_W", "T_PW_B", "T_BN_S", "T_BN_T",
                 "T_RED2_W", "T_RED2_B", "T_RED3_W", "T_RED3_B", "T_HEAD_W", "T_HEAD_B", "T_TC_BASE"):
        assert defs[name] == getattr(P, name), name


@pytest.mark.parametrize("arch", fixtures.ARCHS)
@pytest.mark.parametrize("wset", ("shipped", "w3neg"))
def test_packed_form_matches_oracle(arch, wset):
    sd = fixtures.get_state_dict(arch, wset)
    blob = P.pack_state_dict(sd, arch, "fp32")
    x = fixtures.normal_tensors(2, seed=3)
    logits, taps = packed_eval.forward_packed(blob, x, arch)
    ref = E.forward(sd, x, arch, dtype=np.float64, want_taps=True)
    scale = np.abs(ref["logits"]).max()
    assert np.abs(logits - ref["logits"]).max() <= 2e-6 * scale        # weights were rounded to fp32 once
    for mine, theirs in (("stem", "stem"), ("pool1", "pool1"), ("pool2", "pool2"), ("pool3", "pool3")):
        r = np.transpose(ref["taps"][theirs], (0, 2, 3, 1))
        assert taps[mine].shape == r.shape
        assert np.abs(taps[mine] - r).max() <= 2e-6 * np.abs(r).max(), mine


def test_rejects_bad_state_dicts():
    sd = fixtures.get_state_dict("squeeze-ernet", "w3")
    with pytest.raises(ValueError):
        P.pack_state_dict(sd, "ernet", "fp32")
    with pytest.raises(ValueError):
        P.pack_state_dict(sd, "squeeze-redconv", "fp32")       # wrong key set
    bad = dict(sd)
    bad["conv1.weight"] = bad["conv1.weight"][:8]
    with pytest.raises(ValueError):
        P.pack_state_dict(bad, "squeeze-ernet", "fp32")
    with pytest.raises(ValueError):
        P.pack_state_dict(sd, "squeeze-ernet", "fp64")


def test_blob_layout():
    sd = fixtures.get_state_dict("squeeze-redconv", "w3")
    pz = packed_eval.parse_blob(P.pack_state_dict(sd, "squeeze-redconv", "fp32"))
    assert pz["arch"] == 1 and pz["precision"] == 0
    assert P.T_RED2_W in pz["tensors"] and P.T_HEAD_W in pz["tensors"]
    assert len(pz["tensors"][P.T_STEM_W][1]) == 27 * 8 * 4


# ---------------------------------------------------------------- tensor-core operand images
def test_tap_table_matches_device_header():
    import rtdm_b200.pack_tc as PT
    hdr = open(os.path.join(ROOT, "real-time-disaster-management_b200", "csrc", "tc_common.cuh")).read()
    dy = [int(v) for v in re.search(r"kTapDy\[25\]\s*=\s*\{([^}]*)\}", hdr).group(1).split(",")]
    dx = [int(v) for v in re.search(r"kTapDx\[25\]\s*=\s*\{([^}]*)\}", hdr).group(1).split(",")]
    assert list(zip(dy, dx)) == PT.TAPS


@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_25_tap_fold_is_exact(arch):
    """depthwise trio + concat + 1x1 == one dense 25-tap conv with the folded weights (fp64)."""
    import rtdm_b200.pack_tc as PT
    sd = fixtures.get_state_dict(arch, "w3neg")
    for k, (c, co) in enumerate(P.widths(arch)[:3], start=1):
        rs = np.random.RandomState(k)
        H = {1: 19, 2: 13, 3: 15}[k]
        x = rs.standard_normal((2, c, H, H))
        cat = E.acff_concat(x, {kk: np.asarray(v, np.float64) for kk, v in sd.items() if kk.startswith(f"acff{k}.conv")}, f"acff{k}")
        ref = E.conv2d_pointwise(cat, np.asarray(sd[f"acff{k}.fused_conv.weight"], np.float64),
                                 np.asarray(sd[f"acff{k}.fused_conv.bias"], np.float64))
        weff, beff = PT.fold_block(sd, f"acff{k}", c, max(16, c))
        Ho = H - 2
        xp = np.zeros((2, max(16, c), H + 6, H + 6))
        xp[:, :c, 2:2 + H, 2:2 + H] = x
        z = np.zeros((2, co, Ho, Ho)) + beff.reshape(1, -1, 1, 1)
        for t, (dy, dx) in enumerate(PT.TAPS):
            z += np.einsum("nc,bchw->bnhw", weff[:, t, :], xp[:, :, 2 + dy:2 + dy + Ho, 2 + dx:2 + dx + Ho])
        assert np.abs(z - ref).max() <= 1e-11 * np.abs(ref).max()
        # image layout [tap][chunk][n][8] round-trips
        img = PT.weight_image(weff, "bf16")
        assert img.shape == (25, max(16, c) // 8, co, 8) and img.dtype == np.uint16
        back = PT.from_bits16(img, "bf16").transpose(2, 0, 1, 3).reshape(co, 25, -1)
        assert np.abs(back - weff).max() <= 2 ** -8 * np.abs(weff).max()
        img16 = PT.weight_image(weff, "fp16")
        back16 = PT.from_bits16(img16, "fp16").transpose(2, 0, 1, 3).reshape(co, 25, -1)
        assert np.abs(back16 - weff).max() <= 2 ** -10 * np.abs(weff).max()


def test_bf16_rounding_is_nearest_even():
    import rtdm_b200.pack_tc as PT
    import torch
    x = np.random.RandomState(0).standard_normal(4096).astype(np.float32) * 3
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(PT.to_bits16(x, "bf16"), want)


@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_int8_operand_images(arch):
    """int8 blobs: per-output-channel symmetric weights (max|q| = 127 in every non-zero row), image sizes the kernels
    expect, activation steps at their fixed offsets; Squeeze_RedConv keeps ACFF1 in fp16 (its tap-paired 16-bit form issues
    the int8 MMA count) and carries conv_red2 as a 16-bit 1-tap image padded to 64 outputs."""
    import rtdm_b200.pack_tc as PT
    sd = fixtures.get_state_dict(arch, "w3")
    chans = PT.act_channels(arch)
    rs = np.random.RandomState(3)
    scales = [rs.uniform(0.01, 0.05, c) for c in chans]
    out = PT.derive_tc_int8(sd, arch, scales)
    red = arch == "squeeze-redconv"
    cpad = [(c + 31) // 32 * 32 for c, _ in P.widths(arch)[:3]]
    for k, (c, co) in enumerate(P.widths(arch)[:3]):
        img, _ = out[PT.T_TC_BASE + 4 * k + PT.T_TC_WIMG]
        deq = out[PT.T_TC_BASE + 4 * k + PT.T_TC_DEQ][0]
        assert deq.shape == (co,) and out[PT.T_TC_BASE + 4 * k + PT.T_TC_BIAS][0].shape == (co,)
        if red and k == 0:
            assert img.dtype == np.uint16 and img.shape == (25, 2, co, 8) and np.all(deq == 1.0)      # fp16 block 1
            continue
        assert img.dtype == np.int8 and img.shape == (25, cpad[k] // 16, co, 16)
        rows = img.transpose(2, 0, 1, 3).reshape(co, -1).astype(np.int32)
        assert np.abs(rows).max(axis=1).min() == 127 and np.abs(rows).max() == 127
        assert np.all(rows.reshape(co, 25, -1)[:, :, c:] == 0)                                      # zero-padded channels
        # dequantised weights reproduce the folded, activation-scaled weights to half an int8 step
        weff, _ = PT.fold_block(sd, f"acff{k + 1}", c, cpad[k])
        weff[:, :, :c] *= np.asarray(scales[k]).reshape(1, 1, -1)
        back = rows.reshape(co, 25, -1) * deq.astype(np.float64).reshape(-1, 1, 1)
        assert np.abs(back - weff).max() <= 0.5001 * deq.max()
    qs = out[PT.T_Q_SCALES][0]
    assert qs.shape == (16 + 64 + 96,)
    for off, c, s in zip((0, 16, 80), chans, scales):
        want = np.ones(c) if (red and off == 0) else s
        assert np.allclose(qs[off:off + c], want.astype(np.float32))
    assert (PT.T_TC_RED2_WIMG in out) == red
    if red:
        assert out[PT.T_TC_RED2_WIMG][0].shape == (1, 12, 64, 8) and np.all(out[PT.T_TC_RED2_BIAS][0][48:] == 0)
    with pytest.raises(ValueError):
        PT.derive_tc_int8(sd, arch, scales[:2])


# ------------------------------------------------------------------------------------ numerics of the fp32 engine's splits
def _tf32_trunc(a):
    """What tcgen05 kind::tf32 reads of an fp32 operand: the top 19 bits (sign, exponent, 10 mantissa bits)."""
    return (np.asarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def test_split_tf32_product_is_fp32_accurate():
    """csrc/tc_pw32.cuh: v = hi + lo with hi = v & 0xFFFFE000 (exactly a TF32 number) and lo = v - hi (exact in fp32);
    a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi drops only lo x lo.  Emulated with the operand truncation of the tensor core:
    a K = 288 dot product of trained-like values agrees with the fp64 result to fp32 rounding level (a single TF32 MMA
    does not: ~1e-3), which is what lets the fp32 engine keep the 1e-4 logit bound of north_star."""
    rs = np.random.RandomState(5)
    a = (rs.standard_normal((64, 288)) * np.exp(rs.uniform(-3, 3, (64, 288)))).astype(np.float32)
    b = (rs.standard_normal((288, 32)) * 0.05).astype(np.float32)
    a_hi = _tf32_trunc(a); a_lo = (a - a_hi).astype(np.float32)
    b_hi = _tf32_trunc(b); b_lo = (b - b_hi).astype(np.float32)
    assert np.array_equal(a_hi.astype(np.float64) + a_lo.astype(np.float64), a.astype(np.float64))       # the split is exact
    t = lambda x: _tf32_trunc(x).astype(np.float64)                                                       # noqa: E731
    three = t(a_lo) @ t(b_hi) + t(a_hi) @ t(b_lo) + t(a_hi) @ t(b_hi)
    one = t(a) @ t(b)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    scale = (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))                                 # sum of |terms|
    assert (np.abs(three - ref) / scale).max() <= 2.0 ** -19
    assert (np.abs(one - ref) / scale).max() >= 2.0 ** -13                                                # why one MMA is not enough


@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_split_fp16_stem_weights_are_fp32_accurate(arch):
    """csrc/ingest_fast.cuh (fp32 engine, 5-tap frames): conv1 runs on mma.sync from the uint8 pixels (exact in fp16) against
    two fp16 weight images, w' = hi + lo / 2^16 with hi = fp16(w'), lo = fp16((w' - hi) * 2^16), ToTensor + Normalize folded
    into w' and the bias.  Emulated in numpy on the real conv1 (+conv_red1) weights: every folded weight is represented to
    2^-21 relative (weights below the fp16 normal range: 2^-35 absolute), and conv1 of uint8 windows agrees with the reference's normalise-then-convolve to fp32 rounding level."""
    sd = fixtures.get_state_dict(arch, "shipped")
    w = np.asarray(sd["conv1.weight"], np.float64)                       # (16, 3, 3, 3) OIHW
    bias = np.asarray(sd["conv1.bias"], np.float64) if "conv1.bias" in sd else np.zeros(w.shape[0])
    if arch == "squeeze-redconv":                                        # conv_red1 (1x1, no nonlinearity in between) folds in
        r = np.asarray(sd["conv_red1.weight"], np.float64)[:, :, 0, 0]   # (8, 16)
        bias = r @ bias + np.asarray(sd["conv_red1.bias"], np.float64)
        w = np.einsum("oc,cikl->oikl", r, w)
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    wf = w / (255.0 * std)[None, :, None, None]                          # folded weights
    bf = bias - (w * (mean / std)[None, :, None, None]).sum((1, 2, 3))
    hi = wf.astype(np.float32).astype(np.float16)
    lo = ((wf - hi.astype(np.float64)) * 65536.0).astype(np.float32).astype(np.float16)
    rec = hi.astype(np.float64) + lo.astype(np.float64) / 65536.0
    # 2^-21 relative; below the fp16 normal range (2^-14) hi is subnormal and the error is bounded absolutely instead
    assert (np.abs(rec - wf) <= 2.0 ** -21 * np.maximum(np.abs(wf), 2.0 ** -14)).all()
    rs = np.random.RandomState(9)
    px = rs.randint(0, 256, (500, 3, 3, 3)).astype(np.float64)           # 500 uint8 windows (C, ky, kx)
    got = np.einsum("ockl,nckl->no", rec, px) + bf
    x = (px / 255.0 - mean[None, :, None, None]) / std[None, :, None, None]
    ref = np.einsum("ockl,nckl->no", w, x) + bias
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_pool_first_epilogue_is_exact():
    """csrc/tc_pw32.cuh epilogue: the 2x2 max runs on the RAW accumulators and bias / LeakyReLU / BN on the pooled quarter.
    max commutes with a non-decreasing map; for channels with a negative BN scale the packer negates the weight column
    (accumulator = -acc, exactly) and the epilogue multiplies the pooled value by -1 again.  In fp32, same operation
    order as the kernel: bit-identical to activate-then-pool, including zero and negative scales and ties."""
    rs = np.random.RandomState(11)
    acc = (rs.standard_normal((4, 8, 8, 32)) * 30).astype(np.float32)
    acc[0, :2, :2, :] = acc[0, 0, 0, :]                                   # a window of ties
    bias = rs.standard_normal(32).astype(np.float32)
    s = rs.standard_normal(32).astype(np.float32)
    s[3] = 0.0
    t = rs.standard_normal(32).astype(np.float32)

    def act(z):                                                            # bias added by the caller; LeakyReLU(0.01) then BN affine
        z = np.maximum(z, np.float32(0.01) * z)
        return z * s + t                                                   # numpy: mul then add, both rounded - same for both orders

    def pool(x):
        return np.maximum(np.maximum(x[:, 0::2, 0::2], x[:, 0::2, 1::2]), np.maximum(x[:, 1::2, 0::2], x[:, 1::2, 1::2]))

    want = pool(act(acc + bias))
    sgn = np.where(s < 0, np.float32(-1), np.float32(1))
    got = act(pool(acc * sgn) * sgn + bias)                                # acc * sgn = what the negated weight column accumulates
    assert np.array_equal(got, want)
