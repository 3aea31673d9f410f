"""CPU oracle for the Squeeze-ErNet / Squeeze-ErNet-RedConv forward pass (numpy only).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``real-time-disaster-management_b200``) never
routes through this file and fails loudly when its CUDA library is missing.

It restates, layer by layer, what the reference computes with stock ``torch.nn``
modules (the arithmetic itself lives in third-party PyTorch: pinned by the reference
at torch==1.7.1+cu110, requirements-fyp.txt:192; the conv / pool / linear / softmax
definitions restated here are the published ``torch.nn`` semantics and are stable
across versions).  Citations are relative to ``code/disaster_detection/`` in the
reference repository:

* ``Squeeze_ErNET.forward``        model/squeeze_ernet.py:24-46
* ``Squeeze_RedConv.forward``      model/squeeze_ernet_redconv.py:27-52
* ``ACFF.forward``                 model/acff.py:37-59 (modules built at acff.py:25-35)

Parity pinning: ``tests/golden/make_golden.py`` imports the *real* reference classes
from ``/root/reference`` (in the build container), runs them in fp64 and fp32 on seeded
inputs with the shipped checkpoints and with trained-like random weights, and commits the
outputs as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this file
against every one of those vectors, so the oracle is pinned to outputs of the reference
itself (the reference ships no tests or golden vectors of its own).

All functions work on NCHW numpy arrays in whatever float dtype the input has
(fp64 for error budgeting, fp32 for the like-for-like check).
"""
from __future__ import annotations

import numpy as np

LEAKY_SLOPE = 0.01   # nn.LeakyReLU(0.01), model/acff.py:33
BN_EPS = 1e-5        # nn.BatchNorm2d default eps, model/acff.py:34

ARCH_SQUEEZE = "squeeze-ernet"      # names used by aider-predict.py:25-30
ARCH_REDCONV = "squeeze-redconv"


# --------------------------------------------------------------------------- primitives
def conv2d_dense(x, w, b=None, stride=1):
    """nn.Conv2d(padding=0, dilation=1, groups=1): cross-correlation, no flip.

    x (B,Cin,H,W), w (Cout,Cin,kH,kW).  Used for conv1 (model/squeeze_ernet.py:11)."""
    B, Cin, H, W = x.shape
    Cout, _, kH, kW = w.shape
    Ho = (H - kH) // stride + 1
    Wo = (W - kW) // stride + 1
    out = np.zeros((B, Cout, Ho, Wo), dtype=x.dtype)
    for i in range(kH):
        for j in range(kW):
            xs = x[:, :, i:i + stride * (Ho - 1) + 1:stride, j:j + stride * (Wo - 1) + 1:stride]
            out += np.einsum("oc,bchw->bohw", w[:, :, i, j], xs, optimize=True)
    if b is not None:
        out += b.reshape(1, -1, 1, 1)
    return out


def conv2d_pointwise(x, w, b=None):
    """1x1 nn.Conv2d == per-pixel matrix product (model/acff.py:31, squeeze_ernet.py:19)."""
    out = np.einsum("oc,bchw->bohw", w[:, :, 0, 0], x, optimize=True)
    if b is not None:
        out = out + b.reshape(1, -1, 1, 1)
    return out


def conv2d_depthwise3x3(x, w, b, dilation):
    """nn.Conv2d(C, C, 3, padding=dilation-1, dilation=dilation, groups=C, bias=True).

    model/acff.py:25-30.  Output is (H-2, W-2) for every dilation; taps outside the
    input read the zero padding."""
    B, C, H, W = x.shape
    p = dilation - 1
    xp = np.zeros((B, C, H + 2 * p, W + 2 * p), dtype=x.dtype)
    xp[:, :, p:p + H, p:p + W] = x
    Ho, Wo = H - 2, W - 2
    out = np.zeros((B, C, Ho, Wo), dtype=x.dtype)
    for i in range(3):
        for j in range(3):
            out += w[:, 0, i, j].reshape(1, C, 1, 1) * xp[:, :, i * dilation:i * dilation + Ho,
                                                          j * dilation:j * dilation + Wo]
    return out + b.reshape(1, C, 1, 1)


def leaky_relu(x):
    return np.where(x >= 0, x, x * x.dtype.type(LEAKY_SLOPE))


def batch_norm_eval(x, gamma, beta, mean, var):
    """nn.BatchNorm2d in eval mode: running statistics (model/acff.py:34,53)."""
    t = x.dtype.type
    inv = gamma / np.sqrt(var + t(BN_EPS))
    return (x - mean.reshape(1, -1, 1, 1)) * inv.reshape(1, -1, 1, 1) + beta.reshape(1, -1, 1, 1)


def max_pool_2x2(x):
    """nn.MaxPool2d(2, 2): floor mode drops an odd last row/column (squeeze_ernet.py:13)."""
    Hp, Wp = x.shape[2] // 2, x.shape[3] // 2
    a = x[:, :, 0:2 * Hp:2, 0:2 * Wp:2]
    b = x[:, :, 0:2 * Hp:2, 1:2 * Wp:2]
    c = x[:, :, 1:2 * Hp:2, 0:2 * Wp:2]
    d = x[:, :, 1:2 * Hp:2, 1:2 * Wp:2]
    return np.maximum(np.maximum(a, b), np.maximum(c, d))


def avg_pool_5x5_s1_p1(x):
    """nn.AvgPool2d(5, stride=1, padding=1), count_include_pad=True -> always /25.

    model/squeeze_ernet.py:20."""
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:1 + H, 1:1 + W] = x
    Ho, Wo = H + 2 - 5 + 1, W + 2 - 5 + 1
    out = np.zeros((B, C, Ho, Wo), dtype=x.dtype)
    for i in range(5):
        for j in range(5):
            out += xp[:, :, i:i + Ho, j:j + Wo]
    return out / x.dtype.type(25)


def avg_pool_5x5_s1_p0(x):
    """nn.AvgPool2d(kernel_size=5, stride=1, padding=0) of ErNET (model/ernet.py:21): (B,C,H,W) -> (B,C,H-4,W-4)."""
    B, C, H, W = x.shape
    out = np.zeros((B, C, H - 4, W - 4), dtype=x.dtype)
    for i in range(5):
        for j in range(5):
            out += x[:, :, i:i + H - 4, j:j + W - 4]
    return out / x.dtype.type(25)


def softmax_dim1(z):
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------- blocks
def acff_concat(x, sd, prefix):
    """The three dilated depthwise branches concatenated on channels (acff.py:46)."""
    outs = []
    for k, d in ((1, 1), (2, 2), (3, 3)):
        outs.append(conv2d_depthwise3x3(x, sd[f"{prefix}.conv{k}.weight"], sd[f"{prefix}.conv{k}.bias"], d))
    return np.concatenate(outs, axis=1)


def acff(x, sd, prefix, taps=None):
    """ACFF.forward (acff.py:37-59): concat -> 1x1 -> LeakyReLU -> BN -> Dropout(eval: id)."""
    cat = acff_concat(x, sd, prefix)
    if taps is not None:
        taps[f"{prefix}.cat"] = cat
    z = conv2d_pointwise(cat, sd[f"{prefix}.fused_conv.weight"], sd[f"{prefix}.fused_conv.bias"])
    z = leaky_relu(z)
    z = batch_norm_eval(z, sd[f"{prefix}.batch_norm.weight"], sd[f"{prefix}.batch_norm.bias"],
                        sd[f"{prefix}.batch_norm.running_mean"], sd[f"{prefix}.batch_norm.running_var"])
    if taps is not None:
        taps[prefix] = z
    return z


def _cast_sd(sd, dtype):
    out = {}
    for k, v in sd.items():
        a = np.asarray(v)
        out[k] = a.astype(dtype) if a.dtype.kind == "f" else a
    return out


def forward(sd, x, arch, dtype=None, want_taps=False):
    """Whole network.  Returns dict(logits, probs[, taps]).

    ``sd``: mapping of the reference's state_dict keys to arrays (56 keys for
    squeeze-ernet, 62 for squeeze-redconv).
    ``x``: (B,3,140,140) NCHW.  A non-140 spatial size is rejected the way the
    reference's ``view(-1, 20)`` effectively does (squeeze_ernet.py:39)."""
    if arch == "ernet":
        return forward_ernet(sd, x, dtype, want_taps)
    x = np.asarray(x)
    if dtype is None:
        dtype = x.dtype
    dtype = np.dtype(dtype)
    if x.ndim != 4 or x.shape[1] != 3 or x.shape[2] != 140 or x.shape[3] != 140:
        raise ValueError(f"expected (B,3,140,140), got {x.shape}")
    sd = _cast_sd(sd, dtype)
    x = x.astype(dtype)
    taps = {} if want_taps else None
    red = arch == ARCH_REDCONV
    if not red and arch != ARCH_SQUEEZE:
        raise ValueError(f"Unsupported model: {arch}")          # aider-predict.py:32

    out = conv2d_dense(x, sd["conv1.weight"], None, stride=2)     # squeeze_ernet.py:25
    if red:
        out = conv2d_pointwise(out, sd["conv_red1.weight"], sd["conv_red1.bias"])   # redconv.py:29
    if taps is not None:
        taps["stem"] = out
    out = acff(out, sd, "acff1", taps)
    out = max_pool_2x2(out)
    if taps is not None:
        taps["pool1"] = out
    out = acff(out, sd, "acff2", taps)
    if red:
        out = conv2d_pointwise(out, sd["conv_red2.weight"], sd["conv_red2.bias"])   # redconv.py:33
    out = max_pool_2x2(out)
    if taps is not None:
        taps["pool2"] = out
    out = acff(out, sd, "acff3", taps)
    out = max_pool_2x2(out)
    if red:
        out = conv2d_pointwise(out, sd["conv_red3.weight"], sd["conv_red3.bias"])   # redconv.py:37
    if taps is not None:
        taps["pool3"] = out
    out = acff(out, sd, "acff4", taps)
    out = conv2d_pointwise(out, sd["conv2.weight"], None)         # squeeze_ernet.py:33
    out = avg_pool_5x5_s1_p1(out)                                 # :34
    flat = out.reshape(-1, 2 * 2 * 5)                             # :39  (c*4 + i*2 + j)
    logits = flat @ sd["fc.weight"].T + sd["fc.bias"]             # :40
    probs = softmax_dim1(logits)                                  # :41
    res = {"logits": logits, "probs": probs}
    if taps is not None:
        res["taps"] = taps
    return res


ARCH_ERNET = "ernet"


def forward_ernet(sd, x, dtype=None, want_taps=False):
    """Baseline ErNET (model/ernet.py:6-49): conv1 -> ACFF(16,64) pool ACFF(64,96) pool ACFF(96,128) pool
    ACFF(128,128) ACFF(128,128) ACFF(128,256) -> conv2 1x1 -> AvgPool(5,1,0) -> view(-1,45) -> fc -> softmax.
    ``x``: (B,3,240,240) NCHW (the only size for which view(-1, 5*3*3) keeps one row per sample, ernet.py:42)."""
    x = np.asarray(x)
    dtype = np.dtype(x.dtype if dtype is None else dtype)
    if x.ndim != 4 or x.shape[1:] != (3, 240, 240):
        raise ValueError(f"expected (B,3,240,240), got {x.shape}")
    sd = _cast_sd(sd, dtype)
    taps = {} if want_taps else None
    out = conv2d_dense(x.astype(dtype), sd["conv1.weight"], None, stride=2)     # ernet.py:25
    if taps is not None:
        taps["stem"] = out
    for k in (1, 2, 3):
        out = max_pool_2x2(acff(out, sd, f"acff{k}", taps))                       # :26-31
        if taps is not None:
            taps[f"pool{k}"] = out
    for k in (4, 5, 6):
        out = acff(out, sd, f"acff{k}", taps)                                     # :32-34
    out = conv2d_pointwise(out, sd["conv2.weight"], None)                         # :35
    out = avg_pool_5x5_s1_p0(out)                                                 # :36
    flat = out.reshape(-1, 5 * 3 * 3)                                             # :42  (c*9 + i*3 + j)
    logits = flat @ sd["fc.weight"].T + sd["fc.bias"]                             # :43
    res = {"logits": logits, "probs": softmax_dim1(logits)}
    if taps is not None:
        res["taps"] = taps
    return res


# --------------------------------------------------------------------------- bookkeeping
def expected_keys(arch):
    """state_dict key -> shape, as the reference constructors define them
    (squeeze_ernet.py:8-22, squeeze_ernet_redconv.py:8-25, acff.py:25-35)."""
    red = arch == ARCH_REDCONV
    widths = [(8, 64), (64, 96), (48, 128), (64, 256)] if red else [(16, 64), (64, 96), (96, 128), (128, 256)]
    keys = {"conv1.weight": (16, 3, 3, 3)}
    if red:
        keys["conv_red1.weight"] = (8, 16, 1, 1)
        keys["conv_red1.bias"] = (8,)
    for k, (c, co) in enumerate(widths, start=1):
        p = f"acff{k}"
        for j in (1, 2, 3):
            keys[f"{p}.conv{j}.weight"] = (c, 1, 3, 3)
            keys[f"{p}.conv{j}.bias"] = (c,)
        keys[f"{p}.fused_conv.weight"] = (co, 3 * c, 1, 1)
        keys[f"{p}.fused_conv.bias"] = (co,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            keys[f"{p}.batch_norm.{n}"] = (co,)
        keys[f"{p}.batch_norm.num_batches_tracked"] = ()
        if red and k == 2:
            keys["conv_red2.weight"] = (48, 96, 1, 1)
            keys["conv_red2.bias"] = (48,)
        if red and k == 3:
            keys["conv_red3.weight"] = (64, 128, 1, 1)
            keys["conv_red3.bias"] = (64,)
    keys["conv2.weight"] = (5, 256, 1, 1)
    keys["fc.weight"] = (5, 20)
    keys["fc.bias"] = (5,)
    return keys


def count_params(arch):
    """Trainable parameter count (model_summary/squeeze_ernet.txt:45 = 169,241;
    model_summary/squeeze_redconv.txt:48 = 109,569)."""
    n = 0
    for k, s in expected_keys(arch).items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            continue
        n += int(np.prod(s)) if s else 1
    return n


def count_macs(arch):
    """Multiply-accumulates per 140x140 image."""
    red = arch == ARCH_REDCONV
    widths = [(8, 64), (64, 96), (48, 128), (64, 256)] if red else [(16, 64), (64, 96), (96, 128), (128, 256)]
    macs = 69 * 69 * 16 * 27
    if red:
        macs += 69 * 69 * 16 * 8
    h = 69
    for k, (c, co) in enumerate(widths):
        ho = h - 2
        macs += ho * ho * c * 27 + ho * ho * 3 * c * co
        if red and k == 1:
            macs += ho * ho * 96 * 48
        h = ho // 2 if k < 3 else ho
        if red and k == 2:
            macs += h * h * 128 * 64
    macs += 4 * 4 * 256 * 5 + 20 * 5
    return macs
