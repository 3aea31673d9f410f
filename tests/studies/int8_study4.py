"""int8 study 4 (round 2): the UN-FOLDED TensorRT-style scheme against the 25-tap-folded one.

What TensorRT builds for an ACFF block (model/acff.py:25-35,46-53) when `--quant int8` is asked
(build_tensorrt_model.py:256-259): every convolution is its own int8 layer.
  * depthwise conv d (d = 1, 2, 3):  int8 input (per-TENSOR activation scale s_x), int8 weights with one scale
    per output channel (max|w_c| / 127), int32 accumulate, fp32 bias, output re-quantised to int8 with the
    scale of the CONCAT tensor (all three branches share it: a concat is a no-op only if its inputs agree);
  * fused 1x1 conv: int8 input (s_cat), int8 weights per output channel, int32 accumulate, fp32 bias,
    LeakyReLU, BatchNorm as a per-channel scale layer, 2x2 max-pool, re-quantised with the next block's s_x.
Activation scales come from a calibration set: max |x| or a percentile (99.99 / 99.999), per tensor.
conv1, ACFF4 and the head stay in float ("first conv and head in >= fp16").

Everything is emulated exactly (integer-valued float64 tensors through torch.nn.functional.conv2d), on the
three input sets of SURVEY 8d: I1 uniform-noise frames, I2 smooth frames, I3 the 17 real JPEGs of the
reference tree (only when /root/reference is present).  Prints one line per scheme: top-1 agreement with the
fp64 reference on each input set and the relative logit error.

    python tests/studies/int8_study4.py [N_PER_SET=2048]
"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F

import fixtures
from oracle import ingest_numpy as I
import rtdm_b200.pack as P
import rtdm_b200.pack_tc as PT

torch.set_grad_enabled(False)
ARCH = sys.argv[2] if len(sys.argv) > 2 else "squeeze-ernet"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
DT = torch.float64


def tsd(sd):
    return {k: torch.as_tensor(np.asarray(v)).to(DT) if np.asarray(v).dtype.kind == "f" else torch.as_tensor(np.asarray(v)) for k, v in sd.items()}


def bn_affine(sd, p):
    s = sd[f"{p}.batch_norm.weight"] / torch.sqrt(sd[f"{p}.batch_norm.running_var"] + 1e-5)
    return s, sd[f"{p}.batch_norm.bias"] - sd[f"{p}.batch_norm.running_mean"] * s


def q(x, step, bits=8):
    m = 2 ** (bits - 1) - 1
    return torch.clamp(torch.round(x / step), -m, m)


def qw(w, bits=8):
    """per-output-channel symmetric weight quantisation -> (integer weights, scale[n])"""
    m = 2 ** (bits - 1) - 1
    amax = w.abs().reshape(w.shape[0], -1).max(1).values.clamp_min(1e-30)
    s = amax / m
    return torch.clamp(torch.round(w / s.reshape(-1, 1, 1, 1)), -m, m), s


def acff_float(x, sd, p):
    c = x.shape[1]
    br = [F.conv2d(x, sd[f"{p}.conv{j}.weight"], sd[f"{p}.conv{j}.bias"], 1, j - 1, j, c) for j in (1, 2, 3)]
    z = F.conv2d(torch.cat(br, 1), sd[f"{p}.fused_conv.weight"], sd[f"{p}.fused_conv.bias"])
    s, t = bn_affine(sd, p)
    return F.leaky_relu(z, 0.01) * s.reshape(1, -1, 1, 1) + t.reshape(1, -1, 1, 1)


def acff_unfolded_q(x, sd, p, s_x, s_cat, cat_per_channel=False):
    """TensorRT-style: x float -> int8(s_x) -> 3 depthwise int8 convs -> int8 concat (s_cat) -> int8 1x1 -> float"""
    c = x.shape[1]
    xq = q(x, s_x)
    br = []
    for j in (1, 2, 3):
        wq_, sw = qw(sd[f"{p}.conv{j}.weight"])
        acc = F.conv2d(xq, wq_, None, 1, j - 1, j, c)                      # exact integers
        br.append(acc * (sw * s_x).reshape(1, -1, 1, 1) + sd[f"{p}.conv{j}.bias"].reshape(1, -1, 1, 1))
    cat = torch.cat(br, 1)
    step = s_cat.reshape(1, -1, 1, 1) if cat_per_channel else s_cat
    cq = q(cat, step)
    wf = sd[f"{p}.fused_conv.weight"]
    if cat_per_channel:                                                     # cross-layer equalisation: fold the per-channel step into the 1x1 weights
        wf = wf * s_cat.reshape(1, -1, 1, 1)
        wq_, sw = qw(wf)
        z = F.conv2d(cq, wq_) * sw.reshape(1, -1, 1, 1)
    else:
        wq_, sw = qw(wf)
        z = F.conv2d(cq, wq_) * (sw * s_cat).reshape(1, -1, 1, 1)
    z = z + sd[f"{p}.fused_conv.bias"].reshape(1, -1, 1, 1)
    s, t = bn_affine(sd, p)
    return F.leaky_relu(z, 0.01) * s.reshape(1, -1, 1, 1) + t.reshape(1, -1, 1, 1)


def conv25(xq, weff, hu):
    """25-tap dense form on integer tensors: weff [N][25][C] scattered into a 7x7 kernel (zeros elsewhere)"""
    B, C, H, _ = xq.shape
    w7 = torch.zeros((weff.shape[0], C, 7, 7), dtype=xq.dtype)
    for t, (dy, dx) in enumerate(PT.TAPS):
        w7[:, :, dy + 2, dx + 2] = weff[:, t, :C].to(xq.dtype)
    return F.conv2d(F.pad(xq, (2, 4, 2, 4)), w7)[:, :, :hu, :hu]


def acff_folded_q(x, sd_np, sd, p, c, s_x, hu, per_channel):
    weff, beff = PT.fold_block(sd_np, p, c, c)
    weff = torch.from_numpy(weff).to(x.dtype)
    step = s_x.reshape(-1) if per_channel else s_x.reshape(1).expand(c)
    xq = q(x, step.reshape(1, -1, 1, 1))
    w = weff * step.reshape(1, 1, -1)
    amax = w.abs().reshape(w.shape[0], -1).max(1).values.clamp_min(1e-30)
    sw = amax / 127.0
    wq_ = torch.clamp(torch.round(w / sw.reshape(-1, 1, 1)), -127, 127)
    z = conv25(xq, wq_, hu) * sw.reshape(1, -1, 1, 1) + torch.from_numpy(beff).to(x.dtype).reshape(1, -1, 1, 1)
    s, t = bn_affine(sd, p)
    return F.leaky_relu(z, 0.01) * s.reshape(1, -1, 1, 1) + t.reshape(1, -1, 1, 1)


def head(a, sd):
    out = acff_float(a, sd, "acff4")
    out = F.avg_pool2d(F.conv2d(out, sd["conv2.weight"]), 5, 1, 1)
    return F.linear(out.reshape(-1, 20), sd["fc.weight"], sd["fc.bias"])


def run(sd_np, sd, x, scheme, cal):
    """scheme: 'float' | 'unfolded' | 'unfolded_pc' (per-channel concat step, equalised) | 'folded' | 'folded_pc'"""
    red = ARCH == "squeeze-redconv"
    a = F.conv2d(x, sd["conv1.weight"], None, 2)
    if red:
        a = F.conv2d(a, sd["conv_red1.weight"], sd["conv_red1.bias"])
    taps = {}
    hus = [66, 30, 12]
    for k in range(3):
        p = f"acff{k + 1}"
        c = a.shape[1]
        taps[f"in{k}"] = a
        if scheme == "float":
            if cal is not None:                                    # record the concat tensor for calibration
                br = [F.conv2d(a, sd[f"{p}.conv{j}.weight"], sd[f"{p}.conv{j}.bias"], 1, j - 1, j, c) for j in (1, 2, 3)]
                taps[f"cat{k}"] = torch.cat(br, 1)
            y = acff_float(a, sd, p)
        elif scheme.startswith("unfolded"):
            y = acff_unfolded_q(a, sd, p, cal[f"in{k}"], cal[f"cat{k}_pc" if scheme.endswith("_pc") else f"cat{k}"], scheme.endswith("_pc"))
        else:
            y = acff_folded_q(a, sd_np, sd, p, c, cal[f"in{k}_pc" if scheme.endswith("_pc") else f"in{k}"], a.shape[2] - 2, scheme.endswith("_pc"))
        if red and k == 1:
            y = F.conv2d(y, sd["conv_red2.weight"], sd["conv_red2.bias"])
        a = F.max_pool2d(y, 2, 2)
        if red and k == 2:
            a = F.conv2d(a, sd["conv_red3.weight"], sd["conv_red3.bias"])
    return head(a, sd), taps


_TAPS_CACHE = {}


def calibrate(sd_np, sd, xcal, pct, key):
    if key not in _TAPS_CACHE:
        _TAPS_CACHE.clear()
        parts = [run(sd_np, sd, xcal[i:i + 128], "float", {})[1] for i in range(0, xcal.shape[0], 128)]
        _TAPS_CACHE[key] = {n: torch.cat([p_[n] for p_ in parts], 0) for n in parts[0]}
    taps = _TAPS_CACHE[key]
    cal = {}
    for name, t in taps.items():
        v = t.abs()
        vc = v.permute(1, 0, 2, 3).reshape(v.shape[1], -1)
        if pct >= 100:
            cal[name] = vc.max() / 127.0
            cal[name + "_pc"] = vc.max(1).values.clamp_min(float(vc.max()) * 1e-6) / 127.0
        else:
            flat = vc.flatten()
            flat = flat[:: max(1, flat.numel() // 4_000_000)]
            cal[name] = torch.quantile(flat, pct / 100.0) / 127.0
            sub = vc[:, :: max(1, vc.shape[1] // 200_000)]
            cal[name + "_pc"] = torch.quantile(sub, pct / 100.0, dim=1).clamp_min(float(vc.max()) * 1e-6) / 127.0
    return cal


def batched(fn, x, bs=128):
    return torch.cat([fn(x[i:i + bs]) for i in range(0, x.shape[0], bs)], 0)


def real_frames():
    pats = ["/root/reference/code/victim_localization/yolov3/data/custom/test/images/*.jpg",
            "/root/reference/code/victim_localization/yolov5/dataset/*/images/*.jpg"]
    files = sorted({os.path.basename(f): f for p in pats for f in glob.glob(p)}.values())
    if not files:
        return None
    from PIL import Image
    return [np.asarray(Image.open(f).convert("RGB")) for f in files]


def main():
    sets = {"I1 noise": I.ingest(fixtures.noise_frames(N, seed=61)), "I2 smooth": I.ingest(fixtures.smooth_frames(N, seed=62))}
    rf = real_frames()
    if rf:
        sets[f"I3 real({len(rf)})"] = np.concatenate([I.ingest(f[None]) for f in rf], 0)
    fcal = np.concatenate([fixtures.noise_frames(256, seed=99), fixtures.smooth_frames(256, seed=98)], 0)   # 512 frames, SURVEY 8d config 4
    xcal = torch.from_numpy(I.ingest(fcal)).to(DT)
    for wset in os.environ.get("WSETS", "shipped,w3,w3neg").split(","):
        sd_np = fixtures.get_state_dict(ARCH, wset)
        sd = tsd(sd_np)
        print(f"== {ARCH} / {wset}", flush=True)
        refs = {}
        for name, x in sets.items():
            xt = torch.from_numpy(x).to(DT)
            refs[name] = (xt, batched(lambda b: run(sd_np, sd, b, "float", None)[0], xt))
            lg = refs[name][1]
            srt = torch.sort(lg, 1).values
            margin = (srt[:, -1] - srt[:, -2]) / lg.abs().max()
            print(f"   {name}: reference top-1 histogram {np.bincount(lg.argmax(1).numpy(), minlength=5).tolist()}, "
                  f"top-2 margin/|logit|max: min {margin.min():.2e}, 0.1% {torch.quantile(margin, 0.001):.2e}, 1% {torch.quantile(margin, 0.01):.2e}, median {margin.median():.2e}", flush=True)
        for pct in [float(v) for v in os.environ.get("PCTS", "100,99.999,99.99").split(",")]:
            cal = calibrate(sd_np, sd, xcal, pct, wset)
            schemes = ["unfolded", "unfolded_pc"] + (["folded", "folded_pc"] if ARCH == "squeeze-ernet" else [])
            for scheme in schemes:
                line = f"   calib {pct:>7}% {scheme:<12}"
                for name, (xt, ref) in refs.items():
                    lg = batched(lambda b: run(sd_np, sd, b, scheme, cal)[0], xt)
                    same = lg.argmax(1) == ref.argmax(1)
                    err = float((lg - ref).abs().max() / ref.abs().max())
                    srt = torch.sort(ref, 1).values
                    margin = (srt[:, -1] - srt[:, -2]) / ref.abs().max()
                    big = int(((~same) & (margin > 2e-2)).sum())          # flips that are not near-ties of the reference
                    line += f" | {name}: agree {float(same.double().mean()):.4f} ({int((~same).sum())} flips/{len(same)}, {big} with margin>2e-2) err {err:.2e}"
                print(line, flush=True)


if __name__ == "__main__":
    main()
