"""Which ACFF block can be int8 without losing top-1 agreement?  Exact numpy emulation (int8_study.py) with a subset
of the blocks quantised (weights + input activations), the others in float64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures
from oracle import ernet_numpy as E, ingest_numpy as I
import rtdm_b200.pack as P, rtdm_b200.pack_tc as PT
sys.argv = sys.argv[:1]
import importlib.util
spec = importlib.util.spec_from_file_location("s1", os.path.join(os.path.dirname(__file__), "int8_study.py"))

arch = "squeeze-ernet"


def conv25(xq, w, hu):
    B, C, H, _ = xq.shape
    xp = np.zeros((B, w.shape[2], H + 6, H + 6)); xp[:, :C, 2:2 + H, 2:2 + H] = xq
    out = np.zeros((B, w.shape[0], hu, hu))
    for t, (dy, dx) in enumerate(PT.TAPS):
        out += np.einsum("nc,bchw->bnhw", w[:, t, :], xp[:, :, 2 + dy:2 + dy + hu, 2 + dx:2 + dx + hu], optimize=True)
    return out


def pool(y):
    B, N, H, _ = y.shape
    return y.reshape(B, N, H // 2, 2, H // 2, 2).max(axis=(3, 5))


def run(sd, x, chan_scales, quant, wbits=8, abits=8):
    sd64 = {k: np.asarray(v, np.float64) for k, v in sd.items()}
    a = E.conv2d_dense(x.astype(np.float64), sd64["conv1.weight"], None, 2)
    hus = [66, 30, 12]
    for k in range(3):
        c, co = P.widths(arch)[k]
        weff, beff = PT.fold_block(sd, f"acff{k+1}", c, max(32, c))
        if k in quant:
            amax = 2 ** (abits - 1) - 1
            s_in = chan_scales[k] * 127.0 / amax
            q = np.clip(np.rint(a / s_in.reshape(1, -1, 1, 1)), -amax, amax)
            weff = weff.copy(); weff[:, :, :c] *= s_in.reshape(1, 1, -1)
            wmax = 2 ** (wbits - 1) - 1
            am = np.abs(weff).reshape(weff.shape[0], -1).max(axis=1)
            s_w = np.maximum(am, 1e-30) / wmax
            wq = np.clip(np.rint(weff / s_w[:, None, None]), -wmax, wmax)
            z = conv25(q, wq, hus[k]) * s_w.reshape(1, -1, 1, 1) + beff.reshape(1, -1, 1, 1)
        else:
            z = conv25(a, weff, hus[k]) + beff.reshape(1, -1, 1, 1)
        z = np.maximum(z, 0.01 * z)
        p = f"acff{k+1}.batch_norm"
        s = sd64[f"{p}.weight"] / np.sqrt(sd64[f"{p}.running_var"] + 1e-5)
        t = sd64[f"{p}.bias"] - sd64[f"{p}.running_mean"] * s
        a = pool(z * s.reshape(1, -1, 1, 1) + t.reshape(1, -1, 1, 1))
    z = E.acff(a, sd64, "acff4")
    out = E.conv2d_pointwise(z, sd64["conv2.weight"], None)
    out = E.avg_pool_5x5_s1_p1(out)
    return out.reshape(-1, 20) @ sd64["fc.weight"].T + sd64["fc.bias"]


def calib(sd, xcal):
    taps = E.forward(sd, xcal, arch, dtype=np.float64, want_taps=True)["taps"]
    out = []
    for name in ("stem", "pool1", "pool2"):
        v = np.abs(taps[name]); vc = v.transpose(1, 0, 2, 3).reshape(v.shape[1], -1)
        out.append(np.maximum(vc.max(axis=1), 1e-12) / 127.0)
    return out


n = int(os.environ.get("N", "96"))
for wset in ["shipped"]:
    sd = fixtures.get_state_dict(arch, wset)
    fcal = np.concatenate([fixtures.noise_frames(12, seed=99), fixtures.smooth_frames(4, seed=98)], 0)
    ftest = np.concatenate([fixtures.noise_frames(n // 2, seed=61), fixtures.smooth_frames(n // 2, seed=62)], 0)
    xcal, xtest = I.ingest(fcal), I.ingest(ftest)
    ref = E.forward(sd, xtest, arch, dtype=np.float64)["logits"]
    srt = np.sort(ref, 1); margin = (srt[:, -1] - srt[:, -2]) / np.abs(ref).max()
    print(wset, "ref top1 hist", np.bincount(ref.argmax(1), minlength=5), "margin min/1%/5%/median", margin.min(), np.percentile(margin, 1), np.percentile(margin, 5), np.median(margin))
    cs = calib(sd, xcal)
    for quant, wb, ab in [((0,), 8, 8), ((1,), 8, 8), ((2,), 8, 8), ((0, 1, 2), 8, 8), ((0,), 8, 16), ((0,), 16, 8), ((0, 1, 2), 16, 8), ((0, 1, 2), 8, 16)]:
        lg = run(sd, xtest, cs, quant, wb, ab)
        agree = (lg.argmax(1) == ref.argmax(1)).mean()
        err = np.abs(lg - ref).max() / np.abs(ref).max()
        flips = np.where(lg.argmax(1) != ref.argmax(1))[0]
        print(f"  blocks {quant} w{wb} a{ab}: agreement {agree:.4f}  rel logit err {err:.3e}  flip margins {np.round(margin[flips], 4)[:8]}", flush=True)
