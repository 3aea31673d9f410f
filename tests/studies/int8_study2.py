import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "studies"))
import numpy as np
import fixtures
from oracle import ernet_numpy as E, ingest_numpy as I
import rtdm_b200.pack as P, rtdm_b200.pack_tc as PT
from int8_study import conv25, pool, calib, arch

def run(sd, x, cs, qa, qw):
    """qa / qw: sets of block indices whose input activations / weights are quantised."""
    sd64 = {k: np.asarray(v, np.float64) for k, v in sd.items()}
    a = E.conv2d_dense(x.astype(np.float64), sd64["conv1.weight"], None, 2)
    hus = [66, 30, 12]
    for k in range(3):
        c, co = P.widths(arch)[k]
        s_in = cs[k]
        if k in qa:
            q = np.clip(np.rint(a / s_in.reshape(1, -1, 1, 1)), -127, 127)
        else:
            q = a / s_in.reshape(1, -1, 1, 1)
        weff, beff = PT.fold_block(sd, f"acff{k+1}", c, max(32, c))
        weff = weff.copy(); weff[:, :, :c] *= s_in.reshape(1, 1, -1)
        if k in qw:
            wq, s_w = PT.quantize_weights(weff); w = wq.astype(np.float64) * s_w[:, None, None]
        else:
            w = weff
        z = conv25(q, w, hus[k]) + beff.reshape(1, -1, 1, 1)
        z = np.maximum(z, 0.01 * z)
        p = f"acff{k+1}.batch_norm"
        s = sd64[f"{p}.weight"] / np.sqrt(sd64[f"{p}.running_var"] + 1e-5)
        t = sd64[f"{p}.bias"] - sd64[f"{p}.running_mean"] * s
        a = pool(z * s.reshape(1, -1, 1, 1) + t.reshape(1, -1, 1, 1))
    z = E.acff(a, sd64, "acff4")
    out = E.avg_pool_5x5_s1_p1(E.conv2d_pointwise(z, sd64["conv2.weight"], None))
    return out.reshape(-1, 20) @ sd64["fc.weight"].T + sd64["fc.bias"]

wset = sys.argv[1] if len(sys.argv) > 1 else "shipped"
sd = fixtures.get_state_dict(arch, wset)
fcal = np.concatenate([fixtures.noise_frames(12, seed=99), fixtures.smooth_frames(4, seed=98)], 0)
ftest = np.concatenate([fixtures.noise_frames(8, seed=61), fixtures.smooth_frames(8, seed=62)], 0)
xcal, xtest = I.ingest(fcal), I.ingest(ftest)
ref = E.forward(sd, xtest, arch, dtype=np.float64)["logits"]
cs = calib(sd, xcal, "channel", 100)
for name, qa, qw in [("none", set(), set()), ("W only", set(), {0,1,2}), ("A only", {0,1,2}, set()),
                     ("A0", {0}, set()), ("A1", {1}, set()), ("A2", {2}, set()),
                     ("W0", set(), {0}), ("W1", set(), {1}), ("W2", set(), {2})]:
    lg = run(sd, xtest, cs, qa, qw)
    print(f"{wset} {name:8s} rel logit err {np.abs(lg-ref).max()/np.abs(ref).max():.3e} agree {(lg.argmax(1)==ref.argmax(1)).mean():.3f}", flush=True)
