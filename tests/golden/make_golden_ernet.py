"""Golden vectors for the baseline ErNET from the REAL reference class.

Run in the build container only (needs /root/reference and torch):  python tests/golden/make_golden_ernet.py
Imports model/ernet.py:6-49 and model/acff.py (never copies them), loads weights/ernet-state_dict.pt and two seeded
random weight sets, runs (B,3,240,240) inputs in fp64 and fp32 and stores the fc output (logits), the probabilities and
sub-sampled intermediates in tests/golden/ernet_golden.npz; the checkpoint goes to tests/golden/weights/ernet_shipped.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/code/disaster_detection"
sys.path.insert(0, REF)

import fixtures  # noqa: E402
from model.ernet import ErNET  # noqa: E402

torch.set_num_threads(8)
TAPS = ["conv1", "acff1", "pool1", "acff2", "pool2", "acff3", "pool3", "acff4", "acff5", "acff6", "conv2", "globalpool", "fc"]

sd_ship = torch.load(os.path.join(REF, "weights/ernet-state_dict.pt"), map_location="cpu", weights_only=True)
np.savez_compressed(os.path.join(HERE, "weights", "ernet_shipped.npz"), **{k: v.numpy() for k, v in sd_ship.items()})
print("exported ernet_shipped.npz", len(sd_ship), "keys")


def build(sd_np, dtype):
    m = ErNET()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()})   # strict: proves fixtures.key_shapes("ernet")
    return m.eval().to(dtype)


def run(model, x, taps):
    cap, hooks = {}, []
    for n in (TAPS if taps else ["fc"]):
        hooks.append(getattr(model, n).register_forward_hook(lambda mod, i, o, n=n: cap.__setitem__(n, o.detach().numpy().copy())))
    with torch.no_grad():
        probs = model(x).numpy().copy()
    for h in hooks:
        h.remove()
    return probs, cap


x = fixtures.normal_tensors(3, seed=17, hw=240)
out = {}
for wset in ("shipped", "w3", "w3neg"):
    sd = fixtures.get_state_dict("ernet", wset)
    p64, c64 = run(build(sd, torch.float64), torch.from_numpy(x).double(), True)
    p32, c32 = run(build(sd, torch.float32), torch.from_numpy(x).float(), False)
    tag = f"ernet/{wset}/norm"
    out[f"{tag}/logits64"], out[f"{tag}/probs64"] = c64["fc"], p64
    out[f"{tag}/logits32"], out[f"{tag}/probs32"] = c32["fc"], p32
    for n, v in c64.items():
        if n != "fc":
            out[f"{tag}/tap/{n}/sub"] = v[0, :, ::5, ::5].astype(np.float64)
            out[f"{tag}/tap/{n}/abs_sum"] = np.asarray(np.abs(v).sum())
            out[f"{tag}/tap/{n}/shape"] = np.asarray(v.shape)
    print(tag, "top1", p64.argmax(1), "max|logit|", float(np.abs(c64["fc"]).max()))
np.savez_compressed(os.path.join(HERE, "ernet_golden.npz"), **out)
print("wrote ernet_golden.npz", len(out), "arrays")
