"""Debug driver: frames path with the fast fused transform+conv1 kernel vs the table-lookup kernel vs the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402
from oracle import ernet_numpy as E, ingest_numpy  # noqa: E402

frames = np.concatenate([fixtures.noise_frames(3), fixtures.smooth_frames(3)], 0)
x_ref = ingest_numpy.ingest(frames)
ft = torch.from_numpy(frames).cuda()
for arch in fixtures.ARCHS:
    for wset in ("shipped", "w3"):
        sd = fixtures.get_state_dict(arch, wset)
        ref = E.forward(sd, x_ref, arch, dtype=np.float64)["logits"]
        for prec in ("bf16", "fp16"):
            m = rtdm_b200.from_state_dict(arch, sd, "cuda:0", prec)
            res = {}
            for fast in (False, True):
                m.set_fast_ingest(fast)
                for order in ("rgb", "bgr"):
                    fr = ft if order == "rgb" else ft.flip(-1).contiguous()
                    _, lg = m.forward_frames(fr, return_logits=True, bgr=(order == "bgr"))
                    torch.cuda.synchronize()
                    l = lg.double().cpu().numpy()
                    res[(fast, order)] = float(np.abs(l - ref).max() / np.abs(ref).max())
            print(arch, wset, prec, {f"{'fast' if k[0] else 'slow'}-{k[1]}": f"{v:.2e}" for k, v in res.items()}, flush=True)
