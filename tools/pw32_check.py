"""fp32 engine: split-TF32 tcgen05 1x1 convolutions (tc_pw32.cuh) against the FFMA kernels - taps, logits, stage times."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import fixtures  # noqa: E402
import rtdm_b200  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(7)
frames = torch.randint(0, 256, (batch, 240, 240, 3), dtype=torch.uint8, generator=g).cuda()
res = {}
for mode in ("0", "1", "1s"):                       # FFMA 1x1 kernels | split-TF32 tcgen05 GEMMs | the latter with the table-lookup ingest
    os.environ["ERNET_FP32_TC"] = mode[0]
    m = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "shipped"), "cuda:0", "fp32")
    m.set_chunk(batch)
    m.set_fast_ingest(mode != "1s")
    _, lg = m.forward_frames(frames, return_logits=True)
    torch.cuda.synchronize()
    m.check_watchdog() if hasattr(m, "check_watchdog") else None
    taps = {n: m.tap(n).clone() for n in ("pool1", "pool2", "pool3")}
    for _ in range(3):
        m.forward_frames(frames)
    m.profile(True)
    for _ in range(5):
        m.forward_frames(frames)
    torch.cuda.synchronize()
    prof = {k: round(v[0] / v[1] * 1e3, 1) for k, v in m.profile_read().items()}
    m.profile(False)
    res[mode] = (lg.clone(), taps, prof)
out = {"batch": batch}
for n in ("pool1", "pool2", "pool3"):
    a, b = res["0"][1][n], res["1"][1][n]
    out[n] = {"max_abs": float((a - b).abs().max()), "ref_max": float(a.abs().max()),
              "rel_fro": float((a - b).norm() / a.norm())}
out["logits_max_abs"] = float((res["0"][0] - res["1"][0]).abs().max())
out["stage_us_ffma"] = res["0"][2]
out["stage_us_tf32x3"] = res["1"][2]
out["logits_max_abs_fast_vs_table_ingest"] = float((res["1"][0] - res["1s"][0]).abs().max())
out["logits_abs_max"] = float(res["1s"][0].abs().max())
out["stage_us_tf32x3_table_ingest"] = res["1s"][2]
out["step_us"] = {k: round(sum(res[k][2].values()), 1) for k in res}
print(json.dumps(out))
