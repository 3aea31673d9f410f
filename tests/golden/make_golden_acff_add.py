"""Generates tests/golden/acff_add_golden.npz from the reference's own add-fusion ACFF class.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_acff_add.py

``code/victim_localization/yolov3/models.py`` cannot be imported here (its ``utils.utils`` import needs matplotlib,
absent from the image), so the ``class ACFF`` statement is cut out of that file with ``ast`` at generation time and
executed unmodified against stock ``torch.nn`` - the reference's code runs, nothing of it is copied into this repo.
Each case: seeded weights (trained-like, BN statistics randomised, one case with negative gammas), seeded input,
outputs of the class in eval mode in fp64 and fp32.
"""
import ast
import os

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/code/victim_localization/yolov3/models.py"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [  # name, cin, cout, B, H, W, negative gammas
    ("c128_104", 128, 128, 1, 22, 22, False),      # yolov3-acffx.cfg:91-97 (128 -> 128), spatial size cut down
    ("c256_52", 256, 256, 1, 13, 13, False),       # cfg:234-240 (256 -> 256)
    ("c128_head14", 128, 14, 2, 12, 10, False),    # cfg:349 (filters=14 heads): N padded to 32 by the packer
    ("c64", 64, 64, 2, 17, 15, True),              # cfg:392-398
    ("c40_generic", 40, 32, 1, 9, 11, False),      # a channel count outside the compiled list (run-time C kernel)
]


def reference_class():
    src = open(REF).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ACFF")
    ns = {"nn": nn, "torch": torch, "debug": False}
    exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns["ACFF"]


def main():
    ACFF = reference_class()
    out = {}
    for name, cin, cout, B, H, W, neg in CASES:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        m = ACFF(cin, cout, 3).double()
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float64) * (0.4 if p.dim() == 4 and p.shape[1] == 1 else
                                                                                  (1.0 / np.sqrt(cin) if p.dim() == 4 else 0.1)))
            m.batch_norm.weight.copy_(torch.rand(cout, generator=g, dtype=torch.float64) * 0.4 + 0.8)
            if neg:
                m.batch_norm.weight[::3] *= -1
            m.batch_norm.running_mean.copy_(torch.randn(cout, generator=g, dtype=torch.float64) * 0.5)
            m.batch_norm.running_var.copy_(torch.rand(cout, generator=g, dtype=torch.float64) * 1.5 + 0.5)
        m = m.float().double()                    # weights and input are fp32-representable: fp64 and fp32 runs see the same numbers
        m.eval()
        x = torch.randn(B, cin, H, W, generator=g, dtype=torch.float64).float().double()
        with torch.no_grad():
            y64 = m(x)
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            y32 = m.float()(x.float())
        out[f"{name}/x"] = x.float().numpy()
        out[f"{name}/y64"] = y64.numpy()
        out[f"{name}/y32"] = y32.numpy()
        for k, v in sd.items():
            out[f"{name}/sd/{k}"] = v.float().numpy() if v.is_floating_point() else v.numpy()
    np.savez_compressed(os.path.join(HERE, "acff_add_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
