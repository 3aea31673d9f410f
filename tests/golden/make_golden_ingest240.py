"""Goldens for ErNET's eval transform, ``aider_transforms = get_val_torchvision_transforms(240)``
(code/disaster_detection/dataloaders/aider.py:412-426,430: Resize(273) -> CenterCrop(240) -> ToTensor -> Normalize), produced
with the real torchvision / Pillow in the build container:  python tests/golden/make_golden_ingest240.py
Stores the uint8 crop for every case and the normalised fp32 tensor for one (the ToTensor+Normalize table is already pinned
by ingest_golden.npz)."""
import os
import sys

import numpy as np
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fixtures  # noqa: E402

CASES = [("noise240x240", "noise", 240, 240, 31), ("smooth480x640", "smooth", 480, 640, 32), ("noise300x280", "noise", 300, 280, 33),
         ("noise273x273_id", "noise", 273, 273, 34), ("smooth260x420", "smooth", 260, 420, 35)]


def frame(kind, h, w, seed):
    return (fixtures.noise_frames if kind == "noise" else fixtures.smooth_frames)(1, h, w, seed)[0]


def main():
    tf_u8 = transforms.Compose([transforms.Resize(int(240 * 1.14)), transforms.CenterCrop(240)])
    tf_full = transforms.Compose([transforms.Resize(int(240 * 1.14)), transforms.CenterCrop(240), transforms.ToTensor(),
                                  transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    out = {}
    for name, kind, h, w, seed in CASES:
        pil = Image.fromarray(frame(kind, h, w, seed))
        out[f"{name}/crop_u8"] = np.asarray(tf_u8(pil)).copy()
    out["noise240x240/tensor"] = tf_full(Image.fromarray(frame("noise", 240, 240, 31))).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ingest240_golden.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
