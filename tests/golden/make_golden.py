"""Generate the golden vectors under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference, torch, torchvision, Pillow):

    python tests/golden/make_golden.py

It imports the reference's own model classes (never copies them), runs them on the
seeded inputs of tests/fixtures.py and stores outputs.  The GPU box has no
/root/reference; tests there read only the committed .npz files.

Sources exercised (relative to /root/reference/code/disaster_detection):
  model/squeeze_ernet.py:8-46, model/squeeze_ernet_redconv.py:8-52, model/acff.py:9-59,
  weights/squeeze-ernet-state_dict.pt, weights/squeeze-redconv-state_dict.pt,
  dataloaders/aider.py:421-426 (the eval transform, restated with torchvision because
  dataloaders/aider.py imports albumentations, which is absent).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/code/disaster_detection"
sys.path.insert(0, REF)

import fixtures  # noqa: E402
from model.squeeze_ernet import Squeeze_ErNET  # noqa: E402
from model.squeeze_ernet_redconv import Squeeze_RedConv  # noqa: E402
from PIL import Image  # noqa: E402
from torchvision import transforms  # noqa: E402

torch.set_num_threads(8)
CLASSES = {"squeeze-ernet": Squeeze_ErNET, "squeeze-redconv": Squeeze_RedConv}
SHIPPED = {"squeeze-ernet": "weights/squeeze-ernet-state_dict.pt",
           "squeeze-redconv": "weights/squeeze-redconv-state_dict.pt"}

TAP_MODULES = ["conv1", "conv_red1", "acff1", "pool1", "acff2", "conv_red2", "pool2", "acff3", "pool3",
               "conv_red3", "acff4", "conv2", "globalpool", "fc"]


def export_shipped():
    os.makedirs(os.path.join(HERE, "weights"), exist_ok=True)
    for arch, rel in SHIPPED.items():
        sd = torch.load(os.path.join(REF, rel), map_location="cpu", weights_only=True)
        out = {k: v.numpy() for k, v in sd.items()}
        name = {"squeeze-ernet": "squeeze_ernet_shipped.npz", "squeeze-redconv": "squeeze_redconv_shipped.npz"}[arch]
        np.savez_compressed(os.path.join(HERE, "weights", name), **out)
        print("exported", name, len(out), "keys")


def build(arch, sd_np, dtype):
    m = CLASSES[arch]()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    m.load_state_dict(sd)          # strict: also proves fixtures.key_shapes == the reference's keys
    m.eval()
    return m.to(dtype)


def run(model, x, taps=False):
    cap = {}
    hooks = []
    names = TAP_MODULES if taps else ["fc"]
    for n in names:
        if hasattr(model, n):
            hooks.append(getattr(model, n).register_forward_hook(
                lambda mod, i, o, n=n: cap.__setitem__(n, o.detach().numpy().copy())))
    with torch.no_grad():
        probs = model(x).numpy().copy()
    for h in hooks:
        h.remove()
    return probs, cap


def model_goldens():
    tf = reference_transform()
    frames = np.concatenate([fixtures.noise_frames(2), fixtures.smooth_frames(2)], axis=0)
    x_frames = np.stack([tf(Image.fromarray(f)).numpy() for f in frames], 0)
    x_norm = fixtures.normal_tensors(4, seed=7)
    out = {"x_frames": x_frames.astype(np.float32)}
    for arch in fixtures.ARCHS:
        for wset in ("shipped", "w3", "w3neg"):
            sd = fixtures.get_state_dict(arch, wset)
            m64 = build(arch, sd, torch.float64)
            m32 = build(arch, sd, torch.float32)
            for iname, x in (("norm", x_norm), ("frames", x_frames)):
                p64, c64 = run(m64, torch.from_numpy(x).double(), taps=(iname == "norm"))
                p32, c32 = run(m32, torch.from_numpy(x).float())
                tag = f"{arch}/{wset}/{iname}"
                out[f"{tag}/logits64"] = c64["fc"]
                out[f"{tag}/probs64"] = p64
                out[f"{tag}/logits32"] = c32["fc"]
                out[f"{tag}/probs32"] = p32
                if iname == "norm":
                    for n, v in c64.items():
                        if n == "fc":
                            continue
                        # sub-sampled intermediates of sample 0 + a global checksum
                        out[f"{tag}/tap/{n}/sub"] = v[0, :, ::5, ::5].astype(np.float64)
                        out[f"{tag}/tap/{n}/abs_sum"] = np.asarray(np.abs(v).sum())
                        out[f"{tag}/tap/{n}/shape"] = np.asarray(v.shape)
                print(tag, "top1", p64.argmax(1), "max|logit|", float(np.abs(c64["fc"]).max()))
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)


def reference_transform():
    # get_val_torchvision_transforms(140), dataloaders/aider.py:412-426
    return transforms.Compose([
        transforms.Resize(int(140 * 1.14)),
        transforms.CenterCrop(140),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])


INGEST_CASES = [  # (name, kind, h, w, seed)
    ("noise240", "noise", 240, 240, 1234),
    ("smooth240", "smooth", 240, 240, 1235),
    ("noise480x640", "noise", 480, 640, 11),
    ("smooth350x372", "smooth", 350, 372, 12),
    ("noise372x350", "noise", 372, 350, 13),
    ("noise100x120_up", "noise", 100, 120, 14),
    ("noise159x159_id", "noise", 159, 159, 15),
    ("smooth720x1280", "smooth", 720, 1280, 16),
    ("noise161x300", "noise", 161, 300, 17),
]


def ingest_frame(kind, h, w, seed):
    f = fixtures.noise_frames(1, h, w, seed) if kind == "noise" else fixtures.smooth_frames(1, h, w, seed)
    return f[0]


def ingest_goldens():
    tf_full = reference_transform()
    tf_u8 = transforms.Compose([transforms.Resize(int(140 * 1.14)), transforms.CenterCrop(140)])
    out = {}
    for name, kind, h, w, seed in INGEST_CASES:
        f = ingest_frame(kind, h, w, seed)
        pil = Image.fromarray(f)
        out[f"{name}/crop_u8"] = np.asarray(tf_u8(pil)).copy()
        out[f"{name}/tensor"] = tf_full(pil).numpy().copy()
        print("ingest", name, out[f"{name}/crop_u8"].shape)
    # one real AIDER-named JPEG shipped with the reference (240x240), decoded here
    jpg = "/root/reference/code/victim_localization/yolov3/data/custom/test/images/normal_image2085.jpg"
    if os.path.exists(jpg):
        pil = Image.open(jpg).convert("RGB")
        out["real240/frame"] = np.asarray(pil).copy()
        out["real240/crop_u8"] = np.asarray(tf_u8(pil)).copy()
        out["real240/tensor"] = tf_full(pil).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ingest_golden.npz"), **out)


if __name__ == "__main__":
    export_shipped()
    ingest_goldens()
    model_goldens()
    print("done")
