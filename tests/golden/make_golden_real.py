"""Real-image fixture (SURVEY.md 8d input set I3): the AIDER-named JPEGs that ship in the reference tree
(code/victim_localization/yolov3/data/custom/test/images, yolov5/dataset/*/images; 240x240 ... 1280x720), reduced to what
the classifier sees - the 140x140 uint8 crop of the eval transform (Resize(159) -> CenterCrop(140), dataloaders/aider.py:
421-423), produced by torchvision/Pillow itself - plus the fp64 logits of the real reference classes on the normalised
tensors.  The frames themselves are too large to commit (up to 2.7 MB each); the crops are 58.8 KB each.

    python tests/golden/make_golden_real.py      # needs /root/reference; writes tests/golden/real_golden.npz
"""
import glob
import os
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/code/disaster_detection"
sys.path.insert(0, REF)
from model.squeeze_ernet import Squeeze_ErNET            # noqa: E402
from model.squeeze_ernet_redconv import Squeeze_RedConv  # noqa: E402

pats = ["/root/reference/code/victim_localization/yolov3/data/custom/test/images/*.jpg",
        "/root/reference/code/victim_localization/yolov5/dataset/*/images/*.jpg"]
files = sorted({os.path.basename(f): f for p in pats for f in glob.glob(p)}.values())
crop = transforms.Compose([transforms.Resize(int(140 * 1.14)), transforms.CenterCrop(140)])
norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
crops, xs, names, sizes = [], [], [], []
for f in files:
    im = Image.open(f).convert("RGB")
    c = crop(im)
    crops.append(np.asarray(c))
    xs.append(norm(c))
    names.append(os.path.basename(f))
    sizes.append(im.size[::-1])
x = torch.stack(xs, 0)
out = {"crops_u8": np.stack(crops, 0), "names": np.array(names), "frame_hw": np.array(sizes, np.int32)}
for arch, cls, w in (("squeeze-ernet", Squeeze_ErNET, "squeeze-ernet-state_dict.pt"), ("squeeze-redconv", Squeeze_RedConv, "squeeze-redconv-state_dict.pt")):
    m = cls()
    m.load_state_dict(torch.load(os.path.join(REF, "weights", w), weights_only=True))
    m = m.double().eval()
    cap = {}
    m.fc.register_forward_hook(lambda mod, i, o: cap.__setitem__("l", o))
    with torch.no_grad():
        p = m(x.double())
    out[f"{arch}/shipped/logits64"] = cap["l"].numpy()
    out[f"{arch}/shipped/probs64"] = p.numpy()
np.savez_compressed(os.path.join(HERE, "real_golden.npz"), **out)
print(len(files), "frames;", {k: v.shape for k, v in out.items()})
