"""Single-image prediction (BASELINE.json configs[0] flavour): mirror of ``predict`` / ``main`` in
``code/disaster_detection/aider-predict.py:47-86,127-180``.

``predict(model, image_path, transform, device)`` keeps the reference's signature and return value
``(class name, confidence in percent)``, including its quirk that the confidence is a *second* softmax over the model's
softmax output (aider-predict.py:80).  With ``transform=None`` (the fast path) the BGR frame that ``cv2.imread`` returns goes
straight to ``model.forward_frames(frame, bgr=True)``: channel swap, resize, centre crop, ToTensor and Normalize all run in the
ingest kernel on the GPU.  With a ``transform`` callable the reference's own sequence is followed (RGB -> PIL -> transform ->
reshape -> ``model(data)``).
"""
from __future__ import annotations

import argparse
import logging
import os
from typing import Tuple

import numpy as np
import torch

from .model import load_model

logger = logging.getLogger(__name__)
CLASSES = ['collapsed building', 'fire', 'flooded areas', 'normal', 'traffic incident']     # aider-predict.py:83


def predict(model, image_path: str, transform=None, device=None, use_trt: bool = False, quant: str = 'fp16') -> Tuple[str, float]:
    """Make prediction on a single image (aider-predict.py:47-86)."""
    import cv2
    image = cv2.imread(image_path)                                   # BGR, uint8 (aider-predict.py:57)
    if image is None:
        raise ValueError(f"Could not load image at {image_path}")
    device = torch.device(device) if device is not None else next(model.parameters()).device
    with torch.no_grad():
        if transform is None:
            frame = torch.from_numpy(np.ascontiguousarray(image)[None]).to(device)
            output = model.forward_frames(frame, bgr=True)
        else:
            from PIL import Image
            data = transform(Image.fromarray(cv2.cvtColor(image, cv2.COLOR_BGR2RGB)))
            hw = getattr(model, "IN_HW", 140)
            data = torch.reshape(data, (1, 3, hw, hw)).to(device)     # aider-predict.py:69-70
            if use_trt and quant == 'fp16':
                data = data.half()
            output = model(data)
        predicted_class = output.data.max(1, keepdim=True)[1]
        confidence = torch.nn.functional.softmax(output.float(), dim=1)[0][predicted_class].item() * 100
    return CLASSES[predicted_class.item()], confidence


def main(argv=None):
    parser = argparse.ArgumentParser(description='Predict disaster types from aerial images (B200 engine)')
    parser.add_argument('--model', type=str, default='ernet', choices=['ernet', 'squeeze-ernet', 'squeeze-redconv'])
    parser.add_argument('--image', type=str, required=True)
    parser.add_argument('--weights', type=str, default=None)
    parser.add_argument('--trt', action='store_true', help='also run the packed engine saved by build_engine (TRTModule)')
    parser.add_argument('--quant', type=str, default='fp16', choices=['fp16', 'bf16', 'fp32', 'int8'])
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(name)s - %(levelname)s - %(message)s')
    if not torch.cuda.is_available():
        raise RuntimeError('CUDA is not available: the B200 engine has no CPU path')
    device = torch.device('cuda')
    if not args.weights:                                              # aider-predict.py:139-146
        args.weights = {'ernet': 'weights/ernet-state_dict.pt', 'squeeze-ernet': 'weights/squeeze-ernet-state_dict.pt',
                        'squeeze-redconv': 'weights/squeeze-redconv-state_dict.pt'}[args.model]
    if not os.path.exists(args.weights):
        raise FileNotFoundError(f"No weights found at {args.weights}")
    model = load_model(args.model, args.weights, device)
    prediction, confidence = predict(model, args.image, None, device)
    logger.info(f"Prediction: {prediction} ({confidence:.1f}%)")
    if args.trt:                                                      # aider-predict.py:165-171
        from .build_engine import TRTModule
        trt_model = TRTModule()
        trt_model.load_state_dict(torch.load(f'tensorrt_state_dicts/{args.model}_{args.quant}_trt.pth', weights_only=True))
        trt_model = trt_model.to(device)
        trt_prediction, trt_confidence = predict(trt_model, args.image, None, device, True, args.quant)
        logger.info(f"Engine Prediction: {trt_prediction} ({trt_confidence:.1f}%)")
    return prediction, confidence


if __name__ == '__main__':
    main()
