"""Engine build / export tooling: what ``build_tensorrt_model.py`` does for the reference.

The reference converts a loaded model with torch2trt, verifies one inference and saves ``model.state_dict()`` of the
resulting ``TRTModule`` as ``tensorrt_state_dicts/{model}_{quant}[_{suffix}]_trt.pth``
(code/disaster_detection/build_tensorrt_model.py:79-290); callers later do
``TRTModule().load_state_dict(torch.load(path))`` and ``model(x.half())`` (aider-predict.py:167-171).  Its int8 branch is
two dead lines (build_tensorrt_model.py:256-259).  Here:

* ``build_trt_model(model, input_tensor, args)`` - same signature and file naming; "conversion" = packing the weights
  for the chosen precision (``fp32`` CUDA-core engine, ``fp16`` / ``bf16`` tcgen05 engines, ``int8`` with activation
  scales calibrated on ``args.calibration_frames`` or the default synthetic set), one verification inference, and
  ``torch.save`` of the engine's state_dict;
* ``TRTModule`` - the loader side: an ``nn.Module`` whose state_dict is the packed engine (architecture, precision,
  weight blob, int8 activation scales), runnable without the original checkpoint;
* ``load_model`` / ``get_model_input_shape`` / ``main`` - the rest of the script's surface.

Everything runs through ``libernet_b200.so``; there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import argparse
import ctypes as C
import logging
import os
from types import SimpleNamespace
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .model import ErNET, Squeeze_ErNET, Squeeze_RedConv, _EngineRuntime
from .pack import pack_state_dict

logger = logging.getLogger(__name__)
FORMAT = "ernet_b200_engine/1"
QUANTS = ("fp16", "fp32", "int8", "bf16")           # build_tensorrt_model.py:327 plus bf16


def get_model_input_shape(model_name: str) -> Tuple[int, int, int, int]:
    """build_tensorrt_model.py:310-317."""
    if model_name == 'ernet':
        return (1, 3, 240, 240)
    elif model_name in ['squeeze-ernet', 'squeeze-redconv']:
        return (1, 3, 140, 140)
    raise ValueError(f"Unsupported model: {model_name}")


def load_model(model_name: str, weights_path: str, device) -> nn.Module:
    """build_tensorrt_model.py:20-77: construct by name, accept the three checkpoint layouts, eval mode, one test
    inference on the device."""
    logger.info(f"Loading model {model_name} from {weights_path}")
    if model_name == 'ernet':
        model = ErNET()
    elif model_name == 'squeeze-ernet':
        model = Squeeze_ErNET()
    elif model_name == 'squeeze-redconv':
        model = Squeeze_RedConv()
    else:
        raise ValueError(f"Unsupported model: {model_name}")
    if not os.path.exists(weights_path):
        raise FileNotFoundError(f"Weights file not found: {weights_path}")
    checkpoint = torch.load(weights_path, map_location="cpu", weights_only=True)
    if isinstance(checkpoint, dict) and 'model_state_dict' in checkpoint:
        model.load_state_dict(checkpoint['model_state_dict'])
    elif isinstance(checkpoint, dict) and 'state_dict' in checkpoint:
        model.load_state_dict(checkpoint['state_dict'])
    else:
        model.load_state_dict(checkpoint)
    model = model.to(device)
    model.eval()
    with torch.no_grad():
        output = model(torch.randn(*get_model_input_shape(model_name)).to(device))
        logger.info(f"Model loaded successfully. Test output shape: {output.shape}")
    return model


class TRTModule(_EngineRuntime, nn.Module):
    """Packed engine with the surface callers use on torch2trt's ``TRTModule`` (aider-predict.py:167-171):
    ``TRTModule()``, ``load_state_dict(torch.load(path))``, ``.to(device)``, ``model(x)`` (fp32 / fp16 / bf16 input,
    probabilities out in the input's dtype).  Also everything else an engine can do: ``logits``, ``forward_frames``,
    ``classify_host`` ..."""

    def __init__(self):
        super().__init__()
        self.ARCH = None
        self.IN_HW = None
        self.precision = None
        self.act_scales = None
        self.register_buffer("blob", torch.zeros(0, dtype=torch.uint8))
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)     # so that next(model.parameters()) works
        self._engine = None
        self._loaded_version = None
        self._workspace = None
        self._last_batch = 0

    # ------------------------------------------------------------------ (de)serialisation
    def state_dict(self, *args, **kwargs):
        if self.ARCH is None:
            raise RuntimeError("empty TRTModule: nothing to save")
        return {"format": FORMAT, "arch": self.ARCH, "precision": self.precision, "blob": self.blob.detach().cpu().clone(),
                "act_scales": [list(map(float, a)) for a in self.act_scales] if self.act_scales is not None else None,
                "input_shape": [1, 3, self.IN_HW, self.IN_HW]}

    def load_state_dict(self, sd, strict=True):
        if not isinstance(sd, dict) or sd.get("format") != FORMAT:
            raise ValueError(f"not a {FORMAT} state_dict (build one with build_trt_model / build_engine)")
        if sd["arch"] not in _lib.ARCH or sd["precision"] not in _lib.PRECISION:
            raise ValueError(f"unknown arch / precision {sd['arch']!r} / {sd['precision']!r}")
        blob = torch.as_tensor(sd["blob"])
        if blob.dtype != torch.uint8 or blob.dim() != 1 or blob.numel() == 0:
            raise ValueError("corrupt engine: blob must be a non-empty 1-D uint8 tensor")
        self._release()
        self.ARCH, self.precision = sd["arch"], sd["precision"]
        self.IN_HW = get_model_input_shape(self.ARCH)[2]
        self.act_scales = sd.get("act_scales")
        self.blob = blob.to(self.blob.device).contiguous()
        self._loaded_version = None
        return self

    # ------------------------------------------------------------------ engine plumbing
    def eval(self):
        return super().eval()

    def _release(self):
        if self._engine is not None:
            _lib.load().ernet_destroy(self._engine[0])
            self._engine = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_engine(self):
        if self.ARCH is None:
            raise RuntimeError("empty TRTModule: call load_state_dict first")
        dev = self._anchor.device
        if dev.type != "cuda":
            raise RuntimeError("this engine runs only on a CUDA device (B200, sm_100a); call .to('cuda') first — there is no CPU fallback")
        lib = _lib.load()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._engine is None or self._engine[1] != idx:
            self._release()
            h = C.c_void_p()
            _lib.check(lib.ernet_create(C.byref(h), _lib.ARCH[self.ARCH], _lib.PRECISION[self.precision], idx))
            self._engine = (h, idx)
            self._loaded_version = None
        ver = (self.blob.data_ptr(), self.blob._version)
        if ver != self._loaded_version:
            host = self.blob.detach().cpu().contiguous().numpy()             # ernet_load_packed reads host memory
            _lib.check(lib.ernet_load_packed(self._engine[0], host.ctypes.data, host.nbytes))
            self._loaded_version = ver
        return lib, self._engine[0], idx


def build_engine(model, quant="fp16", calibration_frames=None):
    """Pack ``model`` (a Squeeze_ErNET / Squeeze_RedConv / ErNET of this package, weights loaded, on a CUDA device) for
    precision ``quant`` and return the runnable ``TRTModule``."""
    if quant not in QUANTS:
        raise ValueError(f"Unsupported quantization scheme: {quant}")                 # build_tensorrt_model.py:261
    if not hasattr(model, "ARCH") or model.ARCH not in _lib.ARCH:
        raise ValueError("build_engine expects a model of this package (Squeeze_ErNET, Squeeze_RedConv, ErNET)")
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("CUDA is not available. Cannot build the engine!")         # build_tensorrt_model.py:333
    act_scales = None
    if quant == "int8":
        twin = type(model)(precision="int8")
        twin.load_state_dict(model.state_dict())
        twin = twin.to(dev).eval()
        act_scales = twin.calibrate(calibration_frames)
        twin._release()
    sd = {k: v.detach().float().cpu() if v.is_floating_point() else v.detach().cpu() for k, v in model.state_dict().items()}
    blob = pack_state_dict(sd, model.ARCH, quant, act_scales)
    eng = TRTModule()
    eng.load_state_dict({"format": FORMAT, "arch": model.ARCH, "precision": quant,
                         "blob": torch.frombuffer(bytearray(blob), dtype=torch.uint8),
                         "act_scales": [list(a) for a in act_scales] if act_scales is not None else None})
    return eng.to(dev).eval()


def build_trt_model(model: nn.Module, input_tensor: torch.Tensor, args) -> nn.Module:
    """build_tensorrt_model.py:79-290 with the B200 engine as the target: build for ``args.quant``, verify one inference
    on ``input_tensor``, save the engine's state_dict under the reference's naming scheme, return the engine."""
    logger.info(f'Building B200 engine [{args.quant}]...')
    model = model.cuda()
    input_tensor = input_tensor.cuda()
    if model.training:
        logger.warning("Model was in training mode, switching to eval mode")       # build_tensorrt_model.py:96-98
        model.eval()
    if torch.isnan(input_tensor).any() or torch.isinf(input_tensor).any():
        logger.warning("Input tensor contains NaN or Inf values. Normalizing...")   # build_tensorrt_model.py:105-107
        input_tensor = torch.nan_to_num(input_tensor, nan=0.0, posinf=1.0, neginf=-1.0)
    engine = build_engine(model, args.quant, getattr(args, "calibration_frames", None))
    with torch.no_grad():                                                             # build_tensorrt_model.py:263-275
        test_input = input_tensor.clone()
        if args.quant == 'fp16':
            test_input = test_input.half()                                             # build_tensorrt_model.py:158-162
        output = engine(test_input)
        torch.cuda.synchronize()
        if not torch.isfinite(output.float()).all():
            raise RuntimeError("Engine verification failed: non-finite output")
        logger.info(f"Engine build successful. Output shape: {output.shape}, dtype {output.dtype}")
    output_dir = getattr(args, "output_dir", None) or 'tensorrt_state_dicts'           # build_tensorrt_model.py:277
    os.makedirs(output_dir, exist_ok=True)
    output_path = os.path.join(
        output_dir,
        f"{args.model}_{args.quant}_{args.output}_trt.pth" if getattr(args, "output", None) else f"{args.model}_{args.quant}_trt.pth")
    logger.info(f"Saving engine to {output_path}")
    torch.save(engine.state_dict(), output_path)
    engine.saved_path = output_path
    return engine


def main(argv=None):
    """CLI of build_tensorrt_model.py:319-372 (same flags; --quant int8 works)."""
    parser = argparse.ArgumentParser(description='B200 engine build script (drop-in for the TensorRT conversion script)')
    parser.add_argument('--model', type=str, default='ernet', choices=['ernet', 'squeeze-ernet', 'squeeze-redconv'])
    parser.add_argument('--output', type=str, default=None, help='Suffix to append at the output file name')
    parser.add_argument('--weights', type=str, default=None, help='Path to pre-trained PyTorch weights (.pt) file')
    parser.add_argument('--quant', type=str, default='fp16', choices=list(QUANTS))
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(name)s - %(levelname)s - %(message)s')
    if not torch.cuda.is_available():
        raise RuntimeError('CUDA is not available. Cannot build the engine!')
    if not args.weights:                                                              # build_tensorrt_model.py:343-349
        args.weights = {'ernet': 'weights/ernet.pt', 'squeeze-ernet': 'weights/squeeze-ernet.pt',
                        'squeeze-redconv': 'weights/squeeze-redconv.pt'}[args.model]
    device = torch.device('cuda')
    model = load_model(args.model, args.weights, device)
    input_tensor = torch.randn(*get_model_input_shape(args.model))
    return build_trt_model(model, input_tensor, args)


if __name__ == '__main__':
    main()
