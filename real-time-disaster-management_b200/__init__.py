"""B200-native engine for the Squeeze-ErNet / Squeeze-ErNet-RedConv classification hot path of
qazi0/real-time-disaster-management.

The directory name follows the reference repository; since it is not a valid Python identifier,
import it as ``import rtdm_b200`` (a top-level alias module) or via
``importlib.import_module("real-time-disaster-management_b200")``.
"""
from .model import ErNET, Squeeze_ErNET, Squeeze_RedConv, from_state_dict, load_model  # noqa: F401
from .pack import pack_state_dict  # noqa: F401
from .build_engine import TRTModule, build_trt_model  # noqa: F401  (build_engine.build_engine: the function behind it)
from .evaluate import compute_per_class_metrics, evaluate_model, frame_batches  # noqa: F401
from .acff_add import ACFF  # noqa: F401  (add-fusion block of the detector half, yolov3/models.py:265)

__all__ = ["Squeeze_ErNET", "Squeeze_RedConv", "ErNET", "load_model", "from_state_dict", "pack_state_dict", "ACFF", "TRTModule", "build_trt_model", "evaluate_model", "compute_per_class_metrics",
           "frame_batches"]
