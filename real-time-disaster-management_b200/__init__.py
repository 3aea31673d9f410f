"""B200-native engine for the Squeeze-ErNet / Squeeze-ErNet-RedConv classification hot path of
qazi0/real-time-disaster-management.

The directory name follows the reference repository; since it is not a valid Python identifier,
import it as ``import rtdm_b200`` (a top-level alias module) or via
``importlib.import_module("real-time-disaster-management_b200")``.
"""
from .model import ErNET, Squeeze_ErNET, Squeeze_RedConv, from_state_dict, load_model  # noqa: F401
from .pack import pack_state_dict  # noqa: F401

__all__ = ["Squeeze_ErNET", "Squeeze_RedConv", "ErNET", "load_model", "from_state_dict", "pack_state_dict"]
