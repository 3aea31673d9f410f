"""ctypes binding of include/ernet_b200.h.  The CUDA library is mandatory: importing the compute
entry points without it raises (there is deliberately no CPU or PyTorch fallback)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libernet_b200.so")

OK = 0
ERR_INVALID_ARG, ERR_BAD_SHAPE, ERR_NOT_LOADED, ERR_BAD_BLOB, ERR_WORKSPACE, ERR_CUDA, ERR_UNSUPPORTED = \
    -1, -2, -3, -4, -5, -6, -7

ARCH = {"squeeze-ernet": 0, "squeeze-redconv": 1, "ernet": 2}
PRECISION = {"fp32": 0, "fp16": 1, "bf16": 2, "int8": 3}
DTYPE_F32, DTYPE_F16, DTYPE_BF16, DTYPE_U8 = 0, 1, 2, 3
NCHW, NHWC = 0, 1
RGB, BGR = 0, 1
ENGINE = {"auto": 0, "simt": 1, "tc": 2}
TAPS = {"ingest": 0, "stem": 1, "pool1": 2, "pool2": 3, "pool3": 4, "acff4": 5}

_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
SIGNATURES = {
    "ernet_create": (_i, [C.POINTER(_vp), _i, _i, _i]),
    "ernet_destroy": (None, [_vp]),
    "ernet_load_packed": (_i, [_vp, _vp, _sz]),
    "ernet_workspace_bytes": (_sz, [_vp, _i]),
    "ernet_set_chunk": (_i, [_vp, _i]),
    "ernet_get_chunk": (_i, [_vp]),
    "ernet_set_engine": (_i, [_vp, _i]),
    "ernet_get_engine": (_i, [_vp]),
    "ernet_set_debug_taps": (_i, [_vp, _i]),
    "ernet_set_persistent": (_i, [_vp, _i]),
    "ernet_debug_timeline": (_i, [_vp, _sz]),
    "ernet_debug_chain": (_i, [_vp, _i]),
    "ernet_host_copy_bytes_per_frame": (_sz, [_vp, _i, _i]),
    "ernet_classify_frames_host_submit": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ernet_classify_frames_host_wait": (_i, [_vp, _i]),
    "ernet_set_fast_ingest": (_i, [_vp, _i]),
    "ernet_set_fuse_ingest": (_i, [_vp, _i]),
    "ernet_set_host_gather": (_i, [_vp, _i, _i]),
    "ernet_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ernet_ingest_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "ernet_prepare_ingest": (_i, [_vp, _i, _i]),
    "ernet_forward_frames": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "ernet_classify_frames_host": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ernet_acff_depthwise": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ernet_set_depthwise_form": (_i, [_i]),
    "ernet_acff_add_depthwise": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ernet_confusion_update": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ernet_pointwise": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "ernet_debug_tap": (_i, [_vp, _i, _vp, _i, _vp, _sz, _vp]),
    "ernet_ingest_tables_host": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ernet_debug_device_status": (_i, [_vp, _i]),
    "ernet_check_watchdog": (_i, []),
    "ernet_source_hash": (C.c_char_p, []),
    "ernet_profile_enable": (_i, [_vp, _i]),
    "ernet_profile_read": (_i, [_vp, _vp, _vp, _i]),
    "ernet_launches_per_forward": (_i, [_vp, _i, _i]),
    "ernet_last_error": (C.c_char_p, []),
    "ernet_abi_version": (_i, []),
}

STAGES = ["ingest", "stem", "dw1", "pw1", "dw2", "pw2", "red2", "dw3", "pw3", "red3", "dw4", "pw4", "head",
          "tc_block1", "tc_block2", "tc_block3", "tc_block4"]

_lib = None


def load():
    """Return the loaded CDLL, building nothing: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc). This package has no CPU / PyTorch fallback by design.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().ernet_last_error().decode("utf-8", "replace")


def check(rc):
    """Map ernet_status to the exceptions the reference's callers would see."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_INVALID_ARG, ERR_BAD_SHAPE, ERR_BAD_BLOB):
        raise ValueError(msg)
    raise RuntimeError(f"ernet_b200 error {rc}: {msg}")
