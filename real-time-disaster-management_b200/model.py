"""Host-side mirror of the reference's model plugin surface, backed by the CUDA library.

Drop-in for ``code/disaster_detection/model`` as used by ``aider-predict.py:22-45,76`` and
``evaluate-classification-metrics.py:24-47,77``:

    model = Squeeze_ErNET()                  # zero-arg constructor   (model/squeeze_ernet.py:8)
    model.load_state_dict(torch.load(path))  # same 56 / 62 keys      
    model = model.to(device); model.eval()
    probs = model(x)                         # (B,3,140,140) -> (B,5) softmax probabilities

The ``torch.nn`` leaf modules created here are parameter *containers* only (so ``state_dict``,
``load_state_dict``, ``parameters``, ``to``, ``half``, ``apply`` behave exactly like the
reference's); none of them is ever called.  All arithmetic happens in ``libernet_b200.so``
through the C ABI of ``include/ernet_b200.h``.  Without that library, or without a CUDA device,
``forward`` raises: there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .pack import pack_state_dict, widths

_PREC_OF_DTYPE = {torch.float32: "fp32", torch.float16: "fp16", torch.bfloat16: "bf16"}
_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}
_TAP_SHAPES = {  # name -> lambda(arch) -> (C, H, W)
    "ingest": lambda a: (3, 140, 140),
    "stem": lambda a: (widths(a)[0][0], 69, 69),
    "pool1": lambda a: (64, 33, 33),
    "pool2": lambda a: (widths(a)[2][0], 15, 15),
    "pool3": lambda a: (widths(a)[3][0], 6, 6),
    "acff4": lambda a: (256, 4, 4),
}


class _ACFFParams(nn.Module):
    """Parameter container with the child names of the reference's ACFF block (acff.py:25-35)."""

    def __init__(self, cin, cout):
        super().__init__()
        for j, d in ((1, 1), (2, 2), (3, 3)):
            setattr(self, f"conv{j}", nn.Conv2d(cin, cin, 3, 1, d - 1, d, groups=cin, bias=True))
        self.fused_conv = nn.Conv2d(3 * cin, cout, 1, bias=True)
        self.batch_norm = nn.BatchNorm2d(cout)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: compute happens in libernet_b200.so")


class _EngineRuntime:
    """Everything that drives a loaded engine handle through the C ABI: ``model(x)``, the frames path, the host entry
    points, schedule switches and introspection.  Shared by the weight-holding model classes below (which pack their
    state_dict on demand) and by ``build_engine.TRTModule`` (which loads a previously packed engine).  A subclass
    provides ``ARCH``, ``IN_HW``, ``_workspace``, ``_last_batch`` and ``_ensure_engine() -> (lib, handle, device_index)``."""

    def _get_workspace(self, lib, h, batch, device):
        need = lib.ernet_workspace_bytes(h, batch)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != device:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
        return self._workspace

    def set_engine(self, engine):
        """'auto' (default), 'simt' (CUDA-core kernels) or 'tc' (tcgen05 block kernels)."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_engine(h, _lib.ENGINE[engine]))
        self._workspace = None
        return self

    @property
    def engine(self):
        lib, h, _ = self._ensure_engine()
        return {1: "simt", 2: "tc"}[lib.ernet_get_engine(h)]

    def set_persistent(self, on=True):
        """Schedule of the tensor-core block kernels: 2 / True = persistent + CTA pairs (default), 1 = persistent single
        CTAs, 0 / False = one image per CTA."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_persistent(h, 2 if on is True else int(on)))
        return self

    def set_fast_ingest(self, on=True):
        """Frames path of the 16-bit engines: folded-normalisation fast kernel (default) or the table-lookup kernel that is
        bit-identical to ingest() followed by forward()."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_fast_ingest(h, 1 if on else 0))
        return self

    def set_fuse_ingest(self, on=True):
        """Frames path: run the eval transform + conv1 inside block 1's persistent kernel (helper warps under the tcgen05
        MMAs, csrc/tc_fblock.cuh) instead of as a kernel of their own.  Bit-identical results; off by default."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_fuse_ingest(h, 1 if on else 0))
        return self

    def set_host_gather(self, on=True, ctas=0):
        """Host path (classify_host / classify_host_submit): pull only the footprint of the crop window from PINNED frames
        over PCIe with a small kernel (csrc/host_gather.cuh) instead of a copy-engine transfer of the row range.  Same
        results; pageable frames take the copy-engine path either way."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_host_gather(h, 1 if on else 0, int(ctas)))
        return self

    def set_debug_taps(self, on=True):
        """Also write the intermediates that fused kernels keep on chip (needed for tap('acff4'))."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_debug_taps(h, 1 if on else 0))
        return self

    def set_chunk(self, n):
        """Images processed per internal pass (bounds the workspace; default 1024)."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_set_chunk(h, int(n)))
        return self

    # ------------------------------------------------------------------ model(x)
    def _run(self, x, want_logits):
        if not isinstance(x, torch.Tensor):
            raise TypeError("expected a torch.Tensor")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, self.IN_HW, self.IN_HW):
            # the reference fails (or silently mixes images) for any other size, squeeze_ernet.py:39 / ernet.py:42
            raise ValueError(f"expected input of shape (B,3,{self.IN_HW},{self.IN_HW}), got {tuple(x.shape)}")
        if x.dtype not in _DTYPE_CODE:
            raise ValueError(f"unsupported input dtype {x.dtype}")
        lib, h, idx = self._ensure_engine()
        if x.device.type != "cuda" or (x.device.index or 0) != idx:
            raise RuntimeError(f"input is on {x.device} but the model is on cuda:{idx}")
        B = x.shape[0]
        if B < 1:
            raise ValueError("empty batch")
        layout = _lib.NCHW
        if not x.is_contiguous():
            if x.is_contiguous(memory_format=torch.channels_last):
                layout = _lib.NHWC
            else:
                x = x.contiguous()
        ws = self._get_workspace(lib, h, B, x.device)
        probs = torch.empty((B, 5), dtype=torch.float32, device=x.device)
        logits = torch.empty((B, 5), dtype=torch.float32, device=x.device) if want_logits else None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.ernet_forward(h, x.data_ptr(), _DTYPE_CODE[x.dtype], layout, B, probs.data_ptr(),
                                     logits.data_ptr() if want_logits else None, ws.data_ptr(), ws.numel(), stream))
        self._last_batch = B
        return probs, logits

    def forward(self, x):
        """Softmax probabilities (B,5) in the dtype of ``x`` (squeeze_ernet.py:41)."""
        probs, _ = self._run(x, False)
        return probs if x.dtype == torch.float32 else probs.to(x.dtype)

    def logits(self, x):
        """Pre-softmax ``fc`` output (B,5) fp32 — what parity is judged on."""
        return self._run(x, True)[1]

    def forward_with_logits(self, x):
        return self._run(x, True)

    # ------------------------------------------------------------------ frames -> probabilities
    def forward_frames(self, frames, *, bgr=False, return_logits=False, stream=None):
        """uint8 (B,H,W,3) device frames -> probabilities: the eval transform
        (dataloaders/aider.py:421-426) and the model in one call.  ``stream``: a torch.cuda.Stream to enqueue on
        (default: the current stream); outputs are allocated on it."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError(f"expected uint8 frames of shape (B,H,W,3), got {frames.dtype} {tuple(frames.shape)}")
        lib, h, idx = self._ensure_engine()
        if frames.device.type != "cuda" or (frames.device.index or 0) != idx:
            raise RuntimeError(f"frames are on {frames.device} but the model is on cuda:{idx}")
        frames = frames.contiguous()
        B, H, W, _ = frames.shape
        ws = self._get_workspace(lib, h, B, frames.device)
        probs = torch.empty((B, 5), dtype=torch.float32, device=frames.device)
        logits = torch.empty((B, 5), dtype=torch.float32, device=frames.device) if return_logits else None
        if stream is None:
            stream = torch.cuda.current_stream(frames.device).cuda_stream
        else:
            probs.record_stream(stream)
            if logits is not None:
                logits.record_stream(stream)
            stream = stream.cuda_stream
        _lib.check(lib.ernet_forward_frames(h, frames.data_ptr(), B, H, W, _lib.BGR if bgr else _lib.RGB,
                                            probs.data_ptr(), logits.data_ptr() if return_logits else None,
                                            ws.data_ptr(), ws.numel(), stream))
        self._last_batch = B
        return (probs, logits) if return_logits else probs

    def graph_frames(self, frames, *, bgr=False, return_logits=False):
        """Capture ``forward_frames(frames)`` into a CUDA graph and return a zero-argument callable that replays it.
        ``frames`` is the static input buffer: write the next batch into it (``frames.copy_(...)``) and call the runner; it
        returns the same output tensor(s) every time (clone them to keep a result).  The whole chain - transform + conv1,
        the block kernels with their programmatic-dependent-launch edges, the head - becomes one graph launch, which is
        what a single-frame caller (aider-predict.py:76, real-time-inference.py:97) pays per frame."""
        lib, h, idx = self._ensure_engine()
        # the graph bakes in device pointers: give it a workspace of its own (the shared one is re-allocated when a larger
        # batch comes along) and keep it alive with the runner.  Re-capture after loading new weights.
        own_ws = torch.empty(max(1, lib.ernet_workspace_bytes(h, frames.shape[0])), dtype=torch.uint8, device=frames.device)
        shared_ws, self._workspace = self._workspace, own_ws
        try:
            for _ in range(2):                               # builds the resize tables outside the capture
                self.forward_frames(frames, bgr=bgr, return_logits=return_logits)
            torch.cuda.synchronize(frames.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward_frames(frames, bgr=bgr, return_logits=return_logits)
        finally:
            self._workspace = shared_ws

        def run():
            graph.replay()
            return out
        run.graph, run.frames, run.out, run.workspace = graph, frames, out, own_ws
        return run

    def classify_host(self, frames, *, bgr=False, return_logits=False):
        """Host uint8 frames (numpy array or CPU tensor, (B,H,W,3)) -> numpy probabilities.  Copies in
        and out happen inside the library, overlapped with the kernels (``predict()`` of
        aider-predict.py:47-86 for a batch)."""
        if isinstance(frames, torch.Tensor):
            if frames.device.type != "cpu":
                raise ValueError("classify_host expects host memory; use forward_frames for device tensors")
            arr = frames.contiguous().numpy()
        else:
            arr = np.ascontiguousarray(frames)
        if arr.dtype != np.uint8 or arr.ndim != 4 or arr.shape[3] != 3:
            raise ValueError(f"expected uint8 frames of shape (B,H,W,3), got {arr.dtype} {arr.shape}")
        lib, h, _ = self._ensure_engine()
        B, H, W, _c = arr.shape
        probs = np.empty((B, 5), dtype=np.float32)
        logits = np.empty((B, 5), dtype=np.float32) if return_logits else None
        _lib.check(lib.ernet_classify_frames_host(h, arr.ctypes.data, B, H, W, _lib.BGR if bgr else _lib.RGB,
                                                  probs.ctypes.data, logits.ctypes.data if return_logits else None))
        return (probs, logits) if return_logits else probs

    def classify_host_submit(self, frames, *, bgr=False, return_logits=False):
        """Streaming form of classify_host(): enqueue one batch of host frames and return a ticket whose ``result()``
        blocks for the probabilities.  Keep one ticket outstanding while submitting the next batch and the host->device
        copy of batch i+1 runs under the kernels of batch i (at most two tickets in flight).  Outputs live in pinned
        memory; pass pinned frames (``tensor.pin_memory()``) for a truly asynchronous copy."""
        if isinstance(frames, torch.Tensor):
            if frames.device.type != "cpu":
                raise ValueError("classify_host expects host memory; use forward_frames for device tensors")
            keep = frames.contiguous()
            ptr, shape, dt = keep.data_ptr(), tuple(keep.shape), keep.dtype == torch.uint8
        else:
            keep = np.ascontiguousarray(frames)
            ptr, shape, dt = keep.ctypes.data, keep.shape, keep.dtype == np.uint8
        if not dt or len(shape) != 4 or shape[3] != 3:
            raise ValueError(f"expected uint8 frames of shape (B,H,W,3), got {shape}")
        lib, h, _ = self._ensure_engine()
        B, H, W, _c = shape
        probs = torch.empty((B, 5), dtype=torch.float32).pin_memory()
        logits = torch.empty((B, 5), dtype=torch.float32).pin_memory() if return_logits else None
        tk = C.c_int(0)
        _lib.check(lib.ernet_classify_frames_host_submit(h, ptr, B, H, W, _lib.BGR if bgr else _lib.RGB, probs.data_ptr(),
                                                         logits.data_ptr() if return_logits else None, C.byref(tk)))
        return _HostTicket(self, lib, h, tk.value, keep, probs, logits)

    def prepare_ingest(self, height, width):
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_prepare_ingest(h, int(height), int(width)))

    def ingest(self, frames, *, bgr=False, dtype=torch.float32):
        """The eval transform alone: uint8 (B,H,W,3) device frames -> (B,3,140,140) tensor ((B,3,240,240) for ErNET:
        ``aider_transforms``, dataloaders/aider.py:430)."""
        lib, h, _ = self._ensure_engine()
        frames = frames.contiguous()
        B, H, W, _c = frames.shape
        out = torch.empty((B, 3, self.IN_HW, self.IN_HW), dtype=dtype, device=frames.device)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        _lib.check(lib.ernet_ingest_u8(h, frames.data_ptr(), B, H, W, _lib.BGR if bgr else _lib.RGB,
                                       out.data_ptr(), _DTYPE_CODE[dtype], _lib.NCHW, stream))
        return out

    # ------------------------------------------------------------------ introspection (tests)
    def tap(self, name):
        """fp32 NCHW copy of an intermediate of the most recent forward (B <= chunk)."""
        lib, h, idx = self._ensure_engine()
        Cc, H, W = _TAP_SHAPES[name](self.ARCH)
        B = self._last_batch
        out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=self._workspace.device)
        stream = torch.cuda.current_stream(out.device).cuda_stream
        _lib.check(lib.ernet_debug_tap(h, _lib.TAPS[name], self._workspace.data_ptr(), B, out.data_ptr(),
                                       out.numel(), stream))
        return out

    def profile(self, on=True):
        """Turn per-stage CUDA-event timing on/off (see ernet_profile_enable)."""
        lib, h, _ = self._ensure_engine()
        _lib.check(lib.ernet_profile_enable(h, 1 if on else 0))

    def profile_read(self):
        """{stage: (total_ms, launches)} since the previous read."""
        lib, h, _ = self._ensure_engine()
        n = len(_lib.STAGES)
        ms = (C.c_double * n)()
        cnt = (C.c_int * n)()
        _lib.check(lib.ernet_profile_read(h, ms, cnt, n))
        return {_lib.STAGES[i]: (ms[i], cnt[i]) for i in range(n) if cnt[i]}

    def host_copy_bytes_per_frame(self, height, width):
        """Bytes of one frame that classify_host() sends over PCIe (the rows the crop window reads)."""
        lib, h, _ = self._ensure_engine()
        return int(lib.ernet_host_copy_bytes_per_frame(h, int(height), int(width)))

    def launches_per_forward(self, batch, with_ingest=True):
        lib, h, _ = self._ensure_engine()
        return lib.ernet_launches_per_forward(h, int(batch), 1 if with_ingest else 0)


class _ErnetB200(_EngineRuntime, nn.Module):
    ARCH = None

    def __init__(self, *, precision=None, device=None):
        super().__init__()
        arch = self.ARCH
        red = arch == "squeeze-redconv"
        w = widths(arch)
        self.IN_HW = 240 if arch == "ernet" else 140     # ErNET is the 240x240-native model (model/ernet.py:21-22,42)
        self.conv1 = nn.Conv2d(3, 16, 3, 2, 0, bias=False)
        if red:
            self.conv_red1 = nn.Conv2d(16, 8, 1)
        self.acff1 = _ACFFParams(*w[0])
        self.acff2 = _ACFFParams(*w[1])
        if red:
            self.conv_red2 = nn.Conv2d(96, 48, 1)
        self.acff3 = _ACFFParams(*w[2])
        if red:
            self.conv_red3 = nn.Conv2d(128, 64, 1)
        self.acff4 = _ACFFParams(*w[3])
        if arch == "ernet":
            self.acff5 = _ACFFParams(*w[4])
            self.acff6 = _ACFFParams(*w[5])
        self.conv2 = nn.Conv2d(256, 5, 1, bias=False)
        self.fc = nn.Linear(3 * 3 * 5 if arch == "ernet" else 2 * 2 * 5, 5)
        if precision is not None and precision not in _lib.PRECISION:
            raise ValueError(f"unknown precision {precision!r}; expected one of {sorted(_lib.PRECISION)}")
        self._precision = precision          # None: follow the parameter dtype (fp32, or fp16 after .half())
        self._engine = None                  # (handle, device_index, precision)
        self._fingerprint = None
        self._workspace = None
        self._last_batch = 0
        self._act_scales = None              # int8: calibrated per-tensor activation scales (stem, pool1, pool2)
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------ precision / engine
    @property
    def precision(self):
        if self._precision is not None:
            return self._precision
        return _PREC_OF_DTYPE.get(self.conv1.weight.dtype, "fp32")

    def set_precision(self, precision):
        if precision not in _lib.PRECISION:
            raise ValueError(f"unknown precision {precision!r}")
        self._precision = precision
        return self

    def _weights_fingerprint(self):
        fp = []
        for t in list(self.parameters()) + list(self.buffers()):
            fp.append((t.data_ptr(), t._version))
        return tuple(fp)

    def _release(self):
        if self._engine is not None:
            _lib.load().ernet_destroy(self._engine[0])
            self._engine = None
            self._fingerprint = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_engine(self):
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("this model runs only on a CUDA device (B200, sm_100a); call .to('cuda') first — "
                               "there is no CPU fallback")
        if self.training:
            raise RuntimeError("inference-only engine: call model.eval() first (train mode would need dropout and "
                               "batch statistics, acff.py:34-35)")
        lib = _lib.load()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        prec = self.precision
        if self._engine is None or self._engine[1:] != (idx, prec):
            self._release()
            h = C.c_void_p()
            _lib.check(lib.ernet_create(C.byref(h), _lib.ARCH[self.ARCH], _lib.PRECISION[prec], idx))
            self._engine = (h, idx, prec)
        if prec == "int8" and self._act_scales is None:
            self.calibrate()                 # default synthetic calibration set
        fp = (self._weights_fingerprint(), self._act_scales)
        if fp != self._fingerprint:
            blob = pack_state_dict(self.state_dict(), self.ARCH, prec, self._act_scales)
            buf = (C.c_char * len(blob)).from_buffer_copy(blob)
            _lib.check(lib.ernet_load_packed(self._engine[0], buf, len(blob)))
            self._fingerprint = fp
        return lib, self._engine[0], idx

    # ------------------------------------------------------------------ int8 calibration
    def calibrate(self, frames=None, *, percentile=100.0, per_channel=True, batch=128):
        """Activation scales for the int8 engine (the TensorRT scheme behind the reference's int8 artefacts:
        symmetric int8, activation scales from a calibration set, per-output-channel weight scale
        max|W|/127, int32 accumulate, fp32 dequant + bias).

        ``frames``: uint8 (N,H,W,3) calibration frames (numpy / tensor); default = 512 synthetic 240x240
        frames, seed 99 (half uniform noise, half smooth).  An fp32 twin of this model runs them through the
        CUDA-core engine; at the three quantised tensors (stem, pool1, pool2 outputs) the ``percentile`` of
        |activation| becomes 127 int8 steps - per channel when ``per_channel`` (cross-layer equalisation:
        the per-channel factors are folded into the producer's epilogue constants and the consumer's weights
        at pack time, the runtime tensor keeps one scale), else per tensor.  Returns the three scale arrays."""
        dev = self.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("calibration runs on the GPU; call .to('cuda') first")
        if frames is None:
            frames = default_calibration_frames()
        frames = torch.as_tensor(np.asarray(frames) if not isinstance(frames, torch.Tensor) else frames)
        twin = type(self)(precision="fp32")
        twin.load_state_dict({k: v.detach().float() if v.is_floating_point() else v.detach()
                              for k, v in self.state_dict().items()})
        twin = twin.to(dev)
        twin.eval()
        names = ("stem", "pool1", "pool2")
        amax = {n: None for n in names}
        samples = {n: [] for n in names}
        for i in range(0, frames.shape[0], batch):
            twin.forward_frames(frames[i:i + batch].to(dev))
            for n in names:
                t = twin.tap(n).abs()                                     # (b,C,H,W)
                tc_ = t.permute(1, 0, 2, 3).reshape(t.shape[1], -1)       # (C, b*H*W)
                m = tc_.max(dim=1).values
                amax[n] = m if amax[n] is None else torch.maximum(amax[n], m)
                if percentile < 100:
                    step = max(1, tc_.shape[1] // 20_000)
                    samples[n].append(tc_[:, ::step].clone())
        scales = []
        for n in names:
            if percentile < 100:
                v = torch.cat(samples[n], dim=1).double()
                r = torch.quantile(v, percentile / 100.0, dim=1) if per_channel else \
                    torch.quantile(v.flatten()[: 16_000_000], percentile / 100.0).expand(v.shape[0])
            else:
                r = amax[n].double() if per_channel else amax[n].double().max().expand(amax[n].shape[0])
            floor = float(amax[n].max()) * 1e-6 + 1e-30                   # dead channels: keep the step finite
            scales.append(tuple((torch.clamp(r, min=floor) / 127.0).cpu().tolist()))
        twin._release()
        self._act_scales = tuple(scales)
        return self._act_scales

    @property
    def act_scales(self):
        return self._act_scales

    def set_act_scales(self, scales):
        from .pack_tc import normalize_act_scales
        self._act_scales = tuple(tuple(float(x) for x in a) for a in normalize_act_scales(scales, self.ARCH))
        return self


class _HostTicket:
    """One submitted classify_host call (see classify_host_submit)."""

    def __init__(self, model, lib, handle, ticket, frames, probs, logits):
        self._model, self._lib, self._h, self._tk = model, lib, handle, ticket
        self._frames, self._probs, self._logits, self._done = frames, probs, logits, False

    def result(self):
        if not self._done:
            _lib.check(self._lib.ernet_classify_frames_host_wait(self._h, self._tk))
            self._done, self._frames = True, None
        p = self._probs.numpy()
        return (p, self._logits.numpy()) if self._logits is not None else p


def default_calibration_frames(n=512, seed=99):
    """Synthetic calibration set: half i.i.d. uniform uint8 frames, half smooth frames (coarse noise
    upsampled by pixel replication), 240x240x3."""
    rs = np.random.RandomState(seed)
    noise = rs.randint(0, 256, (n // 2, 240, 240, 3)).astype(np.uint8)
    coarse = rs.randint(0, 256, (n - n // 2, 15, 15, 3)).astype(np.uint8)
    smooth = np.repeat(np.repeat(coarse, 16, axis=1), 16, axis=2)
    return np.concatenate([noise, smooth], 0)


class Squeeze_ErNET(_ErnetB200):
    """B200 engine behind the reference's ``Squeeze_ErNET`` surface (model/squeeze_ernet.py:7-46)."""
    ARCH = "squeeze-ernet"


class Squeeze_RedConv(_ErnetB200):
    """B200 engine behind the reference's ``Squeeze_RedConv`` surface (model/squeeze_ernet_redconv.py:7-52)."""
    ARCH = "squeeze-redconv"


class ErNET(_ErnetB200):
    """Baseline ErNET (model/ernet.py:6-49): six ACFF blocks, (B,3,240,240) inputs, fc 45 -> 5; same 82 state_dict keys.
    SURVEY.md section 8f-1.  bf16 / fp16: all six ACFF blocks on the tcgen05 block kernels (block 1 persistent, blocks
    2-6 on CTA pairs, block 6 with N = 256), conv1 and the collapsed head on CUDA cores; fp32: the layer-wise CUDA-core
    kernels.  ``model(x)`` on (B,3,240,240) tensors; ``forward_frames`` / ``classify_host`` run ``aider_transforms`` (Resize(273) ->
    CenterCrop(240) -> ToTensor -> Normalize, dataloaders/aider.py:430) on the table-driven ingest kernel first (not fused
    into conv1 as for the Squeeze models).  int8 is not wired for it."""
    ARCH = "ernet"


def load_model(model_name, weights_path, device, *, precision=None):
    """Mirror of ``load_model`` in aider-predict.py:22-45 / evaluate-classification-metrics.py:24-47."""
    if model_name == "squeeze-ernet":
        model = Squeeze_ErNET(precision=precision)
    elif model_name == "squeeze-redconv":
        model = Squeeze_RedConv(precision=precision)
    elif model_name == "ernet":                      # aider-predict.py:25
        model = ErNET(precision=precision)
    else:
        raise ValueError(f"Unsupported model: {model_name}")
    checkpoint = torch.load(weights_path, map_location="cpu", weights_only=True)
    if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint:
        model.load_state_dict(checkpoint["model_state_dict"])
    else:
        model.load_state_dict(checkpoint)
    model = model.to(device)
    model.eval()
    return model


def from_state_dict(arch, sd, device="cuda", precision="fp32"):
    """Build an eval-mode engine from a mapping of numpy arrays / tensors."""
    cls = {"squeeze-ernet": Squeeze_ErNET, "squeeze-redconv": Squeeze_RedConv, "ernet": ErNET}.get(arch)
    if cls is None:
        raise ValueError(f"Unsupported model: {arch}")
    m = cls(precision=precision)
    m.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})
    m = m.to(device)
    m.eval()
    return m
