"""ADD-fusion ACFF block of the detector half behind the reference's module surface.

Drop-in for ``class ACFF`` of ``code/victim_localization/yolov3/models.py:265-315`` as constructed by
``create_modules`` (models.py:46-56): ``ACFF(in_channels, out_channels, kernel_size)`` with the same ten state_dict
entries (conv1/2/3 weight+bias, fused_conv weight+bias, batch_norm weight/bias/running_mean/running_var/
num_batches_tracked) and ``module(x)``: (B,C,H,W) -> (B,Cout,H-2,W-2).

Two kernels of ``libernet_b200.so`` do the arithmetic (fp32, eval mode only; there is no CPU or PyTorch fallback):

* ``ernet_acff_add_depthwise`` - conv1(x)+conv2(x)+conv3(x) (models.py:302) on the register-tile depthwise kernel
  (one channel x 4x4 output patch per thread, the 27 tap weights in registers), NHWC in, NHWC out;
* ``ernet_pointwise`` - fused_conv + bias + LeakyReLU(0.01) + BatchNorm (eval: per-channel scale/shift)
  (models.py:307-309); output channel counts that are not a multiple of 32 (the detector's 14-channel heads,
  yolov3-acffx.cfg:349) are zero-padded to the next multiple of 32 at pack time and sliced off the result.

The torch.nn leaf modules are parameter containers; none of them is ever called.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-5


class ACFF(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3):
        super().__init__()
        if kernel_size != 3:
            raise ValueError("the B200 ACFF kernels implement kernel_size=3 (every [acff] entry of yolov3-acffx.cfg)")
        if in_channels % 8:
            raise ValueError("in_channels must be a multiple of 8")
        c = in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv1 = nn.Conv2d(c, c, 3, 1, 0, 1, groups=c, bias=True)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1, 2, groups=c, bias=True)
        self.conv3 = nn.Conv2d(c, c, 3, 1, 2, 3, groups=c, bias=True)
        self.fused_conv = nn.Conv2d(c, out_channels, 1, bias=True)
        self.leaky_relu = nn.LeakyReLU(0.01)
        self.batch_norm = nn.BatchNorm2d(out_channels)
        self.dropout = nn.Dropout(0.2)
        self._packed = None
        self._fingerprint = None

    def _pack(self):
        """fp32 device constants in the layouts the kernels read: depthwise taps [3][9][C], biases [3][C], fused_conv as
        [K][Npad], bias / BN scale / BN shift [Npad]."""
        fp = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._packed is not None and fp == self._fingerprint:
            return self._packed
        dev = self.conv1.weight.device
        c, n = self.in_channels, self.out_channels
        npad = (n + 31) // 32 * 32
        f = lambda t: t.detach().to(torch.float64)
        dw = torch.stack([f(getattr(self, f"conv{j}").weight).reshape(c, 9).t() for j in (1, 2, 3)], 0)      # [3][9][C]
        db = torch.stack([f(getattr(self, f"conv{j}").bias) for j in (1, 2, 3)], 0)                              # [3][C]
        pw = torch.zeros(c, npad, dtype=torch.float64, device=dev)
        pw[:, :n] = f(self.fused_conv.weight).reshape(n, c).t()
        pb = torch.zeros(npad, dtype=torch.float64, device=dev)
        pb[:n] = f(self.fused_conv.bias)
        bn = self.batch_norm
        s = f(bn.weight) / torch.sqrt(f(bn.running_var) + BN_EPS)
        t = f(bn.bias) - f(bn.running_mean) * s
        bs = torch.ones(npad, dtype=torch.float64, device=dev)
        bt = torch.zeros(npad, dtype=torch.float64, device=dev)
        bs[:n], bt[:n] = s, t
        self._packed = tuple(v.to(torch.float32).contiguous() for v in (dw, db, pw, pb, bs, bt)) + (npad,)
        self._fingerprint = fp
        return self._packed

    def forward(self, x):
        if self.training:
            raise RuntimeError("inference-only kernels: call .eval() first (train mode needs dropout and batch statistics, "
                               "models.py:309-310)")
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected a (B,{self.in_channels},H,W) tensor, got {tuple(getattr(x, 'shape', ()))}")
        if x.device.type != "cuda" or x.device != self.conv1.weight.device:
            raise RuntimeError("ACFF runs only on a CUDA device shared by the module and its input (no CPU fallback)")
        if x.dtype != torch.float32:
            raise ValueError("fp32 only")
        B, C, H, W = x.shape
        if B < 1 or H < 3 or W < 3:
            raise ValueError("input too small: the block's convolutions produce (H-2, W-2)")
        lib = _lib.load()
        dw, db, pw, pb, bs, bt, npad = self._pack()
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()                      # no copy for channels_last inputs
        stream = torch.cuda.current_stream(x.device).cuda_stream
        oh, ow = H - 2, W - 2
        s = torch.empty((B, oh, ow, C), dtype=torch.float32, device=x.device)
        _lib.check(lib.ernet_acff_add_depthwise(x_nhwc.data_ptr(), _lib.DTYPE_F32, B, H, W, C, oh, ow, dw.data_ptr(),
                                                db.data_ptr(), s.data_ptr(), stream))
        y = torch.empty((B, oh, ow, npad), dtype=torch.float32, device=x.device)
        _lib.check(lib.ernet_pointwise(s.data_ptr(), _lib.DTYPE_F32, B, oh, ow, C, npad, pw.data_ptr(), pb.data_ptr(),
                                       bs.data_ptr(), bt.data_ptr(), 1, 0, y.data_ptr(), stream))
        return y[..., :self.out_channels].permute(0, 3, 1, 2)           # NCHW view of the NHWC result (channels_last)
