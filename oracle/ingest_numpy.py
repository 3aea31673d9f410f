"""CPU oracle for the ingest path (uint8 HWC frame -> normalised 3x140x140 tensor).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/ernet_numpy.py header).

Restates ``squeeze_transforms = get_val_torchvision_transforms(140)``
(code/disaster_detection/dataloaders/aider.py:412-426,431):

    Resize(int(140*1.14)=159) -> CenterCrop(140) -> ToTensor -> Normalize(mean, std)

The arithmetic lives in third-party code that is absent from /root/reference:
torchvision 0.8.2 (``transforms.Resize`` on a PIL image -> ``Image.resize(BILINEAR)``)
and Pillow 8.1.2 (requirements-fyp.txt:109,194,198).  Pillow's 8-bit resampler
(``ImagingResample`` in src/libImaging/Resample.c) is restated from its published
algorithm: a separable triangle filter widened by the down-scale factor, coefficients
normalised in double then quantised to 22-bit fixed point, horizontal pass first into a
uint8 temporary, then the vertical pass, each with round-half-up and clipping.

Parity pinning: ``tests/golden/make_golden.py`` runs the real torchvision/Pillow
transform in the build container on seeded frames of several sizes (down- and
up-scaling, non-square) and commits inputs+outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` requires bit equality of the uint8 crop and of the
fp32 normalised tensor.
"""
from __future__ import annotations

import math

import numpy as np

RESIZE_SHORT = 159            # int(140 * 1.14), aider.py:422
CROP = 140                    # aider.py:423
MEAN = (0.485, 0.456, 0.406)  # aider.py:425
STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2   # Pillow Resample.c


def resized_size(h, w, short=RESIZE_SHORT):
    """torchvision ``Resize(int)``: shorter side -> ``short``, other = int(short*long/short_side)."""
    if w <= h:
        return int(short * h / w), short        # (new_h, new_w)
    return short, int(short * w / h)


def center_crop_offset(size, crop=CROP):
    """torchvision ``CenterCrop``: int(round((size - crop) / 2.0)) with Python's banker's rounding."""
    return int(round((size - crop) / 2.0))


def resample_coeffs(in_size, out_size):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the BILINEAR filter.

    Returns (bounds_min[out], bounds_len[out], kk[out, ksize] int32)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                     # bilinear filter support = 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, dtype=np.int32)
    xlen_a = np.zeros(out_size, dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            a = -a if a < 0 else a
            w = 1.0 - a if a < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            k[:xmax] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(v * (1 << PRECISION_BITS) - 0.5) if v < 0 else int(v * (1 << PRECISION_BITS) + 0.5)
        xmin_a[xx] = xmin
        xlen_a[xx] = xmax
    return xmin_a, xlen_a, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis(img, out_size, axis):
    """One pass of the 8-bit resampler along ``axis`` (0 = vertical, 1 = horizontal).

    Pillow skips a pass whose in and out sizes are equal."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    xmin, xlen, kk = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)        # (in, other, C)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for o in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(int(xlen[o])):
            acc += src[xmin[o] + t] * int(kk[o, t])
        out[o] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_u8(frame, out_h, out_w):
    """PIL ``Image.resize((out_w, out_h), BILINEAR)`` on an (H,W,3) uint8 RGB array:
    horizontal pass first, then vertical."""
    tmp = _resample_axis(frame, out_w, axis=1)
    return _resample_axis(tmp, out_h, axis=0)


def crop_u8(frame, image_size=CROP):
    """Resize(int(image_size*1.14)) + CenterCrop(image_size) on an (H,W,3) uint8 frame -> (image_size,image_size,3) uint8
    (get_val_torchvision_transforms, aider.py:412-426: image_size 140 -> Resize(159) for the Squeeze models, 240 ->
    Resize(273) for ErNET, aider.py:430-431)."""
    frame = np.asarray(frame)
    if frame.ndim != 3 or frame.shape[2] != 3 or frame.dtype != np.uint8:
        raise ValueError("expected (H,W,3) uint8")
    short = int(image_size * 1.14)                                # aider.py:422
    nh, nw = resized_size(frame.shape[0], frame.shape[1], short)
    if nh < image_size or nw < image_size:
        raise ValueError("resized frame smaller than the crop")
    r = resize_u8(frame, nh, nw)
    top = center_crop_offset(nh, image_size)
    left = center_crop_offset(nw, image_size)
    return r[top:top + image_size, left:left + image_size, :]


def normalise_lut():
    """(256,3) fp32 table of ToTensor+Normalize: ((v/255) - mean)/std, each step in fp32
    exactly as torch does (`.div(255)`, `.sub_(mean)`, `.div_(std)`)."""
    v = np.arange(256, dtype=np.float32)[:, None] / np.float32(255.0)
    mean = np.asarray(MEAN, dtype=np.float32)[None, :]
    std = np.asarray(STD, dtype=np.float32)[None, :]
    return ((v - mean) / std).astype(np.float32)


def ingest(frames, image_size=CROP):
    """frames: (B,H,W,3) uint8 (or a list of (H,W,3) of one size) -> (B,3,image_size,image_size) fp32 NCHW."""
    lut = normalise_lut()
    outs = []
    for f in frames:
        c = crop_u8(f, image_size)                               # (S,S,3)
        t = np.stack([lut[c[:, :, ch], ch] for ch in range(3)], axis=0)
        outs.append(t)
    return np.stack(outs, axis=0).astype(np.float32)
