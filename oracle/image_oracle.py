"""CPU oracle of the image transform in front of the model -- TEST INFRASTRUCTURE ONLY (imported by tests/ and
tests/golden/*; the product path never touches it).

Restates, in numpy, what `PrismaticImageProcessor.apply_transform` does to one PIL image
(prismatic/extern/hf/processing_prismatic.py:128-145, parameters set up at :70-123):

    [letterbox_pad_transform (:23-29)] -> TVF.resize(bicubic, antialias) -> TVF.center_crop -> TVF.to_tensor -> TVF.normalize

The arithmetic lives in two third-party dependencies that are not vendored in the reference tree:
  * Pillow (un-pinned transitive dependency of torchvision==0.17.0, pyproject.toml:48): `Image.resize(size, BICUBIC)` ->
    `ImagingResample` (src/libImaging/Resample.c): separable two-pass convolution, horizontal pass first; per output
    pixel a window [xmin, xmin + n) of the input with bicubic (a = -0.5) weights evaluated in double precision at
    (x + xmin - center + 0.5) / filterscale, normalised to sum 1, then turned into 22-bit fixed point
    (`normalize_coeffs_8bpc`: int(+-0.5 + k * 2^22)); pixels accumulate in int32 from 2^21 and are shifted / clipped to
    uint8 after EACH pass.  That published algorithm is restated here.
  * torchvision.transforms.functional: `resize` of a PIL image with an int size keeps the aspect ratio
    (`_compute_resized_output_size`: short side -> size, long side -> int(size * long / short)); `center_crop` takes
    the window at int(round((h - th) / 2.0)) (Python's round-half-to-even); `to_tensor` = uint8 / 255 in float32,
    `normalize` = (x - mean) / std in float32.
Pinned (tests/golden/make_image_golden.py, tests/test_image_transform.py): against the reference's own, unmodified
`PrismaticImageProcessor` executed with the installed Pillow / torchvision (timm.data.create_transform stubbed by the
four torchvision transforms timm 0.9.10 returns for an eval transform), bit for bit on the uint8 frames and exactly on
the float32 tensors, for all three `image_resize_strategy` values.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c


def _bicubic(x: float) -> float:
    """Resample.c bicubic_filter, a = -0.5."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = [0, in_size)).
    Returns (ksize, bounds int32 [out, 2] = (xmin, n), kk int32 [out, ksize])."""
    support_base = 2.0
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = support_base * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h), Image.BICUBIC)` for a uint8 [H, W, C] array."""
    H, W, C = img.shape
    _, bx, kx = resample_coeffs(W, out_w)
    _, by, ky = resample_coeffs(H, out_h)
    src = img.astype(np.int64)
    tmp = np.empty((H, out_w, C), dtype=np.uint8)
    for xo in range(out_w):
        x0, n = bx[xo]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kx[xo, :n].astype(np.int64), axes=([1], [0]))
        tmp[:, xo, :] = _clip8(acc)
    t = tmp.astype(np.int64)
    out = np.empty((out_h, out_w, C), dtype=np.uint8)
    for yo in range(out_h):
        y0, n = by[yo]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(ky[yo, :n].astype(np.int64), t[y0:y0 + n], axes=([0], [0]))
        out[yo] = _clip8(acc)
    return out


def letterbox_pad(img: np.ndarray, fill: Sequence[int]) -> np.ndarray:
    """processing_prismatic.py:23-29: symmetric border of int((max - side) / 2) on both sides (an odd difference leaves
    the image one pixel short of square, as in the reference)."""
    H, W, C = img.shape
    m = max(H, W)
    hp, vp = int((m - W) / 2), int((m - H) / 2)
    out = np.empty((H + 2 * vp, W + 2 * hp, C), dtype=np.uint8)
    out[...] = np.asarray(fill, dtype=np.uint8)
    out[vp:vp + H, hp:hp + W] = img
    return out


def resized_output_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision `_compute_resized_output_size` for an int size, no max_size: (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_offsets(h: int, w: int, th: int, tw: int) -> Tuple[int, int]:
    """torchvision center_crop: top / left = int(round((side - target) / 2.0)), Python rounding (half to even)."""
    return int(round((h - th) / 2.0)), int(round((w - tw) / 2.0))


def transform_u8(img: np.ndarray, strategy: str, size: int = 224, letterbox_fill: Sequence[int] = (127, 127, 127)) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [size, size, 3]: everything of apply_transform before to_tensor (the same for every
    tower of openvla: both resize to 224 with bicubic)."""
    if strategy == "letterbox":
        img = letterbox_pad(img, letterbox_fill)
    H, W, _ = img.shape
    if strategy == "resize-naive":
        oh, ow = size, size
    elif strategy in ("resize-crop", "letterbox"):
        oh, ow = resized_output_size(H, W, size)
    else:
        raise ValueError(f"Image resize strategy `{strategy}` is not supported!")
    r = resize_bicubic_u8(img, ow, oh)
    top, left = center_crop_offsets(oh, ow, size, size)
    if oh < size or ow < size:
        raise ValueError("center_crop would pad: not reachable with these strategies")
    return np.ascontiguousarray(r[top:top + size, left:left + size])


def to_tensor_normalize(frame_u8: np.ndarray, means, stds) -> np.ndarray:
    """TVF.to_tensor + TVF.normalize per tower, channel-stacked (processing_prismatic.py:136-143): float32 [3 * towers, S, S]."""
    x = frame_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    outs = []
    for m, s in zip(means, stds):
        m32 = np.asarray(m, dtype=np.float32).reshape(3, 1, 1)
        s32 = np.asarray(s, dtype=np.float32).reshape(3, 1, 1)
        outs.append((x - m32) / s32)
    return np.concatenate(outs, 0)
