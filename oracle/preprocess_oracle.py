"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's image transform (the loader side of the hot
path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Reference call site: dataset/__init__.py:127-136 (and :53-62 for the training set)

    transforms.Resize((img_size, img_size), Image.BICUBIC)  ->  PIL.Image.resize (uint8 in, uint8 out)
    transforms.ToTensor()                                   ->  HWC uint8 -> CHW float32 / 255
    transforms.Normalize(mean, std)                         ->  (x - mean) / std in float32

The resize arithmetic lives in an un-vendored third-party dependency: Pillow (pulled in by torchvision>=0.15.0,
requirements.txt:8, itself unpinned; this image has Pillow 12.2.0).  What follows restates Pillow's published
algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
ImagingResampleVertical_8bpc, bicubic_filter) with its exact integer arithmetic:

  * filter: Keys cubic, a = -0.5, support 2; when shrinking, the support and the argument are stretched by the scale
    (antialiasing);  window [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image;
  * coefficients are normalised in double, then rounded to fixed point with 22 fractional bits
    (round-half-away-from-zero); a pixel is  clip8((2^21 + sum(pixel * coeff)) >> 22);
  * two passes, horizontal first, with a uint8 intermediate; a pass whose input and output sizes agree is skipped.

PINNED: oracle/make_preprocess_golden.py checks it bit for bit against PIL + torchvision in the build container and
writes tests/golden/preprocess_*.npz; tests/test_preprocess.py re-checks against PIL wherever PIL is importable.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # dataset/__init__.py:130-133
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c:precompute_coeffs + normalize_coeffs_8bpc for the whole-image box.
    Returns (ksize, bounds[out,2] = (xmin, count), kk[out,ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
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
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One pass along axis 0 of a uint8 array [N, ...]."""
    n = img.shape[0]
    if n == out_size:
        return img
    _, bounds, kk = precompute_coeffs(n, out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for i in range(out_size):
        x0, cnt = int(bounds[i, 0]), int(bounds[i, 1])
        k = kk[i, :cnt].astype(np.int64).reshape((cnt,) + (1,) * (img.ndim - 1))
        acc = (src[x0:x0 + cnt] * k).sum(0) + (1 << (PRECISION_BITS - 1))
        out[i] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, size: int) -> np.ndarray:
    """PIL.Image.resize((size, size), BICUBIC) on an HWC uint8 image: horizontal pass, then vertical."""
    h = _resample_axis0(np.ascontiguousarray(img.transpose(1, 0, 2)), size).transpose(1, 0, 2)   # along W
    return _resample_axis0(np.ascontiguousarray(h), size)                                        # along H


def transform_x(img: np.ndarray, size: int) -> np.ndarray:
    """dataset/__init__.py:127-136: HWC uint8 -> CHW float32, resized and CLIP-normalised."""
    r = resize_bicubic_u8(img, size).transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    mean = np.asarray(CLIP_MEAN, np.float32).reshape(3, 1, 1)
    std = np.asarray(CLIP_STD, np.float32).reshape(3, 1, 1)
    return np.ascontiguousarray(((r - mean) / std).astype(np.float32))


def synth_image(h: int, w: int, seed: int) -> np.ndarray:
    """Deterministic HWC uint8 test image: smooth gradients + texture + noise (exercises clipping at 0 / 255)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = 128 + 100 * np.sin(xx * (0.05 + 0.03 * c) + seed) * np.cos(yy * (0.04 + 0.02 * c)) \
            + 80 * ((xx.astype(np.int64) // (7 + c) + yy.astype(np.int64) // (5 + c)) % 2) - 40
    img += rng.randint(-60, 61, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)
