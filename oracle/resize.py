"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the image resize the reference performs before
its ViT / VAE transforms.

Reference call site: ``MaxLongEdgeMinShortEdgeResize.forward`` -> ``torchvision F.resize(PIL image, (h, w), BICUBIC,
antialias=True)`` (codes/data/transforms.py:60-87), i.e. ``PIL.Image.resize((w, h), Image.BICUBIC)``.  The arithmetic lives
in a third-party dependency that is not under /root/reference: Pillow (``pillow`` unpinned in the reference's
environment.yaml; 12.2 in this image), ``src/libImaging/Resample.c``.  Its published algorithm for 8-bit images, restated:

  * per output coordinate a window of input taps: centre = (xx + 0.5) * scale, support = 2 * max(scale, 1) (bicubic, a = -0.5,
    widened by the scale when shrinking = the antialiasing), xmin = int(centre - support + 0.5) clipped to 0,
    xmax = int(centre + support + 0.5) clipped to the input size; weights = filter((x - centre + 0.5) / max(scale, 1)),
    normalised to sum 1 in double precision (``precompute_coeffs``);
  * weights converted to fixed point with 22 fractional bits, rounded half away from zero (``normalize_coeffs_8bpc``);
  * a HORIZONTAL pass over the rows the vertical pass needs, then a VERTICAL pass, each ``(2^21 + sum(pixel * k)) >> 22``
    clipped to 0..255 -- so the intermediate image is rounded to 8 bits (``ImagingResampleHorizontal_8bpc`` / ``Vertical``).

Pinned against Pillow itself in tests/test_resize.py (bit-exact over up- and down-scaling cases)."""
from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def coefficients(in_size: int, out_size: int):
    """(bounds int32 [out, 2] = (xmin, count), kk int32 [out, ksize]) for one axis."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size          # box is float in the C signature
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _bicubic((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                    # sequential double sum, as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = w * float(1 << PRECISION_BITS)
        kk[xx, :xmax] = np.where(w < 0, (-0.5 + fixed).astype(np.int64), (0.5 + fixed).astype(np.int64)).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """Resample axis 1 of img [rows, in, C] uint8 -> [rows, out, C] uint8."""
    out = np.empty((img.shape[0], bounds.shape[0], img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx, (xmin, n) in enumerate(bounds):
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, xmin:xmin + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-identical to PIL.Image.resize((out_w, out_h), BICUBIC)."""
    H, W, _ = img.shape
    x = img
    if out_w != W:
        bh, kh = coefficients(W, out_w)
        x = _pass(x, bh, kh)
    if out_h != H:
        bv, kv = coefficients(H, out_h)
        x = _pass(x.transpose(1, 0, 2), bv, kv).transpose(1, 0, 2)
    return np.ascontiguousarray(x)
