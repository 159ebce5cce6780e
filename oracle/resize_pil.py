"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the first step of every reference transform, `transforms.Resize((224, 224))` on a PIL image
(notebooks/train_multimodal_fusion.py:174,186,195,204; extended_metrics.py:290-300), i.e. torchvision's
`F.resize` -> `PIL.Image.resize(size, BILINEAR)`.  The arithmetic lives in a third-party dependency that is not vendored
under /root/reference: Pillow (requirements.txt:28 `Pillow>=9.5.0`; 12.2.0 installed here), src/libImaging/Resample.c --
`precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`:
a separable triangle filter whose support is stretched by the down-scaling factor (antialiasing), coefficients in
float64, converted to 22-bit fixed point, a horizontal pass and then a vertical pass, EACH rounding to uint8.

PINNING: tests/test_oracle.py checks this restatement bit for bit against Pillow itself (the same library the reference
calls) on seeded images of many sizes; the CUDA kernel (dfu_resize_u8) is then checked bit for bit against both."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the full image box.
    Returns (bounds int32 [out, 2] = (first input index, tap count), coeffs int32 [out, ksize], ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                       # bilinear filter support = 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
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
        ws = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            ws.append(w)
            ww += w
        for x in range(xmax):
            w = ws[x] / ww if ww != 0.0 else ws[x]
            coeffs[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs, ksize


def _pass(img, bounds, coeffs, axis):
    """One 8-bit resampling pass along `axis` (1 = horizontal over W, 0 = vertical over H) of an HWC uint8 array."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)              # [in, other, C]
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        acc += np.tensordot(coeffs[o, :n].astype(np.int64), src[lo:lo + n], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img_hwc, out_h=224, out_w=224):
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C], bit-identical to PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR).
    Pillow skips a pass whose size does not change; an identity pass is exact in this arithmetic, so running both is equal."""
    img = np.ascontiguousarray(img_hwc)
    h, w = img.shape[:2]
    if w != out_w:
        b, c, _ = precompute_coeffs(w, out_w)
        img = _pass(img, b, c, axis=1)
    if h != out_h:
        b, c, _ = precompute_coeffs(h, out_h)
        img = _pass(img, b, c, axis=0)
    return img
