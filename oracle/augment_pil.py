"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's TRAIN-TIME augmentation (notebooks/train_multimodal_fusion.py:173-199,
train_thermal_only.py:103-112) as torchvision applies it to PIL images, GIVEN the drawn parameters:

    Resize((224,224)) -> RandomHorizontalFlip -> RandomVerticalFlip -> RandomRotation(30)
    -> [ColorJitter(brightness .3, contrast .3, saturation .3)]  (RGB only)
    -> [RandomAffine(degrees 20, translate (.1,.1), scale (.8,1.2))] -> ToTensor -> Normalize

The arithmetic lives in third-party code that is not vendored under /root/reference: torchvision (transforms.functional:
rotate / affine build an inverse affine matrix and call PIL.Image.transform(AFFINE, NEAREST, fill 0); adjust_* call
PIL.ImageEnhance) and Pillow (libImaging Geometry.c `affine_fixed`: 16.16 fixed-point source coordinates; Blend.c
`ImagingBlend`; Convert.c `rgb2l`: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16).
PINNING: tests/test_oracle.py checks every function here bit for bit against torchvision.transforms.functional on PIL
images (the calls the reference's transforms make); the CUDA kernel (dfu_augment_u8) is then checked against both."""
import math

import numpy as np


# ------------------------------------------------------------------------------------------------ geometry
def rotate_matrix(angle, w, h):
    """PIL.Image.rotate(angle, expand=False, center=None, translate=None): the inverse map handed to transform(AFFINE)."""
    angle = angle % 360.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    cx, cy = w / 2.0, h / 2.0
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return m


def inverse_affine_matrix(angle, translate, scale, w, h):
    """torchvision.transforms.functional._get_inverse_affine_matrix with shear (0, 0) and the PIL centre (w/2, h/2), as
    F.affine(img, angle, translate, scale, shear=(0,0)) computes it (RandomAffine without shear)."""
    cx, cy = w * 0.5, h * 0.5
    rot = math.radians(angle)
    sx = sy = 0.0
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / scale for x in m]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty)
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty)
    m[2] += cx
    m[5] += cy
    return m


def fixed_coeffs(m):
    """Geometry.c affine_fixed: the six coefficients in 16.16 fixed point, the half-pixel offset folded into a2 / a5."""
    fix = lambda v: int(math.floor(v * 65536.0 + 0.5))
    return [fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5), fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5)]


def affine_nearest(img, m, fill=0):
    """PIL.Image.transform(size, AFFINE, m, NEAREST, fillcolor=fill) of a uint8 HWC image."""
    h, w = img.shape[:2]
    if m[1] == 0 and m[3] == 0:
        return _scale_affine(img, m, fill)
    a0, a1, a2, a3, a4, a5 = fixed_coeffs(m)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.int64)
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.full_like(img, fill)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def _scale_affine(img, m, fill):
    """Geometry.c ImagingScaleAffine (no rotation / shear terms): per-axis lookup tables, COORD(v) = v < 0 ? -1 : (int)v."""
    h, w = img.shape[:2]
    xo = m[2] + m[0] * 0.5
    yo = m[5] + m[4] * 0.5
    xin, yin = [], []
    for _ in range(w):                             # the C loop ACCUMULATES the coordinate (xo += a[0]) in double precision
        xin.append(-1 if xo < 0.0 else int(xo))
        xo += m[0]
    for _ in range(h):
        yin.append(-1 if yo < 0.0 else int(yo))
        yo += m[4]
    xin, yin = np.array(xin), np.array(yin)
    out = np.full_like(img, fill)
    okx, oky = (xin >= 0) & (xin < w), (yin >= 0) & (yin < h)
    out[np.ix_(oky, okx)] = img[np.ix_(yin[oky], xin[okx])]
    return out


def rotate_nearest(img, angle):
    """F.rotate(pil, angle) as RandomRotation calls it (NEAREST, expand False, fill 0), PIL's fast paths included."""
    h, w = img.shape[:2]
    a = angle % 360.0
    if a == 0:
        return img.copy()
    if a == 180:
        return img[::-1, ::-1].copy()
    if a in (90, 270) and w == h:
        return np.rot90(img, 1 if a == 90 else 3).copy()
    return affine_nearest(img, rotate_matrix(angle, w, h))


# ------------------------------------------------------------------------------------------------ colour
def to_l(img):
    """PIL convert('L') of an RGB image (Convert.c rgb2l)."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(degenerate, img, factor):
    """PIL.Image.blend(degenerate, image, factor) (Blend.c): float32 arithmetic, truncation, clipping when extrapolating."""
    a = np.float32(factor)
    d, x = degenerate.astype(np.int32), img.astype(np.int32)
    t = d.astype(np.float32) + a * (x - d).astype(np.float32)
    if 0.0 <= factor <= 1.0:
        return t.astype(np.uint8)
    return np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img, f):
    return blend(np.zeros_like(img), img, f)


def adjust_contrast(img, f):
    lum = to_l(img)
    mean = int(lum.astype(np.float64).sum() / lum.size + 0.5)
    return blend(np.full_like(img, mean), img, f)


def adjust_saturation(img, f):
    return blend(np.repeat(to_l(img)[..., None], 3, axis=2), img, f)


# ------------------------------------------------------------------------------------------------ the pipeline
def augment(img224, p):
    """uint8 [224,224,3] (already resized) -> uint8 [224,224,3].  p: dict(hflip, vflip, angle, jitter=None | dict(order=[..],
    brightness, contrast, saturation), affine=None | dict(angle, translate=(tx,ty), scale))."""
    x = img224
    if p.get("hflip"):
        x = x[:, ::-1]
    if p.get("vflip"):
        x = x[::-1]
    x = rotate_nearest(np.ascontiguousarray(x), p.get("angle", 0.0))
    j = p.get("jitter")
    if j:
        for op in j["order"]:                      # ColorJitter applies its enabled ops in a drawn order
            if op == 0:
                x = adjust_brightness(x, j["brightness"])
            elif op == 1:
                x = adjust_contrast(x, j["contrast"])
            elif op == 2:
                x = adjust_saturation(x, j["saturation"])
    a = p.get("affine")
    if a:
        h, w = x.shape[:2]
        x = affine_nearest(x, inverse_affine_matrix(a["angle"], a["translate"], a["scale"], w, h))
    return x
