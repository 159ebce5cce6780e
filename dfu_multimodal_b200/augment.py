"""Train-time augmentation on the GPU: the reference's PIL pipeline
(notebooks/train_multimodal_fusion.py:173-199 for RGB -- flips, RandomRotation(30), ColorJitter(.3,.3,.3) with p = AUG_PROB,
RandomAffine(20, translate .1, scale .8-1.2) with p = AUG_PROB --, :190-199 / train_thermal_only.py:103-112 for thermal, the
same without ColorJitter) as ONE gather kernel per batch fused with ToTensor + Normalize.

The host draws the random parameters (same distributions as torchvision's get_params) and turns each draw into the integer
inverse maps Pillow itself would build (Image.rotate / transform(AFFINE) -> Geometry.c affine_fixed: 16.16 fixed point);
`dfu_augment_u8` then produces, for those parameters, exactly the uint8 image torchvision produces, already normalised.
What the reference does in 4 DataLoader worker processes at ~100 images/s each runs here at HBM speed on the device."""
import ctypes
import math
import random

import torch

from . import _lib, ops
from ._lib import AugmentParams, check

AUG_PROB = 0.6            # train_multimodal_fusion.py:41


# ------------------------------------------------------------------------------------------------ inverse maps (host)
def _fixed(m):
    """Pillow Geometry.c affine_fixed coefficients (a0, a1, a2', a3, a4, a5') of the inverse map m."""
    fix = lambda v: int(math.floor(v * 65536.0 + 0.5))
    return (fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5), fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def _rotation_map(angle, w, h):
    """(kind, coeffs) of PIL.Image.rotate(angle) as torchvision's F.rotate calls it (NEAREST, expand=False, centre = w/2, h/2)."""
    a = angle % 360.0
    if a == 0:
        return 0, (0,) * 6
    if a == 180:
        return 2, (0,) * 6
    if a in (90, 270) and w == h:
        return (3 if a == 90 else 4), (0,) * 6
    r = -math.radians(a)
    m = [round(math.cos(r), 15), round(math.sin(r), 15), 0.0, round(-math.sin(r), 15), round(math.cos(r), 15), 0.0]
    cx, cy = w / 2.0, h / 2.0
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return 1, _fixed(m)


def _affine_map(angle, translate, scale, w, h):
    """(kind, coeffs, xtab, ytab) of F.affine(img, angle, translate, scale, shear=(0,0)) on a PIL image."""
    cx, cy = w * 0.5, h * 0.5
    rot = math.radians(angle)
    tx, ty = translate
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)      # shear 0: torchvision's a, b, c, d
    m = [d / scale, -b / scale, 0.0, -c / scale, a / scale, 0.0]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty)
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty)
    m[2] += cx
    m[5] += cy
    if m[1] == 0 and m[3] == 0:
        # Pillow's ImagingScaleAffine: per-axis tables, the source coordinate ACCUMULATED in double precision
        xo, yo = m[2] + m[0] * 0.5, m[5] + m[4] * 0.5
        xtab, ytab = [], []
        for _ in range(w):
            xtab.append(-1 if xo < 0.0 else min(int(xo), 32767))
            xo += m[0]
        for _ in range(h):
            ytab.append(-1 if yo < 0.0 else min(int(yo), 32767))
            yo += m[4]
        return 5, (0,) * 6, xtab, ytab
    return 1, _fixed(m), None, None


# ------------------------------------------------------------------------------------------------ parameters
def draw_params(n, modality="rgb", size=(224, 224), rng=None, aug_prob=AUG_PROB):
    """n parameter dicts with the reference's distributions (torchvision get_params): flips p = 0.5 each, rotation
    U(-30, 30), ColorJitter (RGB only, p = aug_prob): factors U(0.7, 1.3) in a random order, RandomAffine (p = aug_prob):
    angle U(-20, 20), translation round(U(-0.1 W, 0.1 W)) pixels, scale U(0.8, 1.2)."""
    rng = rng or random
    h, w = size
    out = []
    for _ in range(n):
        p = {"hflip": rng.random() < 0.5, "vflip": rng.random() < 0.5, "angle": rng.uniform(-30.0, 30.0)}
        if modality == "rgb" and rng.random() < aug_prob:
            order = [0, 1, 2]
            rng.shuffle(order)
            p["jitter"] = {"order": order, "brightness": rng.uniform(0.7, 1.3), "contrast": rng.uniform(0.7, 1.3),
                           "saturation": rng.uniform(0.7, 1.3)}
        if rng.random() < aug_prob:
            p["affine"] = {"angle": rng.uniform(-20.0, 20.0), "scale": rng.uniform(0.8, 1.2),
                           "translate": (int(round(rng.uniform(-0.1 * w, 0.1 * w))), int(round(rng.uniform(-0.1 * h, 0.1 * h))))}
        out.append(p)
    return out


def pack_params(params, size=(224, 224)):
    """Parameter dicts -> pinned uint8 tensor holding the dfu_augment_params array the kernel reads."""
    h, w = size
    arr = (AugmentParams * len(params))()
    for rec, p in zip(arr, params):
        rec.flags = (1 if p.get("hflip") else 0) | (2 if p.get("vflip") else 0)
        rec.rot_kind, rot = _rotation_map(p.get("angle", 0.0), w, h)
        rec.rot[:] = rot
        aff = p.get("affine")
        if aff:
            rec.aff_kind, coeffs, xtab, ytab = _affine_map(aff["angle"], aff["translate"], aff["scale"], w, h)
            rec.aff[:] = coeffs
            if xtab is not None:
                rec.aff_xtab[:w] = xtab
                rec.aff_ytab[:h] = ytab
        j = p.get("jitter")
        if j:
            names = ("brightness", "contrast", "saturation")
            rec.n_color = len(j["order"])
            for k, op in enumerate(j["order"]):
                rec.color_op[k] = op
                rec.color_f[k] = j[names[op]]
    packed = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return packed.pin_memory() if torch.cuda.is_available() else packed


def augment_u8(img_u8, params, mode="nhwc", kind="imagenet"):
    """img_u8 [B,H,W,3] uint8 on the device (H, W <= 224, e.g. the output of ops.resize_u8(..., 'u8')), params: list of B dicts
    (draw_params) or an already packed tensor.  mode 'u8' -> augmented uint8 images; 'nhwc' -> + ToTensor + Normalize(kind) as
    bf16 (logical (B,3,H,W), NHWC memory); 'patches' -> + Normalize + ViT patch unfold, bf16 [B*196, 768]."""
    if not img_u8.is_cuda:
        raise RuntimeError("augment_u8 runs on the GPU (no CPU path); got %s" % img_u8.device)
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[-1] == 3 and img_u8.is_contiguous()
    B, H, W, _ = img_u8.shape
    packed = params if torch.is_tensor(params) else pack_params(params, (H, W))
    if packed.numel() != B * ctypes.sizeof(AugmentParams):
        raise RuntimeError("augment_u8: %d parameter records for a batch of %d" % (packed.numel() // ctypes.sizeof(AugmentParams), B))
    dparams = packed.to(img_u8.device, non_blocking=True)
    sums = torch.empty(B, dtype=torch.int64, device=img_u8.device)
    code = {"u8": 0, "nhwc": 1, "patches": 2}[mode]
    if code == 0:
        out, mean, std = torch.empty_like(img_u8), None, None
    else:
        out = (torch.empty((B, H, W, 3), dtype=torch.bfloat16, device=img_u8.device) if code == 1 else
               torch.empty((B * 196, 768), dtype=torch.bfloat16, device=img_u8.device))
        mean, std = ops._norm_consts(kind)
    check(_lib.load().dfu_augment_u8(img_u8.data_ptr(), B, H, W, dparams.data_ptr(), sums.data_ptr(), code, out.data_ptr(), mean, std,
                                     torch.cuda.current_stream().cuda_stream), "augment_u8")
    ops._count(2)
    return out.permute(0, 3, 1, 2) if code == 1 else out
