"""fp32 validation mode (north_star: "logits within 2e-2 abs (bf16) or 1e-4 (fp32-accum validation mode)").

`forward_fp32(model, ...)` runs the SAME module tree -- same parameters, same state_dict -- through the dfu_v32_* kernels
of libdfu_b200.so: fp32 activations, the fp32 master weights (not their bf16 shadows), fp32 accumulation.  It is the
instrument that separates "bf16 rounds differently" from "the algorithm differs from the reference": with it the logits
of torchvision's ResNet50 + timm's ViT-B/16 + the fusion head (notebooks/extended_metrics.py:338-367) are reproduced to
1e-4 on BatchNorm-gamma-1 weights, where bf16 storage alone moves them by up to 2e-2.

Inference semantics of the module's current mode: eval() uses BatchNorm running statistics, train() uses batch statistics
(no running-stat update, no dropout: a validation pass has no side effects).  Everything runs on our own kernels; there is
no ATen fallback and no CPU path."""
import ctypes

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import V32GemmArgs, check
from .fusion import MultimodalFusionModel, RGBOnlyModel, ThermalOnlyModel
from .nn import DfuDropout, DfuLinear, DfuReLU
from .resnet import DfuResNet50
from .vit import DfuVisionTransformer

F32 = torch.float32

# forward_fp32(..., emulate_bf16_storage=True): every activation an op of the fast path stores as bf16, and every GEMM weight
# (the bf16 shadows), is rounded to bf16 precision in place -- fp32 arithmetic, bf16 storage.  The distance of THAT result from
# the fp32 reference is what bf16 storage alone costs on the given weights; the fast path must sit next to it.
_EMULATE = [False]


def _q(t):
    """Round an fp32 activation to bf16 precision in place when emulating bf16 storage."""
    if _EMULATE[0]:
        check(_L().dfu_v32_round_bf16(_p(t), t.numel(), _st()), "v32_round_bf16")
        ops._count()
    return t


def _qw(w):
    """A weight as the fast path sees it: its bf16 shadow (a rounded copy) when emulating, the fp32 master otherwise."""
    if not _EMULATE[0]:
        return w.detach()
    return _q(w.detach().clone())


def _L():
    return _lib.load()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t, off=0):
    return None if t is None else t.data_ptr() + off * t.element_size()


def _new(shape, like):
    return torch.empty(shape, dtype=F32, device=like.device)


def gemm(A, B, D, M, N, K, a_strides, b_strides, ldd, bias=None, res=None, ldr=0, alpha=1.0, act=0, nb0=1, nb1=1,
         sa=(0, 0), sb=(0, 0), sd=(0, 0), sr=(0, 0), a_off=0, b_off=0, d_off=0):
    a = V32GemmArgs()
    a.M, a.N, a.K, a.nb0, a.nb1 = M, N, K, nb0, nb1
    a.A, a.a_row_stride, a.a_col_stride, a.sa0, a.sa1 = _p(A, a_off), a_strides[0], a_strides[1], sa[0], sa[1]
    a.B, a.b_row_stride, a.b_col_stride, a.sb0, a.sb1 = _p(B, b_off), b_strides[0], b_strides[1], sb[0], sb[1]
    a.D, a.ldd, a.sd0, a.sd1 = _p(D, d_off), ldd, sd[0], sd[1]
    a.bias, a.res, a.ldr, a.sr0, a.sr1 = _p(bias), _p(res), ldr, sr[0], sr[1]
    a.alpha, a.act = alpha, act
    check(_L().dfu_v32_gemm(ctypes.byref(a), _st()), "dfu_v32_gemm")
    ops._count()
    return D


def linear(x2d, weight, bias=None, act=0, res=None):
    """x2d [M,K] fp32 contiguous; weight: the fp32 parameter viewed as [N, K] in storage order."""
    M, K = x2d.shape
    N = weight.shape[0]
    assert weight.numel() == N * K
    y = _new((M, N), x2d)
    return _q(gemm(x2d, _qw(weight), y, M, N, K, (K, 1), (K, 1), N, bias=None if bias is None else bias.detach(), res=res,
                   ldr=N, act=act))


# ------------------------------------------------------------------------------------------------- inputs
def _image_nhwc(x, kind):
    """uint8 NHWC or fp32 NCHW (already normalised, the reference contract) -> fp32 NHWC."""
    if not x.is_cuda:
        raise RuntimeError("validation mode runs on the GPU kernels (no CPU path); got %s" % x.device)
    if x.dtype == torch.uint8:
        B, H, W, _ = x.shape
        out = _new((B, H, W, 3), x)
        mean, std = ops._norm_consts(kind)
        check(_L().dfu_v32_preprocess_u8(_p(x.contiguous()), _p(out), B * H * W, mean, std, _st()), "v32_preprocess_u8")
        ops._count()
        return _q(out)
    x = x.float().contiguous()
    B, C, H, W = x.shape
    out = _new((B, H, W, C), x)
    check(_L().dfu_v32_nchw_to_nhwc(_p(x), _p(out), B, C, H, W, _st()), "v32_nchw_to_nhwc")
    ops._count()
    return _q(out)


# ------------------------------------------------------------------------------------------------- ResNet50
def _conv(x, conv):
    """x fp32 NHWC [B,H,W,C]; conv.weight is stored channels_last == [Cout, R*S*Cin]."""
    B, H, W, C = x.shape
    w = conv.weight
    Cout, Cin, R, S = w.shape
    stride, pad = conv.stride[0], conv.padding[0]
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    if not w.permute(0, 2, 3, 1).is_contiguous():
        raise RuntimeError("conv weights must be stored channels_last")
    if R == 1 and S == 1 and stride == 1 and pad == 0:
        col = x.view(B * H * W, C)
    else:
        col = _new((B * Ho * Wo, R * S * C), x)
        check(_L().dfu_v32_im2col(_p(x), _p(col), B, H, W, C, R, S, stride, pad, Ho, Wo, _st()), "v32_im2col")
        ops._count()
    K = R * S * C
    y = _new((B * Ho * Wo, Cout), x)
    gemm(col, _qw(w), y, B * Ho * Wo, Cout, K, (K, 1), (K, 1), Cout)
    return _q(y).view(B, Ho, Wo, Cout)


def _bn(y, bn, relu, res=None):
    B, H, W, C = y.shape
    M = B * H * W
    if bn.training:
        mean, var = _new(C, y), _new(C, y)
        check(_L().dfu_v32_bn_batch_stats(_p(y), _p(mean), _p(var), M, C, _st()), "v32_bn_batch_stats")
        ops._count()
    else:
        mean, var = bn.running_mean, bn.running_var
    out = torch.empty_like(y)
    check(_L().dfu_v32_bn_apply(_p(y), _p(mean), _p(var), _p(bn.weight.detach()), _p(bn.bias.detach()), _p(res), _p(out), M, C,
                                1 if relu else 0, bn.eps, _st()), "v32_bn_apply")
    ops._count()
    return _q(out)


def _bottleneck(x, blk):
    out = _bn(_conv(x, blk.conv1), blk.bn1, True)
    out = _bn(_conv(out, blk.conv2), blk.bn2, True)
    identity = x if blk.downsample is None else _bn(_conv(x, blk.downsample[0]), blk.downsample[1], False)
    return _bn(_conv(out, blk.conv3), blk.bn3, True, res=identity)


def _head(f, head):
    """nn.Identity, DfuLinear, or a DfuHead of Linear / ReLU / Dropout (dropout is the identity in a validation pass)."""
    if isinstance(head, nn.Identity):
        return f
    if isinstance(head, DfuLinear):
        return linear(f, head.weight, head.bias)
    mods = list(head)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, DfuLinear):
            fuse_relu = i + 1 < len(mods) and isinstance(mods[i + 1], DfuReLU)
            f = linear(f, m.weight, m.bias, act=1 if fuse_relu else 0)
            i += 2 if fuse_relu else 1
        elif isinstance(m, (DfuDropout, nn.Identity)):
            i += 1
        else:
            raise TypeError("validation mode: unsupported head module %r" % (m,))
    return f


def resnet_fp32(resnet, x, kind="imagenet"):
    """DfuResNet50 forward in fp32: (B,2048) features, or fc's output."""
    assert isinstance(resnet, DfuResNet50)
    x = _image_nhwc(x, kind)
    x = _bn(_conv(x, resnet.conv1), resnet.bn1, True)
    B, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = _new((B, Ho, Wo, C), x)
    check(_L().dfu_v32_maxpool3x3s2(_p(x), _p(y), B, H, W, C, _st()), "v32_maxpool")
    ops._count()
    x = y                     # (a max of bf16 values is a bf16 value)
    for layer in (resnet.layer1, resnet.layer2, resnet.layer3, resnet.layer4):
        for blk in layer:
            x = _bottleneck(x, blk)
    B, H, W, C = x.shape
    f = _new((B, C), x)
    check(_L().dfu_v32_avgpool(_p(x), _p(f), B, H * W, C, _st()), "v32_avgpool")
    ops._count()
    return _head(_q(f), resnet.fc)


# ------------------------------------------------------------------------------------------------- ViT-B/16
def _layernorm(x2d, ln):
    rows, D = x2d.shape
    y = torch.empty_like(x2d)
    check(_L().dfu_v32_layernorm(_p(x2d), _p(ln.weight.detach()), _p(ln.bias.detach()), _p(y), rows, D, ln.eps, _st()), "v32_layernorm")
    ops._count()
    return _q(y)


def _attention(x2d, attn, B, T, res):
    H, hd = attn.num_heads, attn.head_dim
    D = H * hd
    qkv = linear(x2d, attn.qkv.weight, attn.qkv.bias)                      # [B*T, 3*D], columns [3][H][64]
    D3 = 3 * D
    P = _new((B, H, T, T), x2d)
    # S = scale * q k^T per (head, batch)
    gemm(qkv, qkv, P, T, T, hd, (D3, 1), (D3, 1), T, alpha=attn.scale, nb0=H, nb1=B, sa=(hd, T * D3), sb=(hd, T * D3),
         sd=(T * T, H * T * T), b_off=D)
    check(_L().dfu_v32_softmax_rows(_p(P), B * H * T, T, T, _st()), "v32_softmax_rows")
    ops._count()
    _q(P)                     # the fused attention kernel feeds P to the tensor cores as bf16
    out = _new((B * T, D), x2d)
    # O = P v: B operand v[t, d] read as B[n=d][k=t]
    gemm(P, qkv, out, T, hd, T, (T, 1), (1, D3), D, nb0=H, nb1=B, sa=(T * T, H * T * T), sb=(hd, T * D3), sd=(hd, T * D),
         b_off=2 * D)
    return linear(_q(out), attn.proj.weight, attn.proj.bias, res=res)


def vit_fp32(vit, x, kind="half"):
    assert isinstance(vit, DfuVisionTransformer)
    x = _image_nhwc(x, kind)
    B = x.shape[0]
    if tuple(x.shape[1:]) != (224, 224, 3):
        raise RuntimeError("vit_base_patch16_224 takes 224x224x3 images, got %s" % (tuple(x.shape),))
    patches = _new((B * 196, 768), x)
    check(_L().dfu_v32_patchify(_p(x), _p(patches), B, _st()), "v32_patchify")
    ops._count()                # (a permutation of already rounded values)
    proj = vit.patch_embed.proj
    po = linear(patches, proj.weight, proj.bias)
    T, D = vit.pos_embed.shape[1], vit.pos_embed.shape[2]
    tok = _new((B * T, D), x)
    check(_L().dfu_v32_vit_assemble(_p(po), _p(vit.cls_token.detach()), _p(vit.pos_embed.detach()), _p(tok), B, T, D, _st()),
          "v32_vit_assemble")
    ops._count()
    _q(tok)
    for blk in vit.blocks:
        tok = _attention(_layernorm(tok, blk.norm1), blk.attn, B, T, res=tok)
        h = linear(_layernorm(tok, blk.norm2), blk.mlp.fc1.weight, blk.mlp.fc1.bias, act=2)
        tok = linear(h, blk.mlp.fc2.weight, blk.mlp.fc2.bias, res=tok)
    tok = _layernorm(tok, vit.norm)
    f = _new((B, D), x)
    check(_L().dfu_v32_copy2d(_p(tok), T * D, _p(f), D, B, D, _st()), "v32_copy2d")
    ops._count()
    return _head(f, vit.head)


# ------------------------------------------------------------------------------------------------- models
@torch.no_grad()
def forward_fp32(model, *inputs, emulate_bf16_storage=False):
    """Logits (fp32) of MultimodalFusionModel(rgb, thermal), RGBOnlyModel(rgb), ThermalOnlyModel(thermal), a bare
    DfuResNet50 or DfuVisionTransformer, computed in fp32 validation mode.  Inputs: uint8 NHWC or fp32 NCHW normalised.
    emulate_bf16_storage=True keeps the fp32 arithmetic but rounds every stored activation and GEMM weight to bf16."""
    _EMULATE[0] = bool(emulate_bf16_storage)
    try:
        return _forward_fp32(model, *inputs)
    finally:
        _EMULATE[0] = False


def _forward_fp32(model, *inputs):
    if isinstance(model, MultimodalFusionModel):
        rgb, thermal = inputs
        rf, tf = resnet_fp32(model.resnet, rgb), vit_fp32(model.vit, thermal)
        B = rf.shape[0]
        fused = _new((B, rf.shape[1] + tf.shape[1]), rf)
        check(_L().dfu_v32_copy2d(_p(rf), rf.shape[1], _p(fused), fused.shape[1], B, rf.shape[1], _st()), "v32_copy2d")
        check(_L().dfu_v32_copy2d(_p(tf), tf.shape[1], _p(fused, rf.shape[1]), fused.shape[1], B, tf.shape[1], _st()), "v32_copy2d")
        ops._count(2)
        return _head(fused, model.fusion.classifier)
    if isinstance(model, RGBOnlyModel):
        return resnet_fp32(model.resnet, inputs[0])
    if isinstance(model, ThermalOnlyModel):
        return vit_fp32(model.vit, inputs[0])
    if isinstance(model, DfuResNet50):
        return resnet_fp32(model, inputs[0])
    if isinstance(model, DfuVisionTransformer):
        return vit_fp32(model, inputs[0])
    raise TypeError("validation mode knows the late-fusion model and its two branches, not %r" % type(model).__name__)
