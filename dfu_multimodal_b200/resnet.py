"""ResNet50 (torchvision v1.5 layout: stride on the 3x3 conv, Bottleneck expansion 4) on libdfu_b200 kernels.
Module/parameter names equal torchvision's so `resnet.*` / `rgb_branch.*` / `backbone.*` checkpoints load
(SURVEY.md 8b).  Reference constructor: torch.hub.load('pytorch/vision:v0.13.1','resnet50')
(notebooks/train_multimodal_fusion.py:294)."""
import os

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .nn import (DfuAvgPoolFlatten, DfuBatchNorm2d, DfuConv2d, DfuLinear, DfuMaxPool2d, DfuReLU, convert_head,
                 has_hooks, to_act)


def _bn(y, bn, relu, res=None, sums=None):
    return Fn.BatchNormActFn.apply(y, bn.weight, bn.bias, res, bn.running_mean, bn.running_var,
                                   bn.num_batches_tracked, relu, bn.training, bn.eps, bn.momentum, sums)


# eval()-mode inference folds every BatchNorm into the convolution that feeds it (DFU_FOLD_BN=0: separate bn_apply pass)
FOLD_BN = os.environ.get("DFU_FOLD_BN", "1") == "1"


def _folded(conv, bn):
    """(bf16 weights scaled by gamma/sqrt(var+eps), fp32 bias beta - mean*gamma/sqrt(var+eps)) of a conv -> eval-BatchNorm
    pair, cached on the conv and recomputed IN PLACE when anything it was computed from may have changed: torch-visible
    edits (version counters), a train() phase of the BatchNorm, a FusedAdamW step."""
    key = (conv.weight._version, conv.weight.data_ptr(), bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, bn.running_mean.data_ptr(), bn._fold_epoch, ops.PARAM_EPOCH[0])
    cached = conv.__dict__.get("_dfu_folded")
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    if not conv.weight.permute(0, 2, 3, 1).is_contiguous():
        raise RuntimeError("conv weights must be stored channels_last (Cout,R,S,Cin); use DfuConv2d parameters")
    wf, bias = ops.fold_bn_weights(conv.weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                   bn.eps, out=(cached[1], cached[2]) if cached is not None else None)
    conv.__dict__["_dfu_folded"] = (key, wf, bias)
    return wf, bias


def refresh_folded(module):
    """Brings every folded conv/BatchNorm pair under `module` up to date (in place).  train.EvalStep calls it before
    replaying a captured graph, whose kernels read the folded buffers but do not contain the folding itself."""
    if not FOLD_BN:
        return
    for m in module.modules():
        if isinstance(m, DfuBottleneck):
            _folded(m.conv1, m.bn1), _folded(m.conv2, m.bn2), _folded(m.conv3, m.bn3)
            if m.downsample is not None:
                _folded(m.downsample[0], m.downsample[1])
        elif isinstance(m, DfuResNet50):
            _folded(m.conv1, m.bn1)


def _conv_bn_folded(x, conv, bn, relu, res=None):
    """Inference: act(conv(x, W * s) + b [+ res]) in ONE GEMM launch -- scale, shift, residual and ReLU ride in the epilogue
    (torchvision Bottleneck in eval(): conv -> bn -> relu / conv3 -> bn3 -> += identity -> relu)."""
    xn = ops.as_nhwc(x)
    B, H, W, C = xn.shape
    Cout, _, R, S = conv.weight.shape
    stride, pad = conv.stride[0], conv.padding[0]
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    wf, bias = _folded(conv, bn)
    res2 = ops.as_nhwc(res).reshape(B * Ho * Wo, Cout) if res is not None else None
    act = (6 if res2 is not None else 1) if relu else 0
    if R == 1 and S == 1 and stride == 1 and pad == 0:
        y, _ = ops.linear_forward(xn.reshape(B * H * W, C), wf, bias, act=act, res=res2)
    elif ops.conv_implicit_ok(C, Cout, R, S, stride) and res2 is None:
        y = ops.conv_forward(xn, wf, R, S, pad, stride=stride, bias=bias, act=act)
    else:                                    # the 3-channel stem: explicit patch matrix
        y, _ = ops.linear_forward(ops.im2col(xn, R, S, stride, pad, Ho, Wo, wf.shape[1]), wf, bias, act=act, res=res2)
    return y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)


def _conv_bn(x, conv, bn, relu, res=None):
    """conv -> BatchNorm (+res) (+ReLU).  train(): the batch statistics come out of the conv GEMM's epilogue; eval() without
    autograd: BatchNorm is folded into the convolution (one launch); eval() under autograd (Grad-CAM): separate kernels."""
    if bn.training:
        y, sums = Fn.Conv2dFn.apply(x, conv.weight, conv.stride[0], conv.padding[0], False, True)
        return _bn(y, bn, relu, res, sums)
    if FOLD_BN and not torch.is_grad_enabled():
        return _conv_bn_folded(x, conv, bn, relu, res)
    return _bn(Fn.Conv2dFn.apply(x, conv.weight, conv.stride[0], conv.padding[0]), bn, relu, res)


class DfuBottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = DfuConv2d(inplanes, planes, 1)
        self.bn1 = DfuBatchNorm2d(planes)
        self.conv2 = DfuConv2d(planes, planes, 3, stride=stride, padding=1)
        self.bn2 = DfuBatchNorm2d(planes)
        self.conv3 = DfuConv2d(planes, planes * 4, 1)
        self.bn3 = DfuBatchNorm2d(planes * 4)
        self.relu = DfuReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        x = to_act(x)
        if has_hooks(self):
            return self._forward_modules(x)
        if torch.is_grad_enabled() and x.requires_grad:
            # conv1 and the skip connection fork here; the skip gradient is added in conv1's dgrad GEMM epilogue
            if self.bn1.training:
                y1, xs, sums = Fn.Conv2dFn.apply(x, self.conv1.weight, 1, 0, True, True)
            else:
                (y1, xs), sums = Fn.Conv2dFn.apply(x, self.conv1.weight, 1, 0, True), None
            out = _bn(y1, self.bn1, True, None, sums)
        else:
            xs = x
            out = _conv_bn(x, self.conv1, self.bn1, True)
        out = _conv_bn(out, self.conv2, self.bn2, True)
        identity = xs if self.downsample is None else _conv_bn(xs, self.downsample[0], self.downsample[1], False)
        return _conv_bn(out, self.conv3, self.bn3, True, res=identity)

    def _forward_modules(self, x):
        """Unfused path: every submodule is called (and hooked) exactly as in torchvision's Bottleneck.forward."""
        needs_fork = torch.is_grad_enabled() and x.requires_grad
        xm, xs = Fn.Fork.apply(x) if needs_fork else (x, x)
        out = self.relu(self.bn1(self.conv1(xm)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        identity = xs if self.downsample is None else self.downsample(xs)
        out = Fn.AddFn.apply(out, identity)
        return self.relu(out)


class DfuResNet50(nn.Module):
    def __init__(self, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = DfuConv2d(3, 64, 7, stride=2, padding=3)
        self.bn1 = DfuBatchNorm2d(64)
        self.relu = DfuReLU(inplace=True)
        self.maxpool = DfuMaxPool2d()
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 2)
        self.avgpool = DfuAvgPoolFlatten()
        self.fc = DfuLinear(2048, num_classes)

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(DfuConv2d(self.inplanes, planes * 4, 1, stride=stride),
                                       DfuBatchNorm2d(planes * 4))
        layers = [DfuBottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            layers.append(DfuBottleneck(self.inplanes, planes))
        return nn.Sequential(*layers)

    def __setattr__(self, name, value):
        # the reference replaces `.fc` with stock torch modules (nn.Identity / Sequential(Dropout, Linear))
        if name == "fc" and isinstance(value, nn.Module):
            value = convert_head(value)
        super().__setattr__(name, value)

    def forward_features(self, x):
        """(B,3,H,W) fp32 NCHW, or bf16 NHWC-memory activations (e.g. from ops.preprocess_u8_nhwc) -> (B,2048) bf16."""
        x = to_act(x)
        if has_hooks(self.conv1, self.bn1, self.relu):
            x = self.relu(self.bn1(self.conv1(x)))
        else:
            x = _conv_bn(x, self.conv1, self.bn1, True)
        x = self.maxpool(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.layer4(x)
        return self.avgpool(x)

    def forward_bf16(self, x):
        """features -> fc, staying in bf16 (fp32 when fc ends in a Linear producing logits)."""
        f = self.forward_features(x)
        if isinstance(self.fc, nn.Identity):
            return f
        if isinstance(self.fc, DfuLinear):
            return self.fc(f, out_fp32=True)
        return self.fc(f)

    def forward(self, x):
        """nn.Module contract of torchvision's ResNet: fp32 in, fp32 out."""
        y = self.forward_bf16(x)
        return y if y.dtype == torch.float32 else Fn.CastToF32.apply(y)


def resnet50(num_classes=1000, **_):
    return DfuResNet50(num_classes=num_classes)
