"""Leaf nn.Modules of the hot path.  They keep the parameter/buffer names (and therefore state_dict keys) of the
torchvision / timm modules the reference builds, hold ordinary fp32 nn.Parameters, and run
libdfu_b200 kernels in forward/backward.  Activations between modules are bf16: images as logical (B,C,H,W)
tensors in NHWC memory, tokens as (B,T,D)."""
import math

import torch
import torch.nn as nn

from . import functional as Fn
from .ops import BF16, F32


def has_hooks(*mods):
    """True when any of the modules (or the global registry) carries hooks; fused fast paths then step aside so
    every submodule is really called (Grad-CAM registers forward hooks on each `layer4*` / `blocks*` module,
    grad_cam_visualization.py:339-357)."""
    g = nn.modules.module
    if g._global_forward_hooks or g._global_forward_pre_hooks or g._global_backward_hooks:
        return True
    for m in mods:
        if m is None:
            continue
        for sub in m.modules():
            if sub._forward_hooks or sub._forward_pre_hooks or sub._backward_hooks or sub._backward_pre_hooks:
                return True
    return False


def to_act(x):
    """Module-boundary adapter: fp32 tensors coming from outside (images, foreign modules) become bf16."""
    if x.dtype == BF16:
        return x
    if x.dim() == 4:
        return Fn.ToNHWC.apply(x.float() if x.dtype != F32 else x)
    return Fn.CastToBF16.apply(x.float() if x.dtype != F32 else x)


class DfuConv2d(nn.Module):
    """nn.Conv2d(bias=False) replacement; weight is (Cout,Cin,R,S) stored channels_last == GEMM B operand."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = (kernel_size, kernel_size), (stride, stride), (padding, padding)
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size).contiguous(memory_format=torch.channels_last)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w)
        self.bias = None

    def forward(self, x):
        return Fn.Conv2dFn.apply(to_act(x), self.weight, self.stride[0], self.padding[0])

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%s, padding=%s, bias=False" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # copy_ keeps our channels_last storage; nothing else to do


class DfuBatchNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._fold_epoch = 0

    def train(self, mode=True):
        # the train-mode kernels update the running statistics through raw pointers: whatever was folded from them
        # (resnet._folded) is stale once this module has been in train() again
        if mode:
            self._fold_epoch += 1
        return super().train(mode)

    def forward(self, x, res=None, relu=False):
        return Fn.BatchNormActFn.apply(to_act(x), self.weight, self.bias, res, self.running_mean, self.running_var,
                                       self.num_batches_tracked, relu, self.training, self.eps, self.momentum)

    def extra_repr(self):
        return "%d, eps=%g, momentum=%g" % (self.num_features, self.eps, self.momentum)


class DfuReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        return Fn.ReLUFn.apply(to_act(x))


class DfuGELU(nn.Module):
    def forward(self, x):
        return Fn.GeluFn.apply(to_act(x))


class DfuMaxPool2d(nn.Module):
    """MaxPool2d(kernel_size=3, stride=2, padding=1) -- the only pooling window ResNet50 uses."""

    def __init__(self):
        super().__init__()
        self.kernel_size, self.stride, self.padding = 3, 2, 1

    def forward(self, x):
        return Fn.MaxPoolFn.apply(to_act(x))


class DfuAvgPoolFlatten(nn.Module):
    """AdaptiveAvgPool2d((1,1)) followed by flatten(1): (B,C,H,W) -> (B,C)."""

    def forward(self, x):
        return Fn.AvgPoolFn.apply(to_act(x))


class DfuLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True, std=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        if std is None:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            if bias:
                bound = 1 / math.sqrt(in_features)
                nn.init.uniform_(self.bias, -bound, bound)
        else:
            nn.init.trunc_normal_(self.weight, std=std)

    def forward(self, x, res=None, act=0, out_fp32=False):
        return Fn.LinearFn.apply(to_act(x), self.weight, self.bias, res, act, out_fp32)

    def extra_repr(self):
        return "in_features=%d, out_features=%d, bias=%s" % (self.in_features, self.out_features, self.bias is not None)


class DfuLayerNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.normalized_shape, self.eps = (dim,), eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        return Fn.LayerNormFn.apply(to_act(x), self.weight, self.bias, self.eps)


_SITE = [0]


class DfuDropout(nn.Module):
    def __init__(self, p=0.5):
        super().__init__()
        self.p = float(p)
        _SITE[0] += 1
        self.site = _SITE[0]

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        return Fn.DropoutFn.apply(to_act(x), self.p, self.site)

    def extra_repr(self):
        return "p=%g" % self.p


def convert_head(mod):
    """Maps the stock torch modules the reference assigns as classifier heads -- nn.Identity,
    nn.Sequential(nn.Dropout, nn.Linear) (train_rgb_only.py:211-214, train_thermal_only.py:199-202), nn.Linear --
    onto our modules with identical state_dict keys.  Our own modules pass through."""
    if isinstance(mod, (DfuLinear, DfuDropout, DfuReLU)) or mod is None:
        return mod
    if isinstance(mod, nn.Identity):
        return mod
    if isinstance(mod, nn.Linear):
        new = DfuLinear(mod.in_features, mod.out_features, bias=mod.bias is not None)
        with torch.no_grad():
            new.weight.copy_(mod.weight)
            if mod.bias is not None:
                new.bias.copy_(mod.bias)
        return new.to(mod.weight.device)
    if isinstance(mod, nn.Dropout):
        return DfuDropout(mod.p)
    if isinstance(mod, nn.ReLU):
        return DfuReLU()
    if isinstance(mod, nn.Sequential):
        return DfuHead(*[convert_head(m) for m in mod])
    raise TypeError("unsupported head module for the B200 path: %r" % (mod,))


class DfuHead(nn.Sequential):
    """nn.Sequential of Linear / ReLU / Dropout with Linear+ReLU fused into one GEMM epilogue and fp32 logits from
    the last Linear.  Child names are the Sequential indices, so keys stay `...0.weight`, `...3.weight`, ..."""

    def forward(self, x):
        mods = list(self)
        fused = not has_hooks(self)
        i = 0
        x = to_act(x)
        last_linear = max((k for k, m in enumerate(mods) if isinstance(m, DfuLinear)), default=-1)
        while i < len(mods):
            m = mods[i]
            if isinstance(m, DfuLinear):
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if fused and isinstance(nxt, DfuReLU):
                    x = m(x, act=1)
                    i += 2
                    continue
                x = m(x, out_fp32=(i == last_linear and i == len(mods) - 1))
            else:
                x = m(x)
            i += 1
        return x
