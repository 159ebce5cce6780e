"""Compares the BatchNorm batch sums produced by the convolution epilogues with a separate pass over the stored output,
for every ResNet50 convolution shape.  Usage: python tools/gpu_check_bnstats.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import functional as Fn, ops  # noqa: E402
from dfu_multimodal_b200.resnet import DfuResNet50  # noqa: E402
from dfu_multimodal_b200.nn import DfuConv2d  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = DfuResNet50().to(dev)
shapes = {}
hooks = []


def mk(name):
    def hook(m, inp, out):
        shapes[name] = (tuple(inp[0].shape), m)
    return hook


for name, m in net.named_modules():
    if isinstance(m, DfuConv2d):
        hooks.append(m.register_forward_hook(mk(name)))
net.eval()
with torch.no_grad():
    net(torch.randn(B, 3, 224, 224, device=dev))
for h in hooks:
    h.remove()
worst = 0.0
for name, (ishape, m) in shapes.items():
    x = torch.randn(ishape, device=dev)
    xa = Fn.ToNHWC.apply(x)
    with torch.no_grad():
        y, sums = Fn.Conv2dFn.apply(xa, m.weight, m.stride[0], m.padding[0], False, True)
    y2 = ops.as_nhwc(y).reshape(-1, y.shape[1])
    ref = ops.bn_stats(y2)
    rel = ((sums - ref).abs().max() / ref.abs().max()).item()
    worst = max(worst, rel)
    flag = "  <-- MISMATCH" if rel > 1e-3 else ""
    print("%-28s in=%-22s k=%d s=%d Cout=%-5d rel err %.2e%s" % (name, ishape, m.kernel_size[0], m.stride[0], y.shape[1], rel, flag))
print("worst", worst)
