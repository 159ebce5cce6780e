"""BatchNorm / LayerNorm / elementwise kernel micro-benchmark at ResNet50 / ViT (B=128) sizes: achieved GB/s against the
algorithmic bytes of each pass.  Usage: python tools/gpu_bn_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def report(name, us, nbytes):
    print("%-44s %9.1f us %8.0f GB/s (%4.1f%% of 6556)" % (name, us, nbytes / us / 1e3, nbytes / us / 1e3 / 65.56), flush=True)


for M, C in ((401408, 256), (401408, 64), (100352, 512), (25088, 1024), (6272, 2048)):
    y = torch.randn(M, C, device=dev).to(BF)
    res = torch.randn(M, C, device=dev).to(BF)
    dz = torch.randn(M, C, device=dev).to(BF)
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    n2 = M * C * 2
    tag = "M=%d C=%d" % (M, C)
    report("bn_stats " + tag, timeit(lambda: ops.bn_stats(y)), n2)
    z, mi = ops.bn_forward(y, gamma, beta, rm, rv, nbt, res, True, True, 1e-5, 0.1)
    report("bn fwd(stats+apply+res) " + tag, timeit(lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, res, True, True, 1e-5, 0.1)), 5 * n2)
    report("bn fwd(stats+apply) " + tag, timeit(lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, True, True, 1e-5, 0.1)), 4 * n2)
    report("bn bwd(z, dres) " + tag, timeit(lambda: ops.bn_backward(dz, z, y, mi, gamma, True, True, True)), 8 * n2)
    report("bn bwd(mask from y) " + tag, timeit(lambda: ops.bn_backward(dz, None, y, mi, gamma, True, True, False, beta=beta)), 5 * n2)
    del y, res, dz, z

rows, D = 25216, 768
x = torch.randn(rows, D, device=dev).to(BF)
dy = torch.randn(rows, D, device=dev).to(BF)
g, b = torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev)
yy, stats = ops.layernorm(x, g, b, 1e-6)
n2 = rows * D * 2
report("layernorm fwd", timeit(lambda: ops.layernorm(x, g, b, 1e-6)), 2 * n2)
report("layernorm bwd", timeit(lambda: ops.layernorm_bwd(dy, x, g, stats)), 3 * n2)
report("layernorm bwd + skip", timeit(lambda: ops.layernorm_bwd(dy, x, g, stats, dy)), 4 * n2)
h = torch.randn(rows, 3072, device=dev).to(BF)
dh = torch.randn(rows, 3072, device=dev).to(BF)
report("gelu_bwd 25216x3072", timeit(lambda: ops.gelu_bwd(dh, h)), 3 * rows * 3072 * 2)
report("colsum 25216x3072", timeit(lambda: ops.colsum(dh)), rows * 3072 * 2)
report("colsum 25216x768", timeit(lambda: ops.colsum(dy)), n2)
report("add 401408x256", timeit(lambda: ops.add(torch.empty(0) if False else dh, dh)), 3 * rows * 3072 * 2)
