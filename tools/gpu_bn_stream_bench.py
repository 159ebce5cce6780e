"""BatchNorm apply / backward kernels at the layer1-layer4 sizes, 50 iterations each: one line per shape.  The kernel variant is
chosen by the environment the library reads at load time (DFU_BN_STREAM, DFU_STREAM_INTERLEAVED, DFU_STREAM_PER_SM)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


tag = " ".join("%s=%s" % (k, os.environ[k]) for k in ("DFU_BN_STREAM", "DFU_STREAM_INTERLEAVED", "DFU_STREAM_PER_SM") if k in os.environ) or "default"
row = []
for M, C in ((401408, 256), (401408, 64), (100352, 512), (25088, 1024), (6272, 2048)):
    y = torch.randn(M, C, device=dev).to(BF)
    res = torch.randn(M, C, device=dev).to(BF)
    dz = torch.randn(M, C, device=dev).to(BF)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    sums = ops.bn_stats(y)
    a = timeit(lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, True, True, 1e-5, 0.1, sums=sums))
    b = timeit(lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, res, True, True, 1e-5, 0.1, sums=sums, want_mask=True))
    z, mi = ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, True, True, 1e-5, 0.1, sums=sums)
    c = timeit(lambda: ops.bn_backward(dz, None, y, mi, gamma, True, True, False, beta=beta))
    n2 = M * C * 2
    row.append("C=%-4d apply %6.1f us (%4.0f GB/s) +res %6.1f (%4.0f) bwd %6.1f (%4.0f)" % (C, a, 2 * n2 / a / 1e3, b, (3 * n2 + n2 // 16) / b / 1e3, c, 5 * n2 / c / 1e3))
    del y, res, dz, z
print("%-60s\n  %s" % (tag, "\n  ".join(row)), flush=True)
