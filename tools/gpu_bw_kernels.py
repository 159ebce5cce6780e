"""Every bandwidth-bound kernel of the train step at its B=128 sizes: a CUDA-event table (python tools/gpu_bw_kernels.py
[iters]) or exactly one launch of each at its largest shape for `ncu --set full` (python tools/gpu_bw_kernels.py ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import ops  # noqa: E402
from dfu_multimodal_b200.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
NCU = len(sys.argv) > 1 and sys.argv[1] == "ncu"          # one launch of each kernel at its largest shape, no warm-up
iters = 1 if NCU else (int(sys.argv[1]) if len(sys.argv) > 1 else 5)
PEAK = 6556.0


def timeit(fn):
    for _ in range(0 if NCU else 2):
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
    print("%-52s %9.1f us %8.0f GB/s (%5.1f%% of %d)" % (name, us, nbytes / us / 1e3, nbytes / us / 1e3 / PEAK * 100, PEAK), flush=True)


for M, C in (((401408, 256),) if NCU else ((401408, 256), (401408, 64), (100352, 512), (25088, 1024), (6272, 2048))):
    y = torch.randn(M, C, device=dev).to(BF)
    res = torch.randn(M, C, device=dev).to(BF)
    dz = torch.randn(M, C, device=dev).to(BF)
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    n2 = M * C * 2
    tag = "M=%d C=%d" % (M, C)
    sums = ops.bn_stats(y)
    report("bn_apply_train (y -> z) " + tag, timeit(lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, True, True, 1e-5, 0.1, sums=sums)), 2 * n2)
    report("bn_apply_train (+res, +mask) " + tag, timeit(lambda: ops.bn_forward(y, gamma, beta, rm, rv, nbt, res, True, True, 1e-5, 0.1, sums=sums, want_mask=True)), 3 * n2 + n2 // 16)
    z, mi = ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, True, True, 1e-5, 0.1, sums=sums)
    report("bn_backward reduce+apply (mask from y) " + tag, timeit(lambda: ops.bn_backward(dz, None, y, mi, gamma, True, True, False, beta=beta)), 5 * n2)
    del y, res, dz, z

rows, D = 25216, 768
x = torch.randn(rows, D, device=dev).to(BF)
dy = torch.randn(rows, D, device=dev).to(BF)
g, b = torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev)
yy, stats = ops.layernorm(x, g, b, 1e-6)
n2 = rows * D * 2
report("layernorm fwd 25216x768", timeit(lambda: ops.layernorm(x, g, b, 1e-6)), 2 * n2)
report("layernorm bwd + skip 25216x768", timeit(lambda: ops.layernorm_bwd(dy, x, g, stats, dy)), 4 * n2)
report("colsum 25216x768", timeit(lambda: ops.colsum(dy)), n2)

img = torch.randint(0, 256, (128, 224, 224, 3), dtype=torch.uint8, device=dev)
report("preprocess_u8_nhwc B=128", timeit(lambda: ops.preprocess_u8_nhwc(img, "imagenet")), 128 * 451584)
report("preprocess_u8_patches B=128", timeit(lambda: ops.preprocess_u8_patches(img, "half")), 128 * 451584)

ps = [torch.nn.Parameter(torch.randn(n, device=dev)) for n in (2304 * 768, 3072 * 768, 768 * 3072, 2048 * 512, 85_000_000, 768, 64)]
opt = FusedAdamW(ps, lr=1e-4, weight_decay=1e-4)
for p in ps:
    p.grad = torch.randn_like(p)
nparam = sum(p.numel() for p in ps)
report("adamw_multi %.1f M params (no bf16 shadows)" % (nparam / 1e6), timeit(opt.step), 28 * nparam)
