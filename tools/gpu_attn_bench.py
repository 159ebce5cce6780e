"""Attention micro-benchmark (B=128, T=197, H=12): fused tcgen05 kernels vs the batched-GEMM + softmax path.
Usage: python tools/gpu_attn_bench.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T, H = 197, 12
dev = torch.device("cuda:0")
qkv = (torch.randn(B * T, 3 * H * 64, device=dev) * 0.5).to(torch.bfloat16)
do = torch.randn(B * T, H * 64, device=dev).to(torch.bfloat16)


def timeit(fn, iters=20):
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


fl_f = 4.0 * T * T * 64 * B * H
for impl in ("fused", "unfused"):
    ops.set_attention_impl(impl)
    out, saved = ops.attention_forward(qkv, B, T, H, 0.125)
    tf = timeit(lambda: ops.attention_forward(qkv, B, T, H, 0.125))
    tb = timeit(lambda: ops.attention_backward(qkv, saved, do, B, T, H, 0.125, out=out))
    print("%-8s fwd %8.1f us (%6.1f TF/s)   bwd %8.1f us (%6.1f TF/s)" % (impl, tf, fl_f / tf / 1e6, tb, 2.5 * fl_f / tb / 1e6),
          flush=True)
