"""GEMM micro-benchmark over the shapes of the fusion train step (B=128): CUDA-event timing, TFLOP/s per shape.
Usage: python tools/gpu_gemm_bench.py [filter-substring]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 1638.8


def bench(name, M, N, K, a_major=0, b_major=0, fp32=False, acc=False, bias=False, act=0, res=False, aux=False,
          nb0=1, nb1=1, iters=20, colstats=0):
    nb = nb0 * nb1
    A = torch.randn((nb, K, M) if a_major else (nb, M, K), device=dev).to(torch.bfloat16)
    Bm = torch.randn((nb, K, N) if b_major else (nb, N, K), device=dev).to(torch.bfloat16)
    D = torch.zeros((nb, M, N), dtype=torch.float32 if fp32 else torch.bfloat16, device=dev)
    bias_t = torch.randn(N, device=dev) if bias else None
    res_t = torch.randn((nb, M, N), device=dev).to(torch.bfloat16) if res else None
    aux_t = torch.empty((nb, M, N), dtype=torch.bfloat16, device=dev) if aux else None
    cs_t = torch.zeros(N * max(colstats, 1), device=dev) if colstats else None
    lda = M if a_major else K
    ldb = N if b_major else K

    def run():
        ops.gemm(A, Bm, D, M=M, N=N, K=K, lda=lda, ldb=ldb, ldd=N, a_major=a_major, b_major=b_major, nb0=nb0, nb1=nb1,
                 sa=(A.stride(0), A.stride(0) * nb0), sb=(Bm.stride(0), Bm.stride(0) * nb0),
                 sd=(M * N, M * N * nb0), bias=bias_t, act=act, res=res_t, ldr=N, sr=(M * N, M * N * nb0), aux=aux_t,
                 accumulate=acc, colstats=cs_t, colstats_mode=colstats)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * M * N * K * nb
    byt = (A.numel() + Bm.numel()) * 2 + D.numel() * D.element_size() * (2 if acc else 1) + \
        (res_t.numel() * 2 if res else 0) + (aux_t.numel() * 2 if aux else 0)
    print("%-34s M=%-7d N=%-5d K=%-6d nb=%-5d %9.1f us %7.1f TF/s (%4.1f%%) %7.0f GB/s" % (
        name, M, N, K, nb, us, fl / us / 1e6, fl / us / 1e6 / PEAK * 100, byt / us / 1e3), flush=True)


CASES = [
    ("vit.fc1 fwd gelu+dgelu aux", dict(M=25216, N=3072, K=768, bias=True, act=4, aux=True)),
    ("vit.fc2 dgrad *res", dict(M=25216, N=3072, K=768, b_major=1, res=True, act=5)),
    ("vit.fc2 dgrad *res +colsum", dict(M=25216, N=3072, K=768, b_major=1, res=True, act=5, colstats=1)),
    ("vit.fc2 dgrad +gelu'", dict(M=25216, N=3072, K=768, b_major=1, res=True, act=3)),
    ("vit.fc2 dgrad +gelu' +colsum", dict(M=25216, N=3072, K=768, b_major=1, res=True, act=3, colstats=1)),
    ("vit.fc2 dgrad +colsum", dict(M=25216, N=3072, K=768, b_major=1, colstats=1)),
    ("l1 1x1 64->256 +bnstats", dict(M=401408, N=256, K=64, colstats=2)),
    ("l1 1x1 256->64 +bnstats", dict(M=401408, N=64, K=256, colstats=2)),
    ("l1 3x3 64 (K=576) +bnstats", dict(M=401408, N=64, K=576, colstats=2)),
    ("l3 1x1 1024->256 +bnstats", dict(M=25088, N=256, K=1024, colstats=2)),
    ("l4 1x1 512->2048 +bnstats", dict(M=6272, N=2048, K=512, colstats=2)),
    ("l1 1x1 dgrad 256->64 +res", dict(M=401408, N=256, K=64, b_major=1, res=True)),
    ("l1 1x1 dgrad 256->64", dict(M=401408, N=256, K=64, b_major=1)),
    # ViT forward
    ("vit.qkv fwd", dict(M=25216, N=2304, K=768, bias=True)),
    ("vit.proj fwd +res", dict(M=25216, N=768, K=768, bias=True, res=True)),
    ("vit.fc1 fwd gelu+aux", dict(M=25216, N=3072, K=768, bias=True, act=2, aux=True)),
    ("vit.fc2 fwd +res", dict(M=25216, N=768, K=3072, bias=True, res=True)),
    # ViT dgrad
    ("vit.qkv dgrad", dict(M=25216, N=768, K=2304, b_major=1)),
    ("vit.fc1 dgrad", dict(M=25216, N=768, K=3072, b_major=1)),
    ("vit.fc2 dgrad", dict(M=25216, N=3072, K=768, b_major=1)),
    # ViT wgrad
    ("vit.qkv wgrad", dict(M=2304, N=768, K=25216, a_major=1, b_major=1, fp32=True, acc=True)),
    ("vit.fc1 wgrad", dict(M=3072, N=768, K=25216, a_major=1, b_major=1, fp32=True, acc=True)),
    ("vit.fc2 wgrad", dict(M=768, N=3072, K=25216, a_major=1, b_major=1, fp32=True, acc=True)),
    ("vit.proj wgrad", dict(M=768, N=768, K=25216, a_major=1, b_major=1, fp32=True, acc=True)),
    # attention (batched)
    # ResNet
    ("stem 7x7 (im2col K=152)", dict(M=1605632, N=64, K=152)),
    ("l1 1x1 64->256", dict(M=401408, N=256, K=64)),
    ("l1 1x1 256->64", dict(M=401408, N=64, K=256)),
    ("l1 3x3 64 (K=576)", dict(M=401408, N=64, K=576)),
    ("l2 3x3 128 (K=1152)", dict(M=100352, N=128, K=1152)),
    ("l3 1x1 1024->256", dict(M=25088, N=256, K=1024)),
    ("l3 3x3 256 (K=2304)", dict(M=25088, N=256, K=2304)),
    ("l4 1x1 512->2048", dict(M=6272, N=2048, K=512)),
    ("l4 3x3 512 (K=4608)", dict(M=6272, N=512, K=4608)),
    ("l1 3x3 dgrad", dict(M=401408, N=576, K=64, b_major=1)),
    ("l3 3x3 dgrad", dict(M=25088, N=2304, K=256, b_major=1)),
    ("l1 3x3 wgrad", dict(M=64, N=576, K=401408, a_major=1, b_major=1, fp32=True, acc=True)),
    ("l1 1x1 wgrad 256x64", dict(M=256, N=64, K=401408, a_major=1, b_major=1, fp32=True, acc=True)),
    ("l3 3x3 wgrad", dict(M=256, N=2304, K=25088, a_major=1, b_major=1, fp32=True, acc=True)),
    ("l4 3x3 wgrad", dict(M=512, N=4608, K=6272, a_major=1, b_major=1, fp32=True, acc=True)),
    ("square 8192", dict(M=8192, N=8192, K=8192)),
]

def bench_attention(B=128, T=197, H=12, iters=10):
    qkv = (torch.randn(B * T, 3 * H * 64, device=dev) * 0.5).to(torch.bfloat16)
    do = torch.randn(B * T, H * 64, device=dev).to(torch.bfloat16)
    out, P = ops.attention_forward(qkv, B, T, H, 0.125)
    for name, fn in (("attention fwd (%s)" % ops._ATTN_IMPL, lambda: ops.attention_forward(qkv, B, T, H, 0.125)),
                     ("attention bwd (%s)" % ops._ATTN_IMPL, lambda: ops.attention_backward(qkv, P, do, B, T, H, 0.125, out=out))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        fl = 4.0 * B * H * T * T * 64 * (1 if "fwd" in name else 2)
        print("%-34s B=%d %9.1f us %7.1f TF/s" % (name, B, us, fl / us / 1e6), flush=True)


if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    print(torch.cuda.get_device_name(0))
    for name, kw in CASES:
        if flt in name:
            try:
                bench(name, **kw)
            except Exception as e:  # noqa: BLE001
                print("%-34s FAILED %r" % (name, e))
    if flt in "attention":
        bench_attention()
