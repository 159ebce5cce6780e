"""Column-statistics epilogue vs a separate pass over the stored output, with the per-column differences expressed in
units of one 32-row slab's contribution.  python tools/gpu_debug_colstats.py M N K"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfu_multimodal_b200 import ops
dev = torch.device("cuda:0")
M, N, K = (int(v) for v in sys.argv[1:4])
torch.manual_seed(1)
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
B = torch.randn(N, K, device=dev).to(torch.bfloat16)
for trial in range(3):
    D = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    st = torch.zeros(2 * N, dtype=torch.float32, device=dev)
    ops.gemm(A, B, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N, colstats=st, colstats_mode=2)
    torch.cuda.synchronize()
    Df = D.float()
    print("   D max err", (Df - A.float() @ B.float().t()).abs().max().item())
    ref = (Df * Df).sum(0)
    diff = st[N:] - ref
    slab = (Df * Df).reshape(M // 32, 32, N).sum(1)           # [slabs, N]
    print("trial", trial, "sumsq rel err %.3e" % (diff.abs().max() / ref.max()).item(),
          "diff/mean-slab per column (first 8):", [round(v, 2) for v in (diff / slab.mean(0))[:8].tolist()],
          "cols wrong:", int((diff.abs() > 1e-4 * ref).sum()))
    # is the difference explained by whole tiles (128 rows) missing or doubled?
    tile = (Df * Df).reshape(M // 128, 128, N).sum(1)
    d0 = diff[0].item()
    near = (tile[:, 0] - abs(d0)).abs().argmin().item()
    print("   col0 diff %.3f  closest single-tile sum %.3f (tile %d)" % (d0, tile[near, 0].item(), near))
