"""Per-shape table of the GEMM launches of one eager fusion train step, timed in situ with CUDA events on the
launching stream (ops.GEMM_EVENTS).  Usage: python tools/gpu_step_gemms.py [B]"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dfu_multimodal_b200 as D  # noqa: E402
from dfu_multimodal_b200 import ops  # noqa: E402
from dfu_multimodal_b200.optim import FusedAdamW  # noqa: E402
from dfu_multimodal_b200.train import TrainStep  # noqa: E402
from oracle import reference_model as R  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
m = D.MultimodalFusionModel().to(dev)
opt = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
ts = TrainStep(m, opt, None, use_graph=False)
a, b, y = [t.to(dev) for t in R.synthetic_batch(B, 3)]
for _ in range(2):
    ts(a, b, y)
torch.cuda.synchronize()
ops.GEMM_EVENTS = []
torch.cuda._sleep(int(0.15 * 1.9e9))     # GPU busy while the host enqueues: events then bracket back-to-back kernels
ts(a, b, y)
torch.cuda.synchronize()
evs, ops.GEMM_EVENTS = ops.GEMM_EVENTS, None
agg = defaultdict(lambda: [0.0, 0])
for e0, e1, sig in evs:
    agg[sig][0] += e0.elapsed_time(e1)
    agg[sig][1] += 1
tot = sum(v[0] for v in agg.values())
print("B=%d: %d GEMM launches, %.2f ms" % (B, len(evs), tot))
print("%-8s %8s %6s %6s %6s %2s %2s %4s %3s %4s %3s %3s %3s | %4s %9s %9s %8s" % (
    "kind", "M", "N", "K", "nb", "aM", "bM", "out", "acc", "bias", "act", "res", "aux", "n", "total ms", "us/launch", "TF/s"))
for sig, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    M, N, K, nb = sig[:4]
    fl = 2.0 * M * N * K * nb
    print("%-8s %8d %6d %6d %6d %2d %2d %4s %3d %4d %3d %3d %3d | %4d %9.3f %9.1f %8.1f" % (
        sig[12], M, N, K, nb, sig[4], sig[5], sig[6], sig[7], sig[8], sig[9], sig[10], sig[11], n, ms, ms / n * 1e3,
        fl * n / ms / 1e9))
