"""Per-kernel time breakdown of one eager fusion train step (torch.profiler / CUPTI), for choosing what to optimise.
Usage: python tools/gpu_profile_step.py [B] ; under ncu use PROFILE=0 to skip the torch profiler."""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dfu_multimodal_b200 as D  # noqa: E402
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
if os.environ.get("PROFILE", "1") == "1":
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts(a, b, y)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    total = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            t = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            agg[ev.name][0] += t
            agg[ev.name][1] += 1
            total += t
    if os.environ.get("DUMP"):            # every launch in start order: "start_us duration_us name" (single stream: clean)
        evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda ev: ev.time_range.start)
        t0 = evs[0].time_range.start
        with open(os.environ["DUMP"], "w") as f:
            for ev in evs:
                f.write("%10.1f %8.1f %s\n" % (ev.time_range.start - t0, ev.device_time, ev.name[:90]))
    print("total device time %.2f ms over %d kernels" % (total / 1e3, sum(v[1] for v in agg.values())))
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%8.2f ms %5.1f%% n=%4d  %s" % (t / 1e3, 100 * t / total, n, name[:110]))
else:
    ts(a, b, y)
    torch.cuda.synchronize()
