"""Layer-by-layer eval-mode comparison of the CUDA ResNet50 / ViT against the oracle."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dfu_multimodal_b200 as D
from dfu_multimodal_b200 import ops
from oracle import reference_model as R
dev = torch.device("cuda:0")
oracle = R.seeded_init_(R.OracleFusion(), 0).eval()
model = D.MultimodalFusionModel(); model.load_state_dict(oracle.state_dict()); model.to(dev).eval()
rgb_u8, th_u8, labels = R.synthetic_batch(2, 0)
rgb = R.preprocess(rgb_u8, "imagenet")
def rel(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a-b).abs().max().item(), b.abs().max().item()
for impl in (0, 1):
    ops.set_gemm_impl(impl)
    print("GEMM impl", impl)
    with torch.no_grad():
        xo = rgb; xc = rgb.to(dev)
        ro, rc = oracle.resnet, model.resnet
        steps = [("conv1", ro.conv1, rc.conv1), ("bn1", ro.bn1, rc.bn1), ("relu", ro.relu, rc.relu), ("maxpool", ro.maxpool, rc.maxpool)]
        for name, fo, fc in steps:
            xo = fo(xo); xc = fc(xc)
            print("%-22s rel=%.3e maxabs=%.3e refmax=%.3e" % ((name,) + rel(xc, xo)))
        for li in (1, 2, 3, 4):
            lo, lc = getattr(ro, "layer%d" % li), getattr(rc, "layer%d" % li)
            for bi in range(len(lo)):
                xo = lo[bi](xo); xc = lc[bi](xc)
                print("%-22s rel=%.3e maxabs=%.3e refmax=%.3e" % (("layer%d.%d" % (li, bi),) + rel(xc, xo)))
                if li == 1 and bi == 0 and impl == 0:
                    # feed oracle input into the CUDA block to isolate the block error
                    pass
        # block-isolated errors: oracle activations in, one CUDA block
        xo = ro.maxpool(ro.relu(ro.bn1(ro.conv1(rgb))))
        for li in (1, 2, 3, 4):
            lo, lc = getattr(ro, "layer%d" % li), getattr(rc, "layer%d" % li)
            for bi in range(len(lo)):
                yo = lo[bi](xo)
                yc = lc[bi](xo.to(dev))
                print("isolated %-13s rel=%.3e maxabs=%.3e refmax=%.3e  minvar=%.2e" % (("layer%d.%d" % (li, bi),) + rel(yc, yo) + (min(float(lo[bi].bn1.running_var.min()), float(lo[bi].bn2.running_var.min()), float(lo[bi].bn3.running_var.min())),)))
                xo = yo
