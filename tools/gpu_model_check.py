"""Developer harness: model-level parity of the CUDA path against the CPU oracle, plus first timings.
Prints one line per check; never fails fast."""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dfu_multimodal_b200 as D  # noqa: E402
from dfu_multimodal_b200 import ops  # noqa: E402
from dfu_multimodal_b200.functional import cross_entropy  # noqa: E402
from dfu_multimodal_b200.gradcam import GradCAM  # noqa: E402
from dfu_multimodal_b200.optim import FusedAdamW  # noqa: E402
from dfu_multimodal_b200.train import EvalStep, TrainStep  # noqa: E402
from oracle import reference_model as R  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item(), b.abs().max().item()


def show(name, a, b):
    r, m, s = rel(a, b)
    print("%-46s rel=%.3e maxabs=%.3e refmax=%.3e" % (name, r, m, s), flush=True)


def section(name, fn):
    print("=== " + name, flush=True)
    t0 = time.time()
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:  # noqa: BLE001
        traceback.print_exc()
        print("!!! %s raised" % name, flush=True)
    print("--- %s done in %.1fs" % (name, time.time() - t0), flush=True)


def set_dropout(model, p):
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, D.nn.DfuDropout)):
            m.p = p


def main():
    B = int(os.environ.get("CHECK_B", "4"))
    oracle = R.seeded_init_(R.OracleFusion(), 0)
    model = D.MultimodalFusionModel()
    print(model.load_state_dict(oracle.state_dict(), strict=True))
    model.to(dev)
    rgb_u8, th_u8, labels = R.synthetic_batch(B, 0)
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    w = R.class_weights(labels)

    def eval_parity():
        oracle.eval(); model.eval()
        with torch.no_grad():
            ro = oracle.resnet(rgb); vo = oracle.vit(th); lo = oracle(rgb, th)
            rc = model.resnet(rgb.to(dev)); vc = model.vit(th.to(dev)); lc = model(rgb.to(dev), th.to(dev))
            lu = model(rgb_u8.to(dev), th_u8.to(dev))
        show("eval resnet features", rc, ro)
        show("eval vit features", vc, vo)
        show("eval logits (fp32 NCHW in)", lc, lo)
        show("eval logits (uint8 NHWC in)", lu, lo)
        print("logits oracle", lo.tolist(), "\nlogits cuda  ", lc.cpu().tolist())
    section("eval parity", eval_parity)

    def train_parity():
        oracle.train(); model.train()
        set_dropout(oracle, 0.0); set_dropout(model, 0.0)
        sd0 = {k: v.clone() for k, v in oracle.state_dict().items()}
        oracle.zero_grad()
        lo = torch.nn.functional.cross_entropy(oracle(rgb, th), labels, weight=w)
        lo.backward()
        model.zero_grad()
        lc = cross_entropy(model(rgb.to(dev), th.to(dev)), labels.to(dev), w.to(dev))
        lc.backward()
        print("loss oracle %.6f cuda %.6f" % (lo.item(), lc.item()))
        og = dict(oracle.named_parameters()); cg = dict(model.named_parameters())
        for k in ["fusion.classifier.3.weight", "fusion.classifier.0.weight", "vit.norm.weight", "vit.blocks.11.mlp.fc2.weight",
                  "vit.blocks.11.attn.qkv.weight", "vit.blocks.0.attn.qkv.weight", "vit.blocks.0.norm1.bias",
                  "vit.patch_embed.proj.weight", "vit.pos_embed", "vit.cls_token", "resnet.layer4.2.conv3.weight",
                  "resnet.layer4.2.bn3.weight", "resnet.layer4.0.downsample.0.weight", "resnet.layer3.0.conv2.weight",
                  "resnet.layer1.0.conv1.weight", "resnet.bn1.weight", "resnet.conv1.weight"]:
            if cg[k].grad is None:
                print("%-46s NO GRAD" % k)
            else:
                show("grad " + k, cg[k].grad, og[k].grad)
        show("bn1 running_mean", model.resnet.bn1.running_mean, oracle.resnet.bn1.running_mean)
        show("layer4.2.bn3 running_var", model.resnet.layer4[2].bn3.running_var, oracle.resnet.layer4[2].bn3.running_var)
        oracle.load_state_dict(sd0); model.load_state_dict(sd0)
    section("train fwd/bwd parity", train_parity)

    def loss_curve():
        steps = int(os.environ.get("CHECK_STEPS", "6"))
        oracle.train(); model.train()
        set_dropout(oracle, 0.0); set_dropout(model, 0.0)
        batches = []
        for s in range(steps):
            a, b, y = R.synthetic_batch(B, 100 + s)
            batches.append((a, b, y))
        lo = R.train_steps(oracle, [(R.preprocess(a, "imagenet"), R.preprocess(b, "half"), y) for a, b, y in batches], w)
        opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
        ts = TrainStep(model, opt, class_weights=w.to(dev), use_graph=False)
        lc = []
        for a, b, y in batches:
            loss, _ = ts(a.to(dev), b.to(dev), y.to(dev))
            lc.append(loss.item())
        print("oracle losses", ["%.4f" % v for v in lo])
        print("cuda   losses", ["%.4f" % v for v in lc])
        print("launches per eager step:", ts.launches_per_step)
    section("loss curve (eager, FusedAdamW)", loss_curve)

    def graph_step():
        m2 = D.MultimodalFusionModel().to(dev)
        m2.load_state_dict(R.seeded_init_(R.OracleFusion(), 0).state_dict())
        m3 = D.MultimodalFusionModel().to(dev)
        m3.load_state_dict(m2.state_dict())
        set_dropout(m2, 0.0); set_dropout(m3, 0.0)
        a, b, y = [t.to(dev) for t in R.synthetic_batch(B, 7)]
        wdev = w.to(dev)
        t_e = TrainStep(m2, FusedAdamW(m2.parameters(), lr=1e-4, weight_decay=1e-4), wdev, use_graph=False)
        t_g = TrainStep(m3, FusedAdamW(m3.parameters(), lr=1e-4, weight_decay=1e-4), wdev, use_graph=True)
        le, lg = [], []
        for i in range(7):
            le.append(t_e(a, b, y)[0].item())
            lg.append(t_g(a, b, y)[0].item())
        print("eager", ["%.4f" % v for v in le])
        print("graph", ["%.4f" % v for v in lg], "captured:", t_g.graph is not None, "launches/step", t_g.launches_per_step)
    section("CUDA-graph train step == eager", graph_step)

    def gradcam():
        gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_gradcam_seed0.pt"))
        model.load_state_dict(R.seeded_init_(R.OracleFusion(), 0).state_dict())
        model.eval()
        x_rgb = rgb[:1].to(dev); x_th = th[:1].to(dev)
        with torch.no_grad():
            g1 = GradCAM(model.resnet, ["layer4"]); cam_r = g1.generate_cam(x_rgb); g1.close()
            g2 = GradCAM(model.vit, ["blocks"]); cam_v = g2.generate_cam(x_th); g2.close()
        show("gradcam resnet layer4 vs reference class", torch.as_tensor(cam_r), gold["resnet_layer4_cam"])
        show("saliency vit blocks vs reference class", torch.as_tensor(cam_v), gold["vit_blocks_saliency"].float())
    section("Grad-CAM vs reference GradCAM golden", gradcam)

    def timing():
        for Bt, graph in ((32, False), (32, True), (128, True)):
            m = D.MultimodalFusionModel().to(dev)
            opt = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
            ts = TrainStep(m, opt, None, use_graph=graph)
            a, b, y = [t.to(dev) for t in R.synthetic_batch(Bt, 3)]
            for _ in range(5):
                ts(a, b, y)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                ts(a, b, y)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print("train step B=%d graph=%s: %.2f ms -> %.0f pairs/s (%.1f%% of 1638.8 TF)  mem %.1f GB" % (
                Bt, graph, ms, Bt / ms * 1e3, Bt / ms * 1e3 * 129.443e9 / 1638.8e12 * 100,
                torch.cuda.max_memory_allocated() / 2**30), flush=True)
            del m, opt, ts
            torch.cuda.empty_cache()
        m = D.MultimodalFusionModel().to(dev)
        es = EvalStep(m, use_graph=True)
        a, b, y = [t.to(dev) for t in R.synthetic_batch(128, 3)]
        for _ in range(3):
            es(a, b)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            es(a, b)
        torch.cuda.synchronize()
        ms = (time.time() - t0) / 5 * 1e3
        print("eval step B=128 graph: %.2f ms -> %.0f pairs/s" % (ms, 128 / ms * 1e3))
    section("timing", timing)


if __name__ == "__main__":
    main()
