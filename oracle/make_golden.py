"""ORACLE TOOLING -- generates tests/golden/*.pt by running the REFERENCE'S OWN classes from /root/reference
(notebooks/extended_metrics.py MultimodalFusionModel / RGBOnlyModel / ThermalOnlyModel and
notebooks/grad_cam_visualization.py GradCAM) on CPU, with only the two third-party constructors substituted
(torch.hub.load -> torchvision resnet50, timm.create_model -> OracleViT; timm/matplotlib are not in this image).
Run in the build container only (needs /root/reference): `python oracle/make_golden.py`.  The GPU box never runs it;
tests read the committed fixtures."""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference_model as R  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _install_stubs():
    timm = types.ModuleType("timm")
    timm.create_model = lambda name, pretrained=False, num_classes=1000, **kw: R.OracleViT(num_classes=num_classes)
    sys.modules["timm"] = timm
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    mpl.use = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    if "seaborn" not in sys.modules:
        sys.modules["seaborn"] = types.ModuleType("seaborn")
    torch.hub.load = lambda repo, model, *a, **k: R.oracle_resnet50()
    from pathlib import Path
    Path.mkdir = lambda self, *a, **k: None       # the scripts create output dirs at import time


def _import(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference (build container only)")
    _install_stubs()
    os.makedirs(OUT, exist_ok=True)
    em = _import(os.path.join(REF, "notebooks", "extended_metrics.py"), "ref_extended_metrics")
    gc = _import(os.path.join(REF, "notebooks", "grad_cam_visualization.py"), "ref_grad_cam")

    # ---- fusion / single-modality logits from the reference's own model classes
    B = 2
    rgb_u8, th_u8, labels = R.synthetic_batch(B, seed=0)
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    torch.manual_seed(0)
    ref_fusion = em.MultimodalFusionModel()
    oracle = R.seeded_init_(R.OracleFusion(), 0)
    missing = ref_fusion.load_state_dict(oracle.state_dict(), strict=True)
    ref_fusion.eval()
    with torch.no_grad():
        logits = ref_fusion(rgb, th)
        probs, preds, _ = None, None, None
    golden = {"seed": 0, "B": B, "fusion_logits": logits.clone(), "labels": labels.clone()}

    ref_rgb = em.RGBOnlyModel()
    o_rgb = R.seeded_init_(R.OracleRGBOnly(), 1)
    ref_rgb.load_state_dict(o_rgb.state_dict(), strict=True)
    ref_rgb.eval()
    ref_th = em.ThermalOnlyModel()
    o_th = R.seeded_init_(R.OracleThermalOnly(), 2)
    ref_th.load_state_dict(o_th.state_dict(), strict=True)
    ref_th.eval()
    with torch.no_grad():
        golden["rgb_only_logits"] = ref_rgb(rgb).clone()
        golden["thermal_only_logits"] = ref_th(th).clone()
    torch.save(golden, os.path.join(OUT, "reference_logits_seed0.pt"))
    print("fusion logits", logits, str(missing))

    # ---- Grad-CAM from the reference's own GradCAM class (called under no_grad like its main, :831)
    x_rgb, x_th = rgb[:1], th[:1]
    with torch.no_grad():
        pred = int(logits[0].argmax())          # visualize_multimodal passes class_idx=prediction (:576)
        cam_r = gc.GradCAM(ref_fusion.resnet, ["layer4"]).generate_cam(x_rgb, class_idx=pred)
        cam_v = gc.GradCAM(ref_fusion.vit, ["blocks"]).generate_cam(x_th, class_idx=pred)
    cam_r = torch.as_tensor(cam_r).float()
    cam_v = torch.as_tensor(cam_v).float()
    torch.save({"seed": 0, "resnet_layer4_cam": cam_r, "vit_blocks_saliency": cam_v.half()},
               os.path.join(OUT, "reference_gradcam_seed0.pt"))
    print("cam_r", tuple(cam_r.shape), float(cam_r.max()), "cam_v", tuple(cam_v.shape), float(cam_v.max()))


if __name__ == "__main__":
    main()
