"""Pins the oracle (oracle/reference_model.py): against torchvision's independent ViT-B/16 through a pure key remap,
and against outputs of the reference's OWN classes (tests/golden/*.pt, written by oracle/make_golden.py from
/root/reference/notebooks/extended_metrics.py and grad_cam_visualization.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_model as R


def test_oracle_vit_matches_torchvision_vit_b16_through_key_remap():
    torchvision = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    tv = torchvision.models.vit_b_16(weights=None).eval()
    with torch.no_grad():
        for p in tv.parameters():
            if p.dim() > 1:
                p.normal_(0, 0.02)
            else:
                p.uniform_(-0.1, 0.1)
    ours = R.OracleViT(num_classes=1000).eval()
    missing = ours.load_state_dict(R.torchvision_vit_to_timm_keys(tv.state_dict()), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.randn(1, 3, 224, 224)
    with torch.no_grad():
        a, b = tv(x), ours(x)
    assert (a - b).abs().max().item() < 1e-4


def test_oracle_fusion_matches_reference_class_outputs(golden_dir):
    g = torch.load(os.path.join(golden_dir, "reference_logits_seed0.pt"))
    rgb_u8, th_u8, labels = R.synthetic_batch(g["B"], seed=g["seed"])
    assert torch.equal(labels, g["labels"])
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    m = R.seeded_init_(R.OracleFusion(), 0).eval()
    with torch.no_grad():
        out = m(rgb, th)
    assert (out - g["fusion_logits"]).abs().max().item() < 1e-4
    m = R.seeded_init_(R.OracleRGBOnly(), 1).eval()
    with torch.no_grad():
        assert (m(rgb) - g["rgb_only_logits"]).abs().max().item() < 1e-4
    m = R.seeded_init_(R.OracleThermalOnly(), 2).eval()
    with torch.no_grad():
        assert (m(th) - g["thermal_only_logits"]).abs().max().item() < 1e-4


def test_oracle_gradcam_matches_reference_gradcam_class(golden_dir):
    g = torch.load(os.path.join(golden_dir, "reference_gradcam_seed0.pt"))
    rgb_u8, th_u8, _ = R.synthetic_batch(2, seed=0)
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    m = R.seeded_init_(R.OracleFusion(), 0).eval()
    cam = R.oracle_gradcam(m.resnet, ["layer4"], rgb[:1])
    assert cam.shape == (7, 7)
    assert np.abs(cam - g["resnet_layer4_cam"].numpy()).max() < 1e-4
    sal = R.oracle_gradcam(m.vit, ["blocks"], th[:1])
    assert sal.shape == (224, 224)
    assert np.abs(sal - g["vit_blocks_saliency"].float().numpy()).max() < 2e-3     # fixture stored as fp16


def test_state_dict_layout_and_param_count():
    m = R.OracleFusion()
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 110_750_018
    assert len([k for k in sd if k.startswith("resnet.")]) == 318
    assert len([k for k in sd if k.startswith("vit.")]) == 150
    assert sd["fusion.classifier.0.weight"].shape == (512, 2816) and sd["fusion.classifier.3.weight"].shape == (2, 512)


def test_preprocess_and_class_weights_and_tta_vote():
    u8 = torch.tensor([[[[0, 128, 255]]]], dtype=torch.uint8)                      # (1,1,1,3) HWC
    x = R.preprocess(u8, "imagenet")
    assert x.shape == (1, 3, 1, 1)
    assert abs(x[0, 0, 0, 0].item() - (0 - 0.485) / 0.229) < 1e-6
    assert abs(x[0, 2, 0, 0].item() - (1 - 0.406) / 0.225) < 1e-6
    h = R.preprocess(u8, "half")
    assert abs(h[0, 1, 0, 0].item() - (128 / 255 - 0.5) / 0.5) < 1e-6
    w = R.class_weights(torch.tensor([0, 0, 0, 1]))
    assert torch.allclose(w, torch.tensor([4 / 3, 4.0]))
    mp, pred = R.oracle_tta_vote([[0.6, 0.7, 0.2, 0.1], [0.9, 0.8, 0.7, 0.1], [0.4, 0.6, 0.6, 0.4]])
    assert pred.tolist() == [0, 1, 0]                                               # ties vote 0 (mean 0.5 is not > 0.5)
    assert np.allclose(mp, [0.4, 0.625, 0.5])


def test_seeded_init_is_deterministic():
    a = R.seeded_init_(R.OracleThermalOnly(), 5).state_dict()
    b = R.seeded_init_(R.OracleThermalOnly(), 5).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_resize_restatement_is_bit_identical_to_pillow():
    """oracle/resize_pil.py (numpy restatement of Pillow's ImagingResample) against the library the reference really calls:
    torchvision.transforms.Resize((224,224)) on a PIL image (train_multimodal_fusion.py:186)."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from oracle.resize_pil import precompute_coeffs, resize_bilinear_u8
    rng = np.random.RandomState(0)
    tf = transforms.Resize((224, 224))
    for h, w in [(224, 224), (240, 200), (300, 260), (180, 232), (480, 640), (100, 90), (225, 223), (448, 448), (37, 1024), (1, 1)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = np.asarray(tf(Image.fromarray(img)))
        assert np.array_equal(resize_bilinear_u8(img), ref), (h, w)
    b, c, k = precompute_coeffs(448, 224)
    assert k == 5 and int(c.sum(1).min()) >= (1 << 22) - 4 and int(c.sum(1).max()) <= (1 << 22) + 4


def test_augmentation_restatement_is_bit_identical_to_torchvision_on_pil():
    """oracle/augment_pil.py against the calls the reference's train transforms make (train_multimodal_fusion.py:173-199):
    torchvision.transforms.functional hflip / vflip / rotate / adjust_brightness / adjust_contrast / adjust_saturation / affine
    on PIL images, with the same drawn parameters -- including Pillow's fast paths (right angles, scale-only affine)."""
    import random
    import numpy as np
    import torchvision.transforms.functional as TF
    from PIL import Image
    from oracle import augment_pil as A
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (224, 224, 3)).astype(np.uint8)
    pil = Image.fromarray(img)
    random.seed(11)
    cases = []
    for _ in range(12):
        p = dict(hflip=random.random() < 0.5, vflip=random.random() < 0.5, angle=random.uniform(-30, 30))
        if random.random() < 0.7:
            order = [0, 1, 2]
            random.shuffle(order)
            p["jitter"] = dict(order=order, brightness=random.uniform(0.7, 1.3), contrast=random.uniform(0.7, 1.3),
                               saturation=random.uniform(0.7, 1.3))
        if random.random() < 0.7:
            p["affine"] = dict(angle=random.uniform(-20, 20), scale=random.uniform(0.8, 1.2),
                               translate=(int(round(random.uniform(-22.4, 22.4))), int(round(random.uniform(-22.4, 22.4)))))
        cases.append(p)
    cases += [dict(hflip=False, vflip=False, angle=0.0, affine=dict(angle=0.0, translate=(3, -4), scale=1.1)),
              dict(hflip=True, vflip=False, angle=180.0), dict(hflip=False, vflip=True, angle=90.0), dict(hflip=True, vflip=True, angle=270.0)]
    for p in cases:
        x = pil
        if p["hflip"]:
            x = TF.hflip(x)
        if p["vflip"]:
            x = TF.vflip(x)
        x = TF.rotate(x, p["angle"])
        if "jitter" in p:
            for op in p["jitter"]["order"]:
                x = [TF.adjust_brightness, TF.adjust_contrast, TF.adjust_saturation][op](x, p["jitter"][["brightness", "contrast", "saturation"][op]])
        if "affine" in p:
            a = p["affine"]
            x = TF.affine(x, a["angle"], list(a["translate"]), a["scale"], [0.0, 0.0])
        assert np.array_equal(np.asarray(x), A.augment(img, p)), p
