"""Host-side logic that needs no GPU: module tree / state_dict compatibility, checkpoint key remaps for the four
reference schemes, the timm / torch.hub drop-in, head conversion, hook detection, and the
fail-loudly contract (no CPU fallback)."""
import sys

import pytest
import torch
import torch.nn as nn

import dfu_multimodal_b200 as D
from dfu_multimodal_b200 import checkpoint, dropin
from dfu_multimodal_b200.nn import DfuDropout, DfuHead, DfuLinear, has_hooks
from oracle import reference_model as R


def test_fusion_state_dict_equals_oracle_layout():
    ours = D.MultimodalFusionModel().state_dict()
    ref = R.OracleFusion().state_dict()
    assert list(ours.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(ours[k].shape) == tuple(ref[k].shape), k
    m = D.MultimodalFusionModel()
    assert sum(p.numel() for p in m.parameters()) == 110_750_018


def test_load_oracle_weights_roundtrip_keeps_channels_last_conv_storage():
    oracle = R.seeded_init_(R.OracleFusion(), 0, calibrate=False)
    m = D.MultimodalFusionModel()
    res = m.load_state_dict(oracle.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    w = m.resnet.layer1[0].conv2.weight
    assert w.permute(0, 2, 3, 1).is_contiguous()           # storage is [Cout, R, S, Cin]
    assert torch.equal(w, oracle.resnet.layer1[0].conv2.weight)
    back = m.state_dict()
    assert all(torch.equal(back[k], v) for k, v in oracle.state_dict().items())


def test_single_modality_models_use_reference_key_names():
    rgb = D.RGBOnlyModel().state_dict()
    assert "resnet.fc.1.weight" in rgb and rgb["resnet.fc.1.weight"].shape == (2, 2048)
    th = D.ThermalOnlyModel().state_dict()
    assert "vit.head.1.weight" in th and th["vit.head.1.weight"].shape == (2, 768)
    assert list(rgb.keys()) == list(R.OracleRGBOnly().state_dict().keys())
    assert list(th.keys()) == list(R.OracleThermalOnly().state_dict().keys())


def _canon():
    return R.seeded_init_(R.OracleFusion(), 3, calibrate=False).state_dict()


def test_remap_train_scheme_with_three_layer_head():
    sd = R.seeded_init_(R.OracleFusion(head_layers=3), 3, calibrate=False).state_dict()
    train = {}
    for k, v in sd.items():                                # what train_multimodal_fusion.py:294-313 writes
        k = k.replace("resnet.", "rgb_branch.", 1) if k.startswith("resnet.") else k
        k = k.replace("vit.", "thermal_branch.", 1) if k.startswith("vit.") else k
        k = k.replace("fusion.classifier.", "fusion.")
        train[k] = v
    assert checkpoint.detect_scheme(train) == "train"
    model = checkpoint.build_fusion_from_checkpoint({"model_state_dict": train, "epoch": 4})
    assert isinstance(model.fusion.classifier[6], DfuLinear)
    assert torch.equal(model.state_dict()["fusion.classifier.6.weight"], sd["fusion.classifier.6.weight"])
    assert torch.equal(model.state_dict()["resnet.layer3.2.conv2.weight"], sd["resnet.layer3.2.conv2.weight"])


def test_remap_ablation_legacy_and_eval_schemes():
    sd = _canon()
    abl = {k.replace("resnet.", "rgb_branch.", 1).replace("vit.", "thermal_branch.", 1).replace("fusion.classifier.", "fusion."): v
           for k, v in sd.items()}
    assert checkpoint.detect_scheme(abl) == "ablation"
    legacy = {("backbone." + k[7:] if k.startswith("resnet.") else "vit_backbone." + k[4:] if k.startswith("vit.") else k): v
              for k, v in sd.items()}
    for variant in (abl, legacy, sd):
        m = D.MultimodalFusionModel()
        res = checkpoint.load_into(m, {"model_state_dict": variant}, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        assert torch.equal(m.state_dict()["vit.blocks.7.attn.qkv.weight"], sd["vit.blocks.7.attn.qkv.weight"])


def test_remap_single_modality_backbone_prefix():
    rgb = R.OracleRGBOnly().state_dict()
    train = {k.replace("resnet.", "backbone.", 1): v for k, v in rgb.items()}       # train_rgb_only.py:206-214
    m = D.RGBOnlyModel()
    assert not checkpoint.load_into(m, train, target="rgb").missing_keys
    th = R.OracleThermalOnly().state_dict()
    train = {k.replace("vit.", "backbone.", 1): v for k, v in th.items()}           # train_thermal_only.py:194-202
    m = D.ThermalOnlyModel()
    assert not checkpoint.load_into(m, train, target="thermal").missing_keys


def test_strict_load_reports_scheme_mismatch_instead_of_silently_skipping():
    m = D.MultimodalFusionModel()
    with pytest.raises(RuntimeError):
        m.load_state_dict({"rgb_branch.conv1.weight": torch.zeros(64, 3, 7, 7)}, strict=True)


def test_dropin_timm_shim_and_hub_interceptor():
    dropin.install(force_timm_shim=True)
    try:
        import timm
        vit = timm.create_model("vit_base_patch16_224", pretrained=False, num_classes=0)
        assert isinstance(vit, D.DfuVisionTransformer) and isinstance(vit.head, nn.Identity)
        vit2 = timm.create_model("vit_base_patch16_224", pretrained=False, num_classes=2)
        assert vit2.head.in_features == 768
        vit2.head = nn.Sequential(nn.Dropout(p=0.5), nn.Linear(768, 2))             # train_thermal_only.py:199-202
        assert isinstance(vit2.head, DfuHead) and isinstance(vit2.head[1], DfuLinear) and isinstance(vit2.head[0], DfuDropout)
        assert "head.1.weight" in vit2.state_dict()
        with pytest.raises(RuntimeError):
            timm.create_model("resnet18")
        rn = torch.hub.load("pytorch/vision:v0.13.1", "resnet50", pretrained=False)
        assert isinstance(rn, D.DfuResNet50) and rn.fc.in_features == 2048
        rn.fc = nn.Identity()                                                        # train_multimodal_fusion.py:296
        assert isinstance(rn.fc, nn.Identity)
        rn.fc = nn.Sequential(nn.Dropout(0.5), nn.Linear(2048, 2))                   # train_rgb_only.py:211-214
        assert "fc.1.bias" in rn.state_dict()
        names = [n for n, _ in rn.named_modules()]
        assert "layer4.2.relu" in names and "layer4.0.downsample.1" in names
        assert any(n.startswith("blocks.11") for n, _ in vit.named_modules())
    finally:
        dropin.uninstall()
    assert "timm" not in sys.modules or not getattr(sys.modules["timm"], "__dfu_shim__", False)


def test_has_hooks_switches_off_fused_paths():
    m = D.DfuResNet50()
    blk = m.layer4[2]
    assert not has_hooks(blk)
    h = blk.relu.register_forward_hook(lambda mod, i, o: None)
    assert has_hooks(blk)
    h.remove()
    assert not has_hooks(blk)


def test_no_cpu_fallback_product_path_fails_loudly():
    m = D.MultimodalFusionModel()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 224, 224), torch.zeros(1, 3, 224, 224))
    from dfu_multimodal_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.predict(torch.zeros(2, 2))
    import inspect
    import dfu_multimodal_b200.ops as ops_mod
    src = inspect.getsource(ops_mod) + inspect.getsource(D.fusion) + inspect.getsource(D.resnet) + inspect.getsource(D.vit)
    assert "oracle" not in src, "product code must not import the oracle"


def test_head_builder_variants():
    from dfu_multimodal_b200.fusion import make_head
    h2, h3 = make_head(2816, 2, 0.7, 2), make_head(2816, 2, 0.5, 3)
    assert [k for k in h2.state_dict()] == ["0.weight", "0.bias", "3.weight", "3.bias"]
    assert [k for k in h3.state_dict()] == ["0.weight", "0.bias", "3.weight", "3.bias", "6.weight", "6.bias"]


def test_device_prefetcher_has_no_cpu_path():
    from dfu_multimodal_b200.data import DevicePrefetcher
    with pytest.raises(RuntimeError):
        DevicePrefetcher(iter([]), "cpu")


def test_kernel_selection_predicates_cover_every_resnet_convolution():
    """Which convolutions take the implicit-GEMM entry point (dfu_conv_gemm) and which shapes the epilogue statistics
    serve: every 3x3 and every strided convolution of ResNet50 except the 3-channel stem; 1x1/s1 is a plain GEMM."""
    from dfu_multimodal_b200 import ops
    ops.set_conv_impl("implicit")
    m = D.DfuResNet50()
    implicit = direct = explicit = 0
    for name, mod in m.named_modules():
        if not isinstance(mod, D.nn.DfuConv2d):
            continue
        Cout, C, R, S = mod.weight.shape
        stride = mod.stride[0]
        if ops.conv_implicit_ok(C, Cout, R, S, stride):
            implicit += 1
            assert R * S > 1 or stride == 2, name
        elif R == 1 and S == 1 and stride == 1:
            direct += 1
        else:
            explicit += 1
            assert name == "conv1", name             # only the stem goes through an im2col matrix
        assert ops.colstats_ok(Cout, mode=2), name   # BatchNorm sums of every layer come from the GEMM epilogue
    assert (implicit, direct, explicit) == (16 + 3, 33, 1)
    assert ops.colstats_ok(3072, mode=1) and not ops.colstats_ok(3072, mode=2)      # fc1 bias gradient: sums only
    ops.set_conv_impl("im2col")
    try:
        assert not ops.conv_implicit_ok(64, 64, 3, 3, 1)
    finally:
        ops.set_conv_impl("implicit")


def test_abi_rejects_unsupported_convolution_and_epilogue_combinations_without_a_device():
    """Argument validation happens before any CUDA call, so the error contract is testable on the CPU."""
    import ctypes
    from dfu_multimodal_b200 import _lib
    L = _lib.load()
    a = _lib.ConvArgs()
    a.mode, a.B, a.H, a.W, a.C, a.Cout, a.R, a.S, a.stride, a.pad = 1, 1, 8, 8, 48, 64, 3, 3, 1, 1      # C % 64 != 0
    assert L.dfu_conv_gemm(ctypes.byref(a), None) != 0
    assert b"multiples of 64" in L.dfu_last_error()
    a.C, a.stride = 64, 3
    assert L.dfu_conv_gemm(ctypes.byref(a), None) != 0
    assert b"stride" in L.dfu_last_error()
    g = _lib.GemmArgs()
    g.M = g.N = g.K = 64
    g.nb0 = g.nb1 = 1
    g.A = g.B = g.D = 16
    g.act = 5                                            # D *= res needs res
    assert L.dfu_gemm(ctypes.byref(g), None) != 0
    assert b"act=5" in L.dfu_last_error()
    g.act = 4                                            # GELU with derivative output needs aux
    assert L.dfu_gemm(ctypes.byref(g), None) != 0
    assert b"act=4" in L.dfu_last_error()


# ----------------------------------------------------------------------------------------------- round 2: host logic of f1 / f2 / f4
def test_paired_dataset_and_sampler_follow_the_reference(tmp_path):
    """data.PairedFolderDataset == MultimodalDataset's pairing (train_multimodal_fusion.py:108-140) on a synthetic tree."""
    import os
    import random
    from dfu_multimodal_b200.data import PairedFolderDataset, class_weights, weighted_sample_indices
    from tools.synthetic_tree import make_tree
    data, files = make_tree(str(tmp_path), counts={"train": 3, "val": 2, "test": 2}, seed=1)
    assert len(files) == 2 * 2 * (3 + 2 + 2) + 2 * 3            # thermal holds one image more per class and split
    ds = PairedFolderDataset(os.path.join(data, "rgb"), os.path.join(data, "thermal"), "train", random.Random(42))
    assert len(ds) == 8 and ds.labels.count(0) == 4 and ds.labels.count(1) == 4
    healthy = [(os.path.basename(r), os.path.basename(t)) for r, t, y in ds.pairs if y == 0]
    # 3 RGB against 4 thermal images: pair i = (rgb[i % 3], thermal[i % 4])
    assert sorted(healthy) == sorted([("rgb_train_healthy_%03d.png" % (i % 3), "thermal_train_healthy_%03d.png" % i) for i in range(4)])
    rgb, th, y = ds.load(0)
    assert rgb.dtype.name == "uint8" and rgb.ndim == 3 and rgb.shape[2] == 3 and th.shape[2] == 3 and y in (0, 1)
    assert torch.allclose(class_weights([0, 0, 0, 1]), torch.tensor([4 / 3, 4.0]))
    idx = weighted_sample_indices([0] * 30 + [1] * 2, torch.Generator().manual_seed(3))
    assert len(idx) == 32 and 8 <= sum(1 for i in idx if i >= 30) <= 24         # classes are drawn about equally often


def test_augmentation_parameter_packing_matches_the_pillow_maps():
    """augment.pack_params builds the same 16.16 fixed-point inverse maps as the oracle's restatement of Pillow / torchvision."""
    import ctypes
    import random
    from dfu_multimodal_b200 import augment as G
    from dfu_multimodal_b200._lib import AugmentParams
    from oracle import augment_pil as A
    rng = random.Random(3)
    ps = G.draw_params(100, "rgb", rng=rng)
    ps += [{"hflip": False, "vflip": True, "angle": 0.0, "affine": {"angle": 0.0, "translate": (3, -4), "scale": 1.1}},
           {"hflip": True, "vflip": False, "angle": 180.0}, {"hflip": True, "vflip": False, "angle": 90.0}]
    packed = G.pack_params(ps)
    arr = (AugmentParams * len(ps)).from_buffer_copy(packed.numpy().tobytes())
    assert packed.numel() == len(ps) * ctypes.sizeof(AugmentParams)
    for rec, p in zip(arr, ps):
        assert rec.flags == (1 if p["hflip"] else 0) | (2 if p["vflip"] else 0)
        if p["angle"] % 360.0 not in (0, 90, 180, 270):
            assert rec.rot_kind == 1 and list(rec.rot) == A.fixed_coeffs(A.rotate_matrix(p["angle"], 224, 224))
        if "affine" in p and rec.aff_kind == 1:
            a = p["affine"]
            assert list(rec.aff) == A.fixed_coeffs(A.inverse_affine_matrix(a["angle"], a["translate"], a["scale"], 224, 224))
        if "jitter" in p:
            assert rec.n_color == 3 and list(rec.color_op) == p["jitter"]["order"]
    assert [r.rot_kind for r in arr[-3:]] == [0, 2, 3] and arr[-3].aff_kind == 5
    assert list(arr[-3].aff_xtab[:3]) == [int(3 * 0 + (112 - 112 / 1.1 - 3 / 1.1 + 0.5 / 1.1 + k / 1.1)) for k in range(3)]
    assert not any("jitter" in p for p in G.draw_params(50, "thermal", rng=rng))


def test_fused_adamw_state_dict_is_a_torch_adamw_state_dict():
    """The schema half of resume (f4) needs no GPU: FusedAdamW.state_dict() loads into torch.optim.AdamW and steps like it."""
    from dfu_multimodal_b200.optim import FusedAdamW

    def make():
        g = torch.Generator().manual_seed(7)
        return [nn.Parameter(torch.randn(s, generator=g)) for s in [(6, 4, 3, 3), (5,)]]

    def run(opt, ps, steps):
        for st in steps:
            g = torch.Generator().manual_seed(100 + st)
            for p in ps:
                p.grad = torch.randn(p.shape, generator=g)
            opt.step()
    a = make()
    oa = torch.optim.AdamW(a, lr=1e-3, weight_decay=1e-2)
    run(oa, a, range(5))
    c = [nn.Parameter(p.detach().clone()) for p in a]
    oc = FusedAdamW(c, lr=1e-3, weight_decay=1e-2)
    step = torch.tensor(5.0)
    oc._steps[("cpu", None)] = step
    for p, q in zip(c, a):
        oc.state[p] = {"step": step, "exp_avg": oa.state[q]["exp_avg"].clone(), "exp_avg_sq": oa.state[q]["exp_avg_sq"].clone()}
    sd = oc.state_dict()
    assert sd["param_groups"][0]["decoupled_weight_decay"] is True and float(sd["state"][0]["step"]) == 5.0
    assert sd["state"][0]["step"] is not sd["state"][1]["step"]                   # per-parameter copies, like torch's
    d = [nn.Parameter(p.detach().clone()) for p in c]
    od = torch.optim.AdamW(d, lr=1e-3, weight_decay=1e-2)
    od.load_state_dict(sd)
    run(od, d, [5])
    run(oa, a, [5])
    assert all(torch.equal(p, q) for p, q in zip(a, d))
    with pytest.raises(RuntimeError):
        oc.step()                                                                 # ... and the update itself has no CPU path


def test_staged_reference_archive_unpacks_to_the_recorded_hashes(tmp_path):
    import hashlib
    import os
    from oracle import stage_reference as S
    if not os.path.exists(S.ARCHIVE):
        pytest.skip("no staged archive (oracle/_ref is built by __graft_entry__.build() where /root/reference exists)")
    real = S.REFERENCE
    S.REFERENCE = str(tmp_path / "absent")                                        # force the archive path
    try:
        dest = S.materialize(str(tmp_path / "out"))
    finally:
        S.REFERENCE = real
    import json
    with open(S.MANIFEST) as f:
        want = json.load(f)["sha256"]
    for rel, sha in want.items():
        with open(os.path.join(dest, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha
