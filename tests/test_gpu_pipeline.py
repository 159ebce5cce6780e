"""SURVEY.md 8(f) ranks 1-4 working together on the GPU: the ingest pipeline (paired dataset, weighted sampler, device-side
resize / augmentation / normalisation), the training loop with device-side metrics, and the best-by-F1 checkpoint, on the
synthetic data tree (tools/synthetic_tree.py)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    from tools.synthetic_tree import make_tree
    home = str(tmp_path_factory.mktemp("tree"))
    data, _ = make_tree(home, counts={"train": 6, "val": 3, "test": 4}, seed=9)
    return os.path.join(data, "rgb"), os.path.join(data, "thermal")


def test_pairing_follows_the_reference_dataset(tree):
    from dfu_multimodal_b200.data import PairedFolderDataset, class_weights, weighted_sample_indices
    rgb_dir, th_dir = tree
    ds = PairedFolderDataset(rgb_dir, th_dir, "test", random.Random(42))
    # 4 RGB and 5 thermal images per class: the shorter list is cycled (train_multimodal_fusion.py:114-132)
    assert len(ds) == 10 and ds.labels.count(0) == 5 and ds.labels.count(1) == 5
    by_label = {0: [], 1: []}
    for r, t, y in ds.pairs:
        assert ("ulcer" in r) == (y == 1) and ("ulcer" in t) == (y == 1)
        by_label[y].append(os.path.basename(r))
    assert all(len(set(v)) == 4 and len(v) == 5 for v in by_label.values())       # one RGB image of each class is used twice
    assert torch.allclose(class_weights(ds.labels), torch.tensor([2.0, 2.0]))
    idx = weighted_sample_indices([0] * 9 + [1], torch.Generator().manual_seed(0))
    assert len(idx) == 10 and 2 <= sum(1 for i in idx if i == 9) <= 9              # the rare class is drawn about half the time


def test_eval_ingest_equals_the_reference_transforms(tree):
    """GpuIngest(train=False) against Resize((224,224)) -> ToTensor -> Normalize of torchvision on the same files."""
    from PIL import Image
    from torchvision import transforms
    from dfu_multimodal_b200.data import GpuIngest, PairedFolderDataset
    rgb_dir, th_dir = tree
    ds = PairedFolderDataset(rgb_dir, th_dir, "test", random.Random(42))
    tf_rgb = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                                 transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    tf_th = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(), transforms.Normalize([0.5] * 3, [0.5] * 3)])
    loader = GpuIngest(ds, 4, "cuda:0", train=False)
    seen = 0
    for rgb, th, labels in loader:
        idx = loader.last_indices
        B = len(idx)
        assert rgb.shape == (B, 3, 224, 224) and rgb.dtype == torch.bfloat16 and th.shape == (B * 196, 768)
        assert labels.tolist() == [ds.labels[i] for i in idx]
        ref_rgb = torch.stack([tf_rgb(Image.open(ds.pairs[i][0]).convert("RGB")) for i in idx])
        ref_th = torch.stack([tf_th(Image.open(ds.pairs[i][1]).convert("RGB")) for i in idx])
        # the uint8 resize is exact (test_resize_is_bit_identical...), what remains is one bf16 rounding of the normalised value
        assert (rgb.float().cpu() - ref_rgb).abs().max().item() <= 2.0 ** -8 * ref_rgb.abs().max().item() + 1e-6
        ref_patches = ref_th.view(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * 196, 768)
        assert (th.float().cpu() - ref_patches).abs().max().item() <= 2.0 ** -8 + 1e-6
        seen += B
    assert seen == len(ds) and len(loader) == 3


def test_fit_trains_on_the_tree_with_device_side_metrics_and_saves_the_best_model(tree, tmp_path):
    import dfu_multimodal_b200 as D
    from dfu_multimodal_b200 import checkpoint, trainer
    from dfu_multimodal_b200.data import GpuIngest, PairedFolderDataset
    rgb_dir, th_dir = tree
    torch.manual_seed(0)
    model = D.MultimodalFusionModel(num_classes=2, dropout=0.5).cuda()
    lines = []
    history, best = trainer.fit(model, rgb_dir, th_dir, epochs=5, batch_size=6, checkpoint_dir=str(tmp_path), log=lines.append)
    print("\\n".join(lines))
    assert all(len(v) == 5 for v in history.values()) and all(np.isfinite(v).all() for v in history.values())
    assert min(history["train_loss"][2:]) < history["train_loss"][0]              # it learns the synthetic signal
    ck = os.path.join(str(tmp_path), "best_model.pt")
    if best > 0:
        assert os.path.exists(ck)
        ckpt = torch.load(ck, map_location="cpu", weights_only=False)
        assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "val_f1", "history"} and ckpt["epoch"] >= 3
        m2 = checkpoint.build_fusion_from_checkpoint(ckpt, device="cuda:0")
        test_ds = PairedFolderDataset(rgb_dir, th_dir, "test", random.Random(42))
        metrics = trainer.evaluate(m2, GpuIngest(test_ds, 8, "cuda:0", train=False))
        assert metrics["n"] == len(test_ds) and 0.0 <= metrics["accuracy"] <= 1.0 and "auc_roc" in metrics
