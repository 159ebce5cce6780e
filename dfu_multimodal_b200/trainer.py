"""The reference's training script as a loop over the B200 path (notebooks/train_multimodal_fusion.py:355-451): epochs of
train (weighted sampler, augmentation, class-weighted CE, AdamW) and validation, history dict, best-by-validation-F1
checkpoint from epoch SAVE_BEST_AFTER_EPOCH on -- with the pieces the reference runs on the host moved to the device:
ingest (data.GpuIngest: resize, augmentation, normalisation), the step (train.TrainStep: one CUDA graph), the metrics
(metrics.DeviceMetrics: no per-step host sync; the reference syncs three times per step, :383-386)."""
import os
import random

import torch

from . import checkpoint
from .data import GpuIngest, PairedFolderDataset, class_weights
from .metrics import DeviceMetrics
from .optim import FusedAdamW
from .train import EvalStep, TrainStep


def evaluate(model, loader, class_w=None, eval_step=None):
    """extended_metrics.py:623-642 / the validation loop of train_multimodal_fusion.py:399-425: metrics stay on the device."""
    from .functional import cross_entropy
    step = eval_step or EvalStep(model, use_graph=False)
    dm = DeviceMetrics(next(model.parameters()).device, capacity=16384)
    dm.reset()
    for rgb, thermal, labels in loader:
        logits, _, _ = step(rgb, thermal)
        with torch.no_grad():
            loss = cross_entropy(logits, labels, class_w)
        dm.update(logits, labels, loss)
    return dm.compute()


def fit(model, rgb_dir, thermal_dir, epochs=10, batch_size=6, lr=1e-4, weight_decay=1e-4, save_best_after_epoch=3,
        checkpoint_dir=None, seed=42, use_graph=True, log=print):
    """Returns (history, best_val_f1).  Defaults are the reference's constants (train_multimodal_fusion.py:35-44)."""
    dev = next(model.parameters()).device
    rng = random.Random(seed)
    gen = torch.Generator().manual_seed(seed)
    train_ds = PairedFolderDataset(rgb_dir, thermal_dir, "train", rng)
    val_ds = PairedFolderDataset(rgb_dir, thermal_dir, "val", rng)
    cw = class_weights(train_ds.labels).to(dev)
    opt = FusedAdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
    step = TrainStep(model, opt, class_weights=cw, use_graph=use_graph)
    train_metrics = DeviceMetrics(dev, capacity=min(16384, max(len(train_ds), 1)))
    history = {k: [] for k in ("train_loss", "train_acc", "train_f1", "val_loss", "val_acc", "val_f1")}
    best = 0.0
    for epoch in range(1, epochs + 1):
        train_metrics.reset()
        for rgb, thermal, labels in GpuIngest(train_ds, batch_size, dev, train=True, sampler_generator=gen, aug_rng=rng):
            loss, logits = step(rgb, thermal, labels)
            train_metrics.update(logits, labels, loss)              # no host sync: read once per epoch below
        tm = train_metrics.compute()
        vm = evaluate(model, GpuIngest(val_ds, batch_size, dev, train=False), cw)
        for k, v in (("train_loss", tm["mean_loss"]), ("train_acc", tm["accuracy"]), ("train_f1", tm["f1"]),
                     ("val_loss", vm["mean_loss"]), ("val_acc", vm["accuracy"]), ("val_f1", vm["f1"])):
            history[k].append(v)
        log("[Epoch %d/%d] Train Loss: %.4f, Acc: %.4f, F1: %.4f | Val Loss: %.4f, Acc: %.4f, F1: %.4f" % (
            epoch, epochs, tm["mean_loss"], tm["accuracy"], tm["f1"], vm["mean_loss"], vm["accuracy"], vm["f1"]))
        if checkpoint_dir and epoch >= save_best_after_epoch and vm["f1"] > best:
            best = vm["f1"]
            os.makedirs(checkpoint_dir, exist_ok=True)
            checkpoint.save_checkpoint(os.path.join(checkpoint_dir, "best_model.pt"), model, opt, epoch, vm["f1"], history)
    return history, best
