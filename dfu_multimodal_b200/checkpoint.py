"""Checkpoint compatibility with the reference's four key schemes.

The reference writes {'epoch','model_state_dict','optimizer_state_dict','val_f1','history'}
(notebooks/train_multimodal_fusion.py:433-439) and its loaders only rename `backbone.` / `vit_backbone.` before a
strict=False load (extended_metrics.py:803-817), so a checkpoint of the training script silently loads nothing.
`remap_state_dict` maps every scheme onto the canonical (eval-side) names so the weights really load."""
import re

import torch

_PREFIX_RULES = [
    # fusion checkpoints
    (r"^rgb_branch\.", "resnet."),
    (r"^thermal_branch\.", "vit."),
    (r"^vit_backbone\.", "vit."),
]


def detect_scheme(sd):
    keys = list(sd.keys())
    if any(k.startswith("rgb_branch.") for k in keys):
        return "train" if "fusion.6.weight" in sd else "ablation"
    if any(k.startswith("fusion.classifier.") for k in keys):
        return "tta" if sd["fusion.classifier.3.weight"].shape[0] == 1 else "eval"
    if any(k.startswith("backbone.") for k in keys):
        if any(k.startswith("backbone.blocks.") for k in keys):
            return "thermal_only"
        return "rgb_only" if any(k.startswith("backbone.layer1.") for k in keys) and not any(
            k.startswith(("vit.", "vit_backbone.", "thermal_branch.")) for k in keys) else "legacy_fusion"
    if any(k.startswith("resnet.") for k in keys) and not any(k.startswith("vit.") for k in keys):
        return "rgb_only"
    if any(k.startswith("vit.") for k in keys) and not any(k.startswith("resnet.") for k in keys):
        return "thermal_only"
    return "eval"


def remap_state_dict(sd, target="fusion"):
    """Returns a new dict in canonical names.  target: 'fusion' (resnet./vit./fusion.classifier.*),
    'rgb' (resnet.*, head keys resnet.fc.1.*), 'thermal' (vit.*, head keys vit.head.1.*)."""
    out = {}
    for k, v in sd.items():
        nk = k
        for pat, rep in _PREFIX_RULES:
            nk = re.sub(pat, rep, nk)
        if nk.startswith("backbone."):
            rest = nk[len("backbone."):]
            if target == "thermal" or rest.startswith(("blocks.", "cls_token", "pos_embed", "patch_embed.", "norm.", "head.")):
                nk = "vit." + rest
            else:
                nk = "resnet." + rest
        if target == "fusion":
            m = re.match(r"^fusion\.(\d+)\.(weight|bias)$", nk)
            if m:
                nk = "fusion.classifier.%s.%s" % (m.group(1), m.group(2))
        out[nk] = v
    return out


def head_layers_of(sd_canonical):
    return 3 if "fusion.classifier.6.weight" in sd_canonical else 2


def load_into(model, sd_or_ckpt, target="fusion", strict=True):
    """Remaps and loads; returns torch's (missing, unexpected) report.  Unlike the reference it defaults to
    strict=True so a scheme mismatch is an error instead of a silently random model."""
    sd = sd_or_ckpt.get("model_state_dict", sd_or_ckpt) if isinstance(sd_or_ckpt, dict) else sd_or_ckpt
    sd = remap_state_dict(sd, target)
    return model.load_state_dict(sd, strict=strict)


def build_fusion_from_checkpoint(sd_or_ckpt, device=None):
    from .fusion import MultimodalFusionModel
    sd = sd_or_ckpt.get("model_state_dict", sd_or_ckpt)
    sd = remap_state_dict(sd, "fusion")
    w_last = sd["fusion.classifier.%d.weight" % (6 if head_layers_of(sd) == 3 else 3)]
    model = MultimodalFusionModel(num_classes=w_last.shape[0], head_layers=head_layers_of(sd))
    model.load_state_dict(sd, strict=True)
    return model.to(device) if device is not None else model


def save_checkpoint(path, model, optimizer, epoch, val_f1, history):
    """Same container as train_multimodal_fusion.py:433-439."""
    torch.save({"epoch": epoch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
                "val_f1": val_f1, "history": history}, path)


def resume(path_or_ckpt, model, optimizer=None, target="fusion", map_location="cpu"):
    """Continue a run from a checkpoint in the reference's container (train_multimodal_fusion.py:433-439; the reference saves
    `optimizer_state_dict` but never reloads it).  The model weights load through the key remap (any scheme); the optimizer
    state loads into FusedAdamW or a stock torch.optim.AdamW alike -- parameter order is the modules' registration order
    (RGB branch, thermal branch, head), which all schemes share.  Returns {'epoch', 'val_f1', 'history'}."""
    ckpt = torch.load(path_or_ckpt, map_location=map_location, weights_only=False) if isinstance(path_or_ckpt, (str, bytes)) or hasattr(
        path_or_ckpt, "__fspath__") else path_or_ckpt
    load_into(model, ckpt, target=target, strict=True)
    osd = ckpt.get("optimizer_state_dict")
    if optimizer is not None and osd:
        n_ckpt = sum(len(g["params"]) for g in osd["param_groups"])
        n_here = sum(len(g["params"]) for g in optimizer.param_groups)
        if n_ckpt != n_here:
            raise RuntimeError("optimizer state holds %d parameters, the optimizer %d (different head depth?)" % (n_ckpt, n_here))
        optimizer.load_state_dict(osd)
    return {"epoch": ckpt.get("epoch", 0), "val_f1": ckpt.get("val_f1"), "history": ckpt.get("history")}
