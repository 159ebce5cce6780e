"""Drop-in boundary for the reference drivers.

The scripts build their backbones through two third-party constructors:
  torch.hub.load('pytorch/vision:v0.13.1', 'resnet50', pretrained=True)   (train_multimodal_fusion.py:294)
  timm.create_model('vit_base_patch16_224', pretrained=True, num_classes=0|2)  (train_multimodal_fusion.py:299)
`install()` routes both to the B200 modules: a `timm` shim module is placed in sys.modules (timm is not in this
image) and torch.hub.load is intercepted for the torchvision resnet50 entry -- nothing is fetched or executed
from the network.  `pretrained=True` cannot download here: weights are seeded random init unless
DFU_PRETRAINED_DIR holds `resnet50.pt` / `vit_base_patch16_224.pt` state_dicts."""
import os
import sys
import types
import warnings

import torch

_ORIG_HUB_LOAD = None


def _maybe_pretrained(model, name, pretrained):
    if not pretrained:
        return model
    d = os.environ.get("DFU_PRETRAINED_DIR")
    path = os.path.join(d, name + ".pt") if d else None
    if path and os.path.exists(path):
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
    else:
        warnings.warn("pretrained=True for %s: no network / no DFU_PRETRAINED_DIR weights, using random init" % name)
    return model


def create_model(model_name, pretrained=False, num_classes=1000, **kwargs):
    """timm.create_model stand-in (only the model the reference uses)."""
    if model_name != "vit_base_patch16_224":
        raise RuntimeError("dfu_multimodal_b200 timm shim only provides vit_base_patch16_224, not %r" % model_name)
    from .vit import DfuVisionTransformer
    return _maybe_pretrained(DfuVisionTransformer(num_classes=num_classes), model_name, pretrained)


def hub_load(repo_or_dir, model, *args, **kwargs):
    """torch.hub.load stand-in: serves ('pytorch/vision:*', 'resnet50'); anything else goes to the original."""
    if isinstance(repo_or_dir, str) and repo_or_dir.startswith("pytorch/vision") and model == "resnet50":
        from .resnet import DfuResNet50
        pretrained = bool(kwargs.get("pretrained", False) or kwargs.get("weights") is not None)
        return _maybe_pretrained(DfuResNet50(), "resnet50", pretrained)
    if _ORIG_HUB_LOAD is None:
        raise RuntimeError("torch.hub.load(%r, %r) is not available offline" % (repo_or_dir, model))
    return _ORIG_HUB_LOAD(repo_or_dir, model, *args, **kwargs)


def install(force_timm_shim=False):
    """Idempotent. Returns the timm module that `import timm` will now resolve to."""
    global _ORIG_HUB_LOAD
    if _ORIG_HUB_LOAD is None:
        _ORIG_HUB_LOAD = torch.hub.load
        torch.hub.load = hub_load
    have_real = False
    if not force_timm_shim:
        try:
            import importlib.util
            have_real = importlib.util.find_spec("timm") is not None and "timm" not in sys.modules
        except (ImportError, ValueError):
            have_real = False
    if "timm" in sys.modules and getattr(sys.modules["timm"], "__dfu_shim__", False):
        return sys.modules["timm"]
    shim = types.ModuleType("timm")
    shim.__dfu_shim__ = True
    shim.__version__ = "0.9.2+dfu_b200"
    shim.create_model = create_model
    shim.list_models = lambda *a, **k: ["vit_base_patch16_224"]
    if have_real:
        warnings.warn("a real timm is installed; the B200 shim takes precedence while dropin.install() is active")
    sys.modules["timm"] = shim
    return shim


def uninstall():
    global _ORIG_HUB_LOAD
    if _ORIG_HUB_LOAD is not None:
        torch.hub.load = _ORIG_HUB_LOAD
        _ORIG_HUB_LOAD = None
    if "timm" in sys.modules and getattr(sys.modules["timm"], "__dfu_shim__", False):
        del sys.modules["timm"]
