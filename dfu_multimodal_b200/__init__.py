"""dfu_multimodal_b200 -- the late-fusion DFU classifier step (ResNet50 + ViT-B/16 -> MLP head) on hand-written
sm_100a kernels (libdfu_b200.so, C ABI in include/dfu_b200.h).  PyTorch is the host: device memory, streams,
autograd bookkeeping and torch.distributed; every kernel on the path is ours.  There is no CPU fallback."""
from . import _lib  # noqa: F401
from .fusion import MLPFusion, MultimodalFusionModel, RGBOnlyModel, ThermalOnlyModel  # noqa: F401
from .resnet import DfuResNet50, resnet50  # noqa: F401
from .vit import DfuVisionTransformer, vit_base_patch16_224  # noqa: F401

__all__ = ["MultimodalFusionModel", "MLPFusion", "RGBOnlyModel", "ThermalOnlyModel", "DfuResNet50",
           "DfuVisionTransformer", "resnet50", "vit_base_patch16_224"]
