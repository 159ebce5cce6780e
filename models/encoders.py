"""Branch encoders: ResNet50 (RGB, torchvision key names) and ViT-B/16 (thermal, timm key names), plus the
single-modality classifiers of train_rgb_only.py:200-217 / train_thermal_only.py:188-205."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfu_multimodal_b200.resnet import DfuResNet50, resnet50  # noqa: E402,F401
from dfu_multimodal_b200.vit import DfuVisionTransformer, vit_base_patch16_224  # noqa: E402,F401
from dfu_multimodal_b200.fusion import RGBOnlyModel, ThermalOnlyModel  # noqa: E402,F401
