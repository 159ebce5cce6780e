"""nn.Module surface of the late-fusion classifier: forward(rgb, thermal) -> logits, reference state_dict keys
(`resnet.*`, `vit.*`, `fusion.classifier.{0,3}.*`; extended_metrics.py:353-367)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfu_multimodal_b200.fusion import MLPFusion, MultimodalFusionModel  # noqa: E402,F401
from dfu_multimodal_b200.checkpoint import build_fusion_from_checkpoint, load_into, remap_state_dict  # noqa: E402,F401
