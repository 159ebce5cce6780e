"""The late-fusion DFU classifier and its single-modality siblings on libdfu_b200 kernels.

Canonical layout = the reference's eval-side model (notebooks/extended_metrics.py:338-367, README.md:59):
    resnet (ResNet50, fc=Identity) -> 2048  ++  vit (ViT-B/16, num_classes=0) -> 768
    fusion.classifier = Linear(2816,512) -> ReLU -> Dropout -> Linear(512,2)
`head_layers=3` builds the training script's 2816->512->256->2 head (train_multimodal_fusion.py:305-313).
forward(rgb, thermal) accepts the reference contract (fp32 NCHW, already normalised) or raw uint8 NHWC batches of any
size, which are resized to 224x224 (bit-identical to the reference's Resize) and normalised on the device by the a9
kernels (ImageNet mean/std for RGB, 0.5/0.5 for thermal)."""
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .nn import DfuDropout, DfuHead, DfuLinear, DfuReLU
from .resnet import DfuResNet50
from .vit import DfuVisionTransformer


def _prep_rgb(x):
    """uint8 NHWC images of any size: Resize((224,224)) (Pillow-exact, only when the size differs) + ToTensor + Normalize."""
    if x.dtype == torch.uint8:
        if tuple(x.shape[1:3]) != (224, 224):
            return ops.resize_u8(x, (224, 224), "nhwc", "imagenet")
        return ops.preprocess_u8_nhwc(x, "imagenet")
    return x


def _prep_thermal(x):
    if x.dtype == torch.uint8:
        if tuple(x.shape[1:3]) != (224, 224):
            return ops.resize_u8(x, (224, 224), "patches", "half")
        return ops.preprocess_u8_patches(x, "half")
    return x


def make_head(in_dim, num_classes=2, dropout=0.7, head_layers=2):
    if head_layers == 2:
        return DfuHead(DfuLinear(in_dim, 512), DfuReLU(), DfuDropout(dropout), DfuLinear(512, num_classes))
    if head_layers == 3:
        return DfuHead(DfuLinear(in_dim, 512), DfuReLU(), DfuDropout(dropout), DfuLinear(512, 256), DfuReLU(),
                       DfuDropout(dropout), DfuLinear(256, num_classes))
    raise ValueError("head_layers must be 2 or 3")


class MLPFusion(nn.Module):
    """extended_metrics.py:338-350."""

    def __init__(self, rgb_feat_dim=2048, thermal_feat_dim=768, hidden_dim=512, num_classes=2, dropout=0.7,
                 head_layers=2):
        super().__init__()
        assert hidden_dim == 512
        self.classifier = make_head(rgb_feat_dim + thermal_feat_dim, num_classes, dropout, head_layers)

    def forward(self, rgb_feat, thermal_feat):
        fused = Fn.ConcatFn.apply(rgb_feat, thermal_feat)
        return self.classifier(fused)


class MultimodalFusionModel(nn.Module):
    """forward(rgb, thermal) -> (B, num_classes) fp32 logits."""

    def __init__(self, num_classes=2, dropout=0.7, head_layers=2):
        super().__init__()
        self.resnet = DfuResNet50()
        self.resnet.fc = nn.Identity()
        self.vit = DfuVisionTransformer(num_classes=0)
        self.fusion = MLPFusion(2048, 768, 512, num_classes, dropout, head_layers)
        self.side_stream = None
        self.rgb_stream = None

    def features(self, rgb, thermal):
        """Runs the two branches; with `enable_branch_overlap()` they run on their own CUDA streams (forward and,
        through autograd's stream bookkeeping, backward), optionally on complementary SM budgets so the HBM-bound
        ResNet kernels run beside -- not between -- the tensor-core-bound ViT GEMMs."""
        if self.side_stream is None or not rgb.is_cuda:
            rgb, thermal = _prep_rgb(rgb), _prep_thermal(thermal)
            return self.resnet.forward_bf16(rgb), self.vit.forward_bf16(thermal)
        main = torch.cuda.current_stream()
        vs, rs = self.side_stream, self.rgb_stream
        vs.wait_stream(main)
        rs.wait_stream(main)
        with torch.cuda.stream(vs):
            tf = self.vit.forward_bf16(_prep_thermal(thermal))
        with torch.cuda.stream(rs):
            rf = self.resnet.forward_bf16(_prep_rgb(rgb))
        main.wait_stream(vs)
        main.wait_stream(rs)
        tf.record_stream(main)
        rf.record_stream(main)
        return rf, tf

    def enable_branch_overlap(self, enabled=True, vit_sms=0, resnet_sms=0):
        """vit_sms / resnet_sms: SM budgets of the two branch streams (0 = each launch spans the whole device)."""
        if self.side_stream is not None:
            ops.set_stream_budget(self.side_stream, 0)
            ops.set_stream_budget(self.rgb_stream, 0)
        self.side_stream = torch.cuda.Stream() if enabled else None
        self.rgb_stream = torch.cuda.Stream() if enabled else None
        if enabled:
            ops.set_stream_budget(self.side_stream, vit_sms)
            ops.set_stream_budget(self.rgb_stream, resnet_sms)
        return self

    def forward(self, rgb, thermal):
        rf, tf = self.features(rgb, thermal)
        return self.fusion(rf, tf)


class RGBOnlyModel(nn.Module):
    """extended_metrics.py:307-318 / train_rgb_only.py:200-217: ResNet50 with fc = Dropout -> Linear(2048, C)."""

    def __init__(self, num_classes=2, drop_rate=0.5):
        super().__init__()
        self.resnet = DfuResNet50()
        self.resnet.fc = DfuHead(DfuDropout(drop_rate), DfuLinear(2048, num_classes))

    def forward(self, x):
        return self.resnet.forward_bf16(_prep_rgb(x))


class ThermalOnlyModel(nn.Module):
    """extended_metrics.py:321-335 / train_thermal_only.py:188-205: ViT-B/16 with head = Dropout -> Linear(768, C)."""

    def __init__(self, num_classes=2, drop_rate=0.5):
        super().__init__()
        self.vit = DfuVisionTransformer(num_classes=num_classes)
        self.vit.head = DfuHead(DfuDropout(drop_rate), DfuLinear(768, num_classes))

    def forward(self, x):
        return self.vit.forward_bf16(_prep_thermal(x))
