"""ViT-Base/16 in timm's `vit_base_patch16_224` layout (timm 0.9.x VisionTransformer: pre-norm blocks, LayerNorm
eps 1e-6, exact-erf GELU, qkv bias, scale 64^-0.5, cls-token pooling after the final norm, no dropout at defaults)
on libdfu_b200 kernels.  Parameter names equal timm's so `vit.*` / `thermal_branch.*` / `backbone.*` checkpoints
load.  Reference constructor: timm.create_model('vit_base_patch16_224', pretrained=True, num_classes=0|2)
(notebooks/train_multimodal_fusion.py:299-301, extended_metrics.py:324)."""
import torch
import torch.nn as nn

from . import functional as Fn
from .nn import DfuGELU, DfuLayerNorm, DfuLinear, convert_head, has_hooks, to_act
from .ops import BF16


class DfuPatchProj(nn.Module):
    """timm PatchEmbed.proj: Conv2d(3, 768, 16, stride 16) == Linear over flattened (c,i,j) patches."""

    def __init__(self, in_chans=3, embed_dim=768, patch=16):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(embed_dim, in_chans, patch, patch))
        self.bias = nn.Parameter(torch.zeros(embed_dim))
        nn.init.trunc_normal_(self.weight, std=0.02)

    def forward(self, patches):
        return Fn.LinearFn.apply(patches, self.weight, self.bias, None, 0, False)


class DfuPatchEmbed(nn.Module):
    def __init__(self, embed_dim=768):
        super().__init__()
        self.proj = DfuPatchProj(3, embed_dim, 16)
        self.num_patches = 196

    def forward(self, x):
        """fp32 (B,3,224,224) image batch, or an already unfolded bf16 patch matrix [B*196, 768]."""
        if x.dim() == 4:
            x = Fn.Patchify.apply(x.float())
        return self.proj(x)


class DfuAttention(nn.Module):
    def __init__(self, dim=768, num_heads=12):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        assert self.head_dim == 64, "the attention kernels are built for head_dim 64"
        self.scale = self.head_dim ** -0.5
        self.qkv = DfuLinear(dim, dim * 3, std=0.02)
        self.proj = DfuLinear(dim, dim, std=0.02)

    def forward(self, x, res=None):
        B, T, _ = x.shape
        qkv = self.qkv(x)
        a = Fn.AttentionFn.apply(qkv, B, T, self.num_heads, self.scale)
        return self.proj(a, res=res)


class DfuMlp(nn.Module):
    def __init__(self, dim=768, hidden=3072):
        super().__init__()
        self.fc1 = DfuLinear(dim, hidden, std=0.02)
        self.act = DfuGELU()
        self.fc2 = DfuLinear(hidden, dim, std=0.02)

    def forward(self, x, res=None):
        if has_hooks(self.act, self.fc1, self.fc2):
            return self.fc2(self.act(self.fc1(x)), res=res)
        # one autograd node: GELU in the fc1 epilogue, residual in the fc2 epilogue, GELU' in the fc2-dgrad epilogue
        return Fn.MlpFn.apply(to_act(x), self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, res)


class DfuBlock(nn.Module):
    def __init__(self, dim=768, num_heads=12, mlp_ratio=4):
        super().__init__()
        self.norm1 = DfuLayerNorm(dim, eps=1e-6)
        self.attn = DfuAttention(dim, num_heads)
        self.norm2 = DfuLayerNorm(dim, eps=1e-6)
        self.mlp = DfuMlp(dim, dim * mlp_ratio)

    def forward(self, x):
        x = to_act(x)
        fork = torch.is_grad_enabled() and x.requires_grad
        if has_hooks(self.attn, self.mlp, self.norm1, self.norm2):
            # unfused residual adds so hooks on attn / mlp observe the branch output alone (timm Block.forward)
            xa, xs = Fn.Fork.apply(x) if fork else (x, x)
            x = Fn.AddFn.apply(xs, self.attn(self.norm1(xa)))
            xa, xs = Fn.Fork.apply(x) if fork else (x, x)
            return Fn.AddFn.apply(xs, self.mlp(self.norm2(xa)))
        # the skip connection forks inside LayerNormForkFn (its gradient is added by the LayerNorm backward kernel)
        # and rejoins in the proj / fc2 GEMM epilogue
        n1, n2 = self.norm1, self.norm2
        if fork:
            xa, xs = Fn.LayerNormForkFn.apply(x, n1.weight, n1.bias, n1.eps)
        else:
            xa, xs = n1(x), x
        x = self.attn(xa, res=xs)
        if torch.is_grad_enabled() and x.requires_grad:
            xa, xs = Fn.LayerNormForkFn.apply(x, n2.weight, n2.bias, n2.eps)
        else:
            xa, xs = n2(x), x
        return self.mlp(xa, res=xs)


class DfuVisionTransformer(nn.Module):
    def __init__(self, num_classes=1000, embed_dim=768, depth=12, num_heads=12):
        super().__init__()
        self.num_classes = num_classes
        self.embed_dim = self.num_features = embed_dim
        self.patch_embed = DfuPatchEmbed(embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, 197, embed_dim) * 0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        self.pos_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[DfuBlock(embed_dim, num_heads) for _ in range(depth)])
        self.norm = DfuLayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Identity()
        self.head = DfuLinear(embed_dim, num_classes, std=0.02) if num_classes > 0 else nn.Identity()

    def __setattr__(self, name, value):
        # the reference replaces `.head` with Sequential(Dropout, Linear) (train_thermal_only.py:199-202)
        if name == "head" and isinstance(value, nn.Module):
            value = convert_head(value)
        super().__setattr__(name, value)

    def forward_features(self, x):
        """(B,3,224,224) fp32 or patch matrix [B*196,768] bf16 -> normalised token sequence (B,197,768) bf16."""
        B = x.shape[0] if x.dim() == 4 else x.shape[0] // 196
        p = self.patch_embed(x)
        t = Fn.VitAssembleFn.apply(p, self.cls_token, self.pos_embed, B)
        t = self.blocks(t)
        return self.norm(t)

    def forward_head_bf16(self, t):
        f = Fn.ClsTokenFn.apply(t)
        if isinstance(self.head, nn.Identity):
            return f
        if isinstance(self.head, DfuLinear):
            return self.head(f, out_fp32=True)
        return self.head(f)

    def forward_bf16(self, x):
        return self.forward_head_bf16(self.forward_features(x))

    def forward(self, x):
        """nn.Module contract of timm's VisionTransformer: fp32 in, fp32 out ((B,768) features or logits)."""
        y = self.forward_bf16(x)
        return y if y.dtype == torch.float32 else Fn.CastToF32.apply(y)


def vit_base_patch16_224(num_classes=1000, **_):
    return DfuVisionTransformer(num_classes=num_classes)
