"""torch.autograd glue: one Function per fused op of the hot path.  Forward and backward both run only kernels from
libdfu_b200.so (through ops.py); autograd supplies the graph bookkeeping so `loss.backward()` and
`optimizer.step()` behave exactly as they do for the reference's nn.Modules
(notebooks/train_multimodal_fusion.py:374-380)."""
import os

import torch
from torch.autograd import Function

from . import ops
from .ops import BF16, F32


# Epilogue fusions that are implemented and parity-tested but lose on B200 at the benchmark shapes because the GEMMs
# they attach to are epilogue-bound (small K): measured numbers in profiles/r01e_gemm_bench_isolated.txt.
FUSE_GELU_BWD = os.environ.get("DFU_FUSE_GELU_BWD", "0") == "1"     # GELU' in the fc2-dgrad epilogue (act=3)
# fc1's epilogue stores gelu'(h) instead of h; the fc2-dgrad epilogue multiplies by it (act=5) and sums fc1's bias gradient
GELU_GRAD_AUX = os.environ.get("DFU_GELU_GRAD_AUX", "1") == "1"
# measured on B200 (profiles/r01o): the act=5 fc2-dgrad GEMM takes 126 us alone and 139 us with the column sums kept in
# its epilogue's registers; a separate colsum pass over its output takes ~35 us
FUSE_DGELU_COLSUM = os.environ.get("DFU_FUSE_DGELU_COLSUM", "1") == "1"
FUSE_BN_STATS = os.environ.get("DFU_FUSE_BN_STATS", "1") == "1"     # BatchNorm sums in the conv GEMM epilogue (colstats)
# skip-connection gradient added in the conv1 dgrad epilogue (bias-less residual epilogue: 102 us against 49 + 100 us for
# the GEMM and a separate add on layer1's shapes)
FUSE_SKIP_DGRAD = os.environ.get("DFU_FUSE_SKIP_DGRAD", "1") == "1"


def _round8(n):
    return (n + 7) // 8 * 8


# ------------------------------------------------------------------------------------------ bf16 weight shadows
def shadow(p, rows, cols):
    """bf16 copy of an fp32 parameter viewed as [rows, cols] in storage order, row stride padded to 8.
    Cached on the parameter and refreshed when the parameter's version or storage changes."""
    ldo = _round8(cols)
    key = (p._version, p.data_ptr(), rows, cols)
    cached = getattr(p, "_dfu_shadow", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    if cached is not None and cached[1].numel() == rows * ldo and cached[1].device == p.device:
        buf = cached[1]
    else:
        buf = torch.empty((rows, ldo), dtype=BF16, device=p.device)
    ops.cast_f32_bf16(p.detach(), rows, cols, ldo, out=buf)
    p._dfu_shadow = (key, buf)
    return buf


def mark_shadow_fresh(p):
    """Called by the fused optimizer after it wrote the shadow itself."""
    cached = getattr(p, "_dfu_shadow", None)
    if cached is not None:
        k = cached[0]
        p._dfu_shadow = ((p._version, p.data_ptr(), k[2], k[3]), cached[1])


def conv_weight_matrix(weight):
    """(Cout,Cin,R,S) parameter whose memory is channels_last == row-major [Cout, R*S*Cin]."""
    Cout, Cin, R, S = weight.shape
    if not weight.permute(0, 2, 3, 1).is_contiguous():
        raise RuntimeError("conv weights must be stored channels_last (Cout,R,S,Cin); use DfuConv2d parameters")
    return shadow(weight, Cout, R * S * Cin)


def _as_bf16_2d(g, cols):
    """Upstream gradient -> dense bf16 [rows, cols]."""
    g2 = g.reshape(-1, cols)
    if g2.dtype == F32:
        g2 = ops.cast_f32_bf16(g2.contiguous()).view(-1, cols)
    elif not g2.is_contiguous():
        g2 = g2.contiguous()
    return g2


# ------------------------------------------------------------------------------------------ input adapters
class ToNHWC(Function):
    """fp32 NCHW image batch -> bf16 activations (logical NCHW, NHWC memory)."""

    @staticmethod
    def forward(ctx, x):
        return ops.nchw_f32_to_nhwc_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return ops.nhwc_bf16_to_nchw_f32(g)


class Patchify(Function):
    """fp32 (B,3,224,224) -> bf16 patch matrix [B*196, 768] (timm PatchEmbed unfold order c,i,j)."""

    @staticmethod
    def forward(ctx, x):
        ctx.B = x.shape[0]
        return ops.patchify_f32(x)

    @staticmethod
    def backward(ctx, g):
        return ops.unpatchify_f32(_as_bf16_2d(g, 768), ctx.B)


class CastToF32(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16_f32(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.cast_f32_bf16(g.contiguous()).view(g.shape)


class CastToBF16(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_f32_bf16(x.contiguous()).view(x.shape)

    @staticmethod
    def backward(ctx, g):
        return ops.cast_bf16_f32(g.contiguous())


# ------------------------------------------------------------------------------------------ convolution
class Conv2dFn(Function):
    """Implicit-GEMM convolution on NHWC bf16: 1x1/s1 feeds the activation matrix to the GEMM directly, other
    shapes go through im2col.  Replaces torchvision ResNet's nn.Conv2d (cudnn fwd/dgrad/wgrad)."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad, fork=False, want_stats=False):
        """fork=True additionally returns x itself (the Bottleneck's skip connection); its gradient is then added
        in the dgrad GEMM epilogue instead of by a separate pass.  want_stats=True additionally returns the
        per-channel sum / sum of squares of the output ([2*Cout] fp32), accumulated by the GEMM epilogue from the fp32
        accumulators: the BatchNorm that follows needs no statistics pass of its own."""
        ctx.set_materialize_grads(False)       # no zero-filled "gradient" for the statistics output
        xn = ops.as_nhwc(x)
        B, H, W, C = xn.shape
        Cout, Cin, R, S = weight.shape
        assert Cin == C, "channel mismatch %d vs %d" % (Cin, C)
        Ho = (H + 2 * pad - R) // stride + 1
        Wo = (W + 2 * pad - S) // stride + 1
        Kc = R * S * C
        ldc = _round8(Kc)
        wm = conv_weight_matrix(weight)
        implicit = ops.conv_implicit_ok(C, Cout, R, S, stride)
        if implicit:
            # tiles gathered by TMA straight from the activation tensor: no im2col matrix, x itself is what is saved
            sums = ops.zeros(2 * Cout, F32, xn.device) if (want_stats and FUSE_BN_STATS and Cout <= 2048) else None
            y = ops.conv_forward(xn, wm, R, S, pad, colstats=sums, stride=stride).view(B * Ho * Wo, Cout)
            if want_stats and sums is None:
                sums = ops.bn_stats(y)
            ctx.geom = (B, H, W, C, Cout, R, S, stride, pad, Ho, Wo, Kc, ldc, False)
            ctx.implicit = True
            ctx.save_for_backward(xn, weight)
            ctx.fork = fork
            out = [y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)]
            if fork:
                out.append(x.view_as(x))
            if want_stats:
                ctx.mark_non_differentiable(sums)
                out.append(sums)
            return tuple(out) if len(out) > 1 else out[0]
        ctx.implicit = False
        direct = (R == 1 and S == 1 and stride == 1 and pad == 0 and ldc == Kc)
        A = xn.reshape(B * H * W, C) if direct else ops.im2col(xn, R, S, stride, pad, Ho, Wo, ldc)
        sums = None
        if want_stats and FUSE_BN_STATS and ops.colstats_ok(Cout):
            sums = ops.zeros(2 * Cout, F32, xn.device)
        y, _ = ops.linear_forward(A, wm, colstats=sums, colstats_mode=2)
        if want_stats and sums is None:          # shape the epilogue cannot reduce: separate statistics pass
            sums = ops.bn_stats(y)
        ctx.geom = (B, H, W, C, Cout, R, S, stride, pad, Ho, Wo, Kc, ldc, direct)
        ctx.save_for_backward(A, weight)
        ctx.fork = fork
        out = [y.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)]
        if fork:
            out.append(x.view_as(x))
        if want_stats:
            ctx.mark_non_differentiable(sums)
            out.append(sums)
        return tuple(out) if len(out) > 1 else out[0]

    @staticmethod
    def backward(ctx, gy, *rest):
        gskip = rest[0] if ctx.fork else None
        A, weight = ctx.saved_tensors
        B, H, W, C, Cout, R, S, stride, pad, Ho, Wo, Kc, ldc, direct = ctx.geom
        if ctx.implicit:
            gn = ops.as_nhwc(gy)
            dw = dx = None
            if ctx.needs_input_grad[1]:
                dw = ops.conv_wgrad(gn, A, R, S, pad, stride).reshape(Cout, R, S, C).permute(0, 3, 1, 2)
            if ctx.needs_input_grad[0]:
                dx = ops.conv_dgrad(gn, conv_weight_matrix(weight), H, W, C, R, S, pad, stride)
                if gskip is not None:
                    dx = ops.add(dx, ops.as_nhwc(gskip))
                dx = dx.permute(0, 3, 1, 2)
            elif gskip is not None:
                dx = gskip
            return dx, dw, None, None, None, None
        g = ops.as_nhwc(gy).reshape(B * Ho * Wo, Cout)
        dw = None
        if ctx.needs_input_grad[1]:
            dw = ops.linear_wgrad(g, A)                       # [Cout, ldc] fp32
            if ldc != Kc:
                dw = dw[:, :Kc]
            dw = dw.reshape(Cout, R, S, C).permute(0, 3, 1, 2)
        dx = None
        if ctx.needs_input_grad[0]:
            wm = conv_weight_matrix(weight)
            skip = ops.as_nhwc(gskip).reshape(B * H * W, C) if gskip is not None else None
            if direct and FUSE_SKIP_DGRAD:
                dx = ops.linear_dgrad(g, wm, ldc, res=skip).view(B, H, W, C)
            elif direct:
                dx = ops.linear_dgrad(g, wm, ldc).view(B, H, W, C)
                if skip is not None:
                    dx = ops.add(dx, skip.view(B, H, W, C))
            else:
                dx = ops.col2im(ops.linear_dgrad(g, wm, ldc), B, H, W, C, R, S, stride, pad, Ho, Wo, ldc)
                if skip is not None:
                    dx = ops.add(dx, skip.view(B, H, W, C))
            dx = dx.permute(0, 3, 1, 2)
        elif gskip is not None:
            dx = gskip
        return dx, dw, None, None, None, None


class BatchNormActFn(Function):
    """BatchNorm2d (+ residual) (+ ReLU) in two passes over the conv output (stats, apply).  train(): batch
    statistics + running-stat update (momentum 0.1, unbiased variance) like torch; eval(): running statistics."""

    @staticmethod
    def forward(ctx, y, gamma, beta, res, running_mean, running_var, nbt, relu, training, eps, momentum, sums=None):
        yn = ops.as_nhwc(y)
        B, H, W, C = yn.shape
        y2 = yn.reshape(B * H * W, C)
        res2 = ops.as_nhwc(res).reshape(B * H * W, C) if res is not None else None
        need_mask = relu and res is not None and any(ctx.needs_input_grad[:4])
        out = ops.bn_forward(y2, gamma, beta, running_mean, running_var, nbt, res2, relu, training, eps, momentum,
                             sums=sums if training else None, want_mask=need_mask)
        z, mi = out[0], out[1]
        ctx.cfg = (B, H, W, C, relu, training, res is not None)
        # ReLU mask for the backward kernels: recomputed from y when there is no residual input, otherwise the 1-bit-
        # per-element mask the apply kernel wrote (z itself is never re-read)
        ctx.save_for_backward(y2, out[2] if need_mask else None, mi, gamma, beta)
        return z.view(B, H, W, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gz):
        y2, zmask, mi, gamma, beta = ctx.saved_tensors
        B, H, W, C, relu, training, has_res = ctx.cfg
        g = ops.as_nhwc(gz).reshape(B * H * W, C)
        dy, dres, dg, db = ops.bn_backward(g, None, y2, mi, gamma, relu, training, has_res and ctx.needs_input_grad[3],
                                           beta=beta, zmask=zmask)
        dy = dy.view(B, H, W, C).permute(0, 3, 1, 2)
        if dres is not None:
            dres = dres.view(B, H, W, C).permute(0, 3, 1, 2)
        return dy, dg, db, dres, None, None, None, None, None, None, None, None


class ReLUFn(Function):
    @staticmethod
    def forward(ctx, x):
        xn = ops.as_nhwc(x) if x.dim() == 4 else x.contiguous()
        y = ops.relu(xn)
        ctx.save_for_backward(y)
        ctx.is4 = x.dim() == 4
        return y.permute(0, 3, 1, 2) if ctx.is4 else y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        gn = ops.as_nhwc(g) if ctx.is4 else _as_bf16_2d(g, y.shape[-1]).view(y.shape)
        dx = ops.relu_bwd(gn, y)
        return dx.permute(0, 3, 1, 2) if ctx.is4 else dx


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        if a.dim() == 4:
            return ops.add(ops.as_nhwc(a), ops.as_nhwc(b)).permute(0, 3, 1, 2)
        return ops.add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


class GeluFn(Function):
    """Standalone exact-erf GELU (only used when hooks force the unfused ViT MLP path)."""

    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous()
        ctx.save_for_backward(xc)
        return ops.gelu(xc)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.gelu_bwd(_as_bf16_2d(g, x.shape[-1]).view(x.shape), x)


class Fork(Function):
    """y1 = y2 = x.  Its backward sums the two incoming gradients with our add kernel, so autograd never has to
    accumulate (which would launch an ATen add) when an activation feeds both a branch and a skip connection."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None:
            return g2
        if g2 is None:
            return g1
        if g1.dim() == 4:
            return ops.add(ops.as_nhwc(g1), ops.as_nhwc(g2)).permute(0, 3, 1, 2)
        c = g1.shape[-1]
        return ops.add(_as_bf16_2d(g1, c), _as_bf16_2d(g2, c)).view(g1.shape)


class MaxPoolFn(Function):
    @staticmethod
    def forward(ctx, x):
        xn = ops.as_nhwc(x)
        y, idx = ops.maxpool(xn, ctx.needs_input_grad[0])
        ctx.hw = (xn.shape[1], xn.shape[2])
        ctx.save_for_backward(idx)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return ops.maxpool_bwd(idx, ops.as_nhwc(g), ctx.hw[0], ctx.hw[1]).permute(0, 3, 1, 2)


class AvgPoolFn(Function):
    """AdaptiveAvgPool2d(1) + flatten: (B,C,H,W) -> (B,C)."""

    @staticmethod
    def forward(ctx, x):
        xn = ops.as_nhwc(x)
        ctx.hw = (xn.shape[1], xn.shape[2])
        return ops.avgpool(xn)

    @staticmethod
    def backward(ctx, g):
        g2 = _as_bf16_2d(g, g.shape[-1])
        return ops.avgpool_bwd(g2, ctx.hw[0], ctx.hw[1]).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------ linear / ViT
class LinearFn(Function):
    """y = act(x W^T + b) (+ res).  act: 0 none, 1 relu, 2 gelu(erf).  Output bf16, or fp32 for logits."""

    @staticmethod
    def forward(ctx, x, weight, bias, res, act, out_fp32):
        N = weight.shape[0]
        K = weight.numel() // N          # patch-embed weight is (768,3,16,16): a [768,768] matrix in memory
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        wm = shadow(weight, N, K)
        res2 = res.reshape(-1, N) if res is not None else None
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]   # (grad mode is off inside forward)
        y, aux = ops.linear_forward(x2, wm, bias, act, res2, want_aux=(act == 2 and need_grad),
                                    out_dtype=F32 if out_fp32 else BF16)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.has_res = res is not None
        ctx.x_shape = x.shape
        saved_act = aux if act == 2 else (y if act == 1 else None)
        ctx.save_for_backward(x2, weight, saved_act)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        x2, weight, saved_act = ctx.saved_tensors
        N = weight.shape[0]
        K = weight.numel() // N
        g2 = _as_bf16_2d(gy, N)
        dres = gy if ctx.has_res else None
        if ctx.act == 2:
            g2 = ops.gelu_bwd(g2, saved_act)
        elif ctx.act == 1:
            g2 = ops.relu_bwd(g2, saved_act)
        dw = ops.linear_wgrad(g2, x2, _round8(K)) if ctx.needs_input_grad[1] else None
        if dw is not None:
            if dw.shape[1] != K:
                dw = dw[:, :K]
            dw = dw.reshape(weight.shape)
        db = ops.colsum(g2) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear_dgrad(g2, shadow(weight, N, K), K).view(ctx.x_shape)
        return dx, dw, db, dres, None, None


class MlpFn(Function):
    """timm Mlp (fc1 -> GELU -> fc2) plus the residual add, as one autograd node.  Forward: GELU and the pre-activation
    derivative (act=4: aux = gelu'(h), bf16) come out of the fc1 GEMM epilogue, the residual goes into the fc2 epilogue.
    Backward: the fc2 dgrad epilogue multiplies by the stored derivative (act=5) and accumulates fc1's bias gradient
    (colstats), so no elementwise pass runs between the five GEMMs."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, res):
        Hd, D = w1.shape
        x2 = x.reshape(-1, D)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        need_grad = any(ctx.needs_input_grad[:5])
        ctx.grad_aux = GELU_GRAD_AUX and need_grad
        a, h = ops.linear_forward(x2, shadow(w1, Hd, D), b1, act=4 if ctx.grad_aux else 2, want_aux=need_grad)
        res2 = res.reshape(-1, D) if res is not None else None
        y, _ = ops.linear_forward(a, shadow(w2, D, Hd), b2, 0, res2)
        ctx.has_res = res is not None
        ctx.save_for_backward(x2, w1, w2, h, a)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2, w1, w2, h, a = ctx.saved_tensors
        Hd, D = w1.shape
        g2 = _as_bf16_2d(gy, D)
        dw2 = ops.linear_wgrad(g2, a, Hd).reshape(w2.shape) if ctx.needs_input_grad[3] else None
        db2 = ops.colsum(g2) if ctx.needs_input_grad[4] else None
        if ctx.grad_aux or FUSE_GELU_BWD:
            # h holds gelu'(pre-activation) when the forward ran with act=4, the pre-activation itself otherwise
            db1 = ops.zeros(Hd, F32, g2.device) if (ctx.needs_input_grad[2] and ops.colstats_ok(Hd, mode=1) and
                                                    (FUSE_DGELU_COLSUM or not ctx.grad_aux)) else None
            dh = ops.linear_dgrad(g2, shadow(w2, D, Hd), Hd, res=h, act=5 if ctx.grad_aux else 3, colstats=db1)
            if db1 is None and ctx.needs_input_grad[2]:
                db1 = ops.colsum(dh)
        else:
            # measured on B200 (profiles/r01e): the 3072-wide, K=768 dgrad GEMM is epilogue-bound, so GELU' inside
            # its epilogue (276 us) loses to the plain GEMM (118 us) + one elementwise pass, which also produces
            # fc1's bias gradient (column sums of its output) on the way
            dh, db1 = ops.gelu_bwd_colsum(ops.linear_dgrad(g2, shadow(w2, D, Hd), Hd), h)
        dw1 = ops.linear_wgrad(dh, x2, D).reshape(w1.shape) if ctx.needs_input_grad[1] else None
        dx = ops.linear_dgrad(dh, shadow(w1, Hd, D), D).view(gy.shape) if ctx.needs_input_grad[0] else None
        return dx, dw1, db1, dw2, db2, (gy if ctx.has_res else None)


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        y, stats = ops.layernorm(x2, gamma, beta, eps)
        ctx.save_for_backward(x2, gamma, stats)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, gy):
        x2, gamma, stats = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(_as_bf16_2d(gy, x2.shape[1]), x2, gamma, stats)
        return dx.view(gy.shape), dg, db, None


class LayerNormForkFn(Function):
    """(LayerNorm(x), x): the pre-norm residual block's fork.  Backward folds the skip-connection gradient into the
    LayerNorm backward kernel (dx = LN'(g_norm) + g_skip), so no separate add pass runs."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        D = x.shape[-1]
        x2 = x.reshape(-1, D)
        y, stats = ops.layernorm(x2, gamma, beta, eps)
        ctx.save_for_backward(x2, gamma, stats)
        return y.view(x.shape), x.view_as(x)

    @staticmethod
    def backward(ctx, gy, gskip):
        x2, gamma, stats = ctx.saved_tensors
        D = x2.shape[1]
        if gy is None:
            return gskip, None, None, None
        dres = _as_bf16_2d(gskip, D) if gskip is not None else None
        dx, dg, db = ops.layernorm_bwd(_as_bf16_2d(gy, D), x2, gamma, stats, dres)
        return dx.view(gy.shape), dg, db, None


class AttentionFn(Function):
    """softmax(q k^T / 8) v over 197 tokens per (batch, head) on the packed qkv projection (timm Attention)."""

    @staticmethod
    def forward(ctx, qkv, B, T, H, scale):
        q2 = qkv.reshape(B * T, -1)
        out, saved = ops.attention_forward(q2, B, T, H, scale)
        ctx.cfg = (B, T, H, scale)
        ctx.save_for_backward(q2, saved, out)
        return out.view(B, T, H * 64)

    @staticmethod
    def backward(ctx, g):
        q2, saved, out = ctx.saved_tensors
        B, T, H, scale = ctx.cfg
        dqkv = ops.attention_backward(q2, saved, _as_bf16_2d(g, H * 64), B, T, H, scale, out=out)
        return dqkv.view(B, T, -1), None, None, None, None


class VitAssembleFn(Function):
    """cat(cls_token, patch tokens) + pos_embed (timm VisionTransformer._pos_embed)."""

    @staticmethod
    def forward(ctx, patch_out, cls_token, pos_embed, B):
        T = pos_embed.shape[1]
        D = pos_embed.shape[2]
        ctx.cfg = (B, T, D)
        return ops.vit_assemble(patch_out, cls_token.detach(), pos_embed.detach(), B, T, D).view(B, T, D)

    @staticmethod
    def backward(ctx, g):
        B, T, D = ctx.cfg
        dpatch, dcls, dpos = ops.vit_assemble_bwd(_as_bf16_2d(g, D), B, T, D)
        return dpatch, dcls.view(1, 1, D), dpos.view(1, T, D), None


class ClsTokenFn(Function):
    """x[:, 0] of (B,T,D) -> (B,D)."""

    @staticmethod
    def forward(ctx, x):
        B, T, D = x.shape
        out = torch.empty((B, D), dtype=BF16, device=x.device)
        ops.copy2d(x, T * D, out, D, B, D)
        ctx.cfg = (B, T, D)
        return out

    @staticmethod
    def backward(ctx, g):
        B, T, D = ctx.cfg
        dx = ops.zeros((B, T, D), BF16, g.device)
        ops.copy2d(_as_bf16_2d(g, D), D, dx, T * D, B, D)
        return dx


class ConcatFn(Function):
    """cat([a, b], dim=1) for 2-D bf16 feature matrices (train_multimodal_fusion.py:321)."""

    @staticmethod
    def forward(ctx, a, b):
        B, Ca = a.shape
        Cb = b.shape[1]
        out = torch.empty((B, Ca + Cb), dtype=BF16, device=a.device)
        ops.copy2d(a.contiguous(), Ca, out, Ca + Cb, B, Ca)
        ops.copy2d(b.contiguous(), Cb, out, Ca + Cb, B, Cb, dst_off=Ca)
        ctx.cfg = (B, Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, g):
        B, Ca, Cb = ctx.cfg
        g2 = _as_bf16_2d(g, Ca + Cb)
        da = torch.empty((B, Ca), dtype=BF16, device=g.device)
        db = torch.empty((B, Cb), dtype=BF16, device=g.device)
        ops.copy2d(g2, Ca + Cb, da, Ca, B, Ca)
        ops.copy2d(g2, Ca + Cb, db, Cb, B, Cb, src_off=Ca)
        return da, db


_DROPOUT_SITE = [0]


class DropoutFn(Function):
    """nn.Dropout in train(): every call draws a fresh mask.  The device-resident counter is advanced by a 1-thread
    kernel right before the mask kernel of the same call (same stream), so eager loops written like the reference's
    (train_multimodal_fusion.py:374-380) and CUDA-graph replays both see a new mask per call."""

    @staticmethod
    def forward(ctx, x, p, site):
        ops.DROPOUT.tick(x.device)
        y, mask = ops.dropout(x.contiguous(), p, site)
        ctx.p = p
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return ops.dropout_bwd(_as_bf16_2d(g, g.shape[-1]).view(g.shape), mask, ctx.p), None, None


class CrossEntropyFn(Function):
    """nn.CrossEntropyLoss(weight=w), mean reduction (train_multimodal_fusion.py:346,376)."""

    @staticmethod
    def forward(ctx, logits, labels, weight):
        loss, dlogits = ops.cross_entropy(logits, labels, weight, want_grad=True)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return ops.scale_by_scalar(dlogits, g.contiguous()), None, None


def cross_entropy(logits, labels, weight=None):
    return CrossEntropyFn.apply(logits, labels, weight)
