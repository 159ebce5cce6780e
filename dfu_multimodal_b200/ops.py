"""Tensor-level wrappers over the C ABI (include/dfu_b200.h).  Every function here launches hand-written sm_100a
kernels from libdfu_b200.so on torch's current stream; torch only provides device memory and streams.
Nothing in this file has a CPU or ATen fallback: a non-CUDA tensor raises."""
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvArgs, GemmArgs, check

BF16 = torch.bfloat16
F32 = torch.float32

# GEMM implementation selector: 0 = tcgen05/TMEM/TMA kernel, 1 = SIMT fp32-accumulate check kernel.
_GEMM_IMPL = int(os.environ.get("DFU_GEMM_IMPL", "0"))
# number of kernels launched through this module (bench.py reports it as gpu_launches)
LAUNCHES = 0
# when set to a list, gemm() appends (start, end) CUDA events for every launch (bench.py roofline measurement)
GEMM_EVENTS = None


def set_gemm_impl(impl):
    global _GEMM_IMPL
    _GEMM_IMPL = int(impl)


def gemm_impl():
    return _GEMM_IMPL


def _L():
    return _lib.load()


# stream handle -> SM budget (see dfu_set_sm_budget); empty = every launch sizes itself for the whole device
_STREAM_BUDGET = {}
_budget_now = [0]


def set_stream_budget(stream, n_sms):
    """Kernels launched on `stream` size their grids for n_sms SMs (0 / None removes the limit)."""
    key = stream.cuda_stream
    if n_sms:
        _STREAM_BUDGET[key] = int(n_sms)
    else:
        _STREAM_BUDGET.pop(key, None)


def _st():
    s = torch.cuda.current_stream().cuda_stream
    if _STREAM_BUDGET or _budget_now[0]:
        want = _STREAM_BUDGET.get(s, 0)
        if want != _budget_now[0]:
            _L().dfu_set_sm_budget(want)
            _budget_now[0] = want
    return s


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dfu_multimodal_b200 ops need CUDA tensors (no CPU path); got %s" % t.device)


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _p(t, elem_offset=0):
    if t is None:
        return None
    return t.data_ptr() + elem_offset * t.element_size()


# ------------------------------------------------------------------------------------------ zeroed accumulators
class _ZeroArena:
    """Accumulators that must start at zero (BN/LN partial sums, bias gradients, split-K wgrad targets) are carved
    out of one arena that a single dfu_zero launch clears at the start of a train step, instead of one fill kernel
    per buffer (~300 per step).  Only active between begin_step() calls made by train.TrainStep, whose step frees
    every gradient (zero_grad(set_to_none=True)) before the arena is cleared again; outside of it zeros() clears
    each buffer with its own dfu_zero launch."""

    ALIGN = 256

    def __init__(self):
        self.buf = None
        self.off = 0
        self.need = 0
        self.active = False
        self.retired = []       # outgrown buffers stay allocated: a captured CUDA graph may still clear and use them

    def begin_step(self, device):
        want = max(self.need, self.off)
        if self.buf is None or self.buf.numel() < want or self.buf.device != device:
            if self.buf is not None:
                self.retired.append(self.buf)
            self.buf = torch.empty(int(want * 1.25) + (1 << 20), dtype=torch.uint8, device=device)
        self.need = want
        used = min(self.buf.numel(), max(want, 1))
        check(_L().dfu_zero(_p(self.buf), used, _st()), "dfu_zero")
        _count()
        self.cleared = used
        self.off = 0
        self.active = True

    def end_step(self):
        self.active = False

    def take(self, nbytes):
        start = (self.off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        end = start + nbytes
        self.off = end
        if self.buf is None or end > self.cleared:
            return None                      # arena too small this step: grows at the next begin_step
        return self.buf[start:end]


ZERO_ARENA = _ZeroArena()


def zeros(shape, dtype, device):
    """Zero-initialised accumulator (see _ZeroArena)."""
    if isinstance(shape, int):
        shape = (shape,)
    n = 1
    for d in shape:
        n *= d
    esz = torch.empty((), dtype=dtype).element_size()
    if ZERO_ARENA.active and ZERO_ARENA.buf is not None and ZERO_ARENA.buf.device == device:
        raw = ZERO_ARENA.take(n * esz)
        if raw is not None:
            return raw.view(dtype).view(shape)
    out = torch.empty(shape, dtype=dtype, device=device)
    if n:
        check(_L().dfu_zero(_p(out), n * esz, _st()), "dfu_zero")
        _count()
    return out


# ------------------------------------------------------------------------------------------ GEMM
def gemm(A, B, D, *, M, N, K, lda, ldb, ldd, a_major=0, b_major=0, a_off=0, b_off=0, d_off=0, nb0=1, nb1=1,
         sa=(0, 0), sb=(0, 0), sd=(0, 0), accumulate=False, aux=None, bias=None, res=None, ldr=0, sr=(0, 0),
         r_off=0, alpha=1.0, act=0, impl=None, max_ctas=0, colstats=None, colstats_mode=0):
    """D[z][m][n] = act(alpha * A.B^T + bias) + res; see dfu_gemm in include/dfu_b200.h. Offsets/strides in elements."""
    _need_cuda(A, B, D, aux, bias, res)
    assert A.dtype == BF16 and B.dtype == BF16 and D.dtype in (BF16, F32)
    a = GemmArgs()
    a.M, a.N, a.K, a.nb0, a.nb1 = M, N, K, nb0, nb1
    a.A, a.a_major, a.lda, a.sa0, a.sa1 = _p(A, a_off), a_major, lda, sa[0], sa[1]
    a.B, a.b_major, a.ldb, a.sb0, a.sb1 = _p(B, b_off), b_major, ldb, sb[0], sb[1]
    a.D, a.d_dtype, a.ldd, a.sd0, a.sd1 = _p(D, d_off), 1 if D.dtype == F32 else 0, ldd, sd[0], sd[1]
    a.accumulate = 1 if accumulate else 0
    a.aux = _p(aux, d_off)
    a.bias = _p(bias)
    a.res, a.ldr, a.sr0, a.sr1 = _p(res, r_off), ldr, sr[0], sr[1]
    a.alpha = alpha
    a.act = act
    a.impl = _GEMM_IMPL if impl is None else impl
    a.max_ctas = max_ctas
    a.colstats = _p(colstats)
    a.colstats_mode = colstats_mode if colstats is not None else 0
    if GEMM_EVENTS is not None:
        # bench.py's roofline leg: CUDA events around each GEMM launch, on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_L().dfu_gemm(ctypes.byref(a), _st()), "dfu_gemm")
        e1.record()
        GEMM_EVENTS.append((e0, e1, (M, N, K, nb0 * nb1, a_major, b_major, "f32" if D.dtype == F32 else "bf16",
                                     int(bool(accumulate)), int(bias is not None), act, int(res is not None),
                                     int(aux is not None), "gemm")))
    else:
        check(_L().dfu_gemm(ctypes.byref(a), _st()), "dfu_gemm")
    _count()
    return D


def _wgrad_needs_splitk(M, N):
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    return tiles < 148


def colstats_ok(N, out_dtype=BF16, mode=2):
    """The GEMM epilogue can accumulate per-column sums when the output rows are 16-byte aligned and the CTA-local
    accumulators (N x mode floats) fit next to the operand ring (<= 4096)."""
    return N * mode <= 4096 and (N * (2 if out_dtype == BF16 else 4)) % 16 == 0 and N >= 16


def linear_forward(x2d, w, bias=None, act=0, res=None, want_aux=False, out_dtype=BF16, colstats=None, colstats_mode=0):
    """y = act(x2d @ w.T + bias) + res.  x2d [M,K] bf16 (row stride = K), w [N,Kp>=K] bf16 (zero padded).
    colstats ([N] or [2N] fp32, zeroed) += column sums (and sums of squares, mode 2) of y."""
    M, K = x2d.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=out_dtype, device=x2d.device)
    aux = torch.empty((M, N), dtype=BF16, device=x2d.device) if want_aux else None
    gemm(x2d, w, y, M=M, N=N, K=K, lda=x2d.stride(0), ldb=w.stride(0), ldd=N, bias=bias, act=act, res=res,
         ldr=(res.stride(0) if res is not None else 0), aux=aux, colstats=colstats, colstats_mode=colstats_mode)
    return y, aux


def linear_dgrad(dy2d, w, K, res=None, act=0, colstats=None):
    """dx[M,K] = dy2d[M,N] @ w[N,:K] (+ res)  (w row-major [N, ldw]: MN-major B operand).
    act=3: dx = (dy2d @ w) * gelu'(res) -- the GELU backward fused into the dgrad epilogue (res = pre-activation);
    act=5: dx = (dy2d @ w) * res -- same with the derivative already stored by the forward GEMM (its act=4 aux output);
    colstats ([K] fp32, zeroed) += column sums of dx (the bias gradient of the layer below)."""
    M, N = dy2d.shape
    dx = torch.empty((M, K), dtype=BF16, device=dy2d.device)
    gemm(dy2d, w, dx, M=M, N=K, K=N, lda=dy2d.stride(0), ldb=w.stride(0), b_major=1, ldd=K, res=res,
         ldr=(res.stride(0) if res is not None else 0), act=act, colstats=colstats, colstats_mode=1 if colstats is not None else 0)
    return dx


def linear_wgrad(dy2d, x2d, Kp=None):
    """dW[N,Kp] (fp32) = dy2d[M,N].T @ x2d[M,Kp]  (both operands MN-major; split-K when the tile count is small)."""
    M, N = dy2d.shape
    Kx = x2d.shape[1] if Kp is None else Kp
    if _wgrad_needs_splitk(N, Kx):
        dw = zeros((N, Kx), F32, dy2d.device)
        acc = True
    else:
        dw = torch.empty((N, Kx), dtype=F32, device=dy2d.device)
        acc = False
    gemm(dy2d, x2d, dw, M=N, N=Kx, K=M, a_major=1, b_major=1, lda=dy2d.stride(0), ldb=x2d.stride(0), ldd=Kx,
         accumulate=acc)
    return dw


# ------------------------------------------------------------------------------------------ implicit-GEMM convolution
_CONV_IMPL = os.environ.get("DFU_CONV_IMPL", "implicit")       # "implicit" (TMA-gathered tiles) | "im2col"


def set_conv_impl(impl):
    global _CONV_IMPL
    assert impl in ("implicit", "im2col")
    _CONV_IMPL = impl


def conv_implicit_ok(C, Cout, R, S, stride):
    """Shapes dfu_conv_gemm takes: channel counts multiples of 64; stride 1 with a real spatial filter (1x1/s1 is a plain
    GEMM on the activation matrix) or stride 2 (forward and wgrad gather strided pixel boxes; dgrad runs per parity class)."""
    return _CONV_IMPL == "implicit" and C % 64 == 0 and Cout % 64 == 0 and ((stride == 1 and R * S > 1) or stride == 2)


def _conv_call(mode, B, H, W, C, Cout, R, S, pad, x=None, w=None, dy=None, out=None, ldo=0, accumulate=False, colstats=None,
               stride=1, bias=None, act=0):
    a = ConvArgs()
    a.mode, a.B, a.H, a.W, a.C = mode, B, H, W, C
    a.Cout, a.R, a.S, a.stride, a.pad = Cout, R, S, stride, pad
    a.x, a.w, a.ldw, a.dy = _p(x), _p(w), (w.stride(0) if w is not None else 0), _p(dy)
    a.out, a.ldo, a.accumulate, a.max_ctas = _p(out), ldo, 1 if accumulate else 0, 0
    a.colstats = _p(colstats)
    a.bias, a.act = _p(bias), act
    st = _st()
    if GEMM_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_L().dfu_conv_gemm(ctypes.byref(a), st), "dfu_conv_gemm")
        e1.record()
        GEMM_EVENTS.append((e0, e1, (B * H * W, Cout, R * S * C, 1, 0, 0, "bf16", 0, 0, 0, 0, 0, "conv%d" % mode)))
    else:
        check(_L().dfu_conv_gemm(ctypes.byref(a), st), "dfu_conv_gemm")
    _count()
    return out


def conv_forward(x_nhwc, wm, R, S, pad, colstats=None, stride=1, bias=None, act=0):
    """y [B,Ho,Wo,Cout] = act(conv(x [B,H,W,C], wm [Cout, >= R*S*C] bf16, columns ordered (r,s,c)) + bias), stride 1 or 2.
    colstats ([2*Cout] fp32, zeroed) += per-channel sum / sum of squares of y (BatchNorm batch statistics).
    bias ([Cout] fp32) / act (0 none, 1 ReLU): the eval-mode BatchNorm folded into the convolution."""
    _need_cuda(x_nhwc, wm)
    B, H, W, C = x_nhwc.shape
    Cout = wm.shape[0]
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    y = torch.empty((B, Ho, Wo, Cout), dtype=BF16, device=x_nhwc.device)
    return _conv_call(1, B, H, W, C, Cout, R, S, pad, x=x_nhwc, w=wm, out=y, colstats=colstats, stride=stride, bias=bias, act=act)


def conv_dgrad(dy_nhwc, wm, H, W, C, R, S, pad, stride=1):
    """dx [B,H,W,C].  Stride 2: one implicit GEMM per parity class of the dx pixels (each sees a fixed subset of the taps)."""
    B, Ho, Wo, Cout = dy_nhwc.shape
    dx = torch.empty((B, H, W, C), dtype=BF16, device=dy_nhwc.device)
    return _conv_call(2, B, H, W, C, Cout, R, S, pad, w=wm, dy=dy_nhwc, out=dx, stride=stride)


def conv_wgrad(dy_nhwc, x_nhwc, R, S, pad, stride=1):
    """dW [Cout, R*S*C] fp32 (columns ordered (r,s,c))."""
    B, H, W, C = x_nhwc.shape
    Cout = dy_nhwc.shape[3]
    N = R * S * C
    split = _wgrad_needs_splitk(Cout, N)
    dw = zeros((Cout, N), F32, x_nhwc.device) if split else torch.empty((Cout, N), dtype=F32, device=x_nhwc.device)
    return _conv_call(3, B, H, W, C, Cout, R, S, pad, x=x_nhwc, dy=dy_nhwc, out=dw, ldo=N, accumulate=split, stride=stride)


# ------------------------------------------------------------------------------------------ layout
_IMAGENET_MEAN = (ctypes.c_float * 3)(0.485, 0.456, 0.406)
_IMAGENET_STD = (ctypes.c_float * 3)(0.229, 0.224, 0.225)
_HALF_MEAN = (ctypes.c_float * 3)(0.5, 0.5, 0.5)
_HALF_STD = (ctypes.c_float * 3)(0.5, 0.5, 0.5)


def _norm_consts(kind):
    if kind == "imagenet":
        return _IMAGENET_MEAN, _IMAGENET_STD
    if kind == "half":
        return _HALF_MEAN, _HALF_STD
    mean, std = kind
    return (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)


def preprocess_u8_nhwc(img_u8, kind="imagenet"):
    """uint8 [B,H,W,3] -> normalised bf16, logical shape (B,3,H,W) in NHWC memory (reference transforms, a9)."""
    _need_cuda(img_u8)
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and img_u8.shape[-1] == 3
    B, H, W, _ = img_u8.shape
    out = torch.empty((B, H, W, 3), dtype=BF16, device=img_u8.device)
    mean, std = _norm_consts(kind)
    check(_L().dfu_preprocess_u8_nhwc(_p(img_u8), _p(out), B * H * W, mean, std, _st()), "preprocess_u8_nhwc")
    _count()
    return out.permute(0, 3, 1, 2)


def preprocess_u8_patches(img_u8, kind="half"):
    """uint8 [B,224,224,3] -> ViT patch matrix bf16 [B*196, 768] (normalise + PatchEmbed unfold)."""
    _need_cuda(img_u8)
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and tuple(img_u8.shape[1:]) == (224, 224, 3)
    B = img_u8.shape[0]
    out = torch.empty((B * 196, 768), dtype=BF16, device=img_u8.device)
    mean, std = _norm_consts(kind)
    check(_L().dfu_preprocess_u8_patches(_p(img_u8), _p(out), B, mean, std, _st()), "preprocess_u8_patches")
    _count()
    return out


def resize_u8(img_u8, out_hw=(224, 224), mode="u8", kind="imagenet", out=None):
    """transforms.Resize(out_hw) of a uint8 [B,Hs,Ws,3] batch, bit-identical to torchvision / Pillow BILINEAR (a9).
    mode "u8": the resized uint8 images [B,OH,OW,3]; "nhwc": + ToTensor + Normalize(kind) as bf16 (logical (B,3,OH,OW), NHWC
    memory); "patches": + Normalize + ViT patch unfold -> bf16 [B*196, 768]."""
    _need_cuda(img_u8)
    assert img_u8.dtype == torch.uint8 and img_u8.dim() == 4 and img_u8.shape[-1] == 3 and img_u8.is_contiguous()
    B, Hs, Ws, _ = img_u8.shape
    OH, OW = out_hw
    L = _L()
    n_ws = L.dfu_resize_workspace_ints(Hs, Ws, OH, OW)
    if n_ws <= 0:
        raise RuntimeError("dfu_resize_workspace_ints: bad sizes %s -> %s" % ((Hs, Ws), (OH, OW)))
    ws = torch.empty(n_ws, dtype=torch.int32, device=img_u8.device)
    code = {"u8": 0, "nhwc": 1, "patches": 2}[mode]
    if code == 0:
        if out is None:
            out = torch.empty((B, OH, OW, 3), dtype=torch.uint8, device=img_u8.device)
        assert out.dtype == torch.uint8 and tuple(out.shape) == (B, OH, OW, 3) and out.is_contiguous()
        mean = std = None
    else:
        assert out is None, "a preallocated output is supported for mode 'u8'"
        out = (torch.empty((B, OH, OW, 3), dtype=BF16, device=img_u8.device) if code == 1 else
               torch.empty((B * 196, 768), dtype=BF16, device=img_u8.device))
        mean, std = _norm_consts(kind)
    check(L.dfu_resize_u8(_p(img_u8), Hs * Ws * 3, B, Hs, Ws, OH, OW, code, _p(out), _p(ws), mean, std, _st()), "resize_u8")
    _count(3)
    return out.permute(0, 3, 1, 2) if code == 1 else out


def nchw_f32_to_nhwc_bf16(x):
    _need_cuda(x)
    x = x.contiguous()
    assert x.dtype == F32
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=BF16, device=x.device)
    check(_L().dfu_nchw_f32_to_nhwc_bf16(_p(x), _p(out), B, C, H, W, _st()), "nchw_f32_to_nhwc_bf16")
    _count()
    return out.permute(0, 3, 1, 2)


def nhwc_bf16_to_nchw_f32(x_nchw_view):
    """x: logical (B,C,H,W) bf16 in NHWC memory -> contiguous fp32 NCHW."""
    x = as_nhwc(x_nchw_view)
    B, H, W, C = x.shape
    out = torch.empty((B, C, H, W), dtype=F32, device=x.device)
    check(_L().dfu_nhwc_bf16_to_nchw_f32(_p(x), _p(out), B, C, H, W, _st()), "nhwc_bf16_to_nchw_f32")
    _count()
    return out


def patchify_f32(x):
    _need_cuda(x)
    x = x.contiguous()
    assert x.dtype == F32 and tuple(x.shape[1:]) == (3, 224, 224)
    B = x.shape[0]
    out = torch.empty((B * 196, 768), dtype=BF16, device=x.device)
    check(_L().dfu_patchify_f32(_p(x), _p(out), B, _st()), "patchify_f32")
    _count()
    return out


def unpatchify_f32(dpatches, B):
    out = torch.empty((B, 3, 224, 224), dtype=F32, device=dpatches.device)
    check(_L().dfu_unpatchify_f32(_p(dpatches.contiguous()), _p(out), B, _st()), "unpatchify_f32")
    _count()
    return out


def as_nhwc(x):
    """Logical (B,C,H,W) bf16 tensor -> its (B,H,W,C) contiguous view; converts through our kernels if needed."""
    _need_cuda(x)
    v = x.permute(0, 2, 3, 1)
    if x.dtype == BF16 and v.is_contiguous():
        return v
    if x.dtype == F32:
        return nchw_f32_to_nhwc_bf16(x).permute(0, 2, 3, 1)
    if x.dtype == BF16:
        # bf16 NCHW-contiguous (e.g. a gradient produced by a foreign op): go through fp32 staging
        return nchw_f32_to_nhwc_bf16(cast_bf16_f32(x.contiguous()).view(x.shape)).permute(0, 2, 3, 1)
    raise RuntimeError("unsupported activation dtype %s" % x.dtype)


def im2col(x_nhwc, R, S, stride, pad, Ho, Wo, ldc):
    B, H, W, C = x_nhwc.shape
    col = torch.empty((B * Ho * Wo, ldc), dtype=BF16, device=x_nhwc.device)
    check(_L().dfu_im2col_nhwc(_p(x_nhwc), _p(col), B, H, W, C, R, S, stride, pad, Ho, Wo, ldc, _st()), "im2col")
    _count(2 if (ldc > R * S * C and C % 8 == 0) else 1)
    return col


def col2im(dcol, B, H, W, C, R, S, stride, pad, Ho, Wo, ldc):
    dx = torch.empty((B, H, W, C), dtype=BF16, device=dcol.device)
    check(_L().dfu_col2im_nhwc(_p(dcol), _p(dx), B, H, W, C, R, S, stride, pad, Ho, Wo, ldc, _st()), "col2im")
    _count()
    return dx


# ------------------------------------------------------------------------------------------ BatchNorm
# bumped by FusedAdamW.step(): parameters change through raw pointers there, without touching torch's version counters
PARAM_EPOCH = [0]


def fold_bn_weights(weight, gamma, beta, running_mean, running_var, eps, out=None):
    """Eval-mode BatchNorm folded into the convolution weights: (bf16 [Cout, round8(K)] with rows w * gamma/sqrt(var+eps),
    fp32 bias [Cout] = beta - mean * gamma/sqrt(var+eps)).  `weight` is a (Cout,Cin,R,S) parameter stored channels_last."""
    _need_cuda(weight, gamma, beta, running_mean, running_var)
    Cout = weight.shape[0]
    K = weight.numel() // Cout
    ldo = (K + 7) // 8 * 8
    if out is not None and out[0].shape == (Cout, ldo) and out[0].device == weight.device:
        wf, bias = out                  # refreshed in place: captured CUDA graphs keep reading the same buffers
    else:
        wf = torch.empty((Cout, ldo), dtype=BF16, device=weight.device)
        bias = torch.empty(Cout, dtype=F32, device=weight.device)
    check(_L().dfu_fold_bn_weights(_p(weight), Cout, K, _p(gamma), _p(beta), _p(running_mean), _p(running_var), eps, _p(wf), ldo,
                                   _p(bias), _st()), "fold_bn_weights")
    _count()
    return wf, bias


def bn_forward(y2d, gamma, beta, running_mean, running_var, nbt, res2d, relu, training, eps, momentum, sums=None,
               want_mask=False):
    """`sums` ([2C] fp32: per-channel sum and sum of squares of y) may come from the convolution GEMM's epilogue.
    want_mask: also return the ReLU mask as 1 bit per element (u8 [M, C/8]) -> (z, mean_invstd, mask)."""
    M, C = y2d.shape
    dev = y2d.device
    zmask = torch.empty((M, C // 8), dtype=torch.uint8, device=dev) if (want_mask and relu) else None
    mean_invstd = torch.empty(2 * C, dtype=F32, device=dev)
    L = _L()
    z = torch.empty_like(y2d)
    if training:
        if sums is None:
            sums = zeros(2 * C, F32, dev)
            check(L.dfu_bn_stats(_p(y2d), _p(sums), M, C, _st()), "bn_stats")
            _count()
        # finalize (mean/invstd, running statistics) and apply in one launch
        check(L.dfu_bn_apply_train(_p(y2d), _p(sums), _p(gamma), _p(beta), _p(res2d), _p(z), _p(zmask), _p(mean_invstd),
                                   _p(running_mean), _p(running_var), _p(nbt), M, C, 1 if relu else 0, eps, momentum,
                                   _st()), "bn_apply_train")
        _count()
        return (z, mean_invstd, zmask) if want_mask else (z, mean_invstd)
    check(L.dfu_bn_eval_stats(_p(running_mean), _p(running_var), _p(mean_invstd), C, eps, _st()), "bn_eval_stats")
    check(L.dfu_bn_apply(_p(y2d), _p(mean_invstd), _p(gamma), _p(beta), _p(res2d), _p(z), _p(zmask), M, C,
                         1 if relu else 0, _st()), "bn_apply")
    _count(2)
    return (z, mean_invstd, zmask) if want_mask else (z, mean_invstd)


def bn_stats(y2d):
    """Per-channel (sum, sum of squares) of y [M, C] -> [2C] fp32."""
    M, C = y2d.shape
    sums = zeros(2 * C, F32, y2d.device)
    check(_L().dfu_bn_stats(_p(y2d), _p(sums), M, C, _st()), "bn_stats")
    _count()
    return sums


def bn_backward(dz2d, z2d, y2d, mean_invstd, gamma, relu, training, want_dres, beta=None, zmask=None):
    """z2d may be None for a ReLU block without residual input: the mask is then recomputed from y (needs beta)."""
    M, C = y2d.shape
    dev = y2d.device
    L = _L()
    sums = zeros(2 * C, F32, dev)
    check(L.dfu_bn_bwd_reduce(_p(dz2d), _p(z2d), _p(zmask), _p(y2d), _p(mean_invstd), _p(gamma), _p(beta), _p(sums), M, C,
                              1 if relu else 0, _st()), "bn_bwd_reduce")
    dy = torch.empty_like(y2d)
    dres = torch.empty_like(y2d) if want_dres else None
    dgamma = torch.empty(C, dtype=F32, device=dev)
    dbeta = torch.empty(C, dtype=F32, device=dev)
    check(L.dfu_bn_bwd_apply(_p(dz2d), _p(z2d), _p(zmask), _p(y2d), _p(mean_invstd), _p(gamma), _p(beta), _p(sums), _p(dy), _p(dres),
                             _p(dgamma), _p(dbeta), M, C, 1 if relu else 0, 1 if training else 0, _st()),
          "bn_bwd_apply")
    _count(2)
    return dy, dres, dgamma, dbeta


# ------------------------------------------------------------------------------------------ pooling / pointwise
def maxpool(x_nhwc, want_idx):
    B, H, W, C = x_nhwc.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.empty((B, Ho, Wo, C), dtype=BF16, device=x_nhwc.device)
    idx = torch.empty((B, Ho, Wo, C), dtype=torch.uint8, device=x_nhwc.device) if want_idx else None
    check(_L().dfu_maxpool3x3s2(_p(x_nhwc), _p(y), _p(idx), B, H, W, C, _st()), "maxpool")
    _count()
    return y, idx


def maxpool_bwd(idx, dy_nhwc, H, W):
    B, Ho, Wo, C = dy_nhwc.shape
    dx = torch.empty((B, H, W, C), dtype=BF16, device=dy_nhwc.device)
    check(_L().dfu_maxpool3x3s2_bwd(_p(idx), _p(dy_nhwc), _p(dx), B, H, W, C, _st()), "maxpool_bwd")
    _count()
    return dx


def avgpool(x_nhwc, out=None, ldo=None):
    B, H, W, C = x_nhwc.shape
    if out is None:
        out = torch.empty((B, C), dtype=BF16, device=x_nhwc.device)
        ldo = C
    check(_L().dfu_avgpool(_p(x_nhwc), _p(out), B, H * W, C, ldo, _st()), "avgpool")
    _count()
    return out


def avgpool_bwd(dout2d, H, W):
    B, C = dout2d.shape
    dx = torch.empty((B, H, W, C), dtype=BF16, device=dout2d.device)
    check(_L().dfu_avgpool_bwd(_p(dout2d), dout2d.stride(0), _p(dx), B, H * W, C, _st()), "avgpool_bwd")
    _count()
    return dx


def _binary(fn, name, a, b):
    out = torch.empty_like(a, memory_format=torch.preserve_format)
    check(fn(_p(a), _p(b), _p(out), a.numel(), _st()), name)
    _count()
    return out


def _dense_same_layout(a, b):
    sa = [st for st, n in zip(a.stride(), a.shape) if n > 1]
    sb = [st for st, n in zip(b.stride(), b.shape) if n > 1]
    if sa != sb or a.shape != b.shape or a.dtype != b.dtype:
        raise RuntimeError("layout mismatch between %s/%s and %s/%s" % (a.stride(), a.dtype, b.stride(), b.dtype))


def relu(x):
    out = torch.empty_like(x, memory_format=torch.preserve_format)
    check(_L().dfu_relu(_p(x), _p(out), x.numel(), _st()), "relu")
    _count()
    return out


def gelu(x):
    out = torch.empty_like(x, memory_format=torch.preserve_format)
    check(_L().dfu_gelu(_p(x), _p(out), x.numel(), _st()), "gelu")
    _count()
    return out


def relu_bwd(dy, y):
    _dense_same_layout(dy, y)
    return _binary(_L().dfu_relu_bwd, "relu_bwd", dy, y)


def add(a, b):
    _dense_same_layout(a, b)
    return _binary(_L().dfu_add, "add", a, b)


def gelu_bwd(dy, h):
    _dense_same_layout(dy, h)
    return _binary(_L().dfu_gelu_bwd, "gelu_bwd", dy, h)


def gelu_bwd_colsum(dy2d, h2d):
    """(dy * gelu'(h), column sums of that product) in one pass over dense [M, N] bf16 matrices."""
    _dense_same_layout(dy2d, h2d)
    M, N = dy2d.shape
    dx = torch.empty_like(dy2d)
    cs = zeros(N, F32, dy2d.device)
    check(_L().dfu_gelu_bwd_colsum(_p(dy2d), _p(h2d), _p(dx), _p(cs), M, N, _st()), "gelu_bwd_colsum")
    _count()
    return dx, cs


def colsum(x2d):
    M, N = x2d.shape
    out = zeros(N, F32, x2d.device)
    check(_L().dfu_colsum(_p(x2d), _p(out), M, N, x2d.stride(0), _st()), "colsum")
    _count()
    return out


def copy2d(src, lds, dst, ldd, rows, cols, src_off=0, dst_off=0):
    check(_L().dfu_copy2d(_p(src, src_off), lds, _p(dst, dst_off), ldd, rows, cols, _st()), "copy2d")
    _count()


def cast_f32_bf16(src, rows=None, cols=None, ldo=None, out=None):
    """fp32 -> bf16 over the tensor's storage order (dense tensors only); optional zero-padded row stride."""
    n = src.numel()
    if rows is None:
        rows, cols, ldo = 1, n, n
    if out is None:
        out = torch.empty(rows * ldo, dtype=BF16, device=src.device)
    check(_L().dfu_cast_f32_bf16(_p(src), _p(out), rows, cols, ldo, _st()), "cast_f32_bf16")
    _count()
    return out


def cast_bf16_f32(src):
    out = torch.empty(src.shape, dtype=F32, device=src.device)
    check(_L().dfu_cast_bf16_f32(_p(src), _p(out), src.numel(), _st()), "cast_bf16_f32")
    _count()
    return out


class _DropoutState:
    """Device-resident seed so CUDA-graph replays draw fresh masks; stream ids separate call sites."""

    def __init__(self):
        self.seeds = {}
        self.next_id = 1

    def seed(self, device):
        key = (device.type, device.index)
        if key not in self.seeds:
            s = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
            self.seeds[key] = torch.tensor([s], dtype=torch.int64, device=device)
        return self.seeds[key]

    def tick(self, device):
        check(_L().dfu_tick_u64(_p(self.seed(device)), _st()), "tick_u64")
        _count()


DROPOUT = _DropoutState()


def dropout(x, p, stream_id):
    y = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(_L().dfu_dropout(_p(x), _p(y), _p(mask), x.numel(), p, _p(DROPOUT.seed(x.device)), stream_id, _st()),
          "dropout")
    _count()
    return y, mask


def dropout_bwd(dy, mask, p):
    dx = torch.empty_like(dy)
    check(_L().dfu_dropout_bwd(_p(dy), _p(mask), _p(dx), dy.numel(), p, _st()), "dropout_bwd")
    _count()
    return dx


# ------------------------------------------------------------------------------------------ ViT pieces
def vit_assemble(patch_out, cls, pos, B, T, D):
    tokens = torch.empty((B * T, D), dtype=BF16, device=patch_out.device)
    check(_L().dfu_vit_assemble(_p(patch_out), _p(cls), _p(pos), _p(tokens), B, T, D, _st()), "vit_assemble")
    _count()
    return tokens


def vit_assemble_bwd(dtokens, B, T, D):
    dev = dtokens.device
    dpatch = torch.empty((B * (T - 1), D), dtype=BF16, device=dev)
    dcls = zeros(D, F32, dev)
    dpos = zeros(T * D, F32, dev)
    check(_L().dfu_vit_assemble_bwd(_p(dtokens), _p(dpatch), _p(dcls), _p(dpos), B, T, D, _st()), "vit_assemble_bwd")
    _count()
    return dpatch, dcls, dpos


def layernorm(x2d, gamma, beta, eps, want_stats=True):
    rows, D = x2d.shape
    y = torch.empty_like(x2d)
    stats = torch.empty(rows * 2, dtype=F32, device=x2d.device) if want_stats else None
    check(_L().dfu_layernorm(_p(x2d), _p(gamma), _p(beta), _p(y), _p(stats), rows, D, eps, _st()), "layernorm")
    _count()
    return y, stats


def layernorm_bwd(dy2d, x2d, gamma, stats, dres=None):
    """dx = LN'(dy) (+ dres, the skip-connection gradient), dgamma, dbeta."""
    rows, D = x2d.shape
    dx = torch.empty_like(x2d)
    dgb = zeros((2, D), F32, x2d.device)
    dgamma, dbeta = dgb[0], dgb[1]
    check(_L().dfu_layernorm_bwd(_p(dy2d), _p(x2d), _p(gamma), _p(stats), _p(dres), _p(dx), _p(dgamma), _p(dbeta),
                                 rows, D, _st()), "layernorm_bwd")
    _count()
    return dx, dgamma, dbeta


# "fused": one tcgen05 kernel per direction (attention.cu); "unfused": batched GEMM + row-softmax kernels.
_ATTN_IMPL = os.environ.get("DFU_ATTN_IMPL", "fused")


def set_attention_impl(impl):
    global _ATTN_IMPL
    assert impl in ("fused", "unfused")
    _ATTN_IMPL = impl


def attention_forward(qkv, B, T, H, scale):
    """softmax(q k^T * scale) v per (batch, head); qkv [B*T, 3*H*64] bf16 in timm order [3][H][64].
    Returns (out [B*T, H*64], saved) where `saved` is what the backward pass needs: the fp32 row log-sum-exp
    [B,H,T] of the fused kernel, or the bf16 probabilities [B,H,T,ldp] of the unfused path."""
    if _ATTN_IMPL == "fused" and 128 < T <= 208:
        _need_cuda(qkv)
        assert qkv.dtype == BF16 and qkv.is_contiguous() and qkv.shape == (B * T, 3 * H * 64)
        out = torch.empty((B * T, H * 64), dtype=BF16, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=F32, device=qkv.device)
        if GEMM_EVENTS is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(_L().dfu_attention_fwd(_p(qkv), _p(out), _p(lse), B, T, H, scale, _st()), "attention_fwd")
        if GEMM_EVENTS is not None:
            e1.record()
            GEMM_EVENTS.append((e0, e1, (T, T, 128, B * H, 0, 0, "bf16", 0, 0, 0, 0, 0, "attn_fwd")))
        _count()
        return out, lse
    hd = 64
    D3 = 3 * H * hd
    ldp = (T + 7) // 8 * 8
    dev = qkv.device
    P = torch.empty((B, H, T, ldp), dtype=BF16, device=dev)
    gemm(qkv, qkv, P, M=T, N=T, K=hd, lda=D3, ldb=D3, ldd=ldp, a_off=0, b_off=H * hd, nb0=H, nb1=B,
         sa=(hd, T * D3), sb=(hd, T * D3), sd=(T * ldp, H * T * ldp), alpha=scale)
    check(_L().dfu_softmax_rows(_p(P), _p(P), B * H * T, T, ldp, _st()), "softmax_rows")
    _count()
    out = torch.empty((B * T, H * hd), dtype=BF16, device=dev)
    gemm(P, qkv, out, M=T, N=hd, K=T, lda=ldp, ldb=D3, b_major=1, b_off=2 * H * hd, ldd=H * hd, nb0=H, nb1=B,
         sa=(T * ldp, H * T * ldp), sb=(hd, T * D3), sd=(hd, T * H * hd))
    return out, P


def attention_backward(qkv, saved, dout, B, T, H, scale, out=None):
    """dqkv from the upstream gradient of the attention output; `saved` as returned by attention_forward."""
    if saved.dtype == F32:
        assert out is not None, "the fused backward needs the forward output"
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        if GEMM_EVENTS is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(_L().dfu_attention_bwd(_p(qkv), _p(out), _p(dout), _p(saved), _p(dqkv), B, T, H, scale, _st()),
              "attention_bwd")
        if GEMM_EVENTS is not None:
            e1.record()
            GEMM_EVENTS.append((e0, e1, (T, T, 320, B * H, 0, 0, "bf16", 0, 0, 0, 0, 0, "attn_bwd")))
        _count()
        return dqkv
    P = saved
    hd = 64
    D3 = 3 * H * hd
    ldp = P.shape[-1]
    dev = qkv.device
    dqkv = torch.empty_like(qkv)
    sP = (T * ldp, H * T * ldp)
    sQ = (hd, T * D3)
    sO = (hd, T * H * hd)
    # dV = P^T dO
    gemm(P, dout, dqkv, M=T, N=hd, K=T, a_major=1, b_major=1, lda=ldp, ldb=H * hd, ldd=D3, d_off=2 * H * hd, nb0=H,
         nb1=B, sa=sP, sb=sO, sd=sQ)
    # dP = dO V^T
    dP = torch.empty_like(P)
    gemm(dout, qkv, dP, M=T, N=T, K=hd, lda=H * hd, ldb=D3, b_off=2 * H * hd, ldd=ldp, nb0=H, nb1=B, sa=sO, sb=sQ,
         sd=sP)
    # dS = P * (dP - rowsum(dP*P)) * scale   (in place over dP)
    check(_L().dfu_softmax_rows_bwd(_p(P), _p(dP), _p(dP), B * H * T, T, ldp, scale, _st()), "softmax_rows_bwd")
    _count()
    # dQ = dS K
    gemm(dP, qkv, dqkv, M=T, N=hd, K=T, lda=ldp, ldb=D3, b_major=1, b_off=H * hd, ldd=D3, d_off=0, nb0=H, nb1=B,
         sa=sP, sb=sQ, sd=sQ)
    # dK = dS^T Q
    gemm(dP, qkv, dqkv, M=T, N=hd, K=T, a_major=1, b_major=1, lda=ldp, ldb=D3, b_off=0, ldd=D3, d_off=H * hd,
         nb0=H, nb1=B, sa=sP, sb=sQ, sd=sQ)
    return dqkv


# ------------------------------------------------------------------------------------------ loss / predict
def cross_entropy(logits, labels, weight=None, want_grad=True):
    _need_cuda(logits, labels)
    assert logits.dtype == F32 and labels.dtype == torch.int64
    logits = logits.contiguous()
    B, C = logits.shape
    loss = torch.empty((), dtype=F32, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    check(_L().dfu_cross_entropy(_p(logits), _p(labels), _p(weight), _p(loss), _p(dlogits), B, C, _st()),
          "cross_entropy")
    _count()
    return loss, dlogits


def scale_by_scalar(a, scalar):
    out = torch.empty_like(a)
    check(_L().dfu_scale_by_scalar(_p(a), _p(scalar), _p(out), a.numel(), _st()), "scale_by_scalar")
    _count()
    return out


def predict(logits):
    """softmax(logits)[:,1] and argmax (extended_metrics.py:636-637) in one kernel."""
    _need_cuda(logits)
    logits = logits.contiguous()
    B, C = logits.shape
    prob1 = torch.empty(B, dtype=F32, device=logits.device)
    pred = torch.empty(B, dtype=torch.int64, device=logits.device)
    check(_L().dfu_predict(_p(logits), _p(prob1), _p(pred), B, C, _st()), "predict")
    _count()
    return prob1, pred


def gradcam_reduce(A_nhwc2d, G_nhwc2d, nc):
    """A [HW, CA], G [HW, CG] bf16 -> cam [HW] fp32 = relu(sum_c mean_hw(G_c) A_c) / max."""
    HW, CA = A_nhwc2d.shape
    CG = G_nhwc2d.shape[1]
    cam = torch.empty(HW, dtype=F32, device=A_nhwc2d.device)
    check(_L().dfu_gradcam(_p(A_nhwc2d), CA, _p(G_nhwc2d), CG, HW, nc, _p(cam), _st()), "gradcam")
    _count()
    return cam


def saliency(dimg_chw):
    C, H, W = dimg_chw.shape
    d = dimg_chw.contiguous()
    out = torch.empty(H * W, dtype=F32, device=d.device)
    scratch = torch.empty(1, dtype=F32, device=d.device)
    check(_L().dfu_saliency(_p(d), _p(out), _p(scratch), C, H * W, _st()), "saliency")
    _count(2)
    return out.view(H, W)


def tta_vote(probs_nv):
    n, V = probs_nv.shape
    probs_nv = probs_nv.contiguous()
    mean_prob = torch.empty(n, dtype=F32, device=probs_nv.device)
    pred = torch.empty(n, dtype=torch.int64, device=probs_nv.device)
    check(_L().dfu_tta_vote(_p(probs_nv), _p(mean_prob), _p(pred), n, V, _st()), "tta_vote")
    _count()
    return mean_prob, pred
