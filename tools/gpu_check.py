"""Developer harness: runs every kernel of libdfu_b200.so against torch fp32 references on the GPU box and prints
one line per check (no fail-fast), so a single gpurun call gives the whole picture.  Not a test: tests/ holds those."""
import json
import math
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfu_multimodal_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
RESULTS = []


def report(name, got, ref, tol):
    got = got.float().cpu()
    ref = ref.float().cpu()
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    scale = ref.abs().max().item() if ref.numel() else 0.0
    bad = not math.isfinite(err) or err > tol * max(1.0, scale)
    RESULTS.append({"name": name, "max_abs_err": err, "ref_max": scale, "ok": not bad})
    print("%-58s err=%.4g ref_max=%.4g %s" % (name, err, scale, "OK" if not bad else "FAIL"), flush=True)


def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        RESULTS.append({"name": name, "ok": False, "exc": repr(e)})
        print("%-58s EXC %r" % (name, e), flush=True)
        traceback.print_exc()


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def gemm_case(M, N, K, a_major, b_major, impl, bias=False, act=0, res=False, fp32=False, acc=False, aux=False):
    def f():
        A = rnd(K, M) if a_major else rnd(M, K)
        B = rnd(K, N) if b_major else rnd(N, K)
        Af = (A.float().t() if a_major else A.float())
        Bf = (B.float().t() if b_major else B.float())
        ref = Af @ Bf.t()
        bias_t = torch.randn(N, device=dev) if bias else None
        if bias:
            ref = ref + bias_t
        pre = ref.clone()
        if act == 1:
            ref = ref.relu()
        elif act == 2:
            ref = F.gelu(ref)
        res_t = rnd(M, N) if res else None
        if res:
            ref = ref + res_t.float()
        D = (torch.zeros if acc else torch.empty)((M, N), dtype=torch.float32 if fp32 else torch.bfloat16, device=dev)
        if acc:
            D += 1.0
            ref = ref + 1.0
        aux_t = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if aux else None
        ops.gemm(A, B, D, M=M, N=N, K=K, lda=A.stride(0), ldb=B.stride(0), ldd=N, a_major=a_major, b_major=b_major,
                 bias=bias_t, act=act, res=res_t, ldr=N, accumulate=acc, aux=aux_t, impl=impl)
        torch.cuda.synchronize()
        tol = 2e-2 if not fp32 else 2e-3
        report("gemm[%s] M%d N%d K%d a%d b%d bias%d act%d res%d f32%d acc%d" % (
            "tc" if impl == 0 else "simt", M, N, K, a_major, b_major, bias, act, res, fp32, acc), D, ref,
            tol * math.sqrt(K) / 8)
        if aux:
            report("   aux", aux_t, pre, tol * math.sqrt(K) / 8)
    return f


def main():
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), flush=True)
    # ---- GEMM: SIMT first (sanity of harness), then tcgen05 in all major modes
    for impl in (1, 0):
        run("g0", gemm_case(128, 64, 64, 0, 0, impl))
        run("g1", gemm_case(256, 256, 128, 0, 0, impl))
        run("g2", gemm_case(300, 200, 96, 0, 0, impl, bias=True))
        run("g3", gemm_case(1024, 768, 768, 0, 0, impl, bias=True, act=2, aux=True))
        run("g4", gemm_case(777, 512, 2816, 0, 0, impl, bias=True, act=1, res=True))
        run("g5", gemm_case(512, 768, 256, 0, 1, impl))           # dgrad form
        run("g6", gemm_case(768, 3072, 1000, 1, 1, impl, fp32=True))  # wgrad form
        run("g7", gemm_case(64, 576, 5000, 1, 1, impl, fp32=True, acc=True))  # split-K wgrad
        run("g8", gemm_case(200, 64, 152, 0, 0, impl))            # stem-like K tail
        run("g9", gemm_case(128, 2, 512, 0, 0, impl, bias=True, fp32=True))   # classifier (SIMT route)
        run("g10", gemm_case(2048, 2304, 768, 0, 0, impl, bias=True))
        run("g11", gemm_case(197, 197, 64, 1, 0, impl))

    # ---- attention (batched GEMMs + softmax) vs torch SDPA math
    def attn():
        B, T, H = 3, 197, 12
        qkv = rnd(B * T, 3 * H * 64, scale=0.5)
        qkv.requires_grad_(False)
        out, P = ops.attention_forward(qkv, B, T, H, 0.125)
        q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
        q = q.clone().requires_grad_(True); k = k.clone().requires_grad_(True); v = v.clone().requires_grad_(True)
        s = (q @ k.transpose(-1, -2)) * 0.125
        p = s.softmax(-1)
        o = (p @ v).permute(0, 2, 1, 3).reshape(B * T, H * 64)
        report("attention fwd", out, o, 2e-2)
        do = rnd(B * T, H * 64)
        o.backward(do.float())
        dqkv = ops.attention_backward(qkv, P, do, B, T, H, 0.125)
        dref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * H * 64)
        report("attention bwd", dqkv, dref, 3e-2)
    run("attn", attn)

    # ---- layout
    def layout():
        B = 3
        img = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
        mean = torch.tensor([0.485, 0.456, 0.406], device=dev); std = torch.tensor([0.229, 0.224, 0.225], device=dev)
        ref = ((img.float() / 255 - mean) / std).permute(0, 3, 1, 2)
        report("preprocess_u8_nhwc", ops.preprocess_u8_nhwc(img, "imagenet"), ref, 1e-2)
        refh = ((img.float() / 255 - 0.5) / 0.5).permute(0, 3, 1, 2)
        patches_ref = refh.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * 196, 768)
        report("preprocess_u8_patches", ops.preprocess_u8_patches(img, "half"), patches_ref, 1e-2)
        x = torch.randn(B, 3, 224, 224, device=dev)
        report("patchify_f32", ops.patchify_f32(x), x.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * 196, 768), 1e-2)
        dp = rnd(B * 196, 768)
        report("unpatchify_f32", ops.unpatchify_f32(dp, B), dp.float().view(B, 14, 14, 3, 16, 16).permute(0, 3, 1, 4, 2, 5).reshape(B, 3, 224, 224), 1e-6)
        x = torch.randn(2, 70, 9, 11, device=dev)
        y = ops.nchw_f32_to_nhwc_bf16(x)
        report("nchw_f32_to_nhwc_bf16", y, x, 1e-2)
        report("nhwc_bf16_to_nchw_f32", ops.nhwc_bf16_to_nchw_f32(y), y.float(), 1e-6)
        for (C, R, S, st, pad, H) in ((64, 3, 3, 1, 1, 14), (64, 3, 3, 2, 1, 14), (3, 7, 7, 2, 3, 32), (128, 1, 1, 2, 0, 8)):
            xx = rnd(2, C, H, H)
            xn = ops.as_nhwc(ops.nchw_f32_to_nhwc_bf16(xx.float()))
            Ho = (H + 2 * pad - R) // st + 1
            Kc = R * S * C
            ldc = (Kc + 7) // 8 * 8
            col = ops.im2col(xn, R, S, st, pad, Ho, Ho, ldc)
            unf = F.unfold(xx.float(), (R, S), padding=pad, stride=st)  # [B, C*R*S, L] order (c, r, s)
            ref = unf.view(2, C, R * S, Ho * Ho).permute(0, 3, 2, 1).reshape(2 * Ho * Ho, Kc)
            report("im2col C%d R%d s%d" % (C, R, st), col[:, :Kc], ref, 1e-6)
            if ldc > Kc:
                report("  pad cols zero", col[:, Kc:], torch.zeros_like(col[:, Kc:]), 1e-6)
            dcol = rnd(2 * Ho * Ho, ldc)
            dx = ops.col2im(dcol, 2, H, H, C, R, S, st, pad, Ho, Ho, ldc)
            dunf = dcol[:, :Kc].float().view(2, Ho * Ho, R * S, C).permute(0, 3, 2, 1).reshape(2, C * R * S, Ho * Ho)
            ref = F.fold(dunf, (H, H), (R, S), padding=pad, stride=st)
            report("col2im C%d R%d s%d" % (C, R, st), dx.permute(0, 3, 1, 2), ref, 2e-2)
    run("layout", layout)

    # ---- BatchNorm
    def bn():
        for C, M in ((64, 5000), (256, 777), (2048, 98)):
            y = rnd(M, C) * 2 + 0.5
            gamma = torch.rand(C, device=dev) + 0.5; beta = torch.randn(C, device=dev)
            rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev); nbt = torch.zeros((), dtype=torch.int64, device=dev)
            res = rnd(M, C)
            z, mi = ops.bn_forward(y, gamma, beta, rm, rv, nbt, res, True, True, 1e-5, 0.1)
            yf = y.float().requires_grad_(True); g = gamma.clone().requires_grad_(True); b = beta.clone().requires_grad_(True)
            rm2 = torch.zeros(C, device=dev); rv2 = torch.ones(C, device=dev)
            resf = res.float().requires_grad_(True)
            zr = F.relu(F.batch_norm(yf, rm2, rv2, g, b, True, 0.1, 1e-5) + resf)
            report("bn fwd C%d" % C, z, zr, 2e-2)
            report("bn running_mean C%d" % C, rm, rm2, 1e-3)
            report("bn running_var C%d" % C, rv, rv2, 1e-3)
            assert int(nbt.item()) == 1
            dz = rnd(M, C)
            # reference uses the relu mask of OUR z to avoid tie noise
            zr2 = F.batch_norm(yf, rm2.clone(), rv2.clone(), g, b, True, 0.1, 1e-5) + resf
            mask = (z.float() > 0).float()
            (zr2 * mask).backward(dz.float())
            dy, dres, dg, db = ops.bn_backward(dz, z, y, mi, gamma, True, True, True)
            report("bn bwd dy C%d" % C, dy, yf.grad, 3e-2)
            report("bn bwd dres C%d" % C, dres, resf.grad, 1e-2)
            report("bn bwd dgamma C%d" % C, dg, g.grad, 2e-2)
            report("bn bwd dbeta C%d" % C, db, b.grad, 2e-2)
            # eval mode
            z2, mi2 = ops.bn_forward(y, gamma, beta, rm, rv, nbt, None, False, False, 1e-5, 0.1)
            report("bn eval fwd C%d" % C, z2, F.batch_norm(y.float(), rm, rv, gamma, beta, False, 0.1, 1e-5), 2e-2)
    run("bn", bn)

    # ---- LayerNorm
    def ln():
        rows, D = 1000, 768
        x = rnd(rows, D) * 1.5 + 0.3
        g = torch.rand(D, device=dev) + 0.5; b = torch.randn(D, device=dev)
        y, stats = ops.layernorm(x, g, b, 1e-6)
        xf = x.float().requires_grad_(True); gf = g.clone().requires_grad_(True); bf = b.clone().requires_grad_(True)
        yr = F.layer_norm(xf, (D,), gf, bf, 1e-6)
        report("layernorm fwd", y, yr, 2e-2)
        dy = rnd(rows, D)
        yr.backward(dy.float())
        dx, dg, db = ops.layernorm_bwd(dy, x, g, stats)
        report("layernorm bwd dx", dx, xf.grad, 2e-2)
        report("layernorm bwd dgamma", dg, gf.grad, 2e-2)
        report("layernorm bwd dbeta", db, bf.grad, 2e-2)
    run("ln", ln)

    # ---- pooling / pointwise
    def pool():
        x = rnd(2, 64, 20, 20)
        xn = ops.as_nhwc(ops.nchw_f32_to_nhwc_bf16(x.float()))
        y, idx = ops.maxpool(xn, True)
        xf = x.float().requires_grad_(True)
        yr = F.max_pool2d(xf, 3, 2, 1)
        report("maxpool fwd", y.permute(0, 3, 1, 2), yr, 1e-6)
        dy = rnd(2, 10, 10, 64)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        report("maxpool bwd", ops.maxpool_bwd(idx, dy, 20, 20).permute(0, 3, 1, 2), xf.grad, 2e-2)
        x = rnd(4, 7, 7, 2048)
        report("avgpool fwd", ops.avgpool(x), x.float().mean((1, 2)), 1e-2)
        d = rnd(4, 2048)
        report("avgpool bwd", ops.avgpool_bwd(d, 7, 7), (d.float() / 49)[:, None, None, :].expand(4, 7, 7, 2048), 1e-2)
        a = rnd(1001); b = rnd(1001)
        report("relu", ops.relu(a), a.float().relu(), 1e-6)
        report("relu_bwd", ops.relu_bwd(a, b), a.float() * (b.float() > 0), 1e-6)
        report("add", ops.add(a, b), a.float() + b.float(), 1e-2)
        hf = b.float().requires_grad_(True)
        F.gelu(hf).backward(a.float())
        report("gelu_bwd", ops.gelu_bwd(a, b), hf.grad, 1e-2)
        x = rnd(3000, 768)
        report("colsum", ops.colsum(x), x.float().sum(0), 1e-3)
        x = rnd(128, 2)
        report("colsum small", ops.colsum(x), x.float().sum(0), 1e-3)
        w = torch.randn(64, 147, device=dev)
        report("cast pad", ops.cast_f32_bf16(w, 64, 147, 152).view(64, 152)[:, :147], w, 1e-2)
        w = torch.randn(1000, 77, device=dev)
        report("cast flat", ops.cast_f32_bf16(w).view(1000, 77), w, 1e-2)
        x = rnd(100000)
        y, m = ops.dropout(x, 0.5, 1)
        keep = m.float().mean().item()
        report("dropout keep-rate", torch.tensor([keep]), torch.tensor([0.5]), 2e-2)
        report("dropout values", y, x.float() * m.float() * 2, 1e-2)
        report("dropout bwd", ops.dropout_bwd(x, m, 0.5), x.float() * m.float() * 2, 1e-2)
        B, T, D = 3, 197, 768
        po = rnd(B * 196, D); cls = torch.randn(D, device=dev); pos = torch.randn(T * D, device=dev)
        tok = ops.vit_assemble(po, cls, pos, B, T, D)
        ref = torch.cat([cls.view(1, 1, D).expand(B, 1, D), po.float().view(B, 196, D)], 1) + pos.view(1, T, D)
        report("vit_assemble", tok.view(B, T, D), ref, 2e-2)
        dt = rnd(B * T, D)
        dpa, dcls, dpos = ops.vit_assemble_bwd(dt, B, T, D)
        report("vit_assemble_bwd dpatch", dpa.view(B, 196, D), dt.float().view(B, T, D)[:, 1:], 1e-6)
        report("vit_assemble_bwd dcls", dcls, dt.float().view(B, T, D)[:, 0].sum(0), 1e-3)
        report("vit_assemble_bwd dpos", dpos.view(T, D), dt.float().view(B, T, D).sum(0), 1e-3)
    run("pool", pool)

    # ---- loss / optimizer / gradcam
    def misc():
        B = 128
        logits = torch.randn(B, 2, device=dev) * 2
        labels = torch.randint(0, 2, (B,), device=dev)
        w = torch.tensor([2.1, 1.9], device=dev)
        loss, dl = ops.cross_entropy(logits, labels, w)
        lf = logits.clone().requires_grad_(True)
        lr = F.cross_entropy(lf, labels, weight=w)
        lr.backward()
        report("cross_entropy loss", loss.view(1), lr.view(1), 1e-5)
        report("cross_entropy grad", dl, lf.grad, 1e-5)
        p1, pred = ops.predict(logits)
        report("predict prob", p1, logits.softmax(1)[:, 1], 1e-5)
        report("predict argmax", pred, logits.argmax(1), 0)
        A = rnd(49, 2048); G = rnd(49, 512)
        cam = ops.gradcam_reduce(A, G, 512)
        wts = G.float().mean(0)
        ref = (A.float()[:, :512] * wts).sum(1).relu()
        ref = ref / ref.max() if ref.max() > 0 else ref
        report("gradcam", cam, ref, 1e-3)
        d = torch.randn(3, 224, 224, device=dev)
        sal = d.abs().mean(0); sal = sal / sal.max()
        report("saliency", ops.saliency(d), sal, 1e-5)
        probs = torch.rand(276, 8, device=dev)
        mp, pr = ops.tta_vote(probs)
        report("tta mean", mp, probs.mean(1), 1e-6)
        report("tta vote", pr, ((probs > 0.5).float().mean(1) > 0.5).long(), 0)
    run("misc", misc)

    ok = all(r.get("ok") for r in RESULTS)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gpu_check.json", "w") as f:
        json.dump(RESULTS, f, indent=1)
    print("ALL OK" if ok else "SOME FAILED: " + ", ".join(r["name"] for r in RESULTS if not r.get("ok")))


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("elapsed %.1fs" % (time.time() - t0))
