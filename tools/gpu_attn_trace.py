"""Where does one (batch, head) of the fused attention forward spend its cycles?

Builds a SECOND copy of the library with -DDFU_ATTN_TRACE (tools/_build/libdfu_b200_trace.so; the product library is
untouched), runs dfu_attention_fwd at the ViT-B/16 shape and prints, for CTA 0, the clock64() deltas between the phase
boundaries of the MMA-issuing thread and of one thread of each softmax group.

  python tools/gpu_attn_trace.py build      # here (no GPU needed)
  python tools/gpu_attn_trace.py [B]        # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_build")
LIB = os.environ.get("DFU_TRACE_LIB") or os.path.join(OUT, "libdfu_b200_trace.so")


def build():
    from dfu_multimodal_b200 import build as B
    os.makedirs(OUT, exist_ok=True)
    objs, procs = [], []
    for s in B.SOURCES:
        obj = os.path.join(OUT, s.replace(".cu", ".o"))
        objs.append(obj)
        flags = [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
        procs.append(subprocess.Popen([B._nvcc()] + flags + ["-DDFU_ATTN_TRACE"] + os.environ.get("DFU_TRACE_DEFS", "").split() + ["-c", os.path.join(B.CSRC, s), "-o", obj]))
    if any(p.wait() for p in procs):
        raise SystemExit("nvcc failed")
    subprocess.check_call([B._nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    for o in objs:
        os.remove(o)
    print(LIB)


MMA = ["qk_full+o_empty0", "S0 issued", "p_full1(prev)", "O1(prev) issued", "o_empty1", "S1 issued", "v_full+p_full0", "O0 issued"]
SM = ["loop top", "s_full", "pass1 done", "barrier", "pass2 done+P arrive", "barrier", "o_full", "O stored+o_empty"]


def main(B=128):
    import torch
    lib = ctypes.CDLL(LIB)
    T, H = 197, 12
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H * T, device="cuda", dtype=torch.float32)
    lib.dfu_attention_fwd.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int32] * 3 + [ctypes.c_float, ctypes.c_void_p]
    lib.dfu_last_error.restype = ctypes.c_char_p
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(6):
        if i == 1:
            ev[0].record()
        rc = lib.dfu_attention_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, 0.125, st)
        if rc:
            raise SystemExit(lib.dfu_last_error().decode())
    ev[1].record()
    torch.cuda.synchronize()
    print("attn_fwd B=%d: %.1f us per launch (5 back-to-back launches)" % (B, ev[0].elapsed_time(ev[1]) * 200))
    dout = torch.randn(B * T, H * 64, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    lib.dfu_attention_bwd.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int32] * 3 + [ctypes.c_float, ctypes.c_void_p]
    for i in range(6):
        if i == 1:
            ev[0].record()
        rc = lib.dfu_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, T, H, 0.125, st)
        if rc:
            raise SystemExit(lib.dfu_last_error().decode())
    ev[1].record()
    torch.cuda.synchronize()
    print("attn_bwd B=%d: %.1f us per launch (5 back-to-back launches)" % (B, ev[0].elapsed_time(ev[1]) * 200))
    if not hasattr(lib, "dfu_attn_trace_read_host"):      # DFU_TRACE_LIB = the product library: timings only (e.g. under ncu)
        return
    buf = (ctypes.c_longlong * (4 * 16 * 12))()
    assert lib.dfu_attn_trace_read_host(buf) == 0
    cats = ["ops_full + D", "wait sd_full", "tmem ld + exp + dS", "wait tile_free", "write P^T/dS^T + arrive", "wait acc_full",
            "dV/dK store", "wait dq_full", "dQ store", "loop"]
    tot = [(buf[(3 * 16 + 8) * 12 + c] - buf[(3 * 16 + 0) * 12 + c]) / 8.0 for c in range(10)]      # running sums: heads 1..8
    print("backward, exp warp 0 of CTA 0, clocks per head (heads 1-8 of the last launch): total %.0f" % sum(tot))
    print("  " + "  ".join("%s %.0f" % (n, v) for n, v in zip(cats, tot)))
    t = [[[buf[(r * 16 + it) * 12 + p] for p in range(12)] for it in range(16)] for r in range(3)]
    its = min(16, (B * H + 147) // 148)
    t0 = t[0][0][0]
    for it in range(its):
        print("head %d of CTA 0 (clocks since the first qk_full)" % it)
        print("  mma     " + "  ".join("%s %d" % (n, t[0][it][p] - t0) for p, n in enumerate(MMA)))
        for g in range(2):
            print("  group %d " % g + "  ".join("%s %d" % (n, t[1 + g][it][p] - t0) for p, n in enumerate(SM)))
    if its > 3:
        per = (t[0][its - 1][0] - t[0][1][0]) / (its - 2)
        print("steady-state period: %.0f clocks per head" % per)
        for g in range(2):
            d = [sum(t[1 + g][it][p + 1] - t[1 + g][it][p] for it in range(1, its - 1)) / (its - 2) for p in range(7)]
            print("  group %d mean phase clocks: " % g + "  ".join("%s->%s %.0f" % (SM[p], SM[p + 1], d[p]) for p in range(7)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        main(int(sys.argv[1]) if len(sys.argv) > 1 else 128)
