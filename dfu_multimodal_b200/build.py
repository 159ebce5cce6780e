"""Builds libdfu_b200.so (sm_100a only) in-tree with nvcc.  No torch headers are involved: the library is a
plain C-ABI shared object (include/dfu_b200.h) that the host side binds with ctypes."""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdfu_b200.so")
STAMP = os.path.join(HERE, ".libdfu_b200.stamp")

SOURCES = ["api.cu", "tc_gemm.cu", "layout.cu", "norm.cu", "pointwise.cu", "train_misc.cu", "attention.cu", "validate_f32.cu", "resize.cu", "metrics.cu", "augment.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
# no --use_fast_math: erff / expf accuracy matters for GELU and softmax parity; fast intrinsics are called explicitly
# where the kernels want them.


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libdfu_b200.so must be built where the CUDA toolkit is available")


def _digest():
    h = hashlib.sha256()
    # every source and header under csrc/ (tc_common.cuh is included by attention.cu and tc_gemm.cu) plus the public ABI
    paths = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))
    paths.append(os.path.join(HERE, "..", "include", "dfu_b200.h"))
    for p in paths:
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libdfu_b200.so next to this file. Returns the library path."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (s, out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(objdir, "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see %s" % os.path.join(objdir, "nvcc.log"))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
