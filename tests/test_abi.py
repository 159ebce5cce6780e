"""C-ABI surface: the header is the contract; the shared library must export every symbol it declares.
No compute entry point is called here (there is no GPU in the build container)."""
import ctypes
import os
import subprocess

import pytest

from dfu_multimodal_b200 import _lib


def test_header_parses_and_names_the_hot_path():
    protos = _lib.parse_header()
    for name in ("dfu_gemm", "dfu_im2col_nhwc", "dfu_col2im_nhwc", "dfu_bn_stats", "dfu_bn_apply", "dfu_bn_bwd_apply",
                 "dfu_layernorm", "dfu_layernorm_bwd", "dfu_softmax_rows", "dfu_cross_entropy", "dfu_adamw",
                 "dfu_preprocess_u8_nhwc", "dfu_preprocess_u8_patches", "dfu_gradcam", "dfu_tta_vote", "dfu_maxpool3x3s2",
                 "dfu_avgpool", "dfu_dropout", "dfu_last_error"):
        assert name in protos, name
    assert len(protos) >= 45


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _lib.parse_header() if not hasattr(lib, n)]
    assert not missing, missing


def test_load_sets_signatures_and_version():
    lib = _lib.load()
    assert lib.dfu_version() == 100
    assert lib.dfu_gemm.argtypes == [ctypes.c_void_p, ctypes.c_void_p]
    assert lib.dfu_last_error.restype == ctypes.c_char_p


def test_gemm_args_struct_matches_header_layout():
    # field order of dfu_gemm_args in include/dfu_b200.h
    names = [f[0] for f in _lib.GemmArgs._fields_]
    assert names[:5] == ["M", "N", "K", "nb0", "nb1"]
    assert names[-6:] == ["alpha", "act", "impl", "max_ctas", "colstats", "colstats_mode"]
    assert ctypes.sizeof(_lib.AdamWTensor) == 48


def test_conv_args_struct_matches_header_layout():
    import re
    with open(_lib.HEADER) as f:
        text = f.read()
    body = re.search(r"typedef struct \{([^}]*)\} dfu_conv_args;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(part.replace("*", " ").split()[-1])
    assert names == [f[0] for f in _lib.ConvArgs._fields_]


def test_sass_has_blackwell_tensor_and_tma_instructions():
    exe = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in out and "UTMASTG" in out, "TMA load/store missing from SASS"
    assert "LDTM" in out, "tcgen05.ld missing from SASS"
    assert "UTCHMMA.2CTA" in out and "UTMALDG.4D.2CTA" in out and "UTCBAR.2CTA.MULTICAST" in out, \
        "cta_group::2 (CTA-pair) GEMM path missing from SASS"
    assert "UTMAREDG" in out, "TMA reduce-add (split-K) missing from SASS"
    assert "HMMA." not in out.replace("UTCHMMA", ""), "legacy mma.sync path present"


def test_argument_validation_needs_no_device():
    lib = _lib.load()
    assert lib.dfu_gemm(None, None) != 0
    assert b"null" in lib.dfu_last_error()
    assert lib.dfu_relu(None, None, 0, None) != 0


def test_augment_params_struct_matches_header_layout():
    import re
    with open(_lib.HEADER) as f:
        text = f.read()
    body = re.search(r"typedef struct \{([^}]*)\} dfu_augment_params;", text, flags=re.S).group(1)
    names = [re.sub(r"\[.*", "", d.strip().split()[-1]) for d in body.split(";") if d.strip()]
    assert names == [f[0] for f in _lib.AugmentParams._fields_]
    assert ctypes.sizeof(_lib.AugmentParams) == 4 * (2 + 6 + 1 + 6 + 1 + 3 + 3) + 2 * 224 * 2


def test_attention_trace_variant_compiles(tmp_path):
    """tools/gpu_attn_trace.py builds a second library with -DDFU_ATTN_TRACE (phase clocks of the attention kernels); the
    instrumented source must keep compiling for sm_100a -- it is otherwise only built when someone profiles."""
    import subprocess
    from dfu_multimodal_b200 import build as B
    flags = [f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    obj = str(tmp_path / "attention_trace.o")
    subprocess.check_call([B._nvcc()] + flags + ["-DDFU_ATTN_TRACE", "-c", os.path.join(B.CSRC, "attention.cu"), "-o", obj])
    sym = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    assert "dfu_attn_trace_read_host" in sym
