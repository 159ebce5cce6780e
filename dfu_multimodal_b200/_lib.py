"""ctypes binding of libdfu_b200.so.  The prototypes are read from include/dfu_b200.h, so the header is the single
source of truth for the C ABI (tests/test_abi.py checks every declared symbol is exported)."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.normpath(os.path.join(HERE, "..", "include", "dfu_b200.h"))
LIB_PATH = os.path.join(HERE, "libdfu_b200.so")


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("nb0", ctypes.c_int32), ("nb1", ctypes.c_int32),
        ("A", ctypes.c_void_p), ("a_major", ctypes.c_int32), ("lda", ctypes.c_int64), ("sa0", ctypes.c_int64), ("sa1", ctypes.c_int64),
        ("B", ctypes.c_void_p), ("b_major", ctypes.c_int32), ("ldb", ctypes.c_int64), ("sb0", ctypes.c_int64), ("sb1", ctypes.c_int64),
        ("D", ctypes.c_void_p), ("d_dtype", ctypes.c_int32), ("ldd", ctypes.c_int64), ("sd0", ctypes.c_int64), ("sd1", ctypes.c_int64),
        ("accumulate", ctypes.c_int32),
        ("aux", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
        ("res", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("sr0", ctypes.c_int64), ("sr1", ctypes.c_int64),
        ("alpha", ctypes.c_float),
        ("act", ctypes.c_int32),
        ("impl", ctypes.c_int32),
        ("max_ctas", ctypes.c_int32),
        ("colstats", ctypes.c_void_p),
        ("colstats_mode", ctypes.c_int32),
    ]


class ConvArgs(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int32), ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("C", ctypes.c_int32),
        ("Cout", ctypes.c_int32), ("R", ctypes.c_int32), ("S", ctypes.c_int32), ("stride", ctypes.c_int32), ("pad", ctypes.c_int32),
        ("x", ctypes.c_void_p), ("w", ctypes.c_void_p), ("ldw", ctypes.c_int64), ("dy", ctypes.c_void_p),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int64), ("accumulate", ctypes.c_int32), ("max_ctas", ctypes.c_int32),
        ("colstats", ctypes.c_void_p),
        ("bias", ctypes.c_void_p), ("act", ctypes.c_int32),
    ]


class V32GemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("nb0", ctypes.c_int32), ("nb1", ctypes.c_int32),
        ("A", ctypes.c_void_p), ("a_row_stride", ctypes.c_int64), ("a_col_stride", ctypes.c_int64), ("sa0", ctypes.c_int64), ("sa1", ctypes.c_int64),
        ("B", ctypes.c_void_p), ("b_row_stride", ctypes.c_int64), ("b_col_stride", ctypes.c_int64), ("sb0", ctypes.c_int64), ("sb1", ctypes.c_int64),
        ("D", ctypes.c_void_p), ("ldd", ctypes.c_int64), ("sd0", ctypes.c_int64), ("sd1", ctypes.c_int64),
        ("bias", ctypes.c_void_p),
        ("res", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("sr0", ctypes.c_int64), ("sr1", ctypes.c_int64),
        ("alpha", ctypes.c_float),
        ("act", ctypes.c_int32),
    ]


class AugmentParams(ctypes.Structure):
    _fields_ = [("flags", ctypes.c_int32), ("rot_kind", ctypes.c_int32), ("rot", ctypes.c_int32 * 6), ("aff_kind", ctypes.c_int32),
                ("aff", ctypes.c_int32 * 6), ("n_color", ctypes.c_int32), ("color_op", ctypes.c_int32 * 3), ("color_f", ctypes.c_float * 3),
                ("aff_xtab", ctypes.c_int16 * 224), ("aff_ytab", ctypes.c_int16 * 224)]


class AdamWTensor(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("shadow", ctypes.c_void_p), ("numel", ctypes.c_int64)]


class CastTensor(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("numel", ctypes.c_int64)]


_SCALARS = {
    "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "float": ctypes.c_float, "double": ctypes.c_double,
    "int": ctypes.c_int,
}


def parse_header(path=HEADER):
    """Returns {name: (restype, [argtypes])} for every `int dfu_*(...)` / `const char* dfu_*(...)` prototype."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int)\s+(dfu_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const ", "").split()[0]
                    argtypes.append(_SCALARS[ty])
        protos[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes)
    return protos


_lib = None


def load():
    """Loads the library (never builds, never falls back). Raises if it is missing: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libdfu_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "this package has no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, argtypes) in parse_header().items():
        fn = getattr(lib, name)   # AttributeError if the header declares something the library lacks
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().dfu_last_error()
        raise RuntimeError("libdfu_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
