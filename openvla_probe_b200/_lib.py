"""ctypes binding of libovla_b200.so (C ABI declared in include/ovla_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libovla_b200.so"
if os.environ.get("OVLA_B200_LIB"):  # A/B runs of two builds of the same library inside one process tree
    _LIB_PATH = Path(os.environ["OVLA_B200_LIB"])
_lib = None


class OvlaError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias_bf16", C.c_void_p),
        ("scale_bf16", C.c_void_p),
        ("resid_bf16", C.c_void_p),
        ("ld_resid", C.c_longlong),
        ("bias_f32", C.c_void_p),
        ("gelu", C.c_int),
        ("round_bf16", C.c_int),
        ("row_sumsq_out", C.c_void_p),
        ("row_sumsq_in", C.c_void_p),
        ("row_sumsq_ld", C.c_int),
        ("row_sumsq_parts", C.c_int),
        ("norm_eps", C.c_float),
    ]


def lib_path() -> Path:
    return _LIB_PATH


_HEADER = Path(__file__).resolve().parent.parent / "include" / "ovla_b200.h"
_SCALARS = {"int": C.c_int, "float": C.c_float, "double": C.c_double, "long long": C.c_longlong,
            "unsigned long long": C.c_ulonglong, "void": None}


def _ctype_of(decl: str):
    """C parameter / return declaration -> ctypes type.  Every pointer is passed as c_void_p (which accepts ints, None,
    c_void_p, ctypes arrays and byref() results), `const char*` as c_char_p."""
    d = " ".join(decl.replace("*", " * ").split())
    if "*" in d:
        return C.c_char_p if d.startswith("const char *") and d.count("*") == 1 else C.c_void_p
    words = [w for w in d.split() if w != "const"]
    for n in (2, 1):                         # "long long x" / "int x" / bare "void"
        for cut in (len(words) - 1, len(words)):
            key = " ".join(words[:cut][:n]) if cut >= 1 else ""
            if key in _SCALARS and len(words[:cut]) == n:
                return _SCALARS[key]
    raise OvlaError(f"cannot map C declaration {decl!r} from {_HEADER.name}")


def header_prototypes() -> dict:
    """{function name: (restype, [argtypes])} parsed from include/ovla_b200.h -- the single source of truth for the
    ABI, so that a 64-bit argument can never be passed as a 32-bit int by a call site that forgot to wrap it."""
    import re

    txt = re.sub(r"/\*.*?\*/", " ", _HEADER.read_text(), flags=re.S)
    txt = re.sub(r"^\s*#.*$", " ", txt, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(ovla_\w+)\s*\(([^()]*)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        ret = ret.split("{")[-1].split(";")[-1].strip()
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        out[name] = (_ctype_of(ret + " x") if "*" in ret else _SCALARS[" ".join(w for w in ret.split() if w != "const")], argtypes)
    return out


def load() -> C.CDLL:
    """Load the native library (building is the job of ``__graft_entry__.build`` / ``openvla_probe_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise OvlaError(
            f"{_LIB_PATH} is missing: build it with `python -m openvla_probe_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)"
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (restype, argtypes) in header_prototypes().items():
        fn = getattr(lib, name)           # AttributeError = declared in the header but not exported: fail loudly
        fn.restype, fn.argtypes = restype, argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ovla_last_error()
        raise OvlaError(msg.decode() if msg else "libovla_b200 call failed")


def ptr(t) -> C.c_void_p:
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
