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
    ]


def lib_path() -> Path:
    return _LIB_PATH


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
    lib.ovla_last_error.restype = C.c_char_p
    lib.ovla_launch_count.restype = C.c_longlong
    lib.ovla_reset_launch_count.restype = None
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
