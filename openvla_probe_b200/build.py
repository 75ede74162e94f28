"""Build libovla_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m openvla_probe_b200.build [--force]

Objects go to openvla_probe_b200/csrc/_build/, the shared library to openvla_probe_b200/libovla_b200.so
(git-ignored, but shipped to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT = PKG / "libovla_b200.so"
BUILD = CSRC / "_build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path, extra=()) -> str:
    h = hashlib.sha256()
    h.update(" ".join([*NVCC_FLAGS, *extra]).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "ovla_b200.h"]):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, force: bool, bdir: Path = BUILD, extra=()) -> Path:
    obj = bdir / (src.stem + ".o")
    stamp = bdir / (src.stem + ".sha")
    dig = _digest(src, extra)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return obj


def build(force: bool = False, verbose: bool = True, defines=(), variant: str = "") -> Path:
    """`defines` / `variant`: an A/B build with extra -D flags, linked as libovla_b200_<variant>.so (selected at run
    time with OVLA_B200_LIB); the default build takes neither."""
    bdir = BUILD / variant if variant else BUILD
    out = PKG / f"libovla_b200_{variant}.so" if variant else OUT
    extra = [f"-D{d}" for d in defines]
    bdir.mkdir(parents=True, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, bdir, extra), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    OUT_ = out
    if force or not OUT_.exists() or OUT_.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(OUT_), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[ovla build] linked {OUT_} from {len(objs)} objects")
    elif verbose:
        print(f"[ovla build] {OUT_.name} up to date")
    return OUT_


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    build(force="--force" in sys.argv, defines=_defs, variant=_var)
