"""Builds libspectra_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m spectra_b200._build [--force]

Objects go to spectra_b200/csrc/_obj/, the library to spectra_b200/lib/libspectra_b200.so (git-ignored,
but it travels with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libspectra_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "spectra_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, verbose: bool, objdir: str = OBJ, defines=()):
    obj = os.path.join(objdir, src[:-3] + ".o")
    cmd = [NVCC, *ARCH, *CFLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return src, r.stderr


def build(force: bool = False, verbose: bool = False, defines=(), suffix: str = "") -> str:
    """suffix/defines build an experimental variant (lib/libspectra_b200<suffix>.so, selected at run time with
    SB200_LIB_SUFFIX); the default build has neither."""
    objdir = OBJ + suffix
    lib = LIB if not suffix else os.path.join(LIBDIR, f"libspectra_b200{suffix}.so")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hm = _headers_mtime()
    todo = []
    for src in _sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        sm = max(os.path.getmtime(os.path.join(CSRC, src)), hm)
        if force or verbose or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(src)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, log in ex.map(lambda s: _compile(s, verbose, objdir, defines), todo):
                if verbose and log:
                    print(f"== {src}\n{log}")
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in _sources()]
    if todo or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", lib, *objs, "-ldl", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    suf = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--suffix=")), "")
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, suffix=suf))
