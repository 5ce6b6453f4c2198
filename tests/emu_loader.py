"""Loads the kernel-logic emulation build (tools/cuda_emu/, TEST INFRASTRUCTURE ONLY) behind a private copy of the Python mirror.

`load()` returns a module object with the API of `spectra_b200` whose ctypes handle points at tests/_emu/libspectra_b200_emu.so
-- the product's own .cu sources compiled for the CPU against the CUDA execution model of cuda_emu.h.  The product package is not
touched and knows nothing about this library: the copy is made here, by executing spectra_b200/__init__.py under another module
name and re-pointing its library path.  Used by the `-m "not gpu"` emulation tests to check kernel logic (indexing, barriers,
shuffles, host sequencing) on machines without a GPU; parity claims rest on the `-m gpu` tests, not on these.
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_mod = None


def load():
    global _mod
    if _mod is not None:
        return _mod
    sys.path.insert(0, os.path.join(ROOT, "tools", "cuda_emu"))
    import emu_build

    lib = emu_build.build()
    pkg = os.path.join(ROOT, "spectra_b200")
    spec = importlib.util.spec_from_file_location("spectra_b200_emu", os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["spectra_b200_emu"] = mod
    spec.loader.exec_module(mod)
    mod._LIB_PATH = lib
    mod._lib = None
    assert "cuda_emu" not in open(os.path.join(pkg, "__init__.py")).read()
    _mod = mod
    return mod
