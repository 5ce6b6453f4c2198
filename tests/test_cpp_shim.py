"""The header-only C++ shim (include/Spectra/) compiles against the C ABI like the reference's headers
would (CPU check), and its test program — the reference's own test flow — passes on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "test_shim")


def _compile():
    import spectra_b200 as sb

    if not os.path.exists(sb.lib_path()):
        from spectra_b200 import _build

        _build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(sb.lib_path())
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"), "-L",
           libdir, "-lspectra_b200", f"-Wl,-rpath,{libdir}", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_shim_compiles_and_links():
    _compile()


@pytest.mark.gpu
def test_shim_reference_flow_on_gpu(gpu):
    exe = _compile()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout


# ---- complex Hermitian shim: compiled on the CPU, run against the kernel-logic emulator (tests/emu_loader.py) ----
HERM_EXE = os.path.join(ROOT, "tests", "_build", "test_shim_herm")


def _compile_herm(libdir, libname, src="test_shim_herm.cpp", defines=()):
    os.makedirs(os.path.dirname(HERM_EXE), exist_ok=True)
    exe = os.path.join(ROOT, "tests", "_build", src[:-4] + "_" + libname)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", *[f"-D{d}" for d in defines], "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", src),
           "-L", libdir, "-l" + libname, f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_herm_shim_reference_flow_on_emulator(emu):
    # test/HermEigs.cpp's sparse flow through include/Spectra/HermEigsSolver.h; the kernels run on the CPU execution model
    libdir = os.path.join(ROOT, "tests", "_emu")
    exe = _compile_herm(libdir, "spectra_b200_emu", defines=("SB200_SHIM_TEST_SMALL",))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout


def test_float_shim_flow_on_emulator(emu):
    # Scalar = float wrappers / solvers / user operator (float at the boundary, fp64 on the device)
    exe = _compile_herm(os.path.join(ROOT, "tests", "_emu"), "spectra_b200_emu", "test_shim_float.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout


def test_shim_reference_flow_small_on_emulator(emu):
    # the double-precision shim test program (user operators, shift-solve, sym / gen solvers, exception types) with its small cases
    exe = _compile_herm(os.path.join(ROOT, "tests", "_emu"), "spectra_b200_emu", "test_shim.cpp", defines=("SB200_SHIM_TEST_SMALL",))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=1500)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout
