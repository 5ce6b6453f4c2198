"""Writes tests/golden/eigen_interop_reference.txt: the output of tests/cpp/test_eigen_interop.cpp compiled against THE REFERENCE
(/root/reference/include, header-only) over the Eigen stand-in (oracle/eigen_standin).  The same source file compiled against this
repository's include/ must print the same results (tests/test_cpp_shim.py).  Needs /root/reference: development container only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_and_run_reference(exe=os.path.join(ROOT, "tests", "_build", "test_eigen_interop_reference")):
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle", "eigen_standin"), "-I", "/root/reference/include",
           os.path.join(ROOT, "tests", "cpp", "test_eigen_interop.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    return r.stdout


if __name__ == "__main__":
    out = build_and_run_reference()
    path = os.path.join(HERE, "eigen_interop_reference.txt")
    with open(path, "w") as f:
        f.write(out)
    sys.stdout.write(out)
    print("wrote", path)
