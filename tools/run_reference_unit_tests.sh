#!/bin/bash
# The reference's own unit tests, complete runs (no test-name filters): (A) reference headers over the Eigen stand-in, (B) this repository's
# shim on the kernel-logic emulator.  Development container only (/root/reference).  usage: tools/run_reference_unit_tests.sh > profiles/r2_reference_unit_tests.log
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/tests/_build/reference_unit_tests_full
mkdir -p "$B"
python -c "import sys; sys.path.insert(0, '$ROOT/tests'); import emu_loader; emu_loader.load()" >/dev/null
g++ -std=c++17 -O1 -I/root/reference/test -c /root/reference/test/tests-main.cpp -o "$B/main.o"
run() { # side name flags...
  local side=$1 name=$2; shift 2
  g++ -std=c++17 -O2 -ffp-contract=off -I"$ROOT/oracle/eigen_standin" "$@" -I/root/reference/test /root/reference/test/$name.cpp "$B/main.o" -o "$B/${side}_$name" 2>&1 | grep -E "error" | head -3
  local s=$(date +%s); local out=$("$B/${side}_$name" 2>&1 | tail -3 | tr '\n' ' '); local e=$(date +%s)
  echo "[$side] $name ($((e - s)) s): $out"
}
echo "== (A) reference headers over oracle/eigen_standin"
for t in Givens QR Schur Arnoldi SparseSymMatProd SparseGenMatProd Example1 Example2 Example4 SymEigsShift SymEigs GenEigs HermEigs ComplexEigs; do
  run ref $t -I/root/reference/include
done
echo "== (B) this repository's include/ + the emulator build of libspectra_b200"
for t in SparseSymMatProd SparseGenMatProd Example2 Example4 SymEigs GenEigs; do
  run shim $t -I"$ROOT/include" -L"$ROOT/tests/_emu" -lspectra_b200_emu -Wl,-rpath,"$ROOT/tests/_emu"
done
