#!/bin/bash
# The reference's own unit tests, complete runs (no test-name filters): (A) reference headers over the Eigen stand-in, (B) this repository's
# shim on the kernel-logic emulator (slow: the 1000 x 1000 cases take 10-30 minutes each there).  Development container only (/root/reference).
# usage: tools/run_reference_unit_tests.sh [A|B|AB] > profiles/r2_reference_unit_tests.log
set -u
WHICH=${1:-AB}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/tests/_build/reference_unit_tests_full
mkdir -p "$B"
python -c "import sys; sys.path.insert(0, '$ROOT/tests'); import emu_loader; emu_loader.load()" >/dev/null
g++ -std=c++17 -O1 -I/root/reference/test -c /root/reference/test/tests-main.cpp -o "$B/main.o"
run() { # side name include-dir [libs...]
  local side=$1 name=$2 inc=$3; shift 3
  g++ -std=c++17 -O2 -ffp-contract=off -I"$ROOT/oracle/eigen_standin" -I"$inc" -I/root/reference/test /root/reference/test/$name.cpp "$B/main.o" "$@" -o "$B/${side}_$name" 2>&1 | grep -E "error" | head -3
  local s=$(date +%s); local out=$(timeout 3000 "$B/${side}_$name" 2>&1 | tail -3 | tr '\n' ' '); local e=$(date +%s)
  echo "[$side] $name ($((e - s)) s): $out"
}
if [[ $WHICH == *A* ]]; then
  echo "== (A) reference headers over oracle/eigen_standin"
  for t in Givens QR Schur Arnoldi SparseSymMatProd SparseGenMatProd Example1 Example2 Example4 SymEigsShift SymEigs GenEigs HermEigs ComplexEigs; do
    run ref $t /root/reference/include
  done
fi
if [[ $WHICH == *B* ]]; then
  echo "== (B) this repository's include/ + the emulator build of libspectra_b200"
  for t in SparseSymMatProd SparseGenMatProd Example1 Example2 Example4 HermEigs SymEigs SymEigsShift GenEigs ComplexEigs; do
    run shim $t "$ROOT/include" -L"$ROOT/tests/_emu" -lspectra_b200_emu -Wl,-rpath,"$ROOT/tests/_emu"
  done
fi
