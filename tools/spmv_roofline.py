"""SpMV roofline study: the same kernel on three column distributions at equal nnz/row.

  uniform : G_sym(n, 20)      every gathered x entry is a distinct 32 B sector anywhere in x (benchmark workload, BASELINE C2/C4)
  band    : G_band(n, b=10)   21 nnz/row, gathered x entries contiguous (stencil / mesh-like matrices, BASELINE C5's class)

Algorithmic bytes per SpMV: 12 nnz + 4 (n+1) + 16 n.   Prints one JSON line per case.
usage: python tools/spmv_roofline.py [n]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import spectra_b200 as sb
from spectra_b200 import synth

PEAK = 6571.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
for name, make in (("uniform_G_sym_d20", lambda: synth.csr(n, 20, 0, True)), ("band_b10", lambda: synth.band_csr(n, 10, 0, 0.0))):
    rp, ci, v = make()
    op = sb.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    nnz = len(ci)
    del rp, ci, v
    ms = op.spmv_device_time(20)
    b = 12 * nnz + 4 * (n + 1) + 16 * n
    print(json.dumps(dict(case=name, n=n, nnz=nnz, ms=round(ms, 4), gbs=round(b / ms / 1e6, 1), frac=round(b / ms / 1e6 / PEAK, 3), peak_gbs=PEAK)), flush=True)
    op.close()
