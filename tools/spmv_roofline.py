"""SpMV roofline study: every kernel variant on two column distributions at equal nnz/row.

  uniform : G_sym(n, 20)      every gathered x entry is a distinct 32 B sector anywhere in x (benchmark workload, BASELINE C2/C4)
  band    : G_band(n, b=10)   21 nnz/row, gathered x entries contiguous (stencil / mesh-like matrices, BASELINE C5's class)

  variants: csr (sub-warp per row, default), sell256 / sell512 / sell1024 (sliced layout, one lane per row, CTA size)

Algorithmic bytes per SpMV: 12 nnz + 4 (n+1) + 16 n.   Prints one JSON line per case.
usage: python tools/spmv_roofline.py [n] [variants, comma separated]
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
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["csr", "sell256", "sell512", "sell1024"]
for name, make in (("uniform_G_sym_d20", lambda: synth.csr(n, 20, 0, True)), ("band_b10", lambda: synth.band_csr(n, 10, 0, 0.0))):
    rp, ci, v = make()
    nnz = len(ci)
    x = np.random.default_rng(0).standard_normal(n)
    y_ref = None
    for var in variants:
        os.environ.pop("SB200_SPMV_FORMAT", None)
        if var.startswith("sell"):
            os.environ["SB200_SPMV_FORMAT"] = "sell"
            os.environ["SB200_SELL_THREADS"] = var[4:] or "512"
        op = sb.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
        lay = op.spmv_layout()
        ms = op.spmv_device_time(20)
        y = op.perform_op(x)
        if y_ref is None:
            y_ref = y
        err = float(np.abs(y - y_ref).max() / np.abs(y_ref).max())
        b = 12 * nnz + 4 * (n + 1) + 16 * n
        print(json.dumps(dict(case=name, variant=var, layout=lay, fill=round(lay["stored_entries"] / nnz, 4), n=n, nnz=nnz, ms=round(ms, 4),
                              gbs=round(b / ms / 1e6, 1), frac=round(b / ms / 1e6 / PEAK, 3), peak_gbs=PEAK, rel_diff_vs_first=err)), flush=True)
        op.close()
    del rp, ci, v
