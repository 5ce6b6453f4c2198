"""Developer probe (not the contract bench): SpMV + one (possibly truncated) solve with per-kernel-class timing.
env: QB_MAXIT (restart cap), QB_NOPROF=1 (skip the profiled solve), SB200_XSLICE_MB / SB200_SPMV_LANES (library knobs)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import spectra_b200 as sb
from spectra_b200 import synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
maxit = int(os.environ.get("QB_MAXIT", "1000"))
nev, ncv = 20, 60
t = time.time()
rp, ci, v = synth.csr(n, 20, 0, True)
gen_s = time.time() - t
t = time.time()
op = sb.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
up_s = time.time() - t
ms = op.spmv_device_time(20)
bytes_spmv = 12 * len(ci) + 4 * (n + 1) + 16 * n
tag = dict(n=n, xslice=os.environ.get("SB200_XSLICE_MB"), lanes=os.environ.get("SB200_SPMV_LANES"), maxit=maxit)
print(json.dumps(dict(tag, kernel="spmv_plain", gen_s=round(gen_s, 2), upload_s=round(up_s, 2), ms=ms, gbs=bytes_spmv / ms / 1e6, frac=bytes_spmv / ms / 1e6 / 6571.2)), flush=True)
for prof in ((0,) if os.environ.get("QB_NOPROF") else (0, 1)):
    sb.set_profiling(prof)
    eigs = sb.SymEigsSolver(op, nev, ncv)
    t = time.time()
    eigs.init()
    nconv = eigs.compute(sb.SortRule.LargestAlge, maxit)
    wall = time.time() - t
    st = eigs.stats()
    out = dict(tag, profiling=prof, wall_s=round(wall, 3), ms_total=round(st["ms_total"], 1), nconv=nconv, nops=eigs.num_operations(), niter=eigs.num_iterations(),
               info=int(eigs.info()), iters_per_s=round(eigs.num_operations() / (st["ms_total"] / 1e3), 1), launches=st["kernel_launches"],
               reorth=st["reorth_passes"])
    if prof:
        pb = 8 * n * (st["panel_cols"] + 2 * st["panel_launches"])
        out.update(ms_spmv=round(st["ms_spmv"], 1), ms_panel=round(st["ms_panel"], 1), ms_compress=round(st["ms_compress"], 1), ms_small=round(st["ms_small"], 1),
                   panel_gbs=round(pb / st["ms_panel"] / 1e6),
                   spmv_gbs=round(((bytes_spmv + 16 * n) * st["spmv_launches"] + 8 * n * st["fused_dot_cols"]) / st["ms_spmv"] / 1e6),
                   fused=st["fused_dot_launches"], fused_cols=st["fused_dot_cols"], panel_launches=st["panel_launches"], host_syncs=st["host_syncs"],
                   avg_panel_ms=round(st["ms_panel"] / st["panel_launches"], 4), avg_spmv_ms=round(st["ms_spmv"] / st["spmv_launches"], 4),
                   avg_small_ms=round(st["ms_small"] / max(st["restarts"] + 1, 1), 3), avg_compress_ms=round(st["ms_compress"] / max(st["compress_launches"], 1), 3))
    print(json.dumps(out), flush=True)
