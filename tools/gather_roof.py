"""The gather roof of a CSR SpMV with uniformly random column ids (sb200_bench_gather, csrc/microbench.cu).

For each operand-slice size: time of 1e8 independent random 8-byte loads (the gathers of ONE column block of the n = 1e7,
20 nnz/row benchmark matrix), the 32 B-sector rate it corresponds to, and what that alone allows the whole SpMV (2 blocks = 2e8
gathers, 2.6 GB algorithmic) to reach against the measured HBM peak.  usage: python tools/gather_roof.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import spectra_b200 as sb

PEAK = 6571.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

G = 100_000_000
for n in (125_000, 1_250_000, 5_000_000, 10_000_000, 40_000_000):
    r = sb.bench_gather(n, G, 5)
    ms = r["ms"]
    spmv_ms = 2 * ms  # two column blocks of 1e8 gathers each
    print(json.dumps(dict(slice_doubles=n, slice_mb=round(8 * n / 2**20, 1), gathers=G, ms=round(ms, 4), gsectors_per_s=round(G / ms / 1e6, 1),
                          l2_sector_gbs=round(32 * G / ms / 1e6, 1), spmv_gather_only_ms=round(spmv_ms, 4),
                          spmv_frac_ceiling=round(2.6e9 / (spmv_ms * 1e-3) / 1e9 / PEAK, 3), checksum=r["checksum"])), flush=True)

# the same gathers fed by the (index, value) stream of a sliced SpMV (12 B per gather, coalesced), uniform and band columns: the floor of
# ANY SpMV kernel on this access pattern -- two 40 MB slices of 1e8 entries make one n = 1e7, 20 nnz/row operator application
for band in (False, True):
    for n in (1_250_000, 5_000_000):
        r = sb.bench_stream_gather(n, G, band, 5)
        ms = r["ms"]
        print(json.dumps(dict(kind="stream+gather", columns="band" if band else "uniform", slice_doubles=n, entries=G, ms=round(ms, 4),
                              stream_gbs=round(12 * G / ms / 1e6, 1), spmv_floor_ms=round(2 * ms, 4),
                              spmv_frac_ceiling=round(2.6e9 / (2 * ms * 1e-3) / 1e9 / PEAK, 3), checksum=r["checksum"])), flush=True)
