"""One-process probe of the code paths that had not run on a GPU when round 1's budget ended: prints ONE JSON line.

bench.py runs it in a separate process (own CUDA context, hard timeout) after its timed work is finished, and attaches the result
under "experimental"; nothing here enters `value`, `e2e` or `roofline`.  Every block is independent and failure-tolerant.

  gather_roof   time of 1e8 uniformly random 8-byte gathers from a 40 MB slice and from the whole operand (csrc/microbench.cu)
  spmv          CSR (default) vs sliced-CSR kernels (SB200_SPMV_FORMAT=sell, CTA sizes 256/512/1024) on the benchmark matrix G_sym and
                on a band matrix: ms, fraction of the HBM roof (algorithmic bytes), fill, difference of the result to the CSR kernel's
  sell_solve    a truncated Lanczos run (first restarts) through the fused sliced step kernel: operations / s next to the default path
  herm          SparseHermMatProd + HermEigsSolver on a random sparse Hermitian matrix: residuals and time
  complex_gen   complex GenEigsSolver (complex Arnoldi + the one-warp complex restart kernel) on a random sparse complex matrix

usage: python tools/experimental_probe.py [n]
"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np


def main():
    out = {"note": "opt-in code paths, first executed on a GPU by this probe; not part of value / e2e / roofline"}
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    import spectra_b200 as sb
    from spectra_b200 import synth

    peak = 6571.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out["n"] = n
    out["hbm_peak_gbs"] = peak

    def block(name, fn):
        t = time.time()
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}
        out.setdefault("seconds", {})[name] = round(time.time() - t, 2)

    def gather_roof():
        res = []
        for sl in (n // 2, n):
            r = sb.bench_gather(sl, 100_000_000, 3)
            res.append({"slice_doubles": sl, "ms_per_1e8_gathers": round(r["ms"], 4), "gsectors_per_s": round(1e8 / r["ms"] / 1e6, 1),
                        "checksum": r["checksum"]})
        return res

    def spmv_variants():
        res = []
        for case, make in (("uniform_G_sym_d20", lambda: synth.csr(n, 20, 0, True)), ("band_b10", lambda: synth.band_csr(n, 10, 0, 0.0))):
            rp, ci, v = make()
            nnz = len(ci)
            alg = 12 * nnz + 4 * (n + 1) + 16 * n
            x = np.random.default_rng(0).standard_normal(n)
            y_ref = None
            for var in ("csr", "csr_blocks", "sell256", "sell512", "sell1024"):
                os.environ.pop("SB200_SPMV_FORMAT", None)
                os.environ.pop("SB200_SPMV_GRID", None)
                if var == "csr_blocks":
                    os.environ["SB200_SPMV_GRID"] = "blocks"  # one CTA per 256-row block instead of the persistent grid
                if var.startswith("sell"):
                    os.environ["SB200_SPMV_FORMAT"] = "sell"
                    os.environ["SB200_SELL_THREADS"] = var[4:]
                try:
                    op = sb.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
                    lay = op.spmv_layout()
                    ms = op.spmv_device_time(20)
                    y = op.perform_op(x)
                    if y_ref is None:
                        y_ref = y
                    res.append({"case": case, "variant": var, "format": lay["format"], "col_blocks": lay["col_blocks"], "fill": round(lay["stored_entries"] / nnz, 4),
                                "ms": round(ms, 4), "frac_of_hbm_roof": round(alg / ms / 1e6 / peak, 3),
                                "rel_diff_vs_csr": float(np.abs(y - y_ref).max() / np.abs(y_ref).max())})
                    op.close()
                except Exception as e:  # noqa: BLE001
                    res.append({"case": case, "variant": var, "error": f"{type(e).__name__}: {e}"})
            os.environ.pop("SB200_SPMV_FORMAT", None)
            os.environ.pop("SB200_SPMV_GRID", None)
            del rp, ci, v
        # the same comparison at BASELINE config C2's size, where the persistent grid has only 3.3 rounds
        if n > 1_000_000:
            n2 = 1_000_000
            rp, ci, v = synth.csr(n2, 20, 0, True)
            alg = 12 * len(ci) + 4 * (n2 + 1) + 16 * n2
            for var in ("csr", "csr_blocks", "sell512"):
                os.environ.pop("SB200_SPMV_FORMAT", None)
                os.environ.pop("SB200_SPMV_GRID", None)
                if var == "csr_blocks":
                    os.environ["SB200_SPMV_GRID"] = "blocks"
                if var.startswith("sell"):
                    os.environ["SB200_SPMV_FORMAT"] = "sell"
                    os.environ["SB200_SELL_THREADS"] = var[4:]
                try:
                    op = sb.SparseGenMatProd.from_csr_slab(n2, 0, rp, ci, v)
                    ms = op.spmv_device_time(50)
                    res.append({"case": "uniform_G_sym_d20_n1e6", "variant": var, "ms": round(ms, 4), "frac_of_hbm_roof": round(alg / ms / 1e6 / peak, 3)})
                    op.close()
                except Exception as e:  # noqa: BLE001
                    res.append({"case": "uniform_G_sym_d20_n1e6", "variant": var, "error": f"{type(e).__name__}: {e}"})
            os.environ.pop("SB200_SPMV_FORMAT", None)
            os.environ.pop("SB200_SPMV_GRID", None)
        return res

    def sell_solve():
        rp, ci, v = synth.csr(n, 20, 0, True)
        res = {}
        for var in ("csr", "sell"):
            os.environ.pop("SB200_SPMV_FORMAT", None)
            if var == "sell":
                os.environ["SB200_SPMV_FORMAT"] = "sell"
            op = sb.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
            eigs = sb.SymEigsSolver(op, 20, 60)
            eigs.init()
            eigs.compute(sb.SortRule.LargestAlge, 8)  # 8 restarts: ~350 operations
            st = eigs.stats()
            res[var] = {"format": op.spmv_layout()["format"], "nops": eigs.num_operations(), "ms_total": round(st["ms_total"], 1),
                        "ops_per_s": round(eigs.num_operations() / (st["ms_total"] / 1e3), 1)}
            ev = eigs.eigenvalues()
            res[var]["nconv_so_far"] = int(len(ev))
            del eigs
            op.close()
        os.environ.pop("SB200_SPMV_FORMAT", None)
        return res

    def herm():
        import scipy.sparse as sp

        m = min(n, 300_000)
        rng = np.random.default_rng(0)
        cnt = 10 * m
        r, c = rng.integers(0, m, cnt), rng.integers(0, m, cnt)
        vals = (rng.random(cnt) - 0.5) + 1j * (rng.random(cnt) - 0.5)
        lo = r > c
        L = sp.csc_matrix((vals[lo], (r[lo], c[lo])), shape=(m, m))
        L.sum_duplicates()
        D = sp.diags(rng.random(m) - 0.5).astype(np.complex128)
        A = (L + D).tocsc()
        Af = (L + L.conj().T + D).tocsr()
        op = sb.SparseHermMatProd(A)
        x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
        y = op.perform_op(x)
        err = float(np.abs(y - Af @ x).max() / np.abs(y).max())
        eigs = sb.HermEigsSolver(op, 10, 30)
        t = time.time()
        eigs.init()
        nconv = eigs.compute(sb.SortRule.LargestAlge)
        wall = time.time() - t
        ev, U = eigs.eigenvalues(), eigs.eigenvectors()
        resid = float((np.linalg.norm(Af @ U - U * ev, axis=0) / np.abs(ev)).max()) if len(ev) else None
        st = eigs.stats()
        return {"n": m, "nnz": int(Af.nnz), "spmv_rel_err": err, "nconv": int(nconv), "nops": eigs.num_operations(), "wall_s": round(wall, 3),
                "ms_total": round(st["ms_total"], 1), "max_rel_residual": resid}

    def complex_gen():
        import scipy.sparse as sp

        m = min(n, 200_000)
        rng = np.random.default_rng(1)
        cnt = 20 * m
        A = sp.csc_matrix(((rng.random(cnt) - 0.5) + 1j * (rng.random(cnt) - 0.5), (rng.integers(0, m, cnt), rng.integers(0, m, cnt))), shape=(m, m))
        A.sum_duplicates()
        op = sb.SparseHermMatProd(A, uplo="general")
        g = sb.GenEigsSolver(op, 6, 30)
        t = time.time()
        g.init()
        nconv = g.compute(sb.SortRule.LargestMagn, 300)
        wall = time.time() - t
        ev, Z = g.eigenvalues(), g.eigenvectors()
        resid = float((np.linalg.norm(A @ Z - Z * ev, axis=0) / np.abs(ev)).max()) if len(ev) else None
        return {"n": m, "nnz": int(A.nnz), "nconv": int(nconv), "nops": g.num_operations(), "niter": g.num_iterations(), "wall_s": round(wall, 3),
                "max_rel_residual": resid}

    block("gather_roof", gather_roof)
    block("spmv", spmv_variants)
    block("sell_solve", sell_solve)
    block("herm", herm)
    block("complex_gen", complex_gen)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"error": f"{type(e).__name__}: {e}"}), flush=True)
