"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py

Row-sharded SymEigsSolver / GenEigsSolver against the CPU oracle on the same synthetic matrix: eigenvalues within
1e-10 relative, ||A x - lambda x|| / |lambda| <= 1e-10, identical operation counts on every rank.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp

import spectra_b200 as sb
from spectra_b200 import dist, synth


def main():
    rank, local_rank, world = dist.env_rank()
    sb.set_device(local_rank)
    dist.init_process_group("gloo")
    comm = dist.make_comm()
    ok = True
    for n, nev, ncv in ((200_003, 10, 30), (1_000_000, 20, 60)):
        row0, nrows = dist.slab_range(n, rank, world)
        rp, ci, v = synth.csr(n, 20, 0, True, row0=row0, nrows=nrows)
        op = sb.SparseGenMatProd.from_csr_slab(n, row0, rp, ci, v, comm=comm)
        # operator tier: y = A x on the local rows
        x = np.cos(np.arange(n) * 0.37)
        y = op.perform_op(x)
        Aloc = sp.csr_matrix((v, ci, rp), shape=(nrows, n))
        err_op = np.abs(y - Aloc @ x).max() / max(1.0, np.abs(y).max())
        eigs = sb.SymEigsSolver(op, nev, ncv)
        t = time.time()
        eigs.init()
        nconv = eigs.compute(sb.SortRule.LargestAlge)
        wall = time.time() - t
        evals = eigs.eigenvalues()
        Xl = eigs.eigenvectors(local=True)
        Xf = eigs.eigenvectors()
        num = np.array([dist.sum_over_ranks(float(np.sum((Aloc @ Xf[:, c] - Xl[:, c] * evals[c]) ** 2))) for c in range(len(evals))])
        res = np.sqrt(num) / np.abs(evals)
        st = eigs.stats()
        line = dict(n=n, rank=rank, world=world, nconv=nconv, info=int(eigs.info()), nops=eigs.num_operations(), niter=eigs.num_iterations(), err_op=err_op,
                    max_res=float(res.max()), wall=round(wall, 3), ms_total=round(st["ms_total"], 1))
        good = eigs.info() == sb.CompInfo.Successful and nconv == nev and err_op <= 1e-13 and res.max() <= 1e-10
        if rank == 0 and n <= 300_000:
            import oracle as O

            rpf, cif, vf = synth.csr(n, 20, 0, True)
            ref = O.sym_eigs(O.Csr.adopt(n, rpf, cif, vf), nev, ncv, O.LargestAlge, want_vectors=False, threads=O.max_threads())
            rel = float(np.abs(evals - ref.eigenvalues).max() / np.abs(ref.eigenvalues).max())
            line.update(oracle_nops=ref.nops, rel_vs_oracle=rel)
            good = good and rel <= 1e-10 and abs(ref.nops - eigs.num_operations()) <= max(60, ref.nops // 5)
        print(("PASS " if good else "FAIL ") + str(line), flush=True)
        ok = ok and good
        del eigs
        op.close()
    # nonsymmetric sharded
    n = 100_000
    row0, nrows = dist.slab_range(n, rank, world)
    rp, ci, v = synth.csr(n, 20, 1, False, row0=row0, nrows=nrows)
    diag = np.zeros(n)
    diag[:20] = 3.0 + 0.35 * np.arange(20)
    Aloc = sp.csr_matrix((v, ci, rp), shape=(nrows, n)) + sp.diags(diag, 0, shape=(n, n)).tocsr()[row0:row0 + nrows]
    Aloc = Aloc.tocsr()
    Aloc.sort_indices()
    op = sb.SparseGenMatProd.from_csr_slab(n, row0, Aloc.indptr.astype(np.int64), Aloc.indices, Aloc.data, comm=comm)
    g = sb.GenEigsSolver(op, 10, 30)
    g.init()
    g.compute(sb.SortRule.LargestMagn)
    ev, Z = g.eigenvalues(), g.eigenvectors()
    num = np.array([dist.sum_over_ranks(float(np.sum(np.abs(Aloc @ Z[:, c] - Z[row0:row0 + nrows, c] * ev[c]) ** 2))) for c in range(len(ev))])
    res = np.sqrt(num) / np.abs(ev)
    good = g.info() == sb.CompInfo.Successful and res.max() <= 1e-9
    print(("PASS " if good else "FAIL ") + str(dict(gen_n=n, rank=rank, nops=g.num_operations(), max_res=float(res.max()), ev0=complex(ev[0]))), flush=True)
    ok = ok and good
    dist.barrier()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
