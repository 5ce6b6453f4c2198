"""Sweep of the number of partial all-gathers per operator application (SB200_AG_CHUNKS) on a row-sharded solve.
Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_chunks.py [n] [maxit]

Prints one JSON line per setting (rank 0): SpMV-iters/s of a truncated solve (device time, max over ranks)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import spectra_b200 as sb
from spectra_b200 import dist, synth


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rank, local_rank, world = dist.env_rank()
    sb.set_device(local_rank)
    dist.init_process_group("gloo")
    comm = dist.make_comm()
    row0, nrows = dist.slab_range(n, rank, world)
    rp, ci, v = synth.csr(n, 20, 0, True, row0=row0, nrows=nrows)
    settings = [(int(c), int(p)) for p in os.environ.get("PEER_LIST", "1").split(",") for c in os.environ.get("CHUNK_LIST", "1,2,4,8").split(",")]
    for chunks, peer in settings:
        os.environ["SB200_AG_CHUNKS"] = str(chunks)
        os.environ["SB200_PEER"] = str(peer)
        op = sb.SparseGenMatProd.from_csr_slab(n, row0, rp, ci, v, comm=comm)
        out = {"peer_mode": op.peer_mode()}
        for prof in (0, 1):
            sb.set_profiling(prof)
            eigs = sb.SymEigsSolver(op, 20, 60)
            dist.barrier()
            eigs.init()
            eigs.compute(sb.SortRule.LargestAlge, maxit)
            st = eigs.stats()
            ms = dist.max_over_ranks(st["ms_total"])
            if prof == 0:
                out.update(chunks=chunks, world=world, n=n, nops=eigs.num_operations(), ms_total=round(ms, 2), iters_per_s=round(eigs.num_operations() / ms * 1e3, 1))
            else:
                out.update(prof_ms_spmv=round(st["ms_spmv"], 1), prof_ms_panel=round(st["ms_panel"], 1), prof_ms_comm=round(st["ms_comm"], 1),
                           prof_ms_small=round(st["ms_small"], 1), prof_ms_total=round(st["ms_total"], 1), host_syncs=st["host_syncs"])
            del eigs
        sb.set_profiling(0)
        if rank == 0:
            print(json.dumps(out), flush=True)
        op.close()
    dist.barrier()


if __name__ == "__main__":
    main()
