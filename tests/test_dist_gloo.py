"""Host-side logic of the N>1 path on CPU: world_size 2 over gloo.  Covers the rendezvous helper that
distributes the communicator id, the row-slab partition, slab-wise generation of the synthetic matrix
and the max-over-ranks timing reduction used by bench.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as td

        from spectra_b200 import dist, synth

        dist.init_process_group("gloo")
        uid = dist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
        n = 10001
        row0, nrows = dist.slab_range(n, rank, world)
        rp, ci, v = synth.csr(n, 20, 3, True, row0=row0, nrows=nrows)
        tmax = dist.max_over_ranks(1.0 + rank)
        tsum = dist.sum_over_ranks(float(len(ci)))
        dist.barrier()
        q.put((rank, uid, row0, nrows, int(rp[-1]), float(v.sum()), tmax, tsum))
        td.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


def test_two_rank_host_path_gloo():
    import torch.multiprocessing as mp

    from spectra_b200 import dist, synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(len(r) == 8 for r in res), res
    (r0, uid0, a0, n0, z0, s0, tmax0, tsum0), (r1, uid1, a1, n1, z1, s1, tmax1, tsum1) = res
    assert uid0 == uid1 == bytes(range(128))
    n = 10001
    assert (a0, n0) == (0, 5001) and (a1, n1) == (5001, 5000) and n0 + n1 == n
    rp, ci, v = synth.csr(n, 20, 3, True)
    assert z0 + z1 == len(ci) and tsum0 == tsum1 == float(len(ci))
    assert abs((s0 + s1) - float(v.sum())) <= 1e-9 * np.abs(v).sum()
    assert tmax0 == tmax1 == 2.0


@pytest.mark.parametrize("n,P", [(10, 3), (7, 8), (1000, 4), (10_000_000, 8), (5, 1)])
def test_slab_partition(n, P):
    from spectra_b200.dist import slab_range

    rows = [slab_range(n, r, P) for r in range(P)]
    assert rows[0][0] == 0 and sum(nr for _, nr in rows) == n
    for (a, na), (b, nb) in zip(rows, rows[1:]):
        assert a + na == b or nb == 0
    slab = (n + P - 1) // P
    assert all(nr <= slab for _, nr in rows)
