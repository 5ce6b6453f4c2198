"""Launcher-side helpers for row-sharded multi-GPU runs (one process per GPU, torchrun).

torch.distributed is used only for the rendezvous (broadcasting the NCCL unique id) and for the
max-over-ranks timing reduction of the benchmark; the data path of the solver (all-gather of the
SpMV operand, all-reduce of the dot products) runs inside the C++ library on its own NCCL
communicator (csrc/comm.cu).
"""
from __future__ import annotations

import os


def slab_range(n: int, rank: int, nranks: int):
    """Rows [row0, row0 + nrows) owned by `rank`: contiguous blocks of ceil(n / nranks) rows
    (the partition sb200_op_create_* expects)."""
    slab = (n + nranks - 1) // nranks
    row0 = min(n, slab * rank)
    nrows = max(0, min(slab, n - row0))
    return row0, nrows


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str | None = None):
    """Initialises torch.distributed from the torchrun environment (MASTER_ADDR defaults to 127.0.0.1)."""
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    if world == 1:
        return None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if not dist.is_initialized():
        if backend is None:
            import torch

            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def broadcast_bytes(payload: bytes | None, src: int = 0) -> bytes:
    """Broadcasts a small byte string (the 128-byte NCCL unique id) from `src` to every rank."""
    import torch.distributed as dist

    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def make_comm(make_id=None):
    """Creates the library communicator for this rank (None when WORLD_SIZE == 1).  `make_id` overrides
    the id factory (tests)."""
    import spectra_b200 as sb

    rank, _, world = env_rank()
    if world == 1:
        return None
    init_process_group()
    uid = None
    if rank == 0:
        uid = (make_id or sb.Comm.unique_id)()
    uid = broadcast_bytes(uid, 0)
    return sb.Comm(rank, world, uid)


def max_over_ranks(value: float) -> float:
    """MAX all-reduce of a python float (timing: the slowest rank defines the step time)."""
    rank, _, world = env_rank()
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    rank, _, world = env_rank()
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    _, _, world = env_rank()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
