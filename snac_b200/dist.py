"""Bootstrap of the peer communication between the processes of one node.

The library needs one fixed-size blob per process (CUDA IPC handle of its arena and the table of
arrays in it), gathered in process order.  Here the gather uses ``torch.distributed`` (any backend);
a Fortran host would use ``MPI_Allgather`` on the same bytes.
"""
from __future__ import annotations


def gather_blobs(blob: bytes, nproc: int) -> bytes:
    """Concatenation of the blobs of processes 0..nproc-1."""
    if nproc == 1:
        return blob
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed must be initialised to connect several processes")
    if dist.get_world_size() != nproc:
        raise RuntimeError(f"world size {dist.get_world_size()} != nproc {nproc}")
    out = [None] * nproc
    dist.all_gather_object(out, blob)
    if any(len(b) != len(blob) for b in out):
        raise RuntimeError("blob sizes differ between processes")
    return b"".join(out)


def _active():
    try:
        import torch.distributed as dist
    except ImportError:
        return None
    return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None


def allreduce(a, op: str = "sum"):
    """Sum (or max) of a float64 numpy array over the processes; the array itself with one process.
    The reference's MPI_ALLREDUCE on the host side of the path (out1d/out2d over comm_block,
    src/output.f90:79-86; the wall-clock stop criterion, src/main.f90:1043-1046)."""
    import numpy as np
    a = np.ascontiguousarray(a, dtype=np.float64)
    dist = _active()
    if dist is None:
        return a
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(a.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t.cpu().numpy()


def barrier():
    dist = _active()
    if dist is not None:
        dist.barrier()
