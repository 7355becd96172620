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
