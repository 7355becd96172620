"""Multi-process host logic on CPU (gloo, world size 2): the bootstrap gather of the peer
communication blobs, and the ownership bookkeeping of the driver."""
import multiprocessing as mp
import os

import numpy as np


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snac_b200.dist import gather_blobs
    blob = bytes([rank]) * 64 + b"\x00" * 8      # fixed size, embedded NULs
    allb = gather_blobs(blob, world)
    q.put((rank, allb))
    dist.barrier()
    dist.destroy_process_group()


def test_blob_gather_is_ordered_by_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    expect = bytes([0]) * 64 + b"\x00" * 8 + bytes([1]) * 64 + b"\x00" * 8
    assert res[0] == expect and res[1] == expect


def test_weak_scaling_decomposition_matches_gpu_count(tmp_path):
    """bench.py's N-GPU workload: one box per GPU, transform axis whole, 256^3 cells per box."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    for n, dims in bench.DIMS.items():
        case = bench.make_tgv(16, dims, str(tmp_path / f"n{n}"))
        assert case.world.nrank == n
        assert dims[2] == 1                              # z (FFT_AXIS=3) is never split
        assert all(tuple(b.n) == (16, 16, 16) for b in case.world.boxes)
        assert np.all(case.world.periods == np.array(dims) * 16)
