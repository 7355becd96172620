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


class _FakeSim:
    """stands in for snac_b200.sim.Simulation in the file-writing host logic (no GPU): owns some boxes of the
    decomposition and 'downloads' deterministic fields"""

    def __init__(self, case, ranks):
        self.case, self.ranks, self.time, self.istep = case, list(ranks), 0.75, 12

    def shape(self, r):
        return tuple(int(x) + 2 for x in self.case.world.boxes[r].n)

    def download(self, r, field):
        return np.asfortranarray(np.random.default_rng(100 * r + field).standard_normal(self.shape(r)))


def _ckpt_worker(rank, world, port, rundir, datadir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snac_b200 import load
    from snac_b200.case import setup_case
    case = setup_case(rundir, fft_axis=2)
    sim = _FakeSim(case, [r for r in range(case.world.nrank) if r % world == rank])   # rank <-> GPU as in main.py
    load.write_checkpoint(sim, datadir)
    dist.barrier()
    dist.destroy_process_group()


def test_checkpoint_written_by_two_processes_owning_boxes_of_one_block(tmp_path):
    """one block of 2 x 1 x 2 boxes, boxes dealt to two processes as `python -m snac_b200.main` under torchrun
    does (rank % nproc): every process writes its sub-arrays of the same fld_b_001_*.bin (src/load.f90:161-187)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import cases
    rundir, datadir = str(tmp_path / "run"), str(tmp_path / "data")
    case = cases.make("channel", rundir, fft_axis=2)
    ctx = mp.get_context("spawn")
    port = 29650 + os.getpid() % 300
    ps = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, rundir, datadir)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(180)
        assert p.exitcode == 0
    ng = tuple(int(x) for x in case.params.blocks[0].ng)
    lo_g = np.min([b.lo for b in case.world.boxes], axis=0)
    for name, fid in (("u", 0), ("v", 1), ("w", 2), ("p", 3)):
        raw = np.fromfile(os.path.join(datadir, f"fld_b_001_{name}.bin"))
        assert raw.size == np.prod(ng) + 2 and raw[-2] == 0.75 and raw[-1] == 12.0
        full = raw[:-2].reshape(ng, order="F")
        for r, box in enumerate(case.world.boxes):
            exp = _FakeSim(case, [r]).download(r, fid)[1:-1, 1:-1, 1:-1]
            s = [int(box.lo[d] - lo_g[d]) for d in range(3)]
            assert np.array_equal(full[s[0]:s[0] + exp.shape[0], s[1]:s[1] + exp.shape[1], s[2]:s[2] + exp.shape[2]], exp)
