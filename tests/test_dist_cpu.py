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


def test_bench_decompositions_match_gpu_count(tmp_path):
    """bench.py's N-GPU workloads: one box per GPU with the transform axis whole -- the channel strong-scaled
    over block_dims (1,1,1)/(2,1,1)/(2,1,2)/(4,1,2), the TGV weak-scaled, the duct as 4 or 8 friend blocks."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    for n in (1, 2, 4, 8):
        case, owner, scaling, topo = bench.make_case("channel", (32, 16, 16), n, str(tmp_path / f"c{n}"))
        assert case.world.nrank == n and sorted(owner) == list(range(n)) and scaling == "strong"
        assert all(int(b.n[1]) == 16 for b in case.world.boxes)          # y (FFT_AXIS=2) is never split
        assert sum(int(np.prod(b.n)) for b in case.world.boxes) == 32 * 16 * 16
        assert topo == ((0, 0, 1), 2) and tuple(case.world.periods) == (32, 16, 0)
        case, owner, scaling, topo = bench.make_case("tgv", (16, 16, 16), n, str(tmp_path / f"t{n}"))
        assert case.world.nrank == n and scaling == "weak"
        assert all(tuple(b.n) == (16, 16, 16) for b in case.world.boxes)
        assert np.all(case.world.periods == np.array(bench.TGV_DIMS[n]) * 16)
        case, owner, scaling, topo = bench.make_case("duct", (16, 16, 16), n, str(tmp_path / f"d{n}"))
        assert case.world.nrank == (8 if n == 8 else 4) and len(case.params.blocks) == case.world.nrank
        assert set(owner) == set(range(n)) and all(int(b.n[0]) == 16 for b in case.world.boxes)
        assert tuple(case.world.periods) == (16, 0, 0)


def _out1d_worker(rank, world, port, rundir, datadir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, os.path.dirname(__file__))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_output_host import _Sim
    from snac_b200 import output
    from snac_b200.case import setup_case
    case = setup_case(rundir, fft_axis=2)
    sim = _Sim(case)                                                    # same seeded fields in every process
    sim.ranks = [r for r in range(case.world.nrank) if r % world == rank]   # ... of which it owns these boxes
    p1d = output.out1d(os.path.join(datadir, "wmean_z.out"), sim, 0, 2, 2, stag=2)
    p2d = output.block_average(sim, 0, 3, 1)
    q.put((rank, p1d, p2d, os.path.exists(os.path.join(datadir, "wmean_z.out"))))
    dist.barrier()
    dist.destroy_process_group()


def test_out1d_out2d_reduce_over_the_processes_of_a_block(tmp_path):
    """out1d/out2d with the boxes of a block dealt to two processes (MPI_ALLREDUCE over comm_block,
    src/output.f90:79-86,170-177): both get the whole-block profile, the block's first process writes the file."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import cases
    from test_output_host import _Sim
    from snac_b200 import output
    rundir, datadir = str(tmp_path / "run"), str(tmp_path / "data")
    os.makedirs(datadir)
    case = cases.make("channel", rundir, fft_axis=2)
    ref1 = output.block_profile(_Sim(case), 0, 2, 2, 2)
    ref2 = output.block_average(_Sim(case), 0, 3, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 300
    ps = [ctx.Process(target=_out1d_worker, args=(r, 2, port, rundir, datadir, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for rank, p1d, p2d, _ in res:
        assert np.allclose(p1d, ref1, rtol=1e-13, atol=1e-15) and np.allclose(p2d, ref2, rtol=1e-13, atol=1e-15)
    tab = np.loadtxt(os.path.join(datadir, "wmean_z.out"))
    assert np.allclose(tab[:, 1], ref1, rtol=6e-7, atol=1e-9)


class _FakeSim:
    """stands in for snac_b200.sim.Simulation in the file-writing host logic (no GPU): owns some boxes of the
    decomposition and 'downloads' deterministic fields"""

    def __init__(self, case, ranks):
        self.case, self.ranks, self.time, self.istep = case, list(ranks), 0.75, 12

    def shape(self, r):
        return tuple(int(x) + 2 for x in self.case.world.boxes[r].n)

    def download(self, r, field):
        return np.asfortranarray(np.random.default_rng(100 * r + field).standard_normal(self.shape(r)))

    # the asynchronous pair (snac_download_async / snac_download_wait): the array is complete only once its ticket has
    # been waited for -- a timer fills it 50 ms after the request
    def download_async(self, r, field):
        import threading
        if not hasattr(self, "_events"):
            self._events = []
        a = np.full(self.shape(r), np.nan, order="F")
        ev = threading.Event()

        def land():
            a[...] = self.download(r, field)
            ev.set()
        threading.Timer(0.05, land).start()
        self._events.append(ev)
        return a, len(self._events) - 1, (lambda: None)

    def download_wait(self, ticket):
        assert self._events[ticket].wait(30)


def _ckpt_worker(rank, world, port, rundir, datadir, asynchronous=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snac_b200 import load
    from snac_b200.case import setup_case
    case = setup_case(rundir, fft_axis=2)
    sim = _FakeSim(case, [r for r in range(case.world.nrank) if r % world == rank])   # rank <-> GPU as in main.py
    if asynchronous:   # the writer thread of the time loop (main.py): join, then the barrier
        load.write_checkpoint_async(sim, datadir).join()
    else:
        load.write_checkpoint(sim, datadir)
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("asynchronous", [False, True])
def test_checkpoint_written_by_two_processes_owning_boxes_of_one_block(tmp_path, asynchronous):
    """one block of 2 x 1 x 2 boxes, boxes dealt to two processes as `python -m snac_b200.main` under torchrun
    does (rank % nproc): every process writes its sub-arrays of the same fld_b_001_*.bin (src/load.f90:161-187)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import cases
    rundir, datadir = str(tmp_path / "run"), str(tmp_path / "data")
    case = cases.make("channel", rundir, fft_axis=2)
    ctx = mp.get_context("spawn")
    port = 29650 + os.getpid() % 300 + (7 if asynchronous else 0)
    ps = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, rundir, datadir, asynchronous)) for r in range(2)]
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
