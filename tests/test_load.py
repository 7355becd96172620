"""Field/checkpoint files in the reference's layout (snac_b200/load.py; reference src/load.f90:84-187,
src/initgrid.f90:128-165, src/main.f90:428-438): written box by box, read back the way the reference's own
readers do (utils/read_binary_data/python/read_single_field_binary.py: np.fromfile + Fortran-order reshape)."""
import os

import numpy as np
import pytest

import cases
from snac_b200 import load
from snac_b200.initgrid import load_grid


def test_load_one_writes_the_reference_layout(tmp_path):
    ng = (6, 5, 4)
    rng = np.random.default_rng(3)
    full = rng.standard_normal(ng)
    fn = str(tmp_path / "fld_b_001_u.bin")
    # two boxes of one block split in y, one ghost layer each (nh = 1), written independently
    for lo, hi, first in (((1, 1, 1), (6, 2, 4), True), ((1, 3, 1), (6, 5, 4), False)):
        n = [hi[d] - lo[d] + 1 for d in range(3)]
        p = np.full([x + 2 for x in n], np.nan, order="F")
        p[1:-1, 1:-1, 1:-1] = full[lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]]
        load.load_one("w", fn, ng, lo, hi, p, time=1.25, istep=42, write_info=first)
    assert os.path.getsize(fn) == (np.prod(ng) + 2) * 8          # src/load.f90:110
    raw = np.fromfile(fn, dtype="float64")                        # the reference reader's way
    assert np.array_equal(raw[:-2].reshape(ng, order="F"), full)
    assert raw[-2] == 1.25 and raw[-1] == 42.0
    # read a box back
    p = np.zeros((8, 5, 6), order="F")
    t, it = load.load_one("r", fn, ng, (1, 3, 1), (6, 5, 4), p)
    assert (t, it) == (1.25, 42)
    assert np.array_equal(p[1:-1, 1:-1, 1:-1], full[:, 2:5, :])
    assert np.all(p[0] == 0.0)                                    # ghost cells are not touched
    # a file of the wrong size is refused, as in the reference (src/load.f90:112-119)
    with open(fn, "ab") as f:
        f.write(b"\0" * 8)
    with pytest.raises(IOError):
        load.load_one("r", fn, ng, (1, 1, 1), (6, 2, 4), np.zeros((8, 4, 6), order="F"))


def test_geometry_and_grid_files_round_trip(tmp_path):
    case = cases.make("channel", str(tmp_path / "run"), fft_axis=2)
    out = str(tmp_path / "data")
    load.save_geometry(out, case)
    geo = np.loadtxt(os.path.join(out, "geometry_b_001.out"), max_rows=4)   # as read_single_field_binary.py does
    blk = case.params.blocks[0]
    assert np.array_equal((geo[1] - geo[0] + 1).astype(int), np.asarray(blk.ng))
    assert np.allclose(geo[2], blk.lmin) and np.allclose(geo[3], blk.lmax)
    for d, c in enumerate("xyz"):
        g = load_grid(os.path.join(out, f"grid_{c}_b_001"), int(blk.ng[d]))
        ref = case.bgrids[0][d]
        assert np.array_equal(g.rf[1:-1], ref.rf[1:-1]) and np.array_equal(g.drc[1:-1], ref.drc[1:-1])
        raw = np.fromfile(os.path.join(out, f"grid_{c}_b_001.bin")).reshape((int(blk.ng[d]), 4), order="F")
        assert np.array_equal(raw[:, 1], ref.rc[1:-1])            # column 1 = cell centres (reference reader)


def _write_box(args):
    fn, ng, lo, hi, seed, first = args
    import numpy as np
    from snac_b200 import load
    n = [hi[d] - lo[d] + 1 for d in range(3)]
    p = np.zeros([x + 2 for x in n], order="F")
    p[1:-1, 1:-1, 1:-1] = np.random.default_rng(seed).standard_normal(n)
    load.load_one("w", fn, ng, lo, hi, p, time=0.5, istep=7, write_info=first)
    return p[1:-1, 1:-1, 1:-1]


def test_boxes_of_a_block_written_by_separate_processes(tmp_path):
    """one process per GPU writes its boxes of the same block file (the MPI-IO sub-array views of the
    reference, src/load.f90:161-187): disjoint regions through memory maps, any order."""
    import multiprocessing as mp
    ng = (8, 6, 4)
    fn = str(tmp_path / "fld_b_001_p.bin")
    boxes = [((1, 1, 1), (4, 3, 4), 11, True), ((5, 1, 1), (8, 3, 4), 12, False),
             ((1, 4, 1), (4, 6, 4), 13, False), ((5, 4, 1), (8, 6, 4), 14, False)]
    with mp.get_context("spawn").Pool(4) as pool:
        parts = pool.map(_write_box, [(fn, ng, lo, hi, seed, first) for lo, hi, seed, first in boxes])
    raw = np.fromfile(fn)
    full = raw[:-2].reshape(ng, order="F")
    for (lo, hi, _, _), part in zip(boxes, parts):
        assert np.array_equal(full[lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]], part)
    assert raw[-2] == 0.5 and raw[-1] == 7.0


def test_out3d_and_visu_log_follow_the_reference_formats(tmp_path):
    """out3d (src/output.f90:227-257) and write_log_output (:259-285): file = the block array alone; log line
    '(A,A,A,9i5,E15.7,i10)' as parsed by the reference's utils/visualize_fields scripts (whitespace split)."""
    ng = (5, 4, 3)
    full = np.random.default_rng(9).standard_normal(ng)
    fn = str(tmp_path / "vex_fld_0000100_b_001.bin")
    for lo, hi in (((1, 1, 1), (5, 4, 1)), ((1, 1, 2), (5, 4, 3))):   # two boxes split in z
        n = [hi[d] - lo[d] + 1 for d in range(3)]
        p = np.zeros([x + 2 for x in n], order="F")
        p[1:-1, 1:-1, 1:-1] = full[lo[0] - 1:hi[0], lo[1] - 1:hi[1], lo[2] - 1:hi[2]]
        load.out3d(fn, ng, lo, hi, p)
    assert os.path.getsize(fn) == np.prod(ng) * 8
    assert np.array_equal(np.fromfile(fn).reshape(ng, order="F"), full)
    log = str(tmp_path / "log_visu_3d_b_001.out")
    load.write_log_output(log, "vex_fld_0000100_b_001.bin", "Velocity_X", (1, 1, 1), ng, (1, 1, 1), 12.5, 100)
    load.write_log_output(log, "vey_fld_0000100_b_001.bin", "Velocity_Y", (1, 1, 1), ng, (1, 1, 1), 12.5, 100)
    lines = open(log).read().splitlines()
    assert len(lines) == 2
    assert lines[0] == "vex_fld_0000100_b_001.bin Velocity_X" + "    1" * 3 + "    5    4    3" + "    1" * 3 + "  0.1250000E+02       100"
    tok = lines[1].split()
    assert tok[0] == "vey_fld_0000100_b_001.bin" and tok[1] == "Velocity_Y" and [int(t) for t in tok[2:11]] == [1, 1, 1, 5, 4, 3, 1, 1, 1]
    assert float(tok[11]) == 12.5 and int(tok[12]) == 100


def test_a_late_creator_does_not_wipe_boxes_already_written(tmp_path, monkeypatch):
    """The interleaving behind the checkpoint race: process B decided the file was missing, process A then
    created it and wrote its box, and only now B gets to create it.  With open('wb') (O_TRUNC) B wiped A's box;
    creation must never truncate."""
    ng = (4, 4, 2)
    fn = str(tmp_path / "fld_b_001_v.bin")
    a = np.zeros((6, 4, 4), order="F"); a[1:-1, 1:-1, 1:-1] = 1.5
    b = np.zeros((6, 4, 4), order="F"); b[1:-1, 1:-1, 1:-1] = 2.5
    load.load_one("w", fn, ng, (1, 1, 1), (4, 2, 2), a, time=3.0, istep=9, write_info=True)
    # B's stale view of the directory: it still believes the file does not exist
    monkeypatch.setattr(os.path, "exists", lambda p: False)
    load.load_one("w", fn, ng, (1, 3, 1), (4, 4, 2), b, write_info=False)
    monkeypatch.undo()
    raw = np.fromfile(fn)
    full = raw[:-2].reshape(ng, order="F")
    assert np.all(full[:, :2, :] == 1.5) and np.all(full[:, 2:, :] == 2.5)
    assert raw[-2] == 3.0 and raw[-1] == 9.0
    # a stale file of another size (a previous run on another grid) is resized, not appended to
    with open(fn, "ab") as f:
        f.write(b"\0" * 64)
    load.load_one("w", fn, ng, (1, 1, 1), (4, 2, 2), a, time=4.0, istep=10, write_info=True)
    assert os.path.getsize(fn) == (np.prod(ng) + 2) * 8
    # the same for out3d files
    fo = str(tmp_path / "pre_fld_0000001_b_001.bin")
    load.out3d(fo, ng, (1, 1, 1), (4, 2, 2), a)
    monkeypatch.setattr(os.path, "exists", lambda p: False)
    load.out3d(fo, ng, (1, 3, 1), (4, 4, 2), b)
    monkeypatch.undo()
    full = np.fromfile(fo).reshape(ng, order="F")
    assert np.all(full[:, :2, :] == 1.5) and np.all(full[:, 2:, :] == 2.5)


REF_READER = "/root/reference/utils/read_binary_data/python"


@pytest.mark.skipif(not os.path.isdir(REF_READER), reason="the reference tree is only present in the build container")
def test_the_reference_reader_reads_our_files(tmp_path, monkeypatch):
    """The reference's own post-processing reader (utils/read_binary_data/python/read_single_field_binary.py,
    the one its tests/*/test.py use) on the geometry, grid and field files written here."""
    import sys
    monkeypatch.syspath_prepend(REF_READER)
    sys.modules.pop("read_single_field_binary", None)
    from read_single_field_binary import read_single_field_binary
    case = cases.make("channel", str(tmp_path / "run"), fft_axis=2)
    out = tmp_path / "data"
    load.save_geometry(str(out), case)
    blk = case.params.blocks[0]
    ng = tuple(int(x) for x in blk.ng)
    full = np.random.default_rng(21).standard_normal(ng)
    for box in case.world.boxes:   # every box of the block writes its own part
        p = np.zeros(tuple(int(x) + 2 for x in box.n), order="F")
        lo1, hi1 = box.lo, box.hi   # one block starting at (1,1,1): block-local = global
        p[1:-1, 1:-1, 1:-1] = full[lo1[0] - 1:hi1[0], lo1[1] - 1:hi1[1], lo1[2] - 1:hi1[2]]
        load.out3d(str(out / "vex_fld_0000010_b_001.bin"), ng, lo1, hi1, p)
    monkeypatch.chdir(out)
    data, xp, yp, zp, xu, yv, zw = read_single_field_binary("vex_fld_0000010_b_001.bin", 1, np.array([1, 1, 1]))
    assert np.array_equal(data, full)
    g = case.bgrids[0]
    assert np.array_equal(xp, g[0].rc[1:-1]) and np.array_equal(xu, g[0].rf[1:-1])
    assert np.array_equal(zp, g[2].rc[1:-1]) and np.array_equal(zw, g[2].rf[1:-1])
