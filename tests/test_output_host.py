"""Host part of out1d / out2d (snac_b200/output.py; reference src/output.f90:34-225, src/out1d.h90): assembling the
boxes' contributions into block profiles and the text formats, with the device reduction replaced by a numpy
stand-in of the same C-ABI calls (the kernels themselves are checked on the GPU)."""
import ctypes as C
import os

import numpy as np

import cases
from snac_b200 import output


class _Lib:
    """snac_out1d / snac_out2d with the documented semantics, evaluated on the host arrays of a fake handle"""

    def __init__(self, sim):
        self.sim = sim

    def _arr(self, ptr, n):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))

    def snac_out1d(self, h, rank, field, idir, wa, wb, scale, out):
        p = self.sim.fields[rank][field][1:-1, 1:-1, 1:-1]
        n = p.shape
        a, b = [d for d in range(3) if d != idir]
        w = [None, None, None]
        w[a] = self._arr(wa, n[a] + 2)[1:-1]
        w[b] = self._arr(wb, n[b] + 2)[1:-1]
        wgt = [w[d].reshape([-1 if e == d else 1 for e in range(3)]) if w[d] is not None else 1.0 for d in range(3)]
        self._arr(out, n[idir])[:] = (p * wgt[a] * wgt[b]).sum(axis=(a, b)) * scale
        return 0

    def snac_out2d(self, h, rank, field, idir, w, scale, out):
        p = self.sim.fields[rank][field][1:-1, 1:-1, 1:-1]
        n = p.shape
        a, b = [d for d in range(3) if d != idir]
        wgt = self._arr(w, n[idir] + 2)[1:-1].reshape([-1 if e == idir else 1 for e in range(3)])
        self._arr(out, n[a] * n[b])[:] = ((p * wgt).sum(axis=idir) * scale).ravel(order="F")
        return 0


class _Sim:
    def __init__(self, case):
        self.case, self.h = case, None
        self.ranks = list(range(case.world.nrank))
        rng = np.random.default_rng(2)
        self.fields = {r: {f: np.asfortranarray(rng.standard_normal([int(x) + 2 for x in b.n])) for f in range(4)}
                       for r, b in enumerate(case.world.boxes)}
        self.lib = _Lib(self)

    def _chk(self, rc):
        assert rc == 0


def test_block_profiles_and_files(tmp_path):
    case = cases.make("channel", str(tmp_path / "run"), fft_axis=2)   # one block, 2 x 1 x 2 boxes, stretched in z
    sim = _Sim(case)
    blk = case.params.blocks[0]
    l = np.asarray(blk.lmax, float) - np.asarray(blk.lmin, float)
    ng = [int(x) for x in blk.ng]
    lo_g = np.min([b.lo for b in case.world.boxes], axis=0)
    for field, stag in ((0, 0), (2, 2), (3, -1)):
        full = np.zeros(ng)
        wfull = [np.zeros(ng[d]) for d in range(3)]
        for r, box in enumerate(case.world.boxes):
            s = [int(box.lo[d] - lo_g[d]) for d in range(3)]
            n = [int(x) for x in box.n]
            full[s[0]:s[0] + n[0], s[1]:s[1] + n[1], s[2]:s[2] + n[2]] = sim.fields[r][field][1:-1, 1:-1, 1:-1]
            g = case.grids[r]
            for d in range(3):
                wfull[d][s[d]:s[d] + n[d]] = (g.dc(d) if d == stag else g.df(d))[1:-1]
        for idir in range(3):
            a, b = [d for d in range(3) if d != idir]
            wa = wfull[a].reshape([-1 if e == a else 1 for e in range(3)])
            wb = wfull[b].reshape([-1 if e == b else 1 for e in range(3)])
            wi = wfull[idir].reshape([-1 if e == idir else 1 for e in range(3)])
            ref1 = (full * wa * wb).sum(axis=(a, b)) / (l[a] * l[b])          # src/output.f90:71-79
            ref2 = (full * wi).sum(axis=idir) / l[idir]                        # src/output.f90:160-168
            assert np.allclose(output.block_profile(sim, 0, field, idir, stag), ref1, rtol=1e-13, atol=1e-15)
            assert np.allclose(output.block_average(sim, 0, field, idir, stag), ref2, rtol=1e-13, atol=1e-15)
    # files: out1d two columns (2E15.7), out2d three columns (3E15.7) with the first kept direction fastest
    fn = str(tmp_path / "wmean_z.out")
    p1d = output.out1d(fn, sim, 0, 2, 2, stag=2)
    tab = np.loadtxt(fn)
    assert tab.shape == (ng[2], 2)
    assert np.allclose(tab[:, 0], case.bgrids[0][2].rf[1:-1], rtol=6e-7)      # w lives on z faces
    assert np.allclose(tab[:, 1], p1d, rtol=6e-7, atol=1e-9)
    assert all(len(line) == 30 for line in open(fn).read().splitlines())
    fn2 = str(tmp_path / "p_avg_y.out")
    p2d = output.out2d(fn2, sim, 0, 3, 1)
    tab2 = np.loadtxt(fn2)
    assert tab2.shape == (ng[0] * ng[2], 3)
    assert np.allclose(tab2[:ng[0], 0], case.bgrids[0][0].rc[1:-1], rtol=6e-7) and np.all(tab2[:ng[0], 1] == tab2[0, 1])
    assert np.allclose(tab2[:, 2], p2d.ravel(order="F"), rtol=6e-7, atol=1e-9)
    output.out1d_hook(sim, str(tmp_path / "data"), "0000005")
    assert len(os.listdir(str(tmp_path / "data"))) == 12
