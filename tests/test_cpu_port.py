"""The C/OpenMP CPU baseline (oracle/snac_cpu.c) against the numpy oracle on the same inputs: the
port is what bench.py times as ``cpu_baseline`` / ``--impl reference`` and what its ``check`` compares
the GPU fields with, so it has to compute the same thing.  Tolerance: relative L2 <= 1e-8 on u,v,w
after 3 steps with the plane solves converged to 1e-11 (the oracle solves them exactly)."""
import numpy as np
import pytest

from oracle import snac_oracle as orc
from oracle.cpu_port import CpuSim
from snac_b200 import decks
from snac_b200.case import initial_fields


def _compare(case, walls, fft_axis, noise):
    ini = initial_fields(case, noise=noise)
    o = orc.OracleSim(case.params, case.world, case.grids, case.bgrids)
    o.u[0][...] = ini[0]["u"]; o.v[0][...] = ini[0]["v"]; o.w[0][...] = ini[0]["w"]; o.pr[0][...] = ini[0]["p"]
    blk = case.params.blocks[0]
    c = CpuSim(tuple(blk.ng), tuple(blk.lmax - blk.lmin), case.params.visc, tol=1e-11, maxiter=200, walls=walls,
               fft_axis=fft_axis, bforce=case.params.bforce)
    for k, key in enumerate("uvwp"):
        c.set(k, ini[0][key])
    o.init_bc(); c.init_bc()
    for k, a in enumerate((o.u[0], o.v[0], o.w[0], o.pr[0])):   # ghost layers filled the same way
        assert np.array_equal(c.get(k), a), k
    o.set_dt()
    assert abs(c.chkdt() - o.chkdt()) <= 1e-12 * o.chkdt()
    for _ in range(3):
        o.step()
        c.step(o.dt)
    vmag = np.sqrt(sum(np.linalg.norm(x[0]) ** 2 for x in (o.u, o.v, o.w)))
    for k, a in enumerate((o.u[0], o.v[0], o.w[0])):
        got = c.get(k)
        assert np.linalg.norm(got - a) <= 1e-8 * max(np.linalg.norm(a), 1e-6 * vmag), k
    gp, op = c.get(3)[1:-1, 1:-1, 1:-1], o.pr[0][1:-1, 1:-1, 1:-1]
    assert np.linalg.norm((gp - gp.mean()) - (op - op.mean())) <= 1e-8 * np.linalg.norm(op - op.mean())
    assert c.divmax() < 1e-9
    c.close()


def test_cpu_port_matches_oracle_on_tgv(tmp_path):
    _compare(decks.make("tgv", str(tmp_path), fft_axis=3, n=16), (0, 0, 0), 3, False)


def test_cpu_port_matches_oracle_on_channel(tmp_path):
    """the benchmark's headline topology: x,y periodic, walls in z, transforms along y, anisotropic cells"""
    case = decks.bench_channel(str(tmp_path), ng=(16, 8, 24), dims=(1, 1, 1), visci=100.0)
    _compare(case, (0, 0, 1), 2, True)


def test_cpu_port_matches_oracle_on_duct(tmp_path):
    """walls in y and z, transforms along x; one block (the discrete problem is the same however it is cut)"""
    case = decks.bench_duct(str(tmp_path), ng=(16, 8, 12), nby=1, nbz=1, visci=100.0)
    _compare(case, (0, 1, 1), 1, True)


def test_cpu_port_rejects_transform_across_walls():
    with pytest.raises(ValueError):
        CpuSim((8, 8, 8), (1, 1, 1), 0.01, walls=(0, 0, 1), fft_axis=3)
