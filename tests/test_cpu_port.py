"""The C/OpenMP CPU baseline (oracle/snac_cpu.c) against the numpy oracle on the same inputs: the
port is what bench.py times as ``cpu_baseline`` / ``--impl reference``, so it has to compute the
same thing.  Tolerance: relative L2 <= 1e-8 on u,v,w after 3 steps with the plane solves converged
to 1e-11 (the oracle solves them exactly)."""
import numpy as np

import cases
from oracle import snac_oracle as orc
from oracle.cpu_port import CpuSim
from snac_b200.case import initial_fields


def test_cpu_port_matches_oracle_on_tgv(tmp_path):
    n = 16
    case = cases.make("tgv", str(tmp_path), fft_axis=3, n=n)
    ini = initial_fields(case)
    o = orc.OracleSim(case.params, case.world, case.grids, case.bgrids)
    o.u[0][...] = ini[0]["u"]; o.v[0][...] = ini[0]["v"]; o.w[0][...] = ini[0]["w"]; o.pr[0][...] = ini[0]["p"]
    o.init_bc(); o.set_dt()
    L = 2.0 * np.pi
    c = CpuSim((n, n, n), (L, L, L), case.params.visc, tol=1e-11, maxiter=200)
    # the oracle's fields after init_bc carry valid ghost layers
    for k, a in enumerate((o.u[0], o.v[0], o.w[0], o.pr[0])):
        c.set(k, a)
    assert abs(c.chkdt() - o.chkdt()) <= 1e-12 * o.chkdt()
    for _ in range(3):
        o.step()
        c.step(o.dt)
    for k, a in enumerate((o.u[0], o.v[0], o.w[0])):
        got = c.get(k)
        ref = np.linalg.norm(a)
        vmag = np.sqrt(sum(np.linalg.norm(x[0]) ** 2 for x in (o.u, o.v, o.w)))
        assert np.linalg.norm(got - a) <= 1e-8 * max(ref, 1e-6 * vmag), k
    gp, op = c.get(3)[1:-1, 1:-1, 1:-1], o.pr[0][1:-1, 1:-1, 1:-1]
    assert np.linalg.norm((gp - gp.mean()) - (op - op.mean())) <= 1e-8 * np.linalg.norm(op - op.mean())
    assert c.divmax() < 1e-9
    c.close()
