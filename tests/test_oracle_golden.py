"""Pins the CPU oracle against the reference's own golden vector (reference
tests/lid_driven_cavity/test.py:5-16, testit.sh:15-16): same 4-block input, 1500 steps, both
Poisson paths, the reference's own assertion (rtol=1e-7, atol=5e-8)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _run_ldc(fft_axis, nstep, q):
    from snac_b200.case import initial_fields, setup_case
    from oracle.snac_oracle import OracleSim
    case = setup_case(os.path.join(GOLDEN, "ldc"), fft_axis=fft_axis)
    sim = OracleSim(case.params, case.world, case.grids, case.bgrids)
    ini = initial_fields(case)
    for r in range(case.world.nrank):
        sim.u[r][...] = ini[r]["u"]
        sim.v[r][...] = ini[r]["v"]
        sim.w[r][...] = ini[r]["w"]
        sim.pr[r][...] = ini[r]["p"]
    sim.init_bc()
    sim.set_dt()
    for _ in range(nstep):
        sim.step()
    d = [sim.v[r][1:-1, 1:-1, 1:-1][0, :, :] for r in range(4)]
    data = np.block([[d[0], d[2]], [d[1], d[3]]])   # same stitching as the reference's test.py
    q.put((fft_axis, data, sim.divmax, sim.dt))


@pytest.fixture(scope="module")
def ldc_runs():
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run_ldc, args=(fa, 1500, q)) for fa in (0, 1)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        fa, data, divmax, dt = q.get(timeout=900)
        res[fa] = (data, divmax, dt)
    for p in procs:
        p.join()
    return res


@pytest.mark.parametrize("fft_axis", [0, 1])
def test_oracle_reproduces_reference_golden_vector(ldc_runs, fft_axis):
    data, divmax, dt = ldc_runs[fft_axis]
    ref = np.loadtxt(os.path.join(GOLDEN, "ldc", "data_ldc_re1000.txt"))
    islice = int(np.size(data[0, :]) / 2)
    np.testing.assert_allclose(data[islice, :], ref[:, 1], rtol=1e-7, atol=5e-8)
    assert divmax < 1e-12
    assert abs(dt - 0.95 * np.sqrt(3.0) / 64.0) < 1e-15
