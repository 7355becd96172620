"""The direct plane solver (snac_b200/csrc/direct.cu: transform along the uniform direction of the plane +
tridiagonal solves along the other one) against the exact solution of the reference's linear
system (oracle LinearSystem, sparse LU) and, over whole time steps, against the oracle."""
import numpy as np
import pytest

import cases
from oracle import snac_oracle as orc
from snac_b200 import lib as L
from snac_b200.case import initial_fields, setup_case
from test_gpu_parity import _solve_and_compare, cmp_uvw, make_pair

pytestmark = pytest.mark.gpu

# planes that are one rectangle with a uniform direction: x periodic + stretched walls in z (non-cyclic Thomas),
# doubly periodic (cyclic), one block cut into several boxes (the slab layout merges them), odd / non-power-of-two sizes
DIRECT_CASES = [("channel", 2, dict(dims=(1, 1, 1))), ("channel", 1, dict(dims=(1, 1, 1))), ("channel", 2, {}), ("channel", 1, {}),
                ("channel", 2, dict(dims=(1, 1, 1), ng=(96, 8, 80))), ("channel", 2, dict(dims=(1, 1, 1), ng=(64, 4, 66))),
                ("channel", 2, dict(dims=(1, 1, 1), ng=(30, 6, 17))), ("channel", 1, dict(dims=(1, 1, 1), ng=(4, 200, 72))),
                ("tgv", 3, dict(n=16)), ("tgv", 2, dict(n=64)), ("tgv", 1, dict(n=32)), ("tgv", 3, dict(n=12)),
                ("channel", 2, dict(dims=(1, 1, 1), ng=(512, 4, 130))),
                # the warp-per-row transforms (fft_row.cuh): n = 128, 256, 512, 1024 along the unit-stride direction
                ("channel", 2, dict(dims=(1, 1, 1), ng=(128, 4, 33))), ("channel", 2, dict(dims=(1, 1, 1), ng=(256, 3, 20))),
                ("channel", 2, dict(dims=(1, 1, 1), ng=(1024, 2, 24))), ("tgv", 3, dict(n=128))]


@pytest.mark.parametrize("kind,fa,kw", DIRECT_CASES)
def test_direct_solver_against_exact_solution(tmp_path, kind, fa, kw):
    _solve_and_compare(tmp_path, kind, fa, kw, False, solver="direct")


def _bc_case(tmp, bcx, n=(16, 10, 8), gr_y=2.0):
    """x: the given pressure boundary-condition pair on a uniform grid (the transformed direction), y: Neumann
    walls on a stretched grid (the tridiagonal direction), z: periodic, FFT_AXIS = 3."""
    pre = [bcx[0], bcx[1], "N", "N", "F", "F"]
    vel = ["F" if c == "F" else ("D" if c == "N" else "N") for c in pre]
    blk = dict(dims=(1, 1, 1), ng=n, lmin=(0, 0, 0), lmax=(1.0, 0.7, 0.5), gt=(0, 0, 0), gr=(0, gr_y, 0),
               cbcvel=["".join(vel)] * 3, cbcpre="".join(pre), inivel="zer")
    cases.write_case(tmp, [blk])
    return setup_case(tmp, fft_axis=3)


@pytest.mark.parametrize("bcx", ["NN", "DD", "ND", "DN", "FF"])
@pytest.mark.parametrize("n", [(16, 10, 8), (12, 7, 6), (128, 9, 4), (256, 5, 2)])
def test_direct_solver_boundary_condition_pairs(tmp_path, bcx, n):
    """every transform kind of find_fft (src/fft.f90:96-149) along the transformed plane direction"""
    case = _bc_case(str(tmp_path), bcx, n)
    _solve_and_compare(tmp_path, None, 3, None, False, solver="direct", case=case)


def test_direct_solver_transform_along_the_second_plane_direction(tmp_path):
    """x stretched walls, y uniform periodic, FFT along z: only plane direction 1 can be transformed"""
    blk = dict(dims=(1, 1, 1), ng=(14, 16, 8), lmin=(0, 0, 0), lmax=(1.0, 0.7, 0.5), gt=(0, 0, 0), gr=(2.0, 0, 0),
               cbcvel=["DDFFFF"] * 3, cbcpre="NNFFFF", inivel="zer")
    cases.write_case(str(tmp_path), [blk])
    case = setup_case(str(tmp_path), fft_axis=3)
    _solve_and_compare(tmp_path, None, 3, None, False, solver="direct", case=case)


def test_direct_solver_refuses_planes_that_are_not_separable(tmp_path):
    from snac_b200.sim import Simulation
    case = cases.make("lshape", str(tmp_path / "a"), fft_axis=1)      # planes split over three friend blocks (an L)
    with pytest.raises(L.SnacError, match="not separable"):
        Simulation(case, solver="direct")
    g = Simulation(case, solver="auto")
    assert g.solver_name() == "mg"
    g.close()
    case = cases.make("cavity", str(tmp_path / "b"), fft_axis=2, dims=(1, 1, 1))   # both plane directions stretched
    g = Simulation(case, solver="auto")
    assert g.solver_name() == "mg"
    g.close()


def test_friend_blocks_tiling_a_rectangle_are_solved_as_one_block(tmp_path):
    """the reference's 2 x 2-block duct: the union of the blocks is a rectangle, the 2-D problems are those of one
    block (api.cu merge_blocks), so the direct solver applies; the result is the exact solution of the 4-block system"""
    _solve_and_compare(tmp_path, "duct", 1, {}, False, solver="direct")
    _solve_and_compare(tmp_path / "b", "duct", 1, dict(n=(16, 12, 10)), False, solver="direct")


@pytest.mark.parametrize("kind,fa,kw", [("channel", 2, {}), ("tgv", 3, dict(n=16)), ("channel", 1, {}), ("duct", 1, {})])
def test_direct_solver_three_steps_against_the_oracle(tmp_path, kind, fa, kw):
    case = cases.make(kind, str(tmp_path), fft_axis=fa, **kw)
    ini = initial_fields(case, noise=(kind != "tgv"))
    o, g = make_pair(case, ini, solver="direct")
    assert g.solver_name() == "direct"
    o.init_bc(); g.init_bc()
    o.set_dt(); g.set_dt()
    for _ in range(3):
        o.step(); g.step()
    cmp_uvw(o, g, 1e-9, interior=True)
    I = (slice(1, -1),) * 3
    num = den = 0.0
    pm_g = np.mean(np.concatenate([g.download(r, L.FIELD_P)[I].ravel() for r in range(o.nr)]))
    pm_o = np.mean(np.concatenate([o.pr[r][I].ravel() for r in range(o.nr)]))
    for r in range(o.nr):
        num += np.sum(((g.download(r, L.FIELD_P)[I] - pm_g) - (o.pr[r][I] - pm_o)) ** 2)
        den += np.sum((o.pr[r][I] - pm_o) ** 2)
    assert np.sqrt(num / den) < 1e-8
    divtot, divmax = g.chkdiv()
    assert divmax < 1e-10
    g.close()
