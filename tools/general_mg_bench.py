"""Timing of the general multi-box multigrid path (mg.cu / mg_coop.cu) on the workloads it serves:

  ldc128   the reference's examples/_manuscript_lid_driven_cavity (128^3, all walls, FFT_AXIS = 0: one 3-D solve per
           sub-step), in one box and cut into 2 x 1 x 2 boxes on one GPU (performance variant)
  lshape   three friend blocks forming an L (FFT_AXIS = 1): per-wavenumber planes split over blocks that do not tile
           a rectangle, coupled through the block interfaces
  impdiff  a wall-bounded channel with IMPDIFF: three staggered Helmholtz solves + the 3-D pressure solve per sub-step

  python tools/general_mg_bench.py [--steps K]      prints one JSON line per case and per SNAC_MG_COOP setting
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def one(name, steps):
    from snac_b200 import decks
    from snac_b200.case import initial_fields, setup_case
    from snac_b200.sim import Simulation
    d = tempfile.mkdtemp()
    if name.startswith("ldc128"):
        bcv = [[0.0] * 6 for _ in range(3)]
        bcv[0][5] = 1.0
        dims = (2, 1, 2) if name.endswith("x4") else (1, 1, 1)   # ldc128auto: FFT_AXIS chosen by the host (an FFT build the reference accepts)
        decks.write_case(d, [dict(dims=dims, ng=(128, 128, 128), lmin=(0, 0, 0), lmax=(1, 1, 1), cbcvel=["DDDDDD"] * 3, cbcpre="NNNNNN",
                                  bcvel=bcv, inivel="zer")], visci=1000.0)
        case = setup_case(d, fft_axis="auto" if name.endswith("auto") else 0)
    elif name == "lshape":
        decks.lshape_blocks_case(d, n=(64, 48, 48))
        case = setup_case(d, fft_axis=1)
    else:
        decks.channel_case(d, ng=(64, 48, 64), dims=(1, 1, 1))
        case = setup_case(d, fft_axis=0, impdiff=True)
    g = Simulation(case)
    g.set_fields(initial_fields(case, noise=True))
    g.init_bc()
    g.set_dt()
    for _ in range(3):
        g.step()
    g.sync()
    l0 = g.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        g.step()
    g.sync()
    sec = (time.perf_counter() - t0) / steps
    launches = (g.launch_count() - l0) / steps
    it, relres = g.solver_stats()
    divtot, divmax = g.chkdiv()
    pts = float(sum(int(b.n[0]) * int(b.n[1]) * int(b.n[2]) for b in case.world.boxes))
    print(json.dumps({"case": name, "coop": os.environ.get("SNAC_MG_COOP", "1"), "ms_per_step": round(sec * 1e3, 3),
                      "launches_per_step": launches, "Mpts_step_per_s": round(pts / sec / 1e6, 2), "cycles_last_solve": it,
                      "relres": float(f"{relres:.2e}"), "divmax": float(f"{divmax:.2e}"), "hypre_tol": case.params.hypre_tol, "fft_axis": int(case.params.fft_axis)}), flush=True)
    g.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--case", default=None)
    a = ap.parse_args()
    if a.case:
        one(a.case, a.steps)
    else:
        for name in ("ldc128", "ldc128x4", "lshape", "impdiff", "ldc128auto"):
            for coop in ("1", "0") if name != "ldc128auto" else ("1",):
                subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name, "--steps", str(a.steps)],
                               env=dict(os.environ, SNAC_MG_COOP=coop))
