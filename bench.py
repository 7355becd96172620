#!/usr/bin/env python
"""Benchmark of the per-time-step pressure-correction pipeline (BASELINE.json metric:
ms per RK3 step and Mpts*step/s at 1/2/4/8 B200; HBM roofline of the dominant kernel).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                  [--workload channel|tgv|duct] [--grid NX NY NZ] [--solver auto|mg|direct]

Workloads (decks in snac_b200/decks.py, the reference's example cases):
  channel (default, BASELINE.json configs[2], the configuration the metric is quoted on): turbulent-channel
          geometry of examples/_manuscript_turbulent_channel on 1024 x 512 x 512, FFT_AXIS = 2, one block
          STRONG-scaled over block_dims (1,1,1)/(2,1,1)/(2,1,2)/(4,1,2) on 1/2/4/8 GPUs; Poiseuille start
          plus the deterministic +-0.05 noise (a pure Poiseuille start has zero divergence: nothing to solve);
  tgv     (configs[1]): Taylor-Green vortex 256^3, FFT_AXIS = 3 (weak-scaled, 256^3 per GPU, when N > 1);
  duct    (configs[3]): turbulent duct 512 x 128 x 128 cut into 2 x 2 (8 GPUs: 2 x 4) friend blocks, FFT_AXIS = 1.
A step = one pass of the reference's time loop body (src/main.f90:897-1069) without file output:
three RK sub-steps and, every icheck steps, chkdt + chkdiv.

Before anything is timed every run (every N) advances a small case of the same topology on the same
process grid for 3 steps with converged solves and compares u,v,w,p with the CPU restatement of the
pipeline (oracle/snac_cpu.c, pinned by tests/test_cpu_port.py): `check` in the JSON line; a relative L2
error above 1e-8 makes the run exit non-zero.  The timed run asserts a finite, small divergence.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TGV_DIMS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1)}
METRIC = "Mpts*step/s (RK3 step of the pressure-correction pipeline)"

# algorithmic HBM bytes per grid point and launch of each kernel class
ALG_BYTES = {"rk_mom": 96.0, "fillps": 32.0, "fft_fwd": 16.0, "fft_bwd": 16.0, "correc_updt_pressure": 72.0,
             "mg_down_l0": 26.0, "mg_up_l0": 26.0, "direct_transform": 16.0, "direct_tridiag": 48.0}


# ---- workloads ---------------------------------------------------------------------------------
def workload_grid(args):
    if args.grid:
        return tuple(args.grid)
    return {"channel": (1024, 512, 512), "tgv": (args.n,) * 3, "duct": (512, 128, 128)}[args.workload]


def make_case(workload, grid, ngpu, rundir):
    """Returns (case, owner, scaling, topology) with topology = (walls, fft_axis) as the CPU port wants it."""
    from snac_b200 import decks
    if workload == "channel":
        if ngpu not in decks.CHANNEL_DIMS:
            raise SystemExit("--gpus must be 1, 2, 4 or 8")
        case = decks.bench_channel(rundir, ng=grid, dims=decks.CHANNEL_DIMS[ngpu])
        return case, list(range(case.world.nrank)), "strong", ((0, 0, 1), 2)
    if workload == "tgv":
        case = decks.bench_tgv(rundir, n=grid[0], dims=TGV_DIMS[ngpu], weak=True)
        return case, list(range(case.world.nrank)), "weak", ((0, 0, 0), 3)
    if workload == "duct":
        nby, nbz = (2, 4) if ngpu == 8 else (2, 2)
        case = decks.bench_duct(rundir, ng=grid, nby=nby, nbz=nbz)
        nr = case.world.nrank
        return case, [r * ngpu // nr for r in range(nr)], "strong", ((0, 1, 1), 1)
    raise SystemExit(f"unknown workload {workload}")


def workload_name(workload, grid):
    return {"channel": "channel_%dx%dx%d_fft_y", "tgv": "tgv_%dx%dx%d_fft_z", "duct": "duct_%dx%dx%d_fft_x"}[workload] % tuple(grid)


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the CPU arm: oracle/snac_cpu.c on the host cores ---------------------------------------------
def global_fields(case, noise):
    """u,v,w,p of the whole domain as one ghosted array each (the CPU port works on one rectangle)."""
    from snac_b200.case import initial_fields
    W = case.world
    lo = np.min([b.lo for b in W.boxes], axis=0)
    hi = np.max([b.hi for b in W.boxes], axis=0)
    shp = tuple(int(x) + 2 for x in (hi - lo + 1))
    ini = initial_fields(case, noise=noise)
    out = {k: np.zeros(shp, order="F") for k in "uvwp"}
    for r, b in enumerate(W.boxes):
        o = [int(x) for x in (b.lo - lo)]
        sl = tuple(slice(o[d] + 1, o[d] + 1 + int(b.n[d])) for d in range(3))
        for k in "uvwp":
            out[k][sl] = ini[r][k][1:-1, 1:-1, 1:-1]
    return out, lo


def cpu_sim(case, topo, noise, tol=None, maxiter=None):
    from oracle.cpu_port import CpuSim
    P, W = case.params, case.world
    lo = np.min([b.lo for b in W.boxes], axis=0)
    hi = np.max([b.hi for b in W.boxes], axis=0)
    lmin = np.min([b.lmin for b in P.blocks], axis=0)
    lmax = np.max([b.lmax for b in P.blocks], axis=0)
    c = CpuSim(tuple(int(x) for x in hi - lo + 1), tuple(lmax - lmin), P.visc, tol=P.hypre_tol if tol is None else tol,
               maxiter=P.hypre_maxiter if maxiter is None else maxiter, walls=topo[0], fft_axis=topo[1], bforce=P.bforce)
    f, _ = global_fields(case, noise)
    for k, key in enumerate("uvwp"):
        c.set(k, f[key])
    c.init_bc()
    return c


def cpu_run(workload, grid, steps, warmup):
    """The C/OpenMP restatement of the reference's CPU pipeline (oracle/snac_cpu.c: one sweep per Fortran
    routine, radix-2 real FFTs, one geometric-multigrid solve per wavenumber to the same tolerance) on the
    same workload at grid `grid`, all host threads.  Returns (seconds per step, threads, V-cycles per solve)."""
    from oracle import cpu_port
    threads = cpu_port.use_all_cores()
    case, _, _, topo = make_case(workload, grid, 1, tempfile.mkdtemp())
    c = cpu_sim(case, topo, noise=(workload != "tgv"))
    cfl, icheck = case.params.cfl, max(case.params.icheck, 1)
    dt = cfl * c.chkdt()
    for _ in range(warmup):
        c.step(dt)
    t0 = time.perf_counter()
    for i in range(steps):
        c.step(dt)
        if (i + 1) % icheck == 0:   # chkdt + chkdiv as in the time loop (src/main.f90:1048-1069)
            dt = cfl * c.chkdt()
            c.divmax()
    sec = (time.perf_counter() - t0) / steps
    cyc = c.mean_cycles()
    c.close()
    return sec, threads, cyc


def cpu_sample_grid(workload, grid, budget_s=20.0, rate=40e6):
    """A grid of the same aspect ratio whose steps fit the time budget of the CPU sample (halved per
    direction until ten steps at ~`rate` points per second fit)."""
    g = list(grid)
    while (np.prod(g) * 10 / rate > budget_s or np.prod(g) > 70e6) and min(g) >= 32:   # 70 M points ~ 7 GB of host arrays
        g = [x // 2 for x in g]
    return tuple(g)


def cpu_baseline(workload, grid, steps=10):
    g = cpu_sample_grid(workload, grid)
    sec, threads, cyc = cpu_run(workload, g, steps, 1)
    return {"value": round(float(np.prod(g)) / sec / 1e6, 3), "unit": "Mpts*step/s", "cores": threads, "kind": "port",
            "sample": f"{workload_name(workload, g)} ({'the full workload' if tuple(g) == tuple(grid) else 'same case, every direction halved until 10 steps fit ~20 s'}), "
                      f"{steps} RK3 steps after 1 warm-up step, C/OpenMP restatement of the reference's sweeps + FFT + "
                      f"per-wavenumber multigrid at hypre_tol (oracle/snac_cpu.c; the reference binary needs a Fortran "
                      f"compiler, MPI, HYPRE and FFTW, none of which exist in this image), {sec * 1e3:.1f} ms/step, "
                      f"{cyc:.2f} V-cycles per plane solve"}


def run_reference(args, rank):
    if rank != 0:
        return
    grid = workload_grid(args)
    # the whole --steps/--warmup run has to end within a few minutes (~40 Mpts*step/s on 16 threads)
    g = cpu_sample_grid(args.workload, grid, budget_s=1800.0 / (args.steps + args.warmup + 2))
    sec, threads, cyc = cpu_run(args.workload, g, args.steps, args.warmup)
    value = float(np.prod(g)) / sec / 1e6
    base = {"value": round(value, 3), "unit": "Mpts*step/s", "cores": threads, "kind": "port",
            "sample": f"{workload_name(args.workload, g)}, {args.steps} RK3 steps after {args.warmup} warm-up steps, C/OpenMP "
                      f"restatement (oracle/snac_cpu.c), {cyc:.2f} V-cycles per plane solve"}
    scaling = "weak" if args.workload == "tgv" else "strong"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mpts*step/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, grid), "cpu_grid": list(g),
                       "note": "Mpts*step/s is size-normalised; the CPU arm runs the same case on cpu_grid with all host cores "
                               "and does not depend on the number of GPUs"},
            "cpu_baseline": base,
            "e2e": {"value": round(value, 3), "unit": "Mpts*step/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- parity check on the process grid of the run ---------------------------------------------------
CHECK_GRID = {"channel": (64, 32, 64), "tgv": (32, 32, 32), "duct": (32, 32, 32)}


def run_check(args, ngpu, rank, local, allreduce_sum, solver):
    """3 steps of a small case of the same topology on the same process grid, solves converged (1e-12),
    against the CPU restatement; returns {"rel_l2": worst over u,v,w,p, "divmax":, ...}."""
    from snac_b200.case import initial_fields
    from snac_b200.sim import Simulation
    grid = CHECK_GRID[args.workload]
    if args.workload == "tgv":
        grid = tuple(grid[0] for _ in range(3))
    case, owner, _, topo = make_case(args.workload, grid, ngpu, tempfile.mkdtemp())
    noise = args.workload != "tgv"
    g = Simulation(case, device=local, nproc=ngpu, myproc=rank, rank_owner=owner, tol=1e-12, maxiter=200, solver=solver)
    g.set_fields(initial_fields(case, ranks=g.ranks, noise=noise))
    g.init_bc()
    g.set_dt()
    c = cpu_sim(case, topo, noise, tol=1e-12, maxiter=300)
    dts = []
    for _ in range(3):
        dts.append(g.dt)
        g.step()
        c.step(dts[-1])
    lo0 = np.min([b.lo for b in case.world.boxes], axis=0)
    ref = [c.get(k) for k in range(4)]
    num = np.zeros(3)
    pdiff = []
    for r in g.ranks:
        b = case.world.boxes[r]
        o = [int(x) for x in (b.lo - lo0)]
        sl = tuple(slice(o[d] + 1, o[d] + 1 + int(b.n[d])) for d in range(3))
        f = g.get_fields(r)
        for k, key in enumerate("uvw"):
            d = f[key][1:-1, 1:-1, 1:-1] - ref[k][sl]
            num[k] += float(np.sum(d * d))
        pdiff.append(f["p"][1:-1, 1:-1, 1:-1] - ref[3][sl])
    divtot, divmax = g.chkdiv()
    g.close()
    # the pressure is defined up to a constant (all-Neumann/periodic): remove the mean of the difference first, then sum
    # the squares (one pass with sum d, sum d^2 cancels catastrophically when the constant dominates)
    s1 = allreduce_sum(np.array([sum(float(np.sum(d)) for d in pdiff), float(sum(d.size for d in pdiff))]))
    pmean = s1[0] / s1[1]
    tot = allreduce_sum(np.concatenate([num, [sum(float(np.sum((d - pmean) ** 2)) for d in pdiff)]]))
    vmag2 = sum(float(np.sum(ref[k][1:-1, 1:-1, 1:-1] ** 2)) for k in range(3))
    pint = ref[3][1:-1, 1:-1, 1:-1]
    pvar = float(np.sum((pint - pint.mean()) ** 2))
    errs = {}
    for k, key in enumerate("uvw"):
        den_k = max(float(np.sum(ref[k][1:-1, 1:-1, 1:-1] ** 2)), 1e-12 * vmag2)
        errs[key] = float(np.sqrt(tot[k] / den_k))
    errs["p"] = float(np.sqrt(tot[3] / max(pvar, 1e-300)))
    c.close()
    worst = max(errs.values())
    return {"rel_l2": float(f"{worst:.3e}"), "per_field": {k: float(f"{v:.2e}") for k, v in errs.items()},
            "divmax": float(f"{divmax:.3e}"), "case": workload_name(args.workload, grid), "steps": 3, "hypre_tol": 1e-12,
            "against": "oracle/snac_cpu.c (CPU restatement, pinned by tests/test_cpu_port.py)", "tolerance": 1e-8,
            "ok": bool(worst <= 1e-8 and divmax < 1e-8)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--workload", default="channel", choices=["channel", "tgv", "duct"])
    ap.add_argument("--grid", type=int, nargs=3, default=None, help="global grid of the workload")
    ap.add_argument("--n", type=int, default=256, help="tgv: cells per direction per GPU")
    ap.add_argument("--solver", default="auto", choices=["auto", "mg", "direct"],
                    help="plane solver: multigrid, the direct transform+tridiagonal solver where the planes are separable, "
                         "or auto (direct where possible)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the short run with the other plane solver")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ngpu = max(world, 1)
    if ngpu not in TGV_DIMS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    import ctypes as C

    from snac_b200 import lib as L
    from snac_b200.build import build_library
    from snac_b200.case import initial_fields
    from snac_b200.sim import Simulation
    if rank == 0:
        build_library()
    if world > 1:
        dist.barrier()

    def allreduce(a, op="sum"):
        a = np.asarray(a, dtype=np.float64)
        if world == 1:
            return a
        t = torch.tensor(a, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.cpu().numpy()

    # ---- parity on this process grid, before anything is timed
    check = None
    if not args.no_check:
        check = run_check(args, ngpu, rank, local, allreduce, args.solver)
        if not check["ok"]:
            if rank == 0:
                print(json.dumps({"check": check, "error": "parity check failed"}), flush=True)
            raise SystemExit(3)

    grid = workload_grid(args)
    case, owner, scaling, topo = make_case(args.workload, grid, ngpu, tempfile.mkdtemp())
    noise = args.workload != "tgv"

    def new_sim(solver):
        g = Simulation(case, device=local, nproc=ngpu, myproc=rank, rank_owner=owner, solver=solver)
        g.set_fields(initial_fields(case, ranks=g.ranks, noise=noise, mean_reduce=allreduce))
        g.init_bc()
        g.set_dt()
        return g

    g = new_sim(args.solver)
    npts_total = float(sum(np.prod(b.ng) for b in case.params.blocks))

    def barrier():
        g.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(sim, steps, record=None):
        ms = C.c_double()
        sim.lib.snac_timer_start(sim.h)
        for i in range(steps):
            sim.step()
            # V-cycles per plane solve: sampled where the time loop synchronises anyway (chkdt/chkdiv every icheck
            # steps) -- reading it after every step would add a device synchronisation to the timed region
            if record is not None and (sim.istep % max(case.params.icheck, 1) == 0 or i == steps - 1):
                record.append(sim.solver_mean_iters())
        sim.lib.snac_timer_stop(sim.h, C.byref(ms))
        return float(allreduce([ms.value], "max")[0])

    for _ in range(args.warmup):
        g.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = g.launch_count()
    iters = []
    t_ms = timed(g, args.steps, iters)
    barrier()
    launches = g.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = npts_total * args.steps / (t_ms * 1e-3) / 1e6
    # the run must have produced a flow, not garbage: divergence at the solver's level, finite fields
    divtot, divmax = g.chkdiv()
    it_last, relres = g.solver_stats()
    if not (np.isfinite(divmax) and divmax < 1e-2 and np.isfinite(g.dt) and g.dt > 0.0):
        raise SystemExit(f"the timed run diverged: divmax = {divmax}, dt = {g.dt}")
    solver_used = g.solver_name()

    # ---- end to end through the C ABI with host buffers: upload u,v,w,p, step, download u,v,w,p
    host = {}
    for r in g.ranks:
        shp = g.shape(r)
        for fid in (L.FIELD_U, L.FIELD_V, L.FIELD_W, L.FIELD_P):
            t = torch.empty(int(np.prod(shp)), dtype=torch.float64, pin_memory=True)
            a = t.numpy().reshape(shp, order="F")
            g._chk(g.lib.snac_download(g.h, r, fid, a.ctypes.data_as(C.c_void_p)))
            host[(r, fid)] = (t, a)
    nbytes = float(allreduce([sum(a.nbytes for _, a in host.values())])[0])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        for (r, fid), (_, a) in host.items():
            g.lib.snac_upload(g.h, r, fid, a.ctypes.data_as(C.c_void_p))
        g.istep += 1
        g.advance()
        for (r, fid), (_, a) in host.items():
            g.lib.snac_download(g.h, r, fid, a.ctypes.data_as(C.c_void_p))
    barrier()
    e2e_s = float(allreduce([time.perf_counter() - t0], "max")[0])
    e2e_val = npts_total * args.e2e_steps / e2e_s / 1e6
    del host

    # ---- per-kernel-class attribution (serialised, outside the timed region) and the roofline
    g.lib.snac_profile_enable(g.h, 1)
    nprof = 2
    for _ in range(nprof):
        g.advance()
    classes = {}
    c = 0
    while True:
        nm = g.lib.snac_profile_name(c)
        if nm is None:
            break
        m, nl, un = C.c_double(), C.c_longlong(), C.c_double()
        g.lib.snac_profile_get(g.h, c, C.byref(m), C.byref(nl))
        g.lib.snac_profile_units(g.h, c, C.byref(un))
        classes[nm.decode()] = (m.value / nprof, nl.value / nprof, un.value / nprof)
        c += 1
    g.lib.snac_profile_enable(g.h, 0)
    g.close()

    # ---- the other plane solver on the same workload (short run): both are reported
    alt = None
    if not args.no_alt and args.solver == "auto" and solver_used == "direct":
        g2 = new_sim("mg")
        for _ in range(min(args.warmup, 3)):
            g2.step()
        g2.sync()
        if world > 1:
            dist.barrier()
        it2 = []
        nst = max(3, min(args.steps, 10))
        t2 = timed(g2, nst, it2)
        alt = {"solver": "mg", "ms_per_step": round(t2 / nst, 4), "value": round(npts_total * nst / (t2 * 1e-3) / 1e6, 3),
               "steps": nst, "mg_vcycles_per_solve": round(float(np.mean(it2)), 2)}
        g2.close()

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        tot_prof = sum(v[0] for v in classes.values())
        table = {}
        for nm, (m, nl, un) in classes.items():
            if nl == 0:
                continue
            ent = {"ms_per_step": round(m, 4), "launches_per_step": nl, "share": round(m / tot_prof, 4) if tot_prof else None}
            if nm in ALG_BYTES and un > 0 and m > 0:
                ent["points_per_step"] = un
                ent["achieved_gbs"] = round(ALG_BYTES[nm] * un / (m * 1e-3) / 1e9, 1)
            table[nm] = ent
        dom = max((k for k in classes if k in ALG_BYTES and classes[k][1] > 0), key=lambda k: classes[k][0])
        m, nl, un = classes[dom]
        ach = ALG_BYTES[dom] * un / (m * 1e-3) / 1e9
        traffic = None
        try:   # dram__bytes_read+write of the dominant kernel from the committed ncu capture (taken on a smaller grid of
            # the same workload: bytes per grid point there x the points of a launch here)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                t = json.load(f).get(f"{args.workload}:{dom}")
            if t:
                traffic = round(t["dram_bytes_per_point"] * un / nl, 1)
        except (OSError, KeyError, TypeError, ValueError):
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                "bytes_per_launch": round(ALG_BYTES[dom] * un / nl, 1), "avg_launch_ms": round(m / nl, 5),
                "algorithmic_bytes_per_point": ALG_BYTES[dom], "classes": table}
        cfg = {"workload": workload_name(args.workload, grid), "grid": [int(x) for x in grid],
               "blocks": len(case.params.blocks), "block_dims": [int(x) for x in case.params.blocks[0].dims],
               "fft_axis": int(case.params.fft_axis), "hypre_tol": case.params.hypre_tol, "visci": 1.0 / case.params.visc,
               "cfl": case.params.cfl, "icheck": case.params.icheck, "plane_solver": solver_used,
               "l2": "working set (every field array of a GPU is larger than the 126 MB L2 and is streamed once per kernel)",
               "mg_vcycles_per_solve": round(float(np.mean(iters)), 2) if solver_used == "mg" else None,
               "divmax_after_run": float(f"{divmax:.3e}"), "last_solve_relres": float(f"{relres:.3e}"),
               "initial_guess": "mg: the same RK sub-step one time step earlier (the reference passes the previous sub-step's "
                                "solution); at hypre_tol the fields differ from a reference run at the hypre_tol level"}
        line = {"metric": METRIC, "value": round(value, 3), "unit": "Mpts*step/s", "n_gpus": ngpu, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(t_ms / args.steps, 4), "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None,
                "dtype": "f64" + (" (fp32 coarse-grid correction inside the multigrid cycle)" if solver_used == "mg" else ""),
                "data": "synthetic", "config": cfg, "check": check, "gpu_launches": int(launches), "clocks": clocks,
                "e2e": {"value": round(e2e_val, 3), "unit": "Mpts*step/s", "h2d_bytes_per_step": int(nbytes),
                        "d2h_bytes_per_step": int(nbytes), "ms_per_step": round(e2e_s / args.e2e_steps * 1e3, 3),
                        "what": "snac_upload(u,v,w,p) from pinned host + snac_step + snac_download(u,v,w,p); bytes summed over all GPUs"},
                "roofline": roof}
        if alt:
            line["other_solver"] = alt
        if not args.no_cpu and ngpu == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload, grid)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
