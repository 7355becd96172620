#!/usr/bin/env python
"""Benchmark of the per-time-step pressure-correction pipeline (BASELINE.json metric:
ms per RK3 step and Mpts*step/s; HBM roofline of the dominant kernel).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--n 256]

N = 1: the Taylor-Green vortex of the reference (examples/_manuscript_taylor_green_vortex/*.nml:
triperiodic, visci = 1600, cfl = .95, icheck = 10, default solver tolerance 1e-4) on 256^3 with
FFT_AXIS = 3, one block on one GPU.  N > 1: the same case weak-scaled, 256^3 cells per GPU, the
block split over the GPUs in x and y (the transform axis z stays whole, src/sanity.f90:75-79).
A step = one pass of the reference's time loop body (src/main.f90:897-1069) without file output:
three RK sub-steps and, every icheck steps, chkdt + chkdiv.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIMS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1)}

# algorithmic HBM bytes per grid point and launch of each kernel class
ALG_BYTES = {"rk_mom": 96.0, "fillps": 32.0, "fft_fwd": 16.0, "fft_bwd": 16.0, "correc_updt_pressure": 72.0,
             "mg_down_l0": 26.0, "mg_up_l0": 26.0}


def make_tgv(n, dims, rundir):
    import cases
    L = 2.0 * np.pi
    ng = tuple(n * d for d in dims)
    blk = dict(dims=dims, ng=ng, lmin=(0, 0, 0), lmax=tuple(L * d for d in dims), cbcvel=["FFFFFF"] * 3,
               cbcpre="FFFFFF", inivel="tgv")
    cases.write_case(rundir, [blk], visci=1600.0, cfl=0.95, icheck=10)
    from snac_b200.case import setup_case
    return setup_case(rundir, fft_axis=3)


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


def _periodic_ghosts(a):
    a = np.array(a, order="F", copy=True)
    a[0] = a[-2]; a[-1] = a[1]
    a[:, 0] = a[:, -2]; a[:, -1] = a[:, 1]
    a[:, :, 0] = a[:, :, -2]; a[:, :, -1] = a[:, :, 1]
    return a


def cpu_run(n_cpu, steps, warmup):
    """The C/OpenMP restatement of the reference's CPU pipeline (oracle/snac_cpu.c: one sweep per
    Fortran routine, radix-2 real FFTs along z, one geometric-multigrid solve per wavenumber to the
    same tolerance) on the same TGV workload, all host threads.  Returns (seconds per step, sim)."""
    from oracle.cpu_port import CpuSim
    from snac_b200.case import initial_fields
    case = make_tgv(n_cpu, (1, 1, 1), tempfile.mkdtemp())
    ini = initial_fields(case)
    L = 2.0 * np.pi
    c = CpuSim((n_cpu,) * 3, (L, L, L), case.params.visc, tol=case.params.hypre_tol, maxiter=case.params.hypre_maxiter)
    for k, key in enumerate("uvwp"):
        c.set(k, _periodic_ghosts(ini[0][key]))
    dt = case.params.cfl * c.chkdt()
    for i in range(warmup):
        c.step(dt)
    t0 = time.perf_counter()
    for i in range(steps):
        c.step(dt)
        if (i + 1) % 10 == 0:   # icheck = 10: chkdt + chkdiv as in the time loop (src/main.f90:1048-1069)
            dt = case.params.cfl * c.chkdt()
            c.divmax()
    sec = (time.perf_counter() - t0) / steps
    return sec, c


def cpu_baseline(n_cpu=256, steps=5):
    sec, c = cpu_run(n_cpu, steps, 1)
    out = {"value": n_cpu ** 3 / sec / 1e6, "unit": "Mpts*step/s", "cores": c.threads, "kind": "port",
           "sample": f"TGV {n_cpu}^3 (the full workload), {steps} RK3 steps after 1 warm-up step, C/OpenMP restatement of the "
                     f"reference's sweeps + FFT + per-wavenumber multigrid at hypre_tol (oracle/snac_cpu.c; the reference "
                     f"binary needs a Fortran compiler, MPI, HYPRE and FFTW, none of which exist in this image), "
                     f"{sec * 1e3:.1f} ms/step, {c.mean_cycles():.2f} V-cycles per plane solve"}
    c.close()
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    # one untimed step sizes the run: the whole --steps/--warmup run has to end within a few minutes
    n_cpu = args.n
    sec1, c = cpu_run(n_cpu, 1, 1)
    c.close()
    if sec1 * (args.steps + args.warmup) > 240.0:
        n_cpu = args.n // 2
    sec, c = cpu_run(n_cpu, args.steps, args.warmup)
    value = n_cpu ** 3 / sec / 1e6
    base = {"value": value, "unit": "Mpts*step/s", "cores": c.threads, "kind": "port",
            "sample": f"TGV {n_cpu}^3, {args.steps} RK3 steps after {args.warmup} warm-up steps, C/OpenMP restatement "
                      f"(oracle/snac_cpu.c), {c.mean_cycles():.2f} V-cycles per plane solve"}
    c.close()
    line = {"impl": "reference", "metric": "Mpts*step/s (RK3 step of the pressure-correction pipeline)",
            "value": round(value, 3), "unit": "Mpts*step/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"tgv{args.n}_fft_z", "cpu_grid": [n_cpu] * 3},
            "cpu_baseline": base,
            "e2e": {"value": round(value, 3), "unit": "Mpts*step/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--n", type=int, default=256, help="cells per direction per GPU")
    ap.add_argument("--n-cpu", type=int, default=0, help="grid of the cpu_baseline sample (default: --n)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
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
    if ngpu not in DIMS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    from snac_b200 import lib as L
    from snac_b200.build import build_library
    from snac_b200.case import initial_fields
    from snac_b200.sim import Simulation
    if rank == 0:
        build_library()
    if world > 1:
        dist.barrier()
    dims = DIMS[ngpu]
    rundir = tempfile.mkdtemp()
    case = make_tgv(args.n, dims, rundir)
    owner = list(range(case.world.nrank))
    g = Simulation(case, device=local, nproc=ngpu, myproc=rank, rank_owner=owner)
    g.set_fields(initial_fields(case, ranks=g.ranks))
    g.init_bc()
    g.set_dt()
    npts_total = float(np.prod(case.params.blocks[0].ng))

    def barrier():
        g.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        g.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    import ctypes as C
    l0 = g.launch_count()
    ms = C.c_double()
    g.lib.snac_timer_start(g.h)
    iters = []
    for i in range(args.steps):
        g.step()
        # V-cycles per plane solve: sampled where the time loop synchronises anyway (chkdt/chkdiv every icheck
        # steps) -- reading it after every step would add a device synchronisation to the timed region
        if g.istep % max(case.params.icheck, 1) == 0 or i == args.steps - 1:
            iters.append(g.solver_mean_iters())
    g.lib.snac_timer_stop(g.h, C.byref(ms))
    barrier()
    launches = g.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t_ms = ms.value
    if world > 1:
        t = torch.tensor([t_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t.item())
    value = npts_total * args.steps / (t_ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers: upload u,v,w,p, step, download u,v,w,p
    r0 = g.ranks[0]
    shp = g.shape(r0)
    host = {}
    for fid in (L.FIELD_U, L.FIELD_V, L.FIELD_W, L.FIELD_P):
        t = torch.empty(int(np.prod(shp)), dtype=torch.float64, pin_memory=True)
        a = t.numpy().reshape(shp, order="F")
        a[...] = g.download(r0, fid)
        host[fid] = (t, a)
    nbytes = sum(a.nbytes for _, a in host.values())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        for fid, (_, a) in host.items():
            g.lib.snac_upload(g.h, r0, fid, a.ctypes.data_as(C.c_void_p))
        g.istep += 1
        g.advance()
        for fid, (_, a) in host.items():
            g.lib.snac_download(g.h, r0, fid, a.ctypes.data_as(C.c_void_p))
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = npts_total * args.e2e_steps / e2e_s / 1e6

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
            ent = {"ms_per_step": round(m, 4), "launches_per_step": nl, "share": round(m / tot_prof, 4) if tot_prof else None}
            if nm in ALG_BYTES and un > 0 and m > 0:
                ent["points_per_step"] = un
                ent["achieved_gbs"] = round(ALG_BYTES[nm] * un / (m * 1e-3) / 1e9, 1)
            table[nm] = ent
        dom = max((k for k in classes if k in ALG_BYTES), key=lambda k: classes[k][0])
        m, nl, un = classes[dom]
        ach = ALG_BYTES[dom] * un / (m * 1e-3) / 1e9
        traffic = None
        try:   # dram__bytes_read+write per launch of the dominant kernel, from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except OSError:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                "bytes_per_launch": round(ALG_BYTES[dom] * un / nl, 1), "avg_launch_ms": round(m / nl, 5),
                "algorithmic_bytes_per_point": ALG_BYTES[dom], "classes": table}
        line = {"metric": "Mpts*step/s (RK3 step of the pressure-correction pipeline)", "value": round(value, 3),
                "unit": "Mpts*step/s", "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(t_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"tgv{args.n}_fft_z", "grid": list(map(int, case.params.blocks[0].ng)),
                           "block_dims": list(dims), "fft_axis": 3, "hypre_tol": case.params.hypre_tol,
                           "visci": 1600.0, "cfl": 0.95, "icheck": 10,
                           "l2": "working set (>= 1.6 GB of fields per GPU) exceeds the 126 MB L2",
                           "mg_vcycles_per_solve": round(float(np.mean(iters)), 2)},
                "gpu_launches": int(launches), "clocks": clocks,
                "e2e": {"value": round(e2e_val, 3), "unit": "Mpts*step/s", "h2d_bytes_per_step": int(nbytes),
                        "d2h_bytes_per_step": int(nbytes), "ms_per_step": round(e2e_s / args.e2e_steps * 1e3, 3),
                        "what": "snac_upload(u,v,w,p) from pinned host + snac_step + snac_download(u,v,w,p)"},
                "roofline": roof}
        if not args.no_cpu and ngpu == 1:
            line["cpu_baseline"] = cpu_baseline(args.n_cpu or args.n, 5)
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
