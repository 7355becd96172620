"""Multi-process parity check, launched with torchrun (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi_gpu_check.py [case ...]

Every process owns the boxes rank % N == its rank, advances a few RK3 steps through the C ABI with
halos and solver transposes over peer memory, and rank 0 compares all boxes with the CPU oracle."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    import cases
    from oracle import snac_oracle as orc
    from snac_b200 import lib as L
    from snac_b200.case import initial_fields
    from snac_b200.sim import Simulation
    from conftest import GOLDEN
    todo = [("channel", 2, dict(dims=(world, 1, 1), ng=(16 * world, 8, 20))),
            ("channel", 2, dict(dims=(1, 1, world), ng=(32, 8, 12 * world))),
            ("tgv", 3, dict(n=32, dims=(world, 1, 1))),
            ("duct", 1, {}),
            # boxes coupled across processes inside the solver: the 3-D pressure multigrid (FFT_AXIS = 0), the L-shaped
            # multi-block geometry, implicit diffusion (Helmholtz solves on the staggered grids) + a transported scalar
            ("channel", 0, dict(dims=(world, 1, 1), ng=(8 * world, 6, 10))),
            ("channel", 0, dict(dims=(1, 1, world), ng=(12, 6, 10 * world), gr_z=0.0)),
            ("lshape", 1, {}),
            ("dhc", 0, dict(impdiff=True, boussinesq=True)),
            # implicit diffusion in an FFT_AXIS build (solve_impdiff_field_fft): the reference's test case and a channel
            ("dhc", 2, dict(impdiff=True, boussinesq=True)),
            ("channel", 2, dict(dims=(world, 1, 1), ng=(8 * world, 8, 12), impdiff=True))]
    ok = True
    for kind, fa, kw in todo:
        from snac_b200.case import setup_case
        imp, bou = bool(kw.pop("impdiff", False)), bool(kw.pop("boussinesq", False))
        if kind == "dhc":   # the reference's differentially-heated-cavity test case (2 x 1 x 2 boxes)
            case = setup_case(os.path.join(GOLDEN, "dhc"), fft_axis=fa, impdiff=imp, boussinesq=bou)
        else:
            tmp = [tempfile.mkdtemp() if rank == 0 else None]
            dist.broadcast_object_list(tmp, 0)
            if rank == 0:
                case = cases.make(kind, tmp[0], fft_axis=fa, impdiff=imp, boussinesq=bou, **kw)
            dist.barrier()
            if rank != 0:
                case = setup_case(tmp[0], fft_axis=fa, impdiff=imp, boussinesq=bou)
        nr = case.world.nrank
        if nr < world:   # fewer boxes than processes (e.g. the three L-shaped blocks on four GPUs): nothing to own
            if rank == 0:
                print(f"[{world} GPUs] {kind} fft_axis={fa} {kw}: {nr} boxes, skipped", flush=True)
            dist.barrier()
            continue
        owner = [r % world for r in range(nr)]
        g = Simulation(case, device=local, nproc=world, myproc=rank, rank_owner=owner, tol=1e-12, maxiter=200)
        ini = initial_fields(case, noise=(kind in ("channel", "duct", "lshape")))
        g.set_fields(ini)
        g.init_bc(); g.set_dt()
        for _ in range(3):
            g.step()
        mine = {r: g.get_fields(r) for r in g.ranks}
        divtot, divmax = g.chkdiv()
        allf = [None] * world
        dist.all_gather_object(allf, mine)
        if rank == 0:
            o = orc.OracleSim(case.params, case.world, case.grids, case.bgrids)
            for r in range(nr):
                o.u[r][...] = ini[r]["u"]; o.v[r][...] = ini[r]["v"]; o.w[r][...] = ini[r]["w"]; o.pr[r][...] = ini[r]["p"]
                for i in range(case.params.nscal):
                    o.s[i][r][...] = ini[r]["s"][i]
            o.init_bc(); o.set_dt()
            for _ in range(3):
                o.step()
            worst = 0.0
            for d in allf:
                for r, f in d.items():
                    I = (slice(1, -1),) * 3   # interior (ghost cells at re-entrant edges of the L belong to no neighbour)
                    vmag = np.sqrt(sum(np.linalg.norm(a[r][I]) ** 2 for a in (o.u, o.v, o.w)))
                    for k, a in (("u", o.u), ("v", o.v), ("w", o.w)):
                        worst = max(worst, float(np.linalg.norm(f[k][I] - a[r][I]) / max(np.linalg.norm(a[r][I]), 1e-6 * vmag, 1e-300)))
            good = worst < 1e-8 and divmax < 1e-9 and abs(g.dt - o.dt) <= 1e-12 * o.dt
            ok = ok and good
            print(f"[{world} GPUs] {kind} fft_axis={fa} {kw}: max rel L2 err vs oracle {worst:.3e}, divmax {divmax:.3e}, "
                  f"dt {g.dt:.6e} vs {o.dt:.6e} -> {'OK' if good else 'FAIL'}", flush=True)
        g.close()
        dist.barrier()
    res = [ok]
    dist.broadcast_object_list(res, 0)
    dist.destroy_process_group()
    sys.exit(0 if res[0] else 1)


if __name__ == "__main__":
    main()
