"""torchrun --nproc-per-node 4 tools/n4_debug.py: small channel on block_dims (2,1,2), where do the fields differ from the CPU restatement?"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from snac_b200.case import initial_fields
from snac_b200.sim import Simulation
from snac_b200 import lib as L
nsteps = int(os.environ.get("NSTEPS", "1"))
case, owner, _, topo = bench.make_case("channel", (64, 32, 64), world, tempfile.mkdtemp())
g = Simulation(case, device=local, nproc=world, myproc=rank, rank_owner=owner, tol=1e-12, maxiter=200, solver=os.environ.get("SOLVER", "auto"))
g.set_fields(initial_fields(case, ranks=g.ranks, noise=True))
g.init_bc(); g.set_dt()
c = bench.cpu_sim(case, topo, True, tol=1e-12, maxiter=300)
for _ in range(nsteps):
    dt = g.dt; g.step(); c.step(dt)
ref = [c.get(k) for k in range(4)]
lo0 = np.min([b.lo for b in case.world.boxes], axis=0)
for r in g.ranks:
    b = case.world.boxes[r]
    o = [int(x) for x in (b.lo - lo0)]
    sl = tuple(slice(o[d] + 1, o[d] + 1 + int(b.n[d])) for d in range(3))
    f = g.get_fields(r)
    msg = [f"rank {r} lo {b.lo.tolist()}"]
    for k, key in enumerate("uvw"):
        d = np.abs(f[key][1:-1, 1:-1, 1:-1] - ref[k][sl])
        idx = np.unravel_index(np.argmax(d), d.shape)
        # share of the squared error within 2 cells of each face
        tot = float(np.sum(d ** 2)) + 1e-300
        faces = []
        for ax in range(3):
            a = np.moveaxis(d, ax, 0)
            faces.append((round(float(np.sum(a[:2] ** 2)) / tot, 2), round(float(np.sum(a[-2:] ** 2)) / tot, 2)))
        msg.append(f"{key}: max {d.max():.2e} at {idx} of {d.shape}, err share near faces (lo,hi) x,y,z: {faces}")
    print(" | ".join(msg), flush=True)
g.close(); c.close()
dist.barrier(); dist.destroy_process_group()
