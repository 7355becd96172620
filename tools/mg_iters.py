"""Diagnostic: V-cycles per wavenumber plane of the TGV benchmark case (run on a GPU box).  With a library built
with -DSNAC_COARSE_TIMING (tools/build_variant.py, SNAC_B200_LIB=...) the coarse kernel also prints the clock
cycles of its phases for plane q = 1."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["SNAC_MG_DUMP"] = "1"
import bench
from snac_b200.case import initial_fields
from snac_b200.sim import Simulation
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
case = bench.make_tgv(n, (1, 1, 1), tempfile.mkdtemp())
g = Simulation(case)
g.set_fields(initial_fields(case)); g.init_bc(); g.set_dt()
for i in range(12):
    g.step()
print("mean iters", g.solver_mean_iters(), "stats", g.solver_stats())
g.close()
