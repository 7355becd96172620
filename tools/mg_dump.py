import sys, os, tempfile
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, bench
from snac_b200.case import initial_fields
from snac_b200.sim import Simulation
case = bench.make_tgv(256,(1,1,1),tempfile.mkdtemp())
g = Simulation(case); g.set_fields(initial_fields(case)); g.init_bc(); g.set_dt()
for i in range(12): g.step()
os.environ["SNAC_MG_DUMP"]="1"
import ctypes as C
RK=[(32/60,0),(25/60,-17/60),(45/60,-25/60)]
# run the pieces of one more step by hand: after every solve print the per-plane iteration counts
g.advance()
g.solver_stats()
