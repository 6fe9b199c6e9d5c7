"""Development helper: evaluates a small aperiodic system with wide nets, printing each stage with a time stamp."""
import faulthandler, os, sys, time
os.environ.setdefault("TM_NO_GRAPH", "1")
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
faulthandler.dump_traceback_later(int(os.environ.get("DIAG_DUMP", "18")), exit=True)
import numpy as np
t0 = time.perf_counter()
def say(*a):
    print(f"[{time.perf_counter() - t0:7.2f}]", *a, flush=True)
H, mode, which = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
from conftest import load_golden
from tensormol_b200.engine import default_params
from tensormol_b200.engine import Engine, random_weights
g = load_golden(which)
eles = [int(e) for e in g["eles"]]
say("start", which, "H", H, "mode", mode)
eng = Engine(eles, [H] * 3, default_params())
W = random_weights(eng.eles, eng.D, [H] * 3, 1)
say("weights made")
eng.set_weights(W)
say("weights set")
eng.set_gemm_mode(mode)
N = len(g["Z"])
if "lattice" in g:
    fn = lambda f: eng.evaluate_lattice(g["xyz"], g["Z"], g["lattice"], int(g["ntess"]), do_force=f)
else:
    fn = lambda f: eng.evaluate(g["xyz"][None], g["Z"][None], np.array([N]), do_force=f)
r = fn(False)
say("energy only", r["Etotal"], eng.timings())
r = fn(True)
say("energy+force", r["Etotal"], np.abs(r["gradient"]).max(), eng.timings())
