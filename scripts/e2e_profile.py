"""Host-buffer call (tm_eval_lattice) broken down: wall per call vs the device stage timings."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from bench import hot_params, HIDDEN
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
Z, X, lat = water_box(20, spacing=3.1072, seed=3, jitter=0.05)
X = wrap_into_cell(X, lat)
eng = Engine([1, 8], HIDDEN, hot_params())
eng.set_weights(random_weights([1, 8], eng.D, HIDDEN, 0))
for it in range(4):
    eng.evaluate_lattice(X, Z, lat, 1)
ws = []
for it in range(10):
    t0 = time.perf_counter(); r = eng.evaluate_lattice(X, Z, lat, 1); ws.append(time.perf_counter() - t0)
print("wall ms/call", [round(w * 1e3, 3) for w in ws])
print({k: round(v, 3) for k, v in eng.timings().items() if isinstance(v, float)})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for it in range(10): eng.evaluate_lattice(X, Z, lat, 1)
pr.disable(); pstats.Stats(pr).sort_stats('cumtime').print_stats(12)
