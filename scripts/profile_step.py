"""One 24k-atom periodic step in a given GEMM mode, for ncu launch lists / captures."""
import os
os.environ.setdefault("TM_NO_GRAPH", "1")   # per-stage timings need the kernel-by-kernel path
import sys
import numpy as np
sys.path.insert(0, '.')
from bench import hot_params, HIDDEN
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
Z, X, lat = water_box(nx, spacing=3.1072, seed=3)
X = wrap_into_cell(X, lat)
eng = Engine([1, 8], HIDDEN, hot_params())
eng.set_weights(random_weights([1, 8], eng.D, HIDDEN, 0))
eng.set_gemm_mode(mode)
for _ in range(steps):
    r = eng.evaluate_lattice(X, Z, lat, 1)
print("E", r["Etotal"][0], eng.timings())
