"""Stage timings of the 24k-atom step (dev helper)."""
import os
os.environ.setdefault("TM_NO_GRAPH", "1")   # per-stage timings need the kernel-by-kernel path
import sys
sys.path.insert(0, '.')
from bench import hot_params, HIDDEN
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
Z, X, lat = water_box(20, spacing=3.1072, seed=3)
X = wrap_into_cell(X, lat)
eng = Engine([1, 8], HIDDEN, hot_params())
eng.set_weights(random_weights([1, 8], eng.D, HIDDEN, 0))
import numpy as np
acc = {}
for it in range(8):
    try:
        r = eng.evaluate_lattice(X, Z, lat, 1)
    except Exception as ex:
        r = None
    t = eng.timings()
    if it >= 3:
        for k, v in t.items():
            if isinstance(v, float): acc[k] = acc.get(k, 0) + v / 5
print({k: round(v, 3) for k, v in acc.items()}, "E", None if r is None else r["Etotal"][0])
