"""Dev check: CTA-pair GEMM mode (2) against modes 0 and 1."""
import os
os.environ.setdefault("TM_NO_GRAPH", "1")   # per-stage timings need the kernel-by-kernel path
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from tensormol_b200.engine import default_params
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 6
hidden = [500, 500, 500]
Z, X, lat = water_box(nx); X = wrap_into_cell(X, lat)
P = default_params()
eng = Engine([1, 8], hidden, P)
eng.set_weights(random_weights([1, 8], eng.D, hidden, 0))
res = {}
for mode in (1, 2):
    eng.set_gemm_mode(mode)
    for it in range(4):
        r = eng.evaluate_lattice(X, Z, lat, 1)
    res[mode] = r
    print("mode", mode, "E", r["Etotal"][0], {k: round(v, 3) for k, v in eng.timings().items() if isinstance(v, float) and k in ("total", "mlp_fwd", "mlp_bwd")}, flush=True)
print("rel dE", abs(res[2]["Etotal"][0] - res[1]["Etotal"][0]) / abs(res[1]["Etotal"][0]), "max|dgrad|", np.abs(res[2]["gradient"] - res[1]["gradient"]).max(), "max|dq|", np.abs(res[2]["charge"] - res[1]["charge"]).max())
