"""Dev check: GEMM column-tile width (mode 3 = 64 forced, 4 = 128 forced, 1 = chosen by size) on water boxes:
identical results, stage timings."""
import os
os.environ.setdefault("TM_NO_GRAPH", "1")   # per-stage timings need the kernel-by-kernel path
import sys
sys.path.insert(0, '.')
import numpy as np
from bench import hot_params, HIDDEN
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights

for n in [int(a) for a in sys.argv[1:]] or [10, 20]:
    Z, X, lat = water_box(n, spacing=3.1072, seed=3)
    X = wrap_into_cell(X, lat)
    eng = Engine([1, 8], HIDDEN, hot_params())
    eng.set_weights(random_weights([1, 8], eng.D, HIDDEN, 0))
    ref = None
    for mode in (4, 3, 1):
        eng.set_gemm_mode(mode)
        acc = {}
        for it in range(8):
            r = eng.evaluate_lattice(X, Z, lat, 1)
            t = eng.timings()
            if it >= 3:
                for k in ("total", "mlp_fwd", "mlp_bwd"):
                    acc[k] = acc.get(k, 0) + t[k] / 5
        g = r["gradient"][0]
        if ref is None:
            ref = (r["Etotal"][0], g.copy())
        print("atoms", len(Z), "mode", mode, {k: round(v, 4) for k, v in acc.items()}, "E", r["Etotal"][0],
              "dE vs mode 4", r["Etotal"][0] - ref[0], "max |dF|", float(np.abs(g - ref[1]).max()))
    eng.close()
