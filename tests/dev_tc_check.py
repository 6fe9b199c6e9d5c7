"""Dev check (test infrastructure, run by hand: python tests/dev_tc_check.py [nx]): GEMM modes against the fp32 FFMA mode and the oracle."""
import os
os.environ.setdefault("TM_NO_GRAPH", "1")   # per-stage timings need the kernel-by-kernel path
import sys
import time

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from oracle import oracle_graph as og
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 6
hidden = [500, 500, 500]
Z, X, lat = water_box(nx)
X = wrap_into_cell(X, lat)
P = og.default_params()
eng = Engine([1, 8], hidden, P)
W = random_weights([1, 8], eng.D, hidden, 0)
eng.set_weights(W)
res = {}
for mode in (0, 1):
    eng.set_gemm_mode(mode)
    for it in range(3):
        t0 = time.time()
        r = eng.evaluate_lattice(X, Z, lat, 1)
        t1 = time.time()
    res[mode] = r
    print("mode", mode, "natom", len(Z), "wall ms", round((t1 - t0) * 1e3, 3), "E", r["Etotal"][0],
          {k: round(v, 3) for k, v in eng.timings().items() if isinstance(v, float)}, flush=True)
d = res[1]
o = res[0]
print("dE", d["Etotal"][0] - o["Etotal"][0], "rel", abs(d["Etotal"][0] - o["Etotal"][0]) / abs(o["Etotal"][0]))
print("max|dgrad|", np.abs(d["gradient"] - o["gradient"]).max(), "max|grad|", np.abs(o["gradient"]).max())
print("max|dq|", np.abs(d["charge"] - o["charge"]).max())
if nx <= 6:
    from oracle import oracle_np as onp
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])
    oo = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, len(Z))
    for mode in (0, 1):
        r = res[mode]
        print("mode", mode, "vs oracle: rel dE", abs(r["Etotal"][0] - oo["Etotal"][0]) / abs(oo["Etotal"][0]),
              "max|dgrad| Ha/Bohr", np.abs(r["gradient"][0] - oo["gradient"][0, :len(Z)]).max() / 1.889725989)
