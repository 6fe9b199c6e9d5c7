import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import water_box
from oracle import oracle_graph as og
from oracle import oracle_np as onp
from tensormol_b200.engine import Engine, random_weights
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 20
hidden = [500, 500, 500]
Z, X, lat = water_box(nx)
X = onp.modulo_lattice(lat, X)
P = og.default_params()
eng = Engine([1, 8], hidden, P)
eng.set_weights(random_weights([1, 8], eng.D, hidden, 0))
for it in range(4):
    t0 = time.time()
    r = eng.evaluate_lattice(X, Z, lat, 1)
    t1 = time.time()
    print("natom", len(Z), "wall ms", (t1 - t0) * 1e3, "E", r["Etotal"][0], {k: round(v, 3) for k, v in eng.timings().items() if isinstance(v, float)})
