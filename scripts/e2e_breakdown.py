"""Host-buffer call on the 24,000-atom box: wall time per call against the device time of the replayed graph (copies included)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from bench import hot_params
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
Z, X, lat = water_box(20, spacing=3.1072, seed=3)
X = wrap_into_cell(X, lat)
n = len(Z)
eng = Engine([1, 8], [500] * 3, hot_params())
eng.set_weights(random_weights([1, 8], eng.D, [500] * 3, 0))
want = ("Etotal", "gradient")
def run(f, tag, k=300):
    for _ in range(6): f()
    torch.cuda.synchronize(); dev = 0.0
    t0 = time.perf_counter()
    for _ in range(k):
        f(); dev += eng.timings()["total"]
    t1 = time.perf_counter()
    t2 = time.perf_counter()
    for _ in range(k): f()
    t3 = time.perf_counter()
    print(f"{tag}: wall {1e3*(t3-t2)/k:.4f} ms/call, graph on device {dev/k:.4f} ms, host side {(1e3*(t3-t2)-dev)/k:.4f} ms")
run(lambda: eng.evaluate_lattice(X, Z, lat, 1, outputs=want), "pageable evaluate_lattice")
Xp, Zp = eng.pinned(X.shape), eng.pinned(Z.shape, np.int32); Xp[:] = X; Zp[:] = Z
into = {"gradient": eng.pinned((1, n, 3))}
run(lambda: eng.evaluate_lattice(Xp, Zp, lat, 1, outputs=want, into=into), "pinned evaluate_lattice")
run(eng.bind_lattice(Xp, Zp, lat, 1, outputs=want, into=into), "pinned bind_lattice")
run(eng.bind_lattice(Xp, Zp, lat, 1, outputs=("Etotal",), into=None), "pinned bound, energy only out")
