"""One rank's share of an N-way slab step on a single GPU (no collectives): stage timings + wall time per phase."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from bench import hot_params
HIDDEN = [500, 500, 500]
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
from tensormol_b200.parallel import EngineSlabBackend
from tensormol_b200._lib import TM_F_FORCE, TM_F_VDW
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = world // 2
Z, X, lat = water_box(20, spacing=3.1072, seed=3)
X = wrap_into_cell(X, lat)
n = len(Z)
eng = Engine([1, 8], HIDDEN, hot_params())
eng.set_weights(random_weights([1, 8], eng.D, HIDDEN, 0))
b = EngineSlabBackend(eng)
dev = torch.device("cuda", 0)
xt = torch.tensor(X, dtype=torch.float64, device=dev); zt = torch.tensor(Z, dtype=torch.int32, device=dev)
q = torch.zeros(n, dtype=torch.float64, device=dev); e = torch.zeros(6, dtype=torch.float64, device=dev); g = torch.zeros(n, 3, dtype=torch.float64, device=dev)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b.slab_phase_a(xt, zt, n, lat, 1, rank, world, q); eng.sync(); t1 = time.perf_counter()
    b.slab_phase_b(q, e); eng.sync(); t2 = time.perf_counter()
    b.slab_phase_c(e, TM_F_FORCE | TM_F_VDW, g); eng.sync(); t3 = time.perf_counter()
print("world", world, "rank", rank, "wall ms A,B,C:", round((t1-t0)*1e3,3), round((t2-t1)*1e3,3), round((t3-t2)*1e3,3))
print({k: round(v, 3) for k, v in eng.timings().items() if isinstance(v, float)})
# the same three phases replayed from one CUDA graph (what a rank executes per step, minus the exchange waits)
if "--graph" in sys.argv:
    import ctypes as C
    from tensormol_b200.engine import GraphedCall
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(C.c_void_p(stream.cuda_stream))
    def step():
        b.slab_phase_a(xt, zt, n, lat, 1, rank, world, q)
        b.slab_phase_b(q, e)
        b.slab_phase_c(e, TM_F_FORCE | TM_F_VDW, g)
    gc = GraphedCall(step, stream, warmup=2)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5): gc()
    stream.synchronize()
    ev0.record(stream)
    for _ in range(50): gc()
    ev1.record(stream)
    stream.synchronize()
    print("graph replay, ms per rank-step:", round(ev0.elapsed_time(ev1) / 50, 4))
