import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from common import water_box
from tensormol_b200.engine import default_params
from tensormol_b200.engine import Engine, random_weights
from tensormol_b200.parallel import EngineSlabBackend
from tensormol_b200._lib import TM_F_FORCE, TM_F_VDW
from tensormol_b200.SystemBuilders import wrap_into_cell
world = 2
Z, X, lat = water_box(5, jitter=0.03); X = wrap_into_cell(X, lat); n = len(Z)
P = default_params(); hidden = [64, 48]; W = random_weights([1, 8], 256, hidden, 3)
dev = torch.device("cuda", 0)
backs, streams = [], []
for r in range(world):
    e = Engine([1, 8], hidden, P); e.set_weights(W)
    st = torch.cuda.Stream(device=dev); e.set_stream(C.c_void_p(st.cuda_stream))
    backs.append(EngineSlabBackend(e)); streams.append(st)
nbytes = backs[0].p2p_bytes(world, n)
bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
torch.cuda.synchronize()
for r, b in enumerate(backs): b.p2p_setup(world, r, n, [t.data_ptr() for t in bufs])
zt = torch.tensor(Z, dtype=torch.int32, device=dev); xt = torch.tensor(X, dtype=torch.float64, device=dev)
es = [torch.zeros(6, dtype=torch.float64, device=dev) for _ in range(world)]
gs = [torch.zeros(n, 3, dtype=torch.float64, device=dev) for _ in range(world)]
torch.cuda.synchronize()
def flags(tag):
    torch.cuda.synchronize()
    for r in range(world):
        f = bufs[r][nbytes - 1024:].view(torch.int32).cpu().numpy()
        print(tag, "rank", r, "counters", f[0], f[16], f[32], "epochs", f[64], f[80], f[96], flush=True)
for r, b in enumerate(backs): b.slab_phase_a(xt, zt, n, lat, 1, r, world, es[r])
flags("after A")
for r, b in enumerate(backs): b.slab_phase_b(es[r], es[r])
flags("after B")
for r, b in enumerate(backs): b.slab_phase_c(es[r], TM_F_FORCE | TM_F_VDW, gs[r])
flags("after C")
print([e.cpu().numpy()[:4] for e in es])
print("---- async step", flush=True)
import time
t0 = time.time()
for r, b in enumerate(backs): b.slab_phase_a(xt, zt, n, lat, 1, r, world, es[r])
print("A enqueued", round(time.time() - t0, 3), flush=True)
for r, b in enumerate(backs):
    b.slab_phase_b(es[r], es[r]); print("B enqueued rank", r, round(time.time() - t0, 3), flush=True)
for r, b in enumerate(backs):
    b.slab_phase_c(es[r], TM_F_FORCE | TM_F_VDW, gs[r]); print("C enqueued rank", r, round(time.time() - t0, 3), flush=True)
flags("after async")
print([e.cpu().numpy()[:4] for e in es])
