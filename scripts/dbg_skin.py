import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import hot_params
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
skin = float(sys.argv[1]) if len(sys.argv) > 1 else 0.8
Z, X, lat = water_box(10, spacing=3.1072, seed=3, jitter=0.02)
X = wrap_into_cell(X, lat)
eng = Engine([1, 8], [64, 64], hot_params(), device=0)
eng.set_weights(random_weights([1, 8], eng.D, [64, 64], 0))
eng.set_skin(skin)
dev = torch.device("cuda", 0)
xd = torch.tensor(X, dtype=torch.float64, device=dev); zd = torch.tensor(Z, dtype=torch.int32, device=dev)
e = torch.zeros(6, dtype=torch.float64, device=dev); g = torch.zeros(len(Z), 3, dtype=torch.float64, device=dev)
eng.evaluate_lattice_dev(C.c_void_p(xd.data_ptr()), C.c_void_p(zd.data_ptr()), len(Z), lat, 1, C.c_void_p(e.data_ptr()), C.c_void_p(g.data_ptr()))
eng.sync()
print("skin", skin, "E", e.cpu().numpy()[:4])
