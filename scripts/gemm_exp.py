"""Measurement aid for tm_gemm_tc.cu (-DTC_EXP build): MLP stage times of the 24k step with parts of the kernel switched off.
    TM_EXTRA_NVCC_FLAGS=-DTC_EXP python -m tensormol_b200.csrc.build --force; for e in 0 1 2 4 3 5 6 7; do TC_EXP=$e python scripts/gemm_exp.py; done"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import hot_params
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
P = hot_params()
Z, X, lat = water_box(20, spacing=3.1072, seed=3, jitter=0.05)
X = wrap_into_cell(X, lat)
eng = Engine([1, 8], [500, 500, 500], P, device=0)
eng.set_weights(random_weights([1, 8], eng.D, [500, 500, 500], 0))
eng.set_gemm_mode(int(os.environ.get("GEMM_MODE", "1")))
dev = torch.device("cuda", 0)
xyz_t = torch.tensor(X, dtype=torch.float64, device=dev); Z_t = torch.tensor(Z, dtype=torch.int32, device=dev)
e_t = torch.zeros(6, dtype=torch.float64, device=dev); g_t = torch.zeros(len(Z), 3, dtype=torch.float64, device=dev)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
acc = {}
n = 8
for it in range(3 + n):
    flush.zero_()
    eng.evaluate_lattice_dev(C.c_void_p(xyz_t.data_ptr()), C.c_void_p(Z_t.data_ptr()), len(Z), lat, 1, C.c_void_p(e_t.data_ptr()), C.c_void_p(g_t.data_ptr()))
    t = eng.timings()
    if it >= 3:
        for k in ("mlp_fwd", "mlp_bwd", "total"):
            acc[k] = acc.get(k, 0.0) + t[k] / n
print("TC_EXP", os.environ.get("TC_EXP", "0"), "mode", os.environ.get("GEMM_MODE", "1"), {k: round(v, 4) for k, v in acc.items()}, "E", float(e_t[0]))
