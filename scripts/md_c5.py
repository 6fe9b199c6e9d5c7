"""Config C5 (BASELINE.json): the 2evq peptide in its explicit water (1,568 atoms, C/H/N/O, D = 768, bounding-box cell of
samples/test_neb.py:91-113; geometry from tests/golden/evq2_periodic.npz), nets 2000^3 random-init, periodic NVT
Nose-Hoover MD at 300 K, dt 0.2 fs, neighbour list rebuilt every step, on the device driver.  Prints steps/s, the stage
times of one evaluation and the temperature reached.  Multi-GPU: this system has 196 atoms per rank at 8 ranks -- far
below the per-step fixed cost, so it is run on one GPU."""
import os, sys, time
os.environ.setdefault("TM_NO_GRAPH", "1")   # stage timings need the kernel-by-kernel host path; the MD driver captures its own graph
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden
from test_c_gpu_api import _manager
from tensormol_b200 import PARAMS, Mol
from tensormol_b200.PhysicalData import IDEALGASR
from tensormol_b200.Simulations.DeviceMD import DevicePeriodicVelocityVerlet
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
hidden = [int(sys.argv[2])] * 3 if len(sys.argv) > 2 else [2000, 2000, 2000]
g = load_golden("evq2_periodic")
m = Mol(g["Z"].astype(np.uint8), g["xyz"])
manager, W = _manager([m], hidden, 6)
# stage times of one evaluation, kernel by kernel
eng = manager.Instances.engine
for _ in range(4):
    r = eng.evaluate_lattice(g["xyz"], g["Z"], g["lattice"], 1)
print("stages (ms):", {k: round(v, 3) for k, v in eng.timings().items() if isinstance(v, float)}, "E", r["Etotal"][0])
PARAMS["MDMaxStep"] = nsteps; PARAMS["MDdt"] = 0.2; PARAMS["MDV0"] = "Random"; PARAMS["MDThermostat"] = "Nose"; PARAMS["MDTemp"] = 300.0
np.random.seed(0)
dev = DevicePeriodicVelocityVerlet(manager, m, g["lattice"], "c5", sync_every_=250)
t0 = time.perf_counter(); log = dev.Prop(); t1 = time.perf_counter()
teff = (2. / 3.) * log[:, 4] / IDEALGASR
print(f"device NVT MD: {len(g['Z'])} atoms, nets {hidden}, {nsteps} steps in {t1 - t0:.3f} s = {nsteps / (t1 - t0):.1f} steps/s = "
      f"{len(g['Z']) * nsteps / (t1 - t0) / 1e6:.2f} M atom-steps/s; Teff first/last {teff[0]:.1f}/{teff[-1]:.1f} K; finite {bool(np.all(np.isfinite(log)))}")
