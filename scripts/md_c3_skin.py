"""Config C3 on the device MD driver: steps/s with a neighbour rebuild every step vs a Verlet skin (rebuild every k steps)."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_c_gpu_api import _manager
from tensormol_b200 import PARAMS, Mol
from tensormol_b200.PhysicalData import JOULEPERHARTREE
from tensormol_b200.Simulations.DeviceMD import DevicePeriodicVelocityVerlet
from tensormol_b200.SystemBuilders import water_box
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Z, X, lat = water_box(10, spacing=3.1072, seed=3, jitter=0.02)
m = Mol(Z.astype(np.uint8), X)
manager, W = _manager([m], [500, 500, 500], 0)
PARAMS["MDMaxStep"] = nsteps; PARAMS["MDdt"] = 0.2; PARAMS["MDV0"] = None; PARAMS["MDThermostat"] = None; PARAMS["MDTemp"] = 300.0
v0 = 1e-3 * np.random.RandomState(1).randn(len(Z), 3)
for skin, every in ((0.0, 1), (0.5, 5), (0.8, 10)):
    dev = DevicePeriodicVelocityVerlet(manager, m, lat, "c3", v0_=v0.copy(), sync_every_=250, skin_=skin, nl_every_=every)
    t0 = time.perf_counter(); log = dev.Prop(); t1 = time.perf_counter()
    etot = log[:, 4] * len(Z) + log[:, 5] * JOULEPERHARTREE
    print(f"skin {skin} A, rebuild every {every}: {len(Z)} atoms, {nsteps} steps in {t1 - t0:.3f} s = {nsteps / (t1 - t0):.1f} steps/s = "
          f"{len(Z) * nsteps / (t1 - t0) / 1e6:.2f} M atom-steps/s; final EPot {log[nsteps - 1, 5]:.9f}; Etot drift {np.ptp(etot[5:]):.3e} J/mol")
