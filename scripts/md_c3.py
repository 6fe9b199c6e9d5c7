"""Config C3 (BASELINE.json): 3,000-atom periodic water box, NVE velocity Verlet with a neighbour-list rebuild every
step, on-device integrator (DevicePeriodicVelocityVerlet) vs the host driver.  Prints steps/s and the energy drift."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_c_gpu_api import _manager
from tensormol_b200 import PARAMS, Mol, PeriodicForce, PeriodicVelocityVerlet
from tensormol_b200.PhysicalData import JOULEPERHARTREE
from tensormol_b200.Simulations.DeviceMD import DevicePeriodicVelocityVerlet
from tensormol_b200.SystemBuilders import water_box
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Z, X, lat = water_box(10, spacing=3.1072, seed=3, jitter=0.02)
m = Mol(Z.astype(np.uint8), X)
manager, W = _manager([m], [500, 500, 500], 0)
PARAMS["MDMaxStep"] = nsteps; PARAMS["MDdt"] = 0.2; PARAMS["MDV0"] = None; PARAMS["MDThermostat"] = None; PARAMS["MDTemp"] = 300.0
v0 = 1e-3 * np.random.RandomState(1).randn(len(Z), 3)
dev = DevicePeriodicVelocityVerlet(manager, m, lat, "c3", v0_=v0, sync_every_=250)
t0 = time.perf_counter(); log = dev.Prop(); t1 = time.perf_counter()
etot = log[:, 4] * len(Z) + log[:, 5] * JOULEPERHARTREE
print(f"device MD: {len(Z)} atoms, {nsteps} steps in {t1 - t0:.3f} s = {nsteps / (t1 - t0):.1f} steps/s = {len(Z) * nsteps / (t1 - t0) / 1e6:.2f} M atom-steps/s;"
      f" Etot drift {np.ptp(etot[5:]):.3e} J/mol over PE range {np.ptp(log[:, 5]) * JOULEPERHARTREE:.3e}")
nh = min(nsteps, 100)
PARAMS["MDMaxStep"] = nh
pf = PeriodicForce(m, lat); pf.BindLatticeForce(manager.LatticeForce(), 15.0)
host = PeriodicVelocityVerlet(pf, "c3h", v0_=v0.copy())
import logging; logging.disable(logging.CRITICAL)
t0 = time.perf_counter(); host.Prop(); t1 = time.perf_counter()
print(f"host driver: {nh} steps in {t1 - t0:.3f} s = {nh / (t1 - t0):.1f} steps/s")
