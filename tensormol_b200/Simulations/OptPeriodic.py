"""Periodic geometry optimisation on a PeriodicForce (reference: Simulations/OptPeriodic.py:8-135)."""
from __future__ import annotations

import numpy as np

from ..Containers.Mol import Mol
from ..Math.QuasiNewtonTools import ConjGradient, RemoveInvariantForce
from ..Util import *   # noqa: F401,F403
from .Opt import GeomOptimizer


class PeriodicGeomOptimizer(GeomOptimizer):
    def __init__(self, f_):
        """f_: a PeriodicForce object."""
        GeomOptimizer.__init__(self, f_)

    def _wrapped(self, m):
        def WrappedEForce(x_, DoForce=True):
            out = self.EnergyAndForce(x_, DoForce)
            if DoForce:
                energy, frc = out
                return energy, RemoveInvariantForce(x_, frc, m.atoms) / JOULEPERHARTREE   # noqa: F405
            return out[0] if isinstance(out, tuple) else out
        return WrappedEForce

    def _relax(self, m, CG, filename, min_steps):
        rmsdisp, rmsgrad, step = 10.0, 10.0, 0
        prev_m = Mol(m.atoms, m.coords)
        while step < self.max_opt_step and rmsgrad > self.thresh and (rmsdisp > 0.0001 or step < min_steps):
            prev_m = Mol(m.atoms, m.coords)
            m.coords, energy, frc = CG(m.coords)
            rmsgrad = np.sum(np.linalg.norm(frc, axis=1)) / m.coords.shape[0]
            rmsdisp = np.sum(np.linalg.norm(m.coords - prev_m.coords, axis=1)) / m.coords.shape[0]
            m.coords = self.EnergyAndForce.lattice.ModuloLattice(m.coords)
            LOGGER.info("step: %i energy: %.6f density: %.4f rmsgrad %.6f rmsdisp %.6f", step, energy, self.EnergyAndForce.Density(), rmsgrad, rmsdisp)
            prev_m.properties['Lattice'] = self.EnergyAndForce.lattice.lattice.copy()
            prev_m.WriteXYZfile(PARAMS["results_dir"], filename, 'a', True)
            step += 1
        return prev_m

    def Opt(self, m, filename="PdicOptLog", Debug=False):
        PARAMS["OptLatticeStep"] = 0.050
        CG = ConjGradient(self._wrapped(m), m.coords)
        prev_m = self._relax(m, CG, filename, 3)
        self.EnergyAndForce.Save(prev_m.coords, "FinalPeriodicOpt")
        self.EnergyAndForce.mol0.coords = prev_m.coords.copy()
        return prev_m

    def OptToDensity(self, m, rho_target=1.0, filename="PdicOptLog", Debug=False):
        """Squeeze the lattice gently until the target density (g/cm**3) is reached, relaxing at every step."""
        CG = ConjGradient(self._wrapped(m), m.coords)
        prev_m = Mol(m.atoms, m.coords)
        Density = self.EnergyAndForce.Density()
        while abs(Density - rho_target) > 0.001:
            fac = rho_target / Density
            oldlat = self.EnergyAndForce.lattice.lattice.copy()
            newlat = 0.65 * oldlat + 0.35 * (oldlat * pow(1.0 / fac, 1.0 / 3.))
            m.coords = self.EnergyAndForce.AdjustLattice(m.coords, oldlat, newlat)
            self.EnergyAndForce.ReLattice(newlat)
            prev_m = self._relax(m, CG, filename, 0)
            Density = self.EnergyAndForce.Density()
        self.EnergyAndForce.Save(prev_m.coords, "FinalPeriodicOpt")
        self.EnergyAndForce.mol0.coords = prev_m.coords.copy()
        return prev_m
