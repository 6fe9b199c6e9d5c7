"""Geometry optimisation with conjugate gradients on an energy/force callback (reference: Simulations/Opt.py:17-75)."""
from __future__ import annotations

import numpy as np

from ..Containers.Mol import Mol
from ..Math.QuasiNewtonTools import ConjGradient, RemoveInvariantForce
from ..Util import *   # noqa: F401,F403


class GeomOptimizer:
    def __init__(self, f_):
        """f_(x, DoForce=True) -> (E [Hartree], F [J/mol/A])  or  E when DoForce is False."""
        self.thresh = PARAMS["OptThresh"]
        self.maxstep = PARAMS["OptMaxStep"]
        self.fscale = PARAMS["OptStepSize"]
        self.momentum = PARAMS["OptMomentum"]
        self.momentum_decay = PARAMS["OptMomentumDecay"]
        self.max_opt_step = PARAMS["OptMaxCycles"]
        self.step = self.maxstep
        self.EnergyAndForce = f_
        self.m = None

    def WrappedEForce(self, x_, DoForce=True):
        if DoForce:
            energy, frc = self.EnergyAndForce(x_, DoForce)
            frc = RemoveInvariantForce(x_, frc, self.m.atoms)     # atomic numbers as weights, like the reference (Opt.py:39)
            return energy, frc / JOULEPERHARTREE   # noqa: F405
        out = self.EnergyAndForce(x_, False)
        return out[0] if isinstance(out, tuple) else out

    def Opt(self, m_, filename="OptLog", Debug=False):
        m = Mol(m_.atoms, m_.coords)
        self.m = m
        rmsdisp, rmsgrad, step = 10.0, 10.0, 0
        prev_m = Mol(m.atoms, m.coords)
        CG = ConjGradient(self.WrappedEForce, m.coords)
        while step < self.max_opt_step and rmsgrad > self.thresh and (rmsdisp > 0.000001 or step < 5):
            prev_m = Mol(m.atoms, m.coords)
            m.coords, energy, frc = CG(m.coords)
            rmsgrad = np.sum(np.linalg.norm(frc, axis=1)) / m.coords.shape[0]
            rmsdisp = np.sum(np.linalg.norm(m.coords - prev_m.coords, axis=1)) / m.coords.shape[0]
            LOGGER.info(filename + "step: %i energy: %0.5f rmsgrad: %0.5f rmsdisp: %0.5f ", step, energy, rmsgrad, rmsdisp)
            prev_m.properties["Step"] = step
            prev_m.properties["Energy"] = energy
            prev_m.WriteXYZfile(PARAMS["results_dir"], filename, 'a', True)
            step += 1
        return prev_m
