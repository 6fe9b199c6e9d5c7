"""Periodic molecular dynamics on top of a PeriodicForce (reference: Simulations/PeriodicMD.py:21-146)."""
from __future__ import annotations

import numpy as np

from ..Containers.Mol import Mol
from ..Util import *   # noqa: F401,F403
from .SimpleMD import KineticEnergy, NoseThermostat, VelocityVerlet

_ACC = pow(10.0, -10.0)


def PeriodicVelocityVerletStep(pf_, a_, x_, v_, m_, dt_):
    """Velocity Verlet with the positions wrapped into the cell before the force call."""
    x = pf_.lattice.ModuloLattice(x_ + v_ * dt_ + 0.5 * a_ * dt_ * dt_)
    e, f_x_ = pf_(x)
    a = _ACC * np.einsum("ax,a->ax", f_x_, 1.0 / m_)
    v = v_ + 0.5 * (a_ + a) * dt_
    return x, v, a, e


class PeriodicNoseThermostat(NoseThermostat):
    def step(self, pf_, a_, x_, v_, m_, dt_):
        self.kT = IDEALGASR * _ACC * self.T   # noqa: F405
        self.tau = 20.0 * PARAMS["MDdt"] * self.N
        self.Q = self.kT * self.tau * self.tau
        x = pf_.lattice.ModuloLattice(x_ + v_ * dt_ + 0.5 * (a_ - self.eta * v_) * dt_ * dt_)
        vdto2 = v_ + 0.5 * (a_ - self.eta * v_) * dt_
        e, f_x_ = pf_(x)
        a = _ACC * np.einsum("ax,a->ax", f_x_, 1.0 / m_)
        target = ((3. * self.N + 1) / 2.) * self.kT
        ke = 0.5 * np.dot(np.einsum("ia,ia->i", v_, v_), m_)
        etadto2 = self.eta + (dt_ / (2. * self.Q)) * (ke - target)
        kedto2 = 0.5 * np.dot(np.einsum("ia,ia->i", vdto2, vdto2), m_)
        self.eta = etadto2 + (dt_ / (2. * self.Q)) * (kedto2 - target)
        v = (vdto2 + (dt_ / 2.) * a) / (1 + (dt_ / 2.) * self.eta)
        return x, v, a, e


class PeriodicVelocityVerlet(VelocityVerlet):
    def __init__(self, Force_, name_="PdicMD", v0_=None):
        """Force_: a PeriodicForce (energy per cell in Hartree, force J/mol/A on the primitive atoms)."""
        self.PForce = Force_
        VelocityVerlet.__init__(self, None, self.PForce.mol0, name_, self.PForce.__call__)
        if v0_ is not None:
            self.v = v0_
        self.Tstat = PeriodicNoseThermostat(self.m, self.v) if PARAMS["MDThermostat"] == "Nose" else None

    def Density(self):
        return self.PForce.Density()

    def WriteTrajectory(self):
        m = Mol(self.atoms, self.x)
        m.properties["Lattice"] = self.PForce.lattice.lattice.copy()
        m.properties["Time"] = self.t
        m.properties["KineticEnergy"] = self.KE
        m.properties["PotEnergy"] = self.EPot
        m.WriteXYZfile(PARAMS["results_dir"], "MDTrajectory" + self.name, 'a', True)

    def Prop(self):
        step = 0
        self.md_log = np.zeros((self.maxstep, 7))
        while step < self.maxstep:
            self.t = step * self.dt
            self.KE = KineticEnergy(self.v, self.m)
            Teff = (2. / 3.) * self.KE / IDEALGASR   # noqa: F405
            if self.Tstat is None:
                self.x, self.v, self.a, self.EPot = PeriodicVelocityVerletStep(self.PForce, self.a, self.x, self.v, self.m, self.dt)
            else:
                self.x, self.v, self.a, self.EPot = self.Tstat.step(self.PForce, self.a, self.x, self.v, self.m, self.dt)
            self.md_log[step, 0] = self.t
            self.md_log[step, 4] = self.KE
            self.md_log[step, 5] = self.EPot
            self.md_log[step, 6] = self.KE + (self.EPot - self.EPot0) * JOULEPERHARTREE   # noqa: F405
            if PARAMS["PrintTMTimer"]:
                PrintTMTIMER()   # noqa: F405
            if PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            if step % 500 == 0:
                self._save_log()
            step += 1
            LOGGER.info("Step: %i time: %.1f(fs) <KE>(kJ/mol): %.5f <|a|>(m/s2): %.5f <EPot>(Eh): %.5f <Etot>(kJ/mol): %.5f Rho(g/cm**3): %.5f Teff(K): %.5f",
                        step, self.t, self.KE / 1000.0, np.linalg.norm(self.a), self.EPot, self.KE / 1000.0 + self.EPot * KJPERHARTREE, self.Density(), Teff)   # noqa: F405
