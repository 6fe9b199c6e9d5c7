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


class PeriodicBoxingDynamics(PeriodicVelocityVerlet):
    """MD in a cell that is deformed linearly in time from the force's lattice to `BoxingLatp_` over `BoxingT_` fs
    (reference: Simulations/PeriodicMD.py:147-215): positions, velocities and accelerations follow the cell in
    fractional coordinates (PeriodicForce.AdjustLattice), then one ordinary (thermostatted) step is taken."""

    def __init__(self, Force_, BoxingLatp_=np.eye(3), name_="PdicBoxMD", BoxingT_=400):
        PeriodicVelocityVerlet.__init__(self, Force_, name_)
        self.BoxingLat0 = Force_.lattice.lattice.copy()
        self.BoxingLatp = np.array(BoxingLatp_, np.float64)
        self.BoxingT = BoxingT_

    def _deform(self):
        """Move the cell to its shape at time self.t (no-op once BoxingT has passed)."""
        if self.t > self.BoxingT:
            return False
        w = (self.BoxingT - self.t) / self.BoxingT
        newlattice = w * self.BoxingLat0 + (1.0 - w) * self.BoxingLatp
        lat0 = self.PForce.lattice.lattice
        self.x = self.PForce.AdjustLattice(self.x, lat0, newlattice)
        self.v = self.PForce.AdjustLattice(self.v, lat0, newlattice)
        self.a = self.PForce.AdjustLattice(self.a, lat0, newlattice)
        self.PForce.ReLattice(newlattice)
        return True

    def Prop(self):
        step = 0
        self.md_log = np.zeros((self.maxstep, 7))
        while step < self.maxstep:
            self.t = step * self.dt
            self.KE = KineticEnergy(self.v, self.m)
            if not self._deform():
                LOGGER.info("Exceeded Boxtime %s", self.BoxingLatp)   # noqa: F405
            Teff = (2. / 3.) * self.KE / IDEALGASR   # noqa: F405
            if self.Tstat is None:
                self.x, self.v, self.a, self.EPot = PeriodicVelocityVerletStep(self.PForce, self.a, self.x, self.v, self.m, self.dt)
            else:
                self.x, self.v, self.a, self.EPot = self.Tstat.step(self.PForce, self.a, self.x, self.v, self.m, self.dt)
            self.md_log[step, 0] = self.t
            self.md_log[step, 4] = self.KE
            self.md_log[step, 5] = self.EPot
            self.md_log[step, 6] = self.KE + (self.EPot - self.EPot0) * JOULEPERHARTREE   # noqa: F405
            if step % 3 == 0 and PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            if step % 500 == 0:
                self._save_log()
            step += 1
            LOGGER.info("Step: %i time: %.1f(fs) <KE>(kJ/mol): %.5f <EPot>(Eh): %.5f Rho(g/cm**3): %.5f Teff(K): %.5f",
                        step, self.t, self.KE / 1000.0, self.EPot, self.Density(), Teff)   # noqa: F405


class PeriodicAnnealer(PeriodicVelocityVerlet):
    """Nose-thermostatted anneal of a periodic system from MDAnnealT0 to MDAnnealTF over MDAnnealSteps steps of 0.1 fs,
    starting at rest; whenever the potential energy drops below the best so far by more than `AnnealThresh_` the
    geometry is kept in .Minx and the schedule restarts from the current target temperature + MDAnnealKickBack
    (reference: Simulations/PeriodicMD.py:218-285)."""

    def __init__(self, Force_, name_="PdicAnneal", AnnealThresh_=0.000009):
        PeriodicVelocityVerlet.__init__(self, Force_, name_)
        self.dt = 0.1
        self.v = self.v * 0.0
        self.AnnealT0 = PARAMS["MDAnnealT0"]
        self.AnnealSteps = PARAMS["MDAnnealSteps"]
        self.MinS = 0
        self.MinE = 99999999.0
        self.Minx = None
        self.AnnealThresh = AnnealThresh_
        self.Tstat = PeriodicNoseThermostat(self.m, self.v)

    def Prop(self):
        step = 0
        self.md_log = np.zeros((max(self.maxstep, self.AnnealSteps), 7))
        while step < self.AnnealSteps:
            self.t = step * self.dt
            self.KE = KineticEnergy(self.v, self.m)
            Teff = (2. / 3.) * self.KE / IDEALGASR   # noqa: F405
            frac = float(self.AnnealSteps - step) / self.AnnealSteps
            self.Tstat.T = self.AnnealT0 * frac + PARAMS["MDAnnealTF"] * (1.0 - frac) + pow(10.0, -10.0)
            self.x, self.v, self.a, self.EPot = self.Tstat.step(self.PForce, self.a, self.x, self.v, self.m, self.dt)
            if self.EPot < self.MinE and abs(self.EPot - self.MinE) > self.AnnealThresh and step > 1:
                self.MinE = self.EPot
                self.Minx = self.x.copy()
                self.MinS = step
                LOGGER.info("   -- cycling annealer -- ")   # noqa: F405
                if PARAMS["MDAnnealT0"] > PARAMS["MDAnnealTF"]:
                    self.AnnealT0 = self.Tstat.T + PARAMS["MDAnnealKickBack"]
                step = 0
            self.md_log[step, 0] = self.t
            self.md_log[step, 4] = self.KE
            self.md_log[step, 5] = self.EPot
            self.md_log[step, 6] = self.KE + (self.EPot - self.EPot0) * JOULEPERHARTREE   # noqa: F405
            if step % 3 == 0 and PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            if step % 500 == 0:
                self._save_log()
            step += 1
            LOGGER.info("Step: %i time: %.1f(fs) <KE>(kJ/mol): %.5f <EPot>(Eh): %.5f Rho(g/cm**3): %.5f Teff(K): %.5f T_target(K): %.5f",
                        step, self.t, self.KE / 1000.0, self.EPot, self.Density(), Teff, self.Tstat.T)   # noqa: F405
        if self.Minx is not None:
            m = Mol(self.atoms, self.Minx)
            m.properties["Lattice"] = self.PForce.lattice.lattice.copy()
            m.WriteXYZfile(PARAMS["results_dir"], "PAnnealMin", wprop=True)
