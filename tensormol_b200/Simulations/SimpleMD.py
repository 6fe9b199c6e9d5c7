"""Molecular dynamics drivers with the reference's classes and unit conventions (Simulations/SimpleMD.py):
Angstrom, fs, masses in kg/mol; the force callback returns J/mol/Angstrom and the energy Hartree.
They consume  f(x) -> F  or  EandF(x) -> (E, F)  and are independent of how those are computed."""
from __future__ import annotations

import os

import numpy as np

from ..Containers.Mol import Mol
from ..Math.QuasiNewtonTools import RemoveInvariantForce
from ..Math.Statistics import OnlineEstimator
from ..Util import *   # noqa: F401,F403

_ACC = pow(10.0, -10.0)     # (J/mol/A)/(kg/mol) = m^2/s^2 per A  ->  A/fs^2


def _force_and_energy(f_, fande_, x):
    if fande_ is None:
        return 0.0, f_(x)
    return fande_(x)


def VelocityVerletStep(f_, a_, x_, v_, m_, dt_, fande_=None):
    """x(t+dt), v(t+dt), a(t+dt), E (reference :14-38)."""
    x = x_ + v_ * dt_ + 0.5 * a_ * dt_ * dt_
    e, f_x_ = _force_and_energy(f_, fande_, x)
    a = _ACC * np.einsum("ax,a->ax", f_x_, 1.0 / m_)
    v = v_ + 0.5 * (a_ + a) * dt_
    return x, v, a, e


def KineticEnergy(v_, m_):
    """Kinetic energy per atom in J/mol (v in A/fs, m in kg/mol)."""
    return 0.5 * np.dot(np.einsum("ia,ia->i", v_, v_) * pow(10.0, 10.0), m_) / len(m_)


def Dipole_Naive(x_, q_):
    """sum_a q_a x_a in e Bohr for positions in Angstrom (reference ForceModels/Electrostatics.py:28-33)."""
    return np.einsum("ax,a->x", x_, np.asarray(q_)) * BOHRPERA   # noqa: F405


def ElectricFieldForce(q_, E_):
    return np.einsum("a,x->ax", np.asarray(q_), np.asarray(E_))


class Thermostat:
    """Velocity rescaling to PARAMS["MDTemp"] after every step."""

    def __init__(self, m_, v_):
        self.N = len(m_)
        self.m = m_.copy()
        self.T = PARAMS["MDTemp"]
        self.Teff = 0.001
        self.kT = IDEALGASR * _ACC * self.T   # noqa: F405
        self.tau = 30 * PARAMS["MDdt"]
        self.name = "Rescaling"
        self.Rescale(v_)

    def step(self, f_, a_, x_, v_, m_, dt_, fande_=None, frc_=True):
        x, v, a, e = VelocityVerletStep(f_, a_, x_, v_, m_, dt_, fande_)
        self.Teff = (2. / 3.) * KineticEnergy(v, self.m) / IDEALGASR   # noqa: F405
        v = v * np.sqrt(self.T / self.Teff)
        if frc_:
            return x, v, a, e, a * m_[:, None] / _ACC
        return x, v, a, e

    def Rescale(self, v_):
        """Per-atom rescale to the target temperature (in place), so light atoms do not fly off."""
        ke = (2.0 / (3.0 * IDEALGASR)) * pow(10.0, 10.0) * 0.5 * self.m * np.einsum("ai,ai->a", v_, v_)   # noqa: F405
        nz = ke != 0.0
        v_[nz] *= np.sqrt(self.T / ke[nz])[:, None]


class NoseThermostat(Thermostat):
    """Single Nose-Hoover thermostat (http://www2.ph.ed.ac.uk/~dmarendu/MVP/MVP03.pdf; reference :90-129)."""

    def __init__(self, m_, v_):
        self.m = m_.copy()
        self.N = len(m_)
        self.T = PARAMS["MDTemp"]
        self.eta = 0.0
        self.name = "Nose"
        self.Rescale(v_)

    def step(self, f_, a_, x_, v_, m_, dt_, fande_=None, frc_=True):
        self.kT = IDEALGASR * _ACC * self.T   # noqa: F405
        self.tau = 20.0 * PARAMS["MDdt"] * self.N
        self.Q = self.kT * self.tau * self.tau
        x = x_ + v_ * dt_ + 0.5 * (a_ - self.eta * v_) * dt_ * dt_
        vdto2 = v_ + 0.5 * (a_ - self.eta * v_) * dt_
        e, f_x_ = _force_and_energy(f_, fande_, x)
        a = _ACC * np.einsum("ax,a->ax", f_x_, 1.0 / m_)
        target = ((3. * self.N + 1) / 2.) * self.kT
        ke = 0.5 * np.dot(np.einsum("ia,ia->i", v_, v_), m_)
        etadto2 = self.eta + (dt_ / (2. * self.Q)) * (ke - target)
        kedto2 = 0.5 * np.dot(np.einsum("ia,ia->i", vdto2, vdto2), m_)
        self.eta = etadto2 + (dt_ / (2. * self.Q)) * (kedto2 - target)
        v = (vdto2 + (dt_ / 2.) * a) / (1 + (dt_ / 2.) * self.eta)
        if frc_:
            return x, v, a, e, f_x_
        return x, v, a, e


class AndersenThermostat(Thermostat):
    def __init__(self, m_, v_):
        self.m = m_.copy()
        self.N = len(list(m_))
        self.T = PARAMS["MDTemp"]
        self.gamma = 1 / 2.0     # collision frequency (1/fs)
        self.name = "Andersen"
        self.Rescale(v_)

    def step(self, f_, a_, x_, v_, m_, dt_, fande_=None, frc_=True):
        x = x_ + v_ * dt_ + 0.5 * a_ * dt_ * dt_
        e, f_x_ = _force_and_energy(f_, fande_, x)
        a = _ACC * np.einsum("ax,a->ax", f_x_, 1.0 / m_)
        v = v_ + 0.5 * (a_ + a) * dt_
        self.kT = IDEALGASR * _ACC * self.T   # noqa: F405
        s = np.sqrt(2.0 * self.gamma * self.kT / self.m)
        # collisions drawn atom by atom from numpy's global generator, in the reference's order (one uniform per atom,
        # three normals right after a hit), so a seeded run reproduces the reference's trajectory (SimpleMD.py:157-159)
        for i in range(x_.shape[0]):
            if np.random.random() < self.gamma * dt_:
                v[i] = np.random.normal(0.0, s[i], size=(3))
        if frc_:
            return x, v, a, e, f_x_
        return x, v, a, e


class LangevinThermostat(Thermostat):
    """arXiv:1212.1244v4 (flagged 'not working' in the reference, kept for interface completeness)."""

    def __init__(self, m_, v_):
        self.m = m_.copy()
        self.N = len(m_)
        self.T = PARAMS["MDTemp"]
        self.gamma = 0.05
        self.name = "Langevin"
        self.Rescale(v_)

    def step(self, f_, a_, x_, v_, m_, dt_, fande_=None, frc_=True):
        self.kT = IDEALGASR * _ACC * self.T   # noqa: F405
        s = np.sqrt(2.0 * self.gamma * self.kT / dt_)
        beta = np.random.normal(0.0, s, size=x_.shape)
        m = np.tile(self.m[:, np.newaxis], (1, 3))
        ca = (1.0 - self.gamma * dt_ / (2.0 * m)) / (1.0 + self.gamma * dt_ / (2.0 * m))
        cb = 1.0 / (1.0 + self.gamma * dt_ / (2.0 * m))
        x = x_ + cb * dt_ * v_ + (cb * dt_ * dt_) / (2.0 * m) * (a_ * m) + (cb * dt_) / (2.0 * m) * beta
        e, f_x_ = _force_and_energy(f_, fande_, x)
        v = ca * v_ + dt_ / (2.0 * m) * (ca * (a_ * m) + f_x_) + (cb / m) * beta
        a = f_x_ / m
        if frc_:
            return x, v, a, e, f_x_
        return x, v, a, e


class NoseChainThermostat(Thermostat):
    """Nose-Hoover chain, Martyna et al. 1996 appendix A (doi:10.1080/00268979600100761)."""

    def __init__(self, m_, v_):
        self.M = PARAMS.get("MNHChain", 3)
        self.N = len(v_)
        self.Nf = len(v_) * 3
        self.T = PARAMS["MDTemp"]
        self.kT = IDEALGASR * _ACC * self.T   # noqa: F405
        self.GNKT = self.Nf * self.kT
        self.nc = 2
        w = 1. / (2. - np.power(2., 1. / 3.))
        self.wj = np.array([w, 1. - 2. * w, w])
        self.tau = 80.0 * PARAMS["MDdt"]
        self.dt = PARAMS["MDdt"]
        self.Qs = np.ones(max(self.M, 1)) * self.kT * self.tau * self.tau
        self.Qs[0] = 3. * self.N * self.kT * self.tau * self.tau
        self.eta = np.zeros(self.M)
        self.Veta = np.zeros(self.M)
        self.Geta = np.zeros(self.M)
        self.m = m_.copy()
        self.name = "NoseHooverChain"
        self.Rescale(v_)

    def ke(self, v_, m_):
        return 0.5 * np.dot(np.einsum("ia,ia->i", v_, v_), m_)

    def step(self, f_, a_, x_, v_, m_, dt_, fande_=None, frc_=True):
        v = self.IntegrateChain(v_, m_)
        v = v + 0.5 * self.dt * a_
        x = x_ + self.dt * v
        e, f_x_ = _force_and_energy(f_, fande_, x)
        a = _ACC * np.einsum("ax,a->ax", f_x_, 1.0 / m_)
        v = self.IntegrateChain(v + 0.5 * self.dt * a, m_)
        if frc_:
            return x, v, a, e, f_x_
        return x, v, a, e

    def IntegrateChain(self, v_, m_):
        """Half-step of the (twice Trotterised) chain; returns the rescaled velocities."""
        if self.M == 0:
            return v_
        M = self.M
        ake = self.ke(v_, m_)
        self.Geta[0] = (2. * ake - self.GNKT) / self.Qs[0]
        scale = 1.0
        for _ in range(self.nc):
            for w in self.wj:
                h2 = (w * self.dt / self.nc) / 2.
                h4, h8 = h2 / 2., h2 / 4.
                self.Veta[-1] += self.Geta[-1] * h4
                for i in range(M - 1)[::-1]:
                    AA = np.exp(-h8 * self.Veta[i + 1])
                    self.Veta[i] = self.Veta[i] * AA * AA + h4 * self.Geta[i] * AA
                scale *= np.exp(-h2 * self.Veta[0])
                self.Geta[0] = (scale * scale * 2.0 * ake - self.GNKT) / self.Qs[0]
                self.eta += self.Veta * h2
                for i in range(M - 1):
                    AA = np.exp(-h8 * self.Veta[i + 1])
                    self.Veta[i] = self.Veta[i] * AA * AA + h4 * self.Geta[i] * AA
                    self.Geta[i + 1] = (self.Qs[i] * self.Veta[i] * self.Veta[i] - self.kT) / self.Qs[i + 1]
                self.Veta[-1] += self.Geta[-1] * h4
        return v_ * scale


_THERMOSTATS = {"Rescaling": Thermostat, "Nose": NoseThermostat, "Andersen": AndersenThermostat, "Langevin": LangevinThermostat,
                "NoseHooverChain": NoseChainThermostat}


class VelocityVerlet:
    def __init__(self, f_, g0_, name_="", EandF_=None, cellsize_=None):
        """f_: force routine (or None when EandF_ is given); g0_: initial molecule; EandF_: energy, force routine.
        PARAMS: MDMaxStep, MDTemp, MDdt, MDV0 (None | "Random" | "Thermal"), MDThermostat, MDLogTrajectory."""
        self.name = name_
        self.cellsize = cellsize_
        self.maxstep = PARAMS["MDMaxStep"]
        self.T = PARAMS["MDTemp"]
        self.dt = PARAMS["MDdt"]
        self.ForceFunction = f_
        self.EnergyAndForce = EandF_
        self.EPot0 = 0.0
        if EandF_ is not None:
            self.EPot0, self.f0 = self.EnergyAndForce(g0_.coords)
        self.EPot = self.EPot0
        self.EnergyStat = OnlineEstimator(self.EPot0)
        self.t = 0.0
        self.KE = 0.0
        self.atoms = g0_.atoms.copy()
        self.m = np.array([ATOMICMASSES[z - 1] for z in self.atoms])   # noqa: F405
        self.natoms = len(self.atoms)
        self.x = g0_.coords.copy()
        self.v = np.zeros(self.x.shape)
        self.a = np.zeros(self.x.shape)
        self.md_log = None
        self.force = None
        if PARAMS["MDV0"] == "Random":
            self.v = np.random.randn(*self.x.shape)
            Thermostat(self.m, self.v)         # rescales self.v in place
        elif PARAMS["MDV0"] == "Thermal":
            self.v = np.random.normal(size=self.x.shape) * np.sqrt(1.38064852e-23 * self.T / self.m)[:, None]
        self.Tstat = None
        if PARAMS["MDThermostat"] in _THERMOSTATS:
            self.Tstat = _THERMOSTATS[PARAMS["MDThermostat"]](self.m, self.v)

    def WriteTrajectory(self):
        m = Mol(self.atoms, self.x)
        m.properties["Time"] = self.t
        m.properties["KineticEnergy"] = self.KE
        m.properties["PotEnergy"] = self.EPot
        m.WriteXYZfile(PARAMS["results_dir"], "MDTrajectory" + self.name)

    def _save_log(self):
        os.makedirs(PARAMS["results_dir"], exist_ok=True)
        np.savetxt(PARAMS["results_dir"] + "MDLog" + self.name + ".txt", self.md_log)

    def Prop(self):
        step = 0
        self.md_log = np.zeros((self.maxstep, 7))
        while step < self.maxstep:
            self.t = step * self.dt
            if self.Tstat is None:
                self.x, self.v, self.a, self.EPot = VelocityVerletStep(self.ForceFunction, self.a, self.x, self.v, self.m, self.dt, self.EnergyAndForce)
            else:
                self.x, self.v, self.a, self.EPot, self.force = self.Tstat.step(self.ForceFunction, self.a, self.x, self.v, self.m, self.dt, self.EnergyAndForce)
            if self.cellsize is not None:
                self.x = np.mod(self.x, self.cellsize)
            self.md_log[step, 0] = self.t
            self.md_log[step, 4] = self.KE
            self.md_log[step, 5] = self.EPot
            self.md_log[step, 6] = self.KE + (self.EPot - self.EPot0) * JOULEPERHARTREE   # noqa: F405
            self.EnergyStat(self.EPot)
            self.KE = KineticEnergy(self.v, self.m)
            Teff = (2. / 3.) * self.KE / IDEALGASR   # noqa: F405
            if step % 3 == 0 and PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            if step % 500 == 0:
                self._save_log()
            step += 1
            LOGGER.info("%s Step: %i time: %.1f(fs) KE(kJ): %.5f PotE(Eh): %.5f ETot(kJ/mol): %.5f Teff(K): %.5f", self.name, step, self.t,
                        self.KE * len(self.m) / 1000.0, self.EPot, self.KE * len(self.m) / 1000.0 + self.EPot * KJPERHARTREE, Teff)   # noqa: F405


class IRTrajectory(VelocityVerlet):
    """Zero-temperature dynamics logging the dipole mu(t) for IR spectra; optional field pulse (reference :427-556)."""

    def __init__(self, f_, q_, g0_, name_=str(0), v0_=None):
        VelocityVerlet.__init__(self, f_, g0_, name_, f_)
        if v0_ is not None:
            self.v = v0_.copy()
        self.EField = np.zeros(3)
        self.IsOn = False
        self.FieldVec = PARAMS["MDFieldVec"]
        self.FieldAmp = PARAMS["MDFieldAmp"]
        self.FieldFreq = PARAMS["MDFieldFreq"]
        self.Tau = PARAMS["MDFieldTau"]
        self.TOn = PARAMS["MDFieldT0"]
        self.UpdateCharges = PARAMS["MDUpdateCharges"]
        self.ChargeFunction = q_
        self.q0 = 0 * self.m
        self.qs = np.ones(self.m.shape)
        self.Mu0 = np.zeros(3)
        self.Mu = np.zeros(3)
        self.mu_his = None
        if q_ is not None:
            self.q0 = np.asarray(self.ChargeFunction(self.x))
            self.qs = self.q0.copy()
            self.Mu0 = Dipole_Naive(self.x, self.q0)
        else:
            self.UpdateCharges = False
        self.MinS, self.MinE, self.Minx = 0, 0.0, None

    def Pulse(self, t_):
        sin_part = np.sin(2.0 * 3.1415 * self.FieldFreq * t_)
        exp_part = (1.0 / np.sqrt(2.0 * 3.1415 * self.Tau * self.Tau)) * np.exp(-1.0 * np.power(t_ - self.TOn, 2.0) / (2.0 * self.Tau * self.Tau))
        amp = self.FieldAmp * sin_part * exp_part
        if np.abs(amp) > 1e-12:
            return self.FieldVec * amp, True
        return np.zeros(3), False

    def ForcesWithCharge(self, x_):
        e, FFForce = self.EnergyAndForce(x_)
        if self.IsOn:
            FFForce = FFForce + 4184.0 * ElectricFieldForce(self.qs, self.EField)
        return e, RemoveInvariantForce(x_, FFForce, self.m)

    def WriteTrajectory(self):
        m = Mol(self.atoms, self.x)
        m.properties["Energy"] = self.EPot
        m.WriteXYZfile(PARAMS["results_dir"], "MDTrajectory" + self.name)

    def _charges(self):
        if self.UpdateCharges and not self.IsOn:
            self.qs = np.asarray(self.ChargeFunction(self.x))
        else:
            self.qs = self.q0

    def Prop(self):
        self.mu_his = np.zeros((self.maxstep, 7))
        step = 0
        while step < self.maxstep:
            self.t = step * self.dt
            self.EField, self.IsOn = self.Pulse(self.t)
            self._charges()
            self.Mu = Dipole_Naive(self.x, self.qs) - self.Mu0
            self.mu_his[step] = [self.t, self.Mu[0], self.Mu[1], self.Mu[2], self.KE, self.EPot, self.KE + self.EPot]
            if self.Tstat is None:
                self.x, self.v, self.a, self.EPot = VelocityVerletStep(None, self.a, self.x, self.v, self.m, self.dt, self.ForcesWithCharge)
            else:
                self.x, self.v, self.a, self.EPot, self.force = self.Tstat.step(None, self.a, self.x, self.v, self.m, self.dt, self.ForcesWithCharge)
            self.KE = KineticEnergy(self.v, self.m)
            if step % 50 == 0 and PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            step += 1


class Annealer(IRTrajectory):
    """Nose-thermostatted anneal from MDAnnealT0 to MDAnnealTF over MDAnnealSteps; keeps the lowest-energy geometry in .Minx."""

    def __init__(self, f_, q_, g0_, name_="anneal", AnnealThresh_=0.000009):
        PARAMS["MDThermostat"] = None
        IRTrajectory.__init__(self, f_, q_, g0_, name_)
        self.AnnealT0 = PARAMS["MDAnnealT0"]
        self.AnnealSteps = PARAMS["MDAnnealSteps"]
        self.AnnealThresh = AnnealThresh_
        self.Tstat = NoseThermostat(self.m, self.v)

    def Prop(self):
        step = 0
        while step < self.AnnealSteps:
            self.t = step * self.dt
            self._charges()
            frac = float(self.AnnealSteps - step) / self.AnnealSteps
            self.Tstat.T = max(0.1, self.AnnealT0 * frac + PARAMS["MDAnnealTF"] * (1.0 - frac) + 1e-10)
            self.x, self.v, self.a, self.EPot, self.force = self.Tstat.step(self.ForceFunction, self.a, self.x, self.v, self.m, self.dt, self.EnergyAndForce)
            if self.EPot < self.MinE and abs(self.EPot - self.MinE) > self.AnnealThresh:
                self.MinE, self.Minx, self.MinS = self.EPot, self.x.copy(), step
                if PARAMS["MDAnnealT0"] > PARAMS["MDAnnealTF"]:
                    self.AnnealT0 = min(PARAMS["MDAnnealT0"], self.Tstat.T + PARAMS["MDAnnealKickBack"])
                step = 0
            self.KE = KineticEnergy(self.v, self.m)
            if step % 7 == 0 and PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            step += 1
