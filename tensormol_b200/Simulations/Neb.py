"""Nudged elastic band (JCP 113, 9978) over an energy/force callback (reference: Simulations/Neb.py:17-220)."""
from __future__ import annotations

import os

import numpy as np

from ..Containers.Mol import Mol
from ..Math.BFGS import BFGS_WithLinesearch, SteepestDescent, VerletOptimizer
from ..Math.DIIS import DIIS
from ..Math.QuasiNewtonTools import ConjGradient, RemoveInvariantForce
from ..Util import *   # noqa: F401,F403


class NudgedElasticBand:
    def __init__(self, f_, g0_, g1_, name_="Neb", thresh_=None, nbeads_=None, fb_=None):
        """f_(x, DoForce) -> (E [Hartree], F [J/mol/A]) or E; g0_, g1_: end-point molecules.
        fb_ (B200 extension, optional): batched callback fb_(xs[B, N, 3], DoForce) -> (E[B], F[B, N, 3]) or E[B]
        (TFMolManage.BatchForce); when given, every solver iteration evaluates ALL beads in one molecule-batch call
        instead of one call per bead, with the same NEB arithmetic on the results."""
        self.name = name_
        self.fb = fb_
        self.thresh = PARAMS["OptThresh"] if thresh_ is None else thresh_
        self.max_opt_step = PARAMS["OptMaxCycles"]
        self.nbeads = PARAMS["NebNumBeads"] if nbeads_ is None else nbeads_
        self.k = PARAMS["NebK"]
        self.f = f_
        self.atoms = g0_.atoms.copy()
        self.natoms = len(self.atoms)
        self.beads = np.array([(1. - l) * g0_.coords + l * g1_.coords for l in np.linspace(0., 1., self.nbeads)])
        self.Fs = np.zeros(self.beads.shape)     # real forces
        self.Ss = np.zeros(self.beads.shape)     # spring forces
        self.Ts = np.zeros(self.beads.shape)     # tangents
        self.Es = np.zeros(self.nbeads)
        self.Esi = np.zeros(self.nbeads)
        self.Rs = np.zeros(self.nbeads)
        self.step = 0
        self.TSI = 0
        solvers = {"SD": SteepestDescent, "Verlet": VerletOptimizer, "BFGS": BFGS_WithLinesearch, "DIIS": DIIS, "CG": ConjGradient}
        if PARAMS["NebSolver"] not in solvers:
            raise Exception("Missing Neb Solver")
        self.Solver = solvers[PARAMS["NebSolver"]](self.WrappedEForce, self.beads)

    def Tangent(self, beads_, i):
        if i == 0 or i == (self.nbeads - 1):
            return np.zeros(self.beads[0].shape)
        t = beads_[i + 1] - beads_[i - 1]
        return t / np.sqrt(np.einsum('ia,ia', t, t))

    def SpringEnergy(self, beads_):
        d = beads_[1:] - beads_[:-1]
        return 0.5 * self.k * self.nbeads * np.sum(d * d)

    def SpringDeriv(self, beads_, i):
        if i == 0 or i == (self.nbeads - 1):
            return np.zeros(self.beads[0].shape)
        return self.k * self.nbeads * (2.0 * beads_[i] - beads_[i + 1] - beads_[i - 1])

    def Parallel(self, v_, t_):
        return t_ * np.einsum("ia,ia", v_, t_)

    def Perpendicular(self, v_, t_):
        return v_ - t_ * np.einsum("ia,ia", v_, t_)

    def BeadAngleCosine(self, beads_, i):
        v1 = beads_[i + 1] - beads_[i]
        v2 = beads_[i - 1] - beads_[i]
        return np.einsum('ia,ia', v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))

    def _energy(self, x):
        out = self.f(x, False)
        return out[0] if isinstance(out, tuple) else out

    def NebForce(self, beads_, i, DoForce=True):
        """Perpendicular true force + parallel spring force; climbing image on the highest bead after 10 steps."""
        if i == 0 or i == (self.nbeads - 1):
            self.Fs[i] = np.zeros(self.beads[0].shape)
            self.Es[i] = self._energy(beads_[i])
        elif DoForce:
            self.Es[i], self.Fs[i] = self.f(beads_[i], DoForce)
        else:
            self.Es[i] = self._energy(beads_[i])
        if not DoForce:
            return self.Es[i]
        t = self.Tangent(beads_, i)
        self.Ts[i] = t
        Spara = self.Parallel(-1.0 * self.SpringDeriv(beads_, i), t)
        self.Ss[i] = Spara
        Fneb = Spara + self.Perpendicular(self.Fs[i].copy(), t)
        if PARAMS["NebClimbingImage"] and self.step > 10 and i == self.TSI:
            Fneb = self.Fs[i] - 2.0 * np.sum(self.Fs[i] * self.Ts[i]) * self.Ts[i]
        return self.Es[i], Fneb

    def _evaluate_band(self, beads_, DoForce):
        """All beads in one batched call: fills Es (every bead) and Fs (interior beads; the end beads keep zero force,
        as in NebForce)."""
        out = self.fb(np.asarray(beads_), DoForce)
        if DoForce:
            Es, Fs = out
            self.Fs[1:-1] = np.asarray(Fs)[1:-1]
            self.Fs[0] = 0.0
            self.Fs[-1] = 0.0
        else:
            Es = out[0] if isinstance(out, tuple) else out
        self.Es[:] = np.asarray(Es, np.float64).reshape(-1)

    def _neb_force_from_stored(self, beads_, i):
        t = self.Tangent(beads_, i)
        self.Ts[i] = t
        Spara = self.Parallel(-1.0 * self.SpringDeriv(beads_, i), t)
        self.Ss[i] = Spara
        Fneb = Spara + self.Perpendicular(self.Fs[i].copy(), t)
        if PARAMS["NebClimbingImage"] and self.step > 10 and i == self.TSI:
            Fneb = self.Fs[i] - 2.0 * np.sum(self.Fs[i] * self.Ts[i]) * self.Ts[i]
        return Fneb

    def WrappedEForce(self, beads_, DoForce=True):
        if self.fb is not None:
            self._evaluate_band(beads_, DoForce)
            if not DoForce:
                return np.sum(self.Es) + self.SpringEnergy(beads_)
            F = np.zeros(beads_.shape)
            for i, bead in enumerate(beads_):
                F[i] = RemoveInvariantForce(bead, self._neb_force_from_stored(beads_, i), self.atoms) / JOULEPERHARTREE   # noqa: F405
            return np.sum(self.Es) + self.SpringEnergy(beads_), F
        if DoForce:
            F = np.zeros(beads_.shape)
            for i, bead in enumerate(beads_):
                self.Es[i], F[i] = self.NebForce(beads_, i, DoForce)
                F[i] = RemoveInvariantForce(bead, F[i], self.atoms) / JOULEPERHARTREE   # noqa: F405
            return np.sum(self.Es) + self.SpringEnergy(beads_), F
        for i in range(len(beads_)):
            self.Es[i] = self.NebForce(beads_, i, DoForce)
        return np.sum(self.Es) + self.SpringEnergy(beads_)

    def IntegrateEnergy(self):
        """Line integral of the force along the band (midpoint rule)."""
        self.Esi[0] = self.Es[0]
        for i in range(1, self.nbeads):
            dR = self.beads[i] - self.beads[i - 1]
            dV = -1 * (self.Fs[i] + self.Fs[i - 1]) / 2.
            self.Esi[i] = self.Esi[i - 1] + np.einsum("ia,ia", dR, dV)

    def WriteTrajectory(self, nm_):
        for i, bead in enumerate(self.beads):
            m = Mol(self.atoms, bead)
            m.properties["bead"] = i
            m.properties["Energy"] = self.Es[i]
            m.properties["NormNebForce"] = np.linalg.norm(self.Fs[i])
            m.WriteXYZfile(PARAMS["results_dir"], nm_ + "Traj")

    def Opt(self, filename="Neb", Debug=False):
        self.step = 0
        self.Fs = np.ones(self.beads.shape)
        PES = np.zeros((self.max_opt_step, self.nbeads))
        while self.step < self.max_opt_step and np.sqrt(np.mean(self.Fs * self.Fs)) > self.thresh:
            self.beads, energy, self.Fs = self.Solver(self.beads)
            PES[self.step] = self.Es.copy()
            self.IntegrateEnergy()
            self.TSI = int(np.argmax(self.Es))
            beadFperp = [np.linalg.norm(self.Perpendicular(self.Fs[i], self.Ts[i])) for i in range(1, self.nbeads - 1)]
            if self.step % 10 == 0:
                self.WriteTrajectory(filename)
            LOGGER.info(self.name + "Step: %i Objective: %.5f RMS Gradient: %.5f  Max Gradient: %.5f |F_perp| : %.5f |F_spring|: %.5f ", self.step,
                        np.sum(PES[self.step]), np.sqrt(np.mean(self.Fs * self.Fs)), np.max(self.Fs), np.mean(beadFperp), np.linalg.norm(self.Ss))
            self.step += 1
        os.makedirs(PARAMS["results_dir"], exist_ok=True)
        np.savetxt(PARAMS["results_dir"] + "NEB_" + filename + "_Energy.txt", PES)
        return self.beads
