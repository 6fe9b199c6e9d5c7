"""Periodic Metropolis Monte Carlo on a PeriodicForce, energy-only evaluations (reference: Simulations/PeriodicMC.py:17-123).

Trial moves are smooth random displacement fields so that nearby atoms move together: four Gaussian "kicks" at random
points (a random vector each, spread with a random width over the atoms), to which each kick adds, with probability 0.6, a
solenoidal (rotation-like) part around its centre; plus a small uniform jitter. Random numbers are drawn from numpy's
global generator in the reference's order, so a seeded run reproduces the reference's chain.
"""
from __future__ import annotations

import os

import numpy as np

from .. import MolEmb
from ..Math.Statistics import OnlineEstimator
from ..Util import *   # noqa: F401,F403
from .PeriodicMD import PeriodicVelocityVerlet


class PeriodicMonteCarlo(PeriodicVelocityVerlet):
    def __init__(self, Force_, name_="PdicMC"):
        PeriodicVelocityVerlet.__init__(self, Force_, name_)
        e0, f0 = self.PForce(self.PForce.mol0.coords)
        self.eold = e0
        self.Estat = OnlineEstimator(e0)
        self.RDFold = self.PForce.RDF(self.PForce.mol0.coords)
        self.RDFstat = OnlineEstimator(self.RDFold)
        self.Xstat = OnlineEstimator(self.x)
        self.PACCstat = OnlineEstimator(1.0)
        self.kbt = KAYBEETEE * (PARAMS["MDTemp"] / 300.0)    # Hartree   # noqa: F405
        self.Eav = self.dE2 = self.Xav = self.dX2 = None
        self.Pacc = 0.0

    def RandomVectorField(self, x_):
        """V(x_j) = sum_i v_i N(|x_j - p_i|; sigma) + solenoidal parts; p_i, v_i, sigma random."""
        hi, lo = np.max(x_), np.min(x_)
        npts = 4
        pts = np.random.uniform(hi - lo, size=(npts, 3)) + lo       # numpy reads the lone argument as `low` (high = 1): kept
        rmagn = np.random.normal(scale=0.07, size=(npts, 1))
        theta = np.random.uniform(3.1415, size=(npts, 1))
        phi = np.random.uniform(2.0 * 3.1415, size=(npts, 1))
        magn = np.concatenate([rmagn * np.sin(theta) * np.cos(phi), rmagn * np.sin(theta) * np.sin(phi), rmagn * np.cos(theta)], axis=1)
        D = MolEmb.Make_DistMat_ForReal(np.concatenate([pts, x_]), npts)[:, npts:]
        sigma = np.random.uniform(2.2) + 0.02
        weight = (1.0 / np.sqrt(6.2831 * sigma * sigma)) * np.exp(-1.0 * D * D / (2 * sigma * sigma))
        field = np.einsum('jk,ji->ik', magn, weight)
        for i in range(npts):
            if np.random.random() < 0.6:
                vs = x_ - pts[i]
                vs = vs / np.linalg.norm(vs, axis=1)[:, None]
                field += np.cross(vs, 4.0 * magn[i]) * weight[i, :, np.newaxis]
        return field

    def MetropolisHastings(self, x_):
        dx = self.RandomVectorField(x_)
        dx += np.random.uniform(size=x_.shape) * 0.0005
        edx = self.PForce(x_ + dx, DoForce=False)[0]
        with np.errstate(over="ignore"):
            PMove = min(1.0, np.exp(-(edx - self.eold) / self.kbt))
        if np.random.random() < PMove:
            self.x = self.PForce.lattice.ModuloLattice(x_ + dx)
            self.eold = edx
            self.RDFold = self.PForce.RDF(self.x)
            self.Pacc, _ = self.PACCstat(1.0)
        else:
            self.Pacc, _ = self.PACCstat(0.0)
        self.Eav, self.dE2 = self.Estat(self.eold)
        self.Xav, self.dX2 = self.Xstat(self.x)

    def Prop(self):
        step = 0
        self.md_log = np.zeros((self.maxstep, 7))
        os.makedirs(PARAMS["results_dir"], exist_ok=True)
        while step < self.maxstep:
            self.t = step
            self.MetropolisHastings(self.x)
            rdf, rdf2 = self.RDFstat(self.RDFold)
            self.md_log[step, 0] = self.t
            self.md_log[step, 5] = self.EPot          # as the reference: EPot is not updated by the chain (stays EPot0)
            if step % 3 == 0 and PARAMS["MDLogTrajectory"]:
                self.WriteTrajectory()
            if step % 500 == 0:
                self._save_log()
                np.savetxt(PARAMS["results_dir"] + "MCRDF" + self.name + ".txt", rdf)
                np.savetxt(PARAMS["results_dir"] + "MCRDF2" + self.name + ".txt", rdf2)
            step += 1
            LOGGER.info("Step: %i <E>(kJ/mol): %.5f sqrt(<dE2>): %.5f sqrt(<dX2>): %.5f Paccept %.5f Rho(g/cm**3): %.5f ", step, self.Eav,
                        np.sqrt(self.dE2), np.sqrt(np.linalg.norm(self.dX2)), self.Pacc, self.Density())
