"""Finite-difference tools, invariant-force projection and the conjugate-gradient optimiser the geometry
drivers use (reference: Math/QuasiNewtonTools.py).  All consume the f(x[,DoForce]) -> (E, F) callback."""
from __future__ import annotations

import numpy as np

from ..Util import *   # noqa: F401,F403
from .LinearOperations import PseudoInverse


def RmsForce(f_):
    return np.mean(np.linalg.norm(f_, axis=1))


def CenterOfMass(x_, m_):
    return np.einsum("m,mx->x", m_, x_) / np.sum(m_)


def InertiaTensor(x_, m_):
    m = np.asarray(m_, np.float64)
    r2 = np.sum(x_ * x_, axis=1)
    return np.einsum("a,a->", m, r2) * np.eye(3) - np.einsum("a,ai,aj->ij", m, x_, x_)


def FdiffGradient(f_, x_, eps_=0.0001):
    """One-sided (forward) finite-difference gradient of a scalar or array valued function, (f(x + eps e_k) - f(x)) / eps
    as the reference's debugging helper (:43-58); output shape x_.shape + f(x).shape."""
    x0 = np.array(x_, dtype=np.float64)
    f0 = np.asarray(f_(x0))
    tore = np.zeros(x0.shape + f0.shape)
    for k in np.ndindex(*x0.shape):
        xk = x0.copy()
        xk[k] += eps_
        tore[k] = (np.asarray(f_(xk)) - f0) / eps_
    return tore


def FdiffHessian(f_, x_, eps_=0.001, mode_="forward", grad_=None):
    """Finite-difference Hessian: 'forward' / 'central' differences of f_, or 'gradient' differences of grad_."""
    x0 = np.array(x_, dtype=np.float64)
    n = x0.size
    shp = x0.shape
    flat = lambda v: v.reshape(shp)   # noqa: E731
    H = np.zeros((n, n))
    e = np.eye(n) * eps_
    if mode_ == "gradient" and grad_ is not None:
        gp = np.array([np.asarray(grad_(flat(x0.reshape(-1) + e[i]))).reshape(-1) for i in range(n)])
        gm = np.array([np.asarray(grad_(flat(x0.reshape(-1) - e[i]))).reshape(-1) for i in range(n)])
        D = (gp - gm) / (2.0 * eps_)
        H = 0.5 * (D + D.T)
    elif mode_ == "forward":
        f0 = float(f_(x0))
        fi = np.array([float(f_(flat(x0.reshape(-1) + e[i]))) for i in range(n)])
        for i in range(n):
            for j in range(i, n):
                fij = float(f_(flat(x0.reshape(-1) + e[i] + e[j])))
                H[i, j] = H[j, i] = (fij - fi[i] - fi[j] + f0) / eps_ / eps_
    else:
        for i in range(n):
            for j in range(i, n):
                v = x0.reshape(-1)
                H[i, j] = H[j, i] = (float(f_(flat(v + e[i] + e[j]))) - float(f_(flat(v + e[i] - e[j])))
                                     - float(f_(flat(v - e[i] + e[j]))) + float(f_(flat(v - e[i] - e[j])))) / (4.0 * eps_ * eps_)
    return H.reshape(shp + shp)


def HarmonicSpectra(f_, x_, at_, grad_=None, eps_=0.001, WriteNM_=False, Mu_=None):
    """Finite-difference normal-mode analysis (http://gaussian.com/vib/).  f_: energy in Hartree of coordinates in A.
    Returns wavenumbers (cm^-1, negative for imaginary modes) and the mass-weighted Cartesian modes."""
    n = x_.shape[0]
    n3 = 3 * n
    m_ = np.array([ATOMICMASSESAMU[z - 1] * ELECTRONPERPROTONMASS for z in np.asarray(at_).tolist()])   # noqa: F405
    if grad_ is not None:
        cHess = FdiffHessian(f_, x_, 0.0005, "gradient", grad_).reshape((n3, n3))
    else:
        cHess = FdiffHessian(f_, x_, 0.0005).reshape((n3, n3))
    cHess = cHess / (BOHRPERA * BOHRPERA)   # noqa: F405
    w3 = np.repeat(m_, 3)
    cHess = cHess / np.sqrt(np.outer(w3, w3))
    w, v = np.linalg.eigh(cHess)
    wave = np.sign(w) * np.sqrt(np.abs(w)) * WAVENUMBERPERHARTREE   # noqa: F405
    if WriteNM_:
        from ..Containers.Mol import Mol
        for i in range(n3):
            nm = (v[:, i] / np.sqrt(w3 / ELECTRONPERPROTONMASS)).reshape((n, 3))   # noqa: F405
            if Mu_ is not None:
                dmudq = (Mu_(x_ + 0.01 * nm) - Mu_(x_)) / 0.01
                print("|dmu/dQ|^2 ", np.dot(dmudq, dmudq.T))
            for alpha in np.append(np.linspace(0.1, -0.1, 30), np.linspace(0.1, -0.1, 30)):
                Mol(at_, x_ + alpha * nm).WriteXYZfile("./results/", "NormalMode_" + str(i))
    return wave, v


def RemoveInvariantForce(x_, f_, m_):
    """Removes net force and torque from f_ (weights m_; the reference passes atomic numbers from the optimisers)."""
    if PARAMS["RemoveInvariant"] is False:
        return f_
    m = np.asarray(m_, np.float64)
    fnet = np.sum(f_, axis=0)
    fnew = f_ - np.einsum("m,f->mf", m, fnet) / np.sum(m)
    torque = np.sum(np.cross(x_, fnew), axis=0)
    dwdt = np.dot(PseudoInverse(InertiaTensor(x_, m)), torque)
    fcorr = m[:, None] * np.cross(dwdt[None, :], x_)
    return fnew - fcorr


class ConjGradient:
    """Polak-Ribiere conjugate gradient with a golden-section line search (reference :350-465)."""

    def __init__(self, f_, x0_, thresh_=0.0001):
        self.EForce = f_
        self.Energy = lambda x: self.EForce(x, False)
        self.x0 = x0_.copy()
        self.xold = x0_.copy()
        self.e, self.gold = self.EForce(x0_)
        self.s = self.gold.copy()
        self.thresh = thresh_
        self.alpha = PARAMS["GSSearchAlpha"]

    def Reset(self, x0_):
        self.xold = x0_.copy()
        self.e, self.gold = self.EForce(x0_)
        self.s = self.gold.copy()
        self.alpha = PARAMS["GSSearchAlpha"]

    def BetaPR(self, g):
        betapr = np.sum(g * (g - self.gold)) / np.sum(self.gold * self.gold)
        self.gold = g.copy()
        return max(0, betapr)

    def __call__(self, x0):
        e, g = self.EForce(x0)
        self.s = g + self.BetaPR(g) * self.s
        self.xold = self.LineSearch(x0, self.s, self.thresh)
        return self.xold, e, g

    def _bracket(self, x0_, p_):
        a = x0_
        b = x0_ + self.alpha * p_
        c = b - (b - a) / GOLDENRATIO   # noqa: F405
        d = a + (b - a) / GOLDENRATIO   # noqa: F405
        return a, b, c, d, self.Energy(a), self.Energy(b), self.Energy(c), self.Energy(d)

    def LineSearch(self, x0_, p_, thresh=0.0001):
        rmsdist = 10.0
        a, b, c, d, fa, fb, fc, fd = self._bracket(x0_, p_)
        while rmsdist > thresh:
            if fa < fc and fa < fd and fa < fb:        # overstep: shrink the bracket
                if self.alpha > 0.00001:
                    self.alpha /= 1.8001
                else:
                    return a
                a, b, c, d, fa, fb, fc, fd = self._bracket(x0_, p_)
            elif fb < fc and fb < fd and fb < fa:      # understep: accept and grow the trial step
                if self.alpha < 100.0:
                    self.alpha *= 1.8
                return (x0_ + self.alpha * p_ + x0_) / 2
            elif fc < fd:
                b = d
                c = b - (b - a) / GOLDENRATIO   # noqa: F405
                d = a + (b - a) / GOLDENRATIO   # noqa: F405
                fb = fd
                fc, fd = self.Energy(c), self.Energy(d)
            else:
                a = c
                c = b - (b - a) / GOLDENRATIO   # noqa: F405
                d = a + (b - a) / GOLDENRATIO   # noqa: F405
                fa = fc
                fc, fd = self.Energy(c), self.Energy(d)
            rmsdist = np.sum(np.linalg.norm(a - b, axis=1)) / a.shape[0]     # axis 1 also for (bead, atom, 3) arrays, as the reference (:462)
        return (b + a) / 2
