"""First-order and quasi-Newton solvers with the common solver interface  x, e, g = solver(x)
(reference: Math/BFGS.py).  `g` is the force (descent direction), as returned by the wrapped callbacks."""
from __future__ import annotations

import numpy as np

from ..Util import *   # noqa: F401,F403


class SteepestDescent:
    def __init__(self, ForceAndEnergy_, x0_):
        self.step = 0
        self.x0 = x0_.copy()
        self.natom = self.x0.shape[0] if len(self.x0.shape) == 2 else self.x0.shape[0] * self.x0.shape[1]
        self.EForce = ForceAndEnergy_

    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        self.step += 1
        return new_vec_ + PARAMS["SDStep"] * g, e, g


class VerletOptimizer:
    """Damped-dynamics minimiser: velocity is zeroed whenever it points against the force."""

    def __init__(self, ForceAndEnergy_, x0_):
        self.step = 0
        self.x0 = x0_.copy()
        self.v = np.zeros(x0_.shape)
        self.a = np.zeros(x0_.shape)
        self.dt = 0.1
        self.EForce = ForceAndEnergy_

    def __call__(self, x_):
        x = x_ + self.v * self.dt + 0.5 * self.a * self.dt * self.dt
        e, f_x_ = self.EForce(x)
        self.v += 0.5 * (self.a + f_x_) * self.dt
        if np.sum(self.v * f_x_) < 0:
            self.v *= 0.0
            self.a *= 0.0
        self.step += 1
        return x, e, f_x_


class BFGS(SteepestDescent):
    """The reference's limited-memory quasi-Newton step (Math/BFGS.py:78-148), restated so that its iterates are
    reproduced exactly: a window of the last PARAMS["MaxBFGS"] points R_k and forces F_k; curvature pairs are the
    differences of CONSECUTIVE window entries, s_i = R_i - R_(i-1), y_i = F_i - F_(i-1) (y is a force difference, so
    s.y < 0 on a convex surface and the returned vector -z points along the force); the two-loop recursion runs over
    the pairs i = n-1 .. 1 with n = min(MaxBFGS, number of earlier calls), i.e. the newest pair is only used once the
    window is full; the initial scaling H = s.y / y.y comes from the pair ending at entry min(MaxBFGS-1, calls); when
    |H| < 1e-3 the step falls back to -0.001 F. `__call__` moves by 0.005 times the returned vector."""

    def __init__(self, ForceAndEnergy_, x0_):
        SteepestDescent.__init__(self, ForceAndEnergy_, x0_)
        self.m_max = PARAMS["MaxBFGS"]
        self.R_Hist = np.zeros([self.m_max] + list(self.x0.shape))
        self.F_Hist = np.zeros([self.m_max] + list(self.x0.shape))

    def _pair(self, i):
        return self.R_Hist[i] - self.R_Hist[i - 1], self.F_Hist[i] - self.F_Hist[i - 1]

    def BFGSstep(self, new_vec_, new_residual_):
        if self.step >= self.m_max:                       # window full: drop the oldest entry
            self.R_Hist = np.roll(self.R_Hist, -1, axis=0)
            self.F_Hist = np.roll(self.F_Hist, -1, axis=0)
        slot = min(self.step, self.m_max - 1)
        self.R_Hist[slot] = new_vec_
        self.F_Hist[slot] = new_residual_
        n = min(self.m_max, self.step)
        q = np.array(new_residual_, dtype=np.float64)
        alpha = np.zeros(n)
        for i in range(n - 1, 0, -1):
            s, y = self._pair(i)
            alpha[i] = (1.0 / np.sum(y * s)) * np.sum(s * q)
            q -= alpha[i] * y
        H = 1.0
        if self.step >= 1:
            s, y = self._pair(min(self.m_max - 1, self.step))
            H = np.sum(s * y) / np.sum(y * y)
            if abs(H) < 0.001:
                self.step += 1
                return -0.001 * new_residual_
        z = H * q
        for i in range(1, n):
            s, y = self._pair(i)
            beta = (1.0 / np.sum(y * s)) * np.sum(y * z)
            z += s * (alpha[i] - beta)
        self.step += 1
        return -1.0 * z

    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        z = self.BFGSstep(new_vec_, g)
        return new_vec_ + 0.005 * z, e, g


class BFGS_WithLinesearch(BFGS):
    """BFGS direction followed by the reference's golden-section search along it (Math/BFGS.py:150-244): bracket
    [x, x + alpha z]; an overstep (f(x) lowest) shrinks alpha by 1.71 and restarts (returns x once alpha <= 1e-4), an
    understep (f(x + alpha z) lowest) grows alpha by 1.7 and restarts, otherwise the bracket is narrowed until its mean
    width per atom is below `thresh`."""

    def __init__(self, ForceAndEnergy_, x0_):
        BFGS.__init__(self, ForceAndEnergy_, x0_)
        self.alpha = PARAMS["GSSearchAlpha"]
        self.Energy = lambda x: self.EForce(x, False)

    def _bracket(self, x0_, p_):
        a = x0_
        b = x0_ + self.alpha * p_
        c = b - (b - a) / GOLDENRATIO   # noqa: F405
        d = a + (b - a) / GOLDENRATIO   # noqa: F405
        return a, b, c, d, self.Energy(a), self.Energy(b), self.Energy(c), self.Energy(d)

    def LineSearch(self, x0_, p_, thresh=0.0001):
        width = 10.0
        a, b, c, d, fa, fb, fc, fd = self._bracket(x0_, p_)
        while width > thresh:
            if fa < fc and fa < fd and fa < fb:
                if self.alpha <= 0.0001:
                    return a
                self.alpha /= 1.71
                a, b, c, d, fa, fb, fc, fd = self._bracket(x0_, p_)
            elif fb < fc and fb < fd and fb < fa:
                if self.alpha < 100.0:
                    self.alpha *= 1.7
                a, b, c, d, fa, fb, fc, fd = self._bracket(x0_, p_)
            elif fc < fd:
                b, fb = d, fd
                c = b - (b - a) / GOLDENRATIO   # noqa: F405
                d = a + (b - a) / GOLDENRATIO   # noqa: F405
                fc, fd = self.Energy(c), self.Energy(d)
            else:
                a, fa = c, fc
                c = b - (b - a) / GOLDENRATIO   # noqa: F405
                d = a + (b - a) / GOLDENRATIO   # noqa: F405
                fc, fd = self.Energy(c), self.Energy(d)
            width = np.sum(np.linalg.norm(a - b, axis=1)) / self.natom     # axis 1 also for bead arrays, as the reference
        return (b + a) / 2

    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        z = self.BFGSstep(new_vec_, g)
        return self.LineSearch(new_vec_, z), e, g
