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
    """Limited-memory BFGS (two-loop recursion) on the force; memory PARAMS["MaxBFGS"]."""

    def __init__(self, ForceAndEnergy_, x0_):
        SteepestDescent.__init__(self, ForceAndEnergy_, x0_)
        self.m_max = PARAMS["MaxBFGS"]
        self.S, self.Y = [], []
        self.xlast = None
        self.glast = None

    def Direction(self, x, g):
        """g is the force (= -gradient)."""
        grad = -g.reshape(-1)
        if self.xlast is not None:
            s = (x - self.xlast).reshape(-1)
            y = grad - self.glast
            if np.dot(s, y) > 1e-12:
                self.S.append(s)
                self.Y.append(y)
                if len(self.S) > self.m_max:
                    self.S.pop(0)
                    self.Y.pop(0)
        self.xlast, self.glast = x.copy(), grad.copy()
        q = grad.copy()
        al = []
        for s, y in zip(reversed(self.S), reversed(self.Y)):
            a = np.dot(s, q) / np.dot(y, s)
            al.append(a)
            q -= a * y
        if self.S:
            q *= np.dot(self.S[-1], self.Y[-1]) / np.dot(self.Y[-1], self.Y[-1])
        else:
            q *= PARAMS["SDStep"]
        for (s, y), a in zip(zip(self.S, self.Y), reversed(al)):
            b = np.dot(y, q) / np.dot(y, s)
            q += s * (a - b)
        return -q.reshape(x.shape)

    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        d = self.Direction(new_vec_, g)
        nrm = np.max(np.abs(d))
        if nrm > PARAMS["OptMaxStep"]:
            d *= PARAMS["OptMaxStep"] / nrm
        self.step += 1
        return new_vec_ + d, e, g


class BFGS_WithLinesearch(BFGS):
    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        d = self.Direction(new_vec_, g)
        t = 1.0
        for _ in range(8):     # backtracking on the energy
            if self.EForce(new_vec_ + t * d, False) <= e:
                break
            t *= 0.5
        self.step += 1
        return new_vec_ + t * d, e, g
