"""Direct inversion in the iterative subspace with the solver interface x, e, g = solver(x) (reference: Math/DIIS.py)."""
from __future__ import annotations

import numpy as np

from ..Util import *   # noqa: F401,F403


class DIIS:
    def __init__(self, ForceAndEnergy_, x0_=None):
        self.m_max = PARAMS["DiisSize"]
        self.Vs, self.Rs = [], []
        self.EForce = ForceAndEnergy_

    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        return self.NextStep(new_vec_, g), e, g

    def NextStep(self, new_vec_, new_residual_):
        self.Vs.append(new_vec_.copy())
        self.Rs.append(new_residual_.copy())
        if len(self.Vs) > self.m_max:
            self.Vs.pop(0)
            self.Rs.pop(0)
        n = len(self.Vs)
        if n < 2:
            return new_vec_ + 0.02 * new_residual_
        R = np.array([r.reshape(-1) for r in self.Rs])
        M = -np.ones((n + 1, n + 1))
        M[:n, :n] = R @ R.T
        M[n, n] = 0.0
        rhs = np.zeros(n + 1)
        rhs[n] = -1.0
        U, s, V = np.linalg.svd(M)
        sinv = np.where(np.abs(s) > 1e-7, 1.0 / np.where(s == 0, 1.0, s), 0.0)
        c = (U @ np.diag(sinv) @ V) @ rhs
        nxt = sum(c[i] * (self.Vs[i] + 0.02 * self.Rs[i]) for i in range(n))
        return nxt
