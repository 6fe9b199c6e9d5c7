"""Direct inversion in the iterative subspace with the solver interface x, e, g = solver(x) (reference: Math/DIIS.py),
restated so that the iterates equal the reference's."""
from __future__ import annotations

import numpy as np

from ..Util import *   # noqa: F401,F403


class DIIS:
    """Keeps the last PARAMS["DiisSize"] points V_k and residuals (forces) R_k and returns sum_k c_k V_k with the
    coefficients of the bordered overlap system [[S, -1], [-1, 0]] c = (0, .., 0, -1), solved through an SVD
    pseudo-inverse with singular values <= 1e-7 dropped; the first call is a plain step of 0.02 R.

    Reproduced quirk (DIIS.py:43-53): once the window is full the reference rolls its (DiisSize+1)-square overlap table
    and writes the new residual's overlaps into the BORDER row / column (index DiisSize), so the DiisSize-square block
    the system is built from lags one call behind the stored vectors (its last row / column are the overlaps of the
    previous newest residual; zero the first time)."""

    def __init__(self, ForceAndEnergy_, x0_=None):
        self.m_max = PARAMS["DiisSize"]
        self.n_now = 0
        self.Vs = None
        self.Rs = None
        self.S = np.zeros((self.m_max + 1, self.m_max + 1))
        self.EForce = ForceAndEnergy_

    def __call__(self, new_vec_):
        e, g = self.EForce(new_vec_)
        return self.NextStep(new_vec_, g), e, g

    def NextStep(self, new_vec_, new_residual_):
        if self.Vs is None:
            self.Vs = np.zeros([self.m_max] + list(new_vec_.shape))
            self.Rs = np.zeros([self.m_max] + list(new_vec_.shape))
        r = np.asarray(new_residual_, np.float64).reshape(-1)
        if self.n_now < self.m_max:
            slot = col = self.n_now
            self.n_now += 1
        else:
            self.Vs = np.roll(self.Vs, -1, axis=0)
            self.Rs = np.roll(self.Rs, -1, axis=0)
            self.S = np.roll(self.S, (-1, -1), axis=(0, 1))
            slot, col = self.m_max - 1, self.m_max
        self.Vs[slot] = new_vec_
        self.Rs[slot] = new_residual_
        k = col if col < self.m_max else self.m_max
        ov = self.Rs[:k].reshape(k, r.size) @ r
        self.S[:k, col] = ov
        self.S[col, :k] = ov
        self.S[col, col] = r @ r
        n = self.n_now
        if n < 2:
            return new_vec_ + 0.02 * new_residual_
        M = -np.ones((n + 1, n + 1))
        M[:n, :n] = self.S[:n, :n]
        M[n, n] = 0.0
        rhs = np.zeros(n + 1)
        rhs[n] = -1.0
        U, s, V = np.linalg.svd(M)
        keep = np.abs(s) > 0.0000001
        sinv = np.zeros_like(s)
        sinv[keep] = 1.0 / s[keep]
        c = np.dot(np.dot(np.dot(U, np.diag(sinv)), V), rhs)
        return np.tensordot(c[:n], self.Vs[:n], axes=(0, 0))
