"""Small dense linear-algebra helpers used by the drivers (reference: Math/LinearOperations.py)."""
from __future__ import annotations

import numpy as np


def MovingAverage(a, n=3):
    ret = np.cumsum(a, dtype=float)
    ret[n:] = ret[n:] - ret[:-n]
    return ret[n - 1:] / n


def PseudoInverse(mat_):
    U, s, V = np.linalg.svd(mat_)
    sinv = np.where(np.abs(s) > 0.0000001, 1.0 / np.where(s == 0.0, 1.0, s), 0.0)
    return np.dot(np.dot(U, np.diag(sinv)), V)


def MatrixPower(A, p, PrintCondition=False):
    """Raise a Hermitian matrix to a possibly fractional power (singular values floored at 1e-14)."""
    u, s, v = np.linalg.svd(A)
    if PrintCondition:
        print("MatrixPower: Minimal Eigenvalue =", np.min(s))
    s = np.where(np.abs(s) < 1e-14, 1e-14, s)
    return np.dot(u, np.dot(np.diag(np.power(s, p)), v))


def Normalize(x_):
    return x_ / np.sqrt(np.sum(x_ * x_))


def RotationMatrix(axis, theta):
    """Counter-clockwise rotation about `axis` by theta radians (Euler-Rodrigues)."""
    axis = np.asarray(axis, np.float64)
    axis = axis / np.sqrt(np.dot(axis, axis))
    a = np.cos(theta / 2.0)
    b, c, d = -axis * np.sin(theta / 2.0)
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c + a * d), 2 * (b * d - a * c)],
                     [2 * (b * c - a * d), a * a + c * c - b * b - d * d, 2 * (c * d + a * b)],
                     [2 * (b * d + a * c), 2 * (c * d - a * b), a * a + d * d - b * b - c * c]])
