"""Drop-in for the hot entries of the reference's CPython extension `MolEmb` (C_API/MolEmb.cpp): the
neighbour search runs on the B200 through libtmolb200 (tm_nlist); there is no CPU fallback.

    Make_NListNaive(xyz, rng, nreal, DoPerms) -> list[nreal] of list[int]     (MolEmb.cpp:1180-1247)
    Make_NListLinear(same)                    -> same lists                    (MolEmb.cpp:1253-1383)

`Make_NListLinear` deliberately returns the *Naive* semantics: the reference's Linear variant drops pairs in
boxes larger than ~2 Rc^2, and both are meant to produce the same sets.
The small helpers the periodic wrapper uses off the hot path (Make_DistMat, Make_DistMat_ForReal, CountInRange)
are plain numpy.
"""
from __future__ import annotations

import numpy as np

_default_engine = None


def _engine():
    global _default_engine
    if _default_engine is None:
        from .engine import Engine
        from .Util import PARAMS
        P = dict(PARAMS)
        P["EECutoffOn"] = 0.0
        P["NeuronType"] = "sigmoid_with_param"
        _default_engine = Engine([1], [8], P, device=int(PARAMS.get("B200Device", 0)))
    return _default_engine


def nlist_csr(xyz, rng, nreal, DoPerms):
    """CSR (offsets, indices) form of Make_NListNaive, rows unsorted."""
    return _engine().nlist(np.ascontiguousarray(xyz, np.float64), float(rng), int(nreal), int(DoPerms))


def Make_NListNaive(xyz, rng, nreal, DoPerms):
    off, idx = nlist_csr(xyz, rng, nreal, DoPerms)
    return [idx[off[i]:off[i + 1]].tolist() for i in range(int(nreal))]


def Make_NListLinear(xyz, rng, nreal, DoPerms):
    return Make_NListNaive(xyz, rng, nreal, DoPerms)


def Make_DistMat(xyz):
    x = np.asarray(xyz, np.float64)
    d = x[:, None, :] - x[None, :, :]
    return np.sqrt((d * d).sum(-1))


def Make_DistMat_ForReal(xyz, nreal):
    """Distances from the first nreal points to all points (MolEmb.cpp: used by Lattice.__init__, Periodic.py:23)."""
    x = np.asarray(xyz, np.float64)
    d = x[:int(nreal), None, :] - x[None, :, :]
    return np.sqrt((d * d).sum(-1))


def CountInRange(zt, xt, natoms, z1, z2, rng, dx):
    """Cumulative pair-count histogram used by PeriodicForce.RDF (Periodic.py:425)."""
    zt = np.asarray(zt)
    xt = np.asarray(xt, np.float64)
    nbin = int(np.arange(0.0, rng, dx).shape[0])
    ni = np.zeros(nbin)
    centres = np.where(zt[:int(natoms)] == z1)[0]
    others = np.where(zt == z2)[0]
    for i in centres:
        d = np.linalg.norm(xt[others] - xt[i], axis=1)
        d = d[(others != i) & (d < rng)]
        b = (d / dx).astype(int)
        ni += np.cumsum(np.bincount(b, minlength=nbin)[:nbin])
    return ni / max(len(centres), 1)
