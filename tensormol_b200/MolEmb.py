"""Drop-in for the hot entries of the reference's CPython extension `MolEmb` (C_API/MolEmb.cpp): the
neighbour search runs on the B200 through libtmolb200 (tm_nlist); there is no CPU fallback.

    Make_NListNaive(xyz, rng, nreal, DoPerms) -> list[nreal] of list[int]     (MolEmb.cpp:1180-1247)
    Make_NListLinear(same)                    -> same lists                    (MolEmb.cpp:1253-1383)

`Make_NListLinear` deliberately returns the *Naive* semantics: the reference's Linear variant drops pairs in
boxes larger than ~2 Rc^2, and both are meant to produce the same sets.
The small helpers the periodic wrapper uses off the hot path (Make_DistMat, Make_DistMat_ForReal, CountInRange, GetRDF_Bin)
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
    """Cumulative pair-count histogram used by PeriodicForce.RDF (MolEmb.cpp:1082-1119, caller Periodic.py:425):
    out[k] = mean over the centres i < natoms of element z1 of the number of partners j != i of element z2 with
    int(d_ij / dx) <= k, k < int(rng / dx)."""
    zt = np.asarray(zt)
    xt = np.asarray(xt, np.float64)
    nbin = int(float(rng) / float(dx))
    ni = np.zeros(nbin)
    centres = np.where(zt[:int(natoms)] == z1)[0]
    others = np.where(zt == z2)[0]
    for i in centres:
        d = np.linalg.norm(xt[others] - xt[i], axis=1)[others != i]
        b = (d / dx).astype(np.int64)
        ni += np.cumsum(np.bincount(b[b < nbin], minlength=nbin)[:nbin])
    with np.errstate(invalid="ignore", divide="ignore"):
        return ni / float(len(centres))          # the reference divides by zero too when no centre matches


def GetRDF_Bin(xyz, Zs, cut, dr, cellsize, ele1, ele2):
    """Bin indices int(d / dr) of every (centre of element ele1, partner of element ele2 in the cubic cell of edge
    `cellsize` or one of its images) distance with d + 1e-11 < cut (MolEmb.cpp:1121-1178; callers Periodic.py:452,
    samples/test_h2o_peri.py:404). Order: centres ascending, partners by image block (cell first, then the blocks in
    i, j, k order without (0,0,0)) and atom index, as the reference's list."""
    x = np.asarray(xyz, np.float64)
    z = np.asarray(Zs)
    nat = x.shape[0]
    ntess = int(float(cut) / float(cellsize)) + 1
    r = range(-ntess, ntess + 1)
    shifts = [(0, 0, 0)] + [(i, j, k) for i in r for j in r for k in r if (i, j, k) != (0, 0, 0)]
    xp = (x[None, :, :] + float(cellsize) * np.asarray(shifts, np.float64)[:, None, :]).reshape(-1, 3)
    partner = np.where(np.tile(z == ele2, len(shifts)))[0]
    out = []
    for i in np.where(z == ele1)[0]:
        p = partner[partner != i]
        d = np.sqrt(((x[i] - xp[p]) ** 2).sum(1)) + 0.00000000001
        out.extend((d[d < cut] / dr).astype(np.int64).tolist())
    return out
