"""Atom neighbour lists with the reference's classes and array layouts (ForceModifiers/Neighbors.py:23-489).
The search (and, for the element-channel tables, the whole assembly) runs in libtmolb200 on the GPU:
  NeighborList.buildPairs / buildPairsAndTriples           <- tm_nlist        (replaces MolEmb + Python loops :75-201)
  NeighborListSet.buildPairsAndTriplesWithEleIndex(...)    <- tm_pairs_triples_ele  (replaces :344-467)
Returned dtypes follow the reference: uint64 for buildPairs/buildPairsAndTriples, float64-typed integers for the
element-index tables (quirk Q16)."""
from __future__ import annotations

import numpy as np

from .. import MolEmb
from ..Util import *   # noqa: F401,F403


def _table_engine(eles):
    """An engine whose element list covers `eles` (tm_pairs_triples_ele needs the element order only)."""
    from ..engine import Engine
    key = tuple(sorted(int(e) for e in np.asarray(eles).reshape(-1)))
    if key not in _table_engine.cache:
        P = dict(PARAMS)
        P["EECutoffOn"] = 0.0
        P["NeuronType"] = "sigmoid_with_param"
        _table_engine.cache[key] = Engine(list(key), [8], P, device=int(PARAMS.get("B200Device", 0)))
    return _table_engine.cache[key]


_table_engine.cache = {}


class NeighborList:
    def __init__(self, x_, DoTriples_=False, DoPerms_=False, ele_=None, alg_=None, sort_=False):
        self.natom = x_.shape[0]
        self.x = x_.copy()
        self.pairs = None
        self.triples = None
        self.DoTriples = DoTriples_
        self.DoPerms = DoPerms_
        self.ele = ele_
        self.npairs = None
        self.ntriples = None
        self.alg = 0 if alg_ is None else alg_
        self.sort = sort_

    def Update(self, x_, rcut_pairs=5.0, rcut_triples=5.0, molind_=None, nreal_=None):
        self.x = x_.copy()
        if self.DoTriples:
            self.pairs, self.triples = self.buildPairsAndTriples(rcut_pairs, rcut_triples, molind_, nreal_=nreal_)
            self.npairs = self.pairs.shape[0]
            self.ntriples = self.triples.shape[0]
        else:
            self.pairs = self.buildPairs(rcut_pairs, molind_, nreal_=nreal_)
            self.npairs = self.pairs.shape[0]

    def _rows(self, rcut, nreal):
        off, idx = MolEmb.nlist_csr(self.x, rcut, nreal, int(self.DoPerms))
        i = np.repeat(np.arange(nreal, dtype=np.int64), np.diff(off))
        return off, idx, i

    @TMTiming("NeighborList::BuildPairs")
    def buildPairs(self, rcut=5.0, molind_=None, nreal_=None):
        """(npair x 2) i,j  or (npair x 3) mol,i,j  uint64 (Neighbors.py:75-115)."""
        ntodo = self.natom if nreal_ is None else int(nreal_)
        off, idx, i = self._rows(rcut, ntodo)
        if molind_ is not None:
            return np.stack([np.full_like(i, molind_), i, idx], axis=1).astype(np.uint64).reshape(-1, 3)
        return np.stack([i, idx], axis=1).astype(np.uint64).reshape(-1, 2)

    def buildPairsAndTriples(self, rcut_pairs=5.0, rcut_triples=5.0, molind_=None, nreal_=None):
        """pairs as above; triples (i,j,k) with k>j by index, then the atom with the smaller atomic number first
        (Neighbors.py:160-198)."""
        ntodo = self.natom if nreal_ is None else int(nreal_)
        p = self.buildPairs(rcut_pairs, molind_, nreal_)
        off, idx, _ = self._rows(rcut_triples, ntodo)
        cnt = np.diff(off)
        ti, tj, tk = [], [], []
        for c in np.unique(cnt):
            if c < 2:
                continue
            centres = np.where(cnt == c)[0]
            a, b = np.triu_indices(int(c), 1)
            nb = np.sort(idx[off[centres][:, None] + np.arange(c)[None, :]], axis=1)
            ti.append(np.repeat(centres, a.size))
            tj.append(nb[:, a].reshape(-1))
            tk.append(nb[:, b].reshape(-1))
        if ti:
            i, j, k = np.concatenate(ti), np.concatenate(tj), np.concatenate(tk)
        else:
            i = j = k = np.zeros(0, np.int64)
        if self.ele is not None and i.size:
            swap = np.asarray(self.ele)[j] > np.asarray(self.ele)[k]
            j, k = np.where(swap, k, j), np.where(swap, j, k)
        order = np.lexsort((k, j, i))
        i, j, k = i[order], j[order], k[order]
        if molind_ is not None:
            t = np.stack([np.full_like(i, molind_), i, j, k], axis=1).astype(np.uint64).reshape(-1, 4)
        else:
            t = np.stack([i, j, k], axis=1).astype(np.uint64).reshape(-1, 3)
        return p, t


class NeighborListSet:
    def __init__(self, x_, nnz_, DoTriples_=False, DoPerms_=False, ele_=None, alg_=None, sort_=False):
        """x_: NMol x MaxNAtom x 3; nnz_: atoms per molecule; ele_: NMol x MaxNAtom atomic numbers."""
        self.nmol = x_.shape[0]
        self.maxnatom = x_.shape[1]
        self.alg = 0 if alg_ is None else alg_
        self.x = x_.copy()
        self.nnz = np.asarray(nnz_).copy()
        self.nreal = np.asarray(nnz_).copy()
        self.ele = ele_
        self.sort = sort_
        self.pairs = None
        self.DoTriples = DoTriples_
        self.DoPerms = DoPerms_
        self.triples = None
        self.UpdateInterval = 1
        self.UpdateCounter = 0
        self.nlist = []
        for i in range(self.nmol):
            e = None if self.ele is None else self.ele[i, :self.nnz[i]]
            self.nlist.append(NeighborList(x_[i, :self.nnz[i]], DoTriples_, DoPerms_, e, self.alg, self.sort))

    @TMTiming("NLSetUpdate")
    def Update(self, x_, rcut_pairs=5.0, rcut_triples=5.0):
        self.x = x_.copy()
        if self.DoTriples:
            self.pairs, self.triples = self.buildPairsAndTriples(rcut_pairs, rcut_triples)
        else:
            self.pairs = self.buildPairs(rcut_pairs)

    def buildPairs(self, rcut=5.0):
        """(nnz pairs x 3) mol, I, J  uint64 (Neighbors.py:262-282)."""
        out = []
        for i, mol in enumerate(self.nlist):
            mol.Update(self.x[i, :self.nnz[i]], rcut, rcut, i, self.nreal[i])
            out.append(mol.pairs)
        return np.concatenate(out, axis=0) if out else np.zeros((0, 3), np.uint64)

    @TMTiming("SetbuildPairsAndTriples")
    def buildPairsAndTriples(self, rcut_pairs=5.0, rcut_triples=5.0):
        ps, ts = [], []
        for i, mol in enumerate(self.nlist):
            mol.DoTriples = True
            mol.Update(self.x[i, :self.nnz[i]], rcut_pairs, rcut_triples, i, nreal_=self.nreal[i])
            ps.append(mol.pairs)
            ts.append(mol.triples)
        return np.concatenate(ps, axis=0), np.concatenate(ts, axis=0)

    @TMTiming("buildPairsWithBothEleIndex")
    def buildPairsWithBothEleIndex(self, rcut=5.0, ele=None, sort_=False):
        """rows mol,i,j,e_i,e_j (Neighbors.py:323-342)."""
        trp = self.buildPairs(rcut).astype(np.int64)
        el = np.asarray(ele).reshape(-1)
        e1 = np.searchsorted(el, self.ele[trp[:, 0], trp[:, 1]])
        e2 = np.searchsorted(el, self.ele[trp[:, 0], trp[:, 2]])
        out = np.concatenate([trp, e1.reshape(-1, 1), e2.reshape(-1, 1)], axis=-1)
        if sort_:
            sw = out[:, 3] > out[:, 4]
            out[sw] = out[sw][:, [0, 2, 1, 4, 3]]
        return out

    def _tables(self, rcut_pairs, rcut_triples, ele):
        eng = _table_engine(ele)
        Zs = np.ascontiguousarray(self.ele, np.int32)
        return eng.pairs_triples_ele(self.x, Zs, np.asarray(self.nnz, np.int64), np.asarray(self.nreal, np.int64), rcut_pairs, rcut_triples)

    def buildPairsAndTriplesWithEleIndex(self, rcut_pairs=5.0, rcut_triples=5.0, ele=None, elep=None):
        """(P x 4) mol,I,J,L sorted by (mol,i,l,j); (T x 5) mol,I,J,K,L sorted by (mol,i,l,k,j); mil_jk; jk_max
        (Neighbors.py:344-423)."""
        if not self.sort:
            print("Warning! Triples need to be sorted")
        rad, ang, mil_j, mil_jk = self._tables(rcut_pairs, rcut_triples, ele)
        jk_max = float(np.max(mil_jk[:, 3])) if mil_jk.shape[0] else 0
        return rad.astype(np.float64), ang.astype(np.float64), mil_jk.astype(np.float64), jk_max

    @TMTiming("buildPairsAndTriplesWithEleIndexPeriodic")
    def buildPairsAndTriplesWithEleIndexPeriodic(self, rcut_pairs=5.0, rcut_triples=5.0, ele=None, elep=None):
        rad, ang, mil_j, mil_jk = self._tables(rcut_pairs, rcut_triples, ele)
        return rad.astype(np.float64), ang.astype(np.float64), mil_j.astype(np.float64), mil_jk.astype(np.float64)

    def buildPairsAndTriplesWithEleIndexLinear(self, rcut_pairs=5.0, rcut_triples=5.0, ele=None, elep=None):
        return self.buildPairsAndTriplesWithEleIndexPeriodic(rcut_pairs, rcut_triples, ele, elep)


class NeighborListSetWithImages(NeighborListSet):
    """Rows only for the first nreal_ atoms of each molecule; the rest are periodic images (Neighbors.py:473-489)."""

    def __init__(self, x_, nnz_, nreal_, DoTriples_=False, DoPerms_=False, ele_=None, alg_=None, sort_=False):
        NeighborListSet.__init__(self, x_, nnz_, DoTriples_, DoPerms_, ele_, alg_, sort_)
        self.nreal = np.asarray(nreal_)
