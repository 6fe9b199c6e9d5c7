"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's host-side index logic.

Nothing in the product package (`tensormol_b200/`) may import this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do, and
there only as the checker or as the reported CPU baseline.

Parity pinning: the reference's tests hold NO golden vectors for this path.
What pins this restatement is executable reference code: `MolEmb.Make_NListNaive`
(C_API/MolEmb.cpp:1180-1247) compiled from /root/reference into oracle/_ref by
oracle/Makefile; `tests/test_oracle.py` checks `make_nlist_naive` below against it.

Each function cites the reference file:line it restates ("NBR" =
TensorMol/ForceModifiers/Neighbors.py, "PER" = TensorMol/ForceModifiers/Periodic.py).
"""
from __future__ import annotations

import numpy as np

try:  # scipy is only a candidate generator; the accept test below is the reference's own.
    from scipy.spatial import cKDTree
except Exception:  # pragma: no cover
    cKDTree = None


# --------------------------------------------------------------------------------------
# Neighbour search  (C_API/MolEmb.cpp:1180-1247  Make_NListNaive)
# --------------------------------------------------------------------------------------
def _accept(x, I, J, rng):
    """The reference accept test, evaluated with the same operation order and roundings:
    dij = sqrt(dx*dx+dy*dy+dz*dz) + 1e-13 ; keep if dij < rng   (MolEmb.cpp:1213-1218).
    numpy evaluates each elementwise op with one IEEE rounding (no FMA contraction), which is
    what gcc -O2 emits for x86-64 (setup.py:22 passes no -march / -ffast-math)."""
    dx = x[I, 0] - x[J, 0]
    dy = x[I, 1] - x[J, 1]
    dz = x[I, 2] - x[J, 2]
    dij = np.sqrt(dx * dx + dy * dy + dz * dz) + 0.0000000000001
    return dij < rng


def candidate_pairs(x, rng, nreal):
    """All unordered index pairs (I<J) with at least one index < nreal that pass the accept
    test.  Returns (I, J) int64 arrays, sorted by (I, J)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    if n < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    if cKDTree is not None and n > 64:
        tree = cKDTree(x)
        if nreal < n:
            # only pairs touching a real atom matter
            treer = cKDTree(x[:nreal])
            sp = treer.sparse_distance_matrix(tree, rng + 1e-6, output_type="coo_matrix")
            I = sp.row.astype(np.int64)
            J = sp.col.astype(np.int64)
            keep = I != J
            I, J = I[keep], J[keep]
            lo = np.minimum(I, J)
            hi = np.maximum(I, J)
            key = np.unique(lo * n + hi)
            I, J = key // n, key % n
        else:
            pr = tree.query_pairs(rng + 1e-6, output_type="ndarray")
            I = pr[:, 0].astype(np.int64)
            J = pr[:, 1].astype(np.int64)
    else:
        I, J = np.triu_indices(n, 1)
        I = I.astype(np.int64)
        J = J.astype(np.int64)
        keep = (I < nreal) | (J < nreal)
        I, J = I[keep], J[keep]
    ok = _accept(x, I, J, rng)
    I, J = I[ok], J[ok]
    order = np.lexsort((J, I))
    return I[order], J[order]


def nlist_csr(x, rng, nreal, do_perms):
    """CSR form of Make_NListNaive: for every i<nreal the SORTED neighbour indices.
    DoPerms=0: row min(I,J) gets max(I,J) only (MolEmb.cpp:1220-1231); DoPerms=1: the reverse
    entry is added when the larger index is also < nreal."""
    I, J = candidate_pairs(x, rng, nreal)
    # I<J always; min index is I and must be < nreal (one of the two is real, so I is)
    ci, cj = I, J
    if do_perms:
        m = J < nreal
        ci = np.concatenate([I, J[m]])
        cj = np.concatenate([J, I[m]])
    order = np.lexsort((cj, ci))
    ci, cj = ci[order], cj[order]
    counts = np.bincount(ci, minlength=nreal)[:nreal]
    offsets = np.zeros(nreal + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    return offsets, cj.astype(np.int64)


def make_nlist_naive(x, rng, nreal, do_perms):
    """list-of-lists like the reference (each row sorted; the reference's row order is sweep
    order, parity is on sets)."""
    off, idx = nlist_csr(x, rng, nreal, do_perms)
    return [idx[off[i]:off[i + 1]].tolist() for i in range(nreal)]


# --------------------------------------------------------------------------------------
# Pair / triple assembly   (NBR:75-115, 117-201)
# --------------------------------------------------------------------------------------
def build_pairs(x, rng, nreal, do_perms, molind=None):
    """NeighborList.buildPairs (NBR:75-115): rows [mol,] i, j   (uint64)."""
    off, idx = nlist_csr(x, rng, nreal, do_perms)
    i = np.repeat(np.arange(nreal, dtype=np.int64), np.diff(off))
    if molind is None:
        return np.stack([i, idx], axis=1).astype(np.uint64).reshape(-1, 2)
    m = np.full_like(i, molind)
    return np.stack([m, i, idx], axis=1).astype(np.uint64).reshape(-1, 3)


def build_triples(x, rng, nreal, do_perms, ele, molind=None):
    """Triples of NeighborList.buildPairsAndTriples (NBR:160-198): for centre i all j,k in the
    rcut_triples list with k>j (atom index), then the atom with the smaller atomic number goes
    first (`ele[j] > ele[k]` swap, NBR:178-180).  Rows [mol,] i, j, k (uint64)."""
    off, idx = nlist_csr(x, rng, nreal, do_perms)
    cnt = np.diff(off)
    rows_i, rows_j, rows_k = [], [], []
    # group centres by neighbour count so each group is one vectorised triu
    for c in np.unique(cnt):
        if c < 2:
            continue
        centres = np.where(cnt == c)[0]
        a, b = np.triu_indices(int(c), 1)
        nb = idx[off[centres][:, None] + np.arange(c)[None, :]]  # (ncent, c) sorted ascending
        j = nb[:, a]
        k = nb[:, b]  # k>j because rows are sorted
        rows_i.append(np.repeat(centres, a.size))
        rows_j.append(j.reshape(-1))
        rows_k.append(k.reshape(-1))
    if rows_i:
        i = np.concatenate(rows_i)
        j = np.concatenate(rows_j)
        k = np.concatenate(rows_k)
    else:
        i = j = k = np.zeros(0, np.int64)
    if ele is not None and i.size:
        swap = ele[j] > ele[k]
        j, k = np.where(swap, k, j), np.where(swap, j, k)
    order = np.lexsort((k, j, i))
    i, j, k = i[order], j[order], k[order]
    if molind is None:
        return np.stack([i, j, k], axis=1).astype(np.uint64).reshape(-1, 3)
    m = np.full_like(i, molind)
    return np.stack([m, i, j, k], axis=1).astype(np.uint64).reshape(-1, 4)


def set_build_pairs(xyzs, nnz, nreal, rng, do_perms):
    """NeighborListSet.buildPairs (NBR:262-282): concatenation over molecules, rows mol,i,j."""
    out = [build_pairs(xyzs[m, :nnz[m]], rng, int(nreal[m]), do_perms, m) for m in range(xyzs.shape[0])]
    return np.concatenate(out, axis=0) if out else np.zeros((0, 3), np.uint64)


def set_build_pairs_and_triples(xyzs, nnz, nreal, Zs, rr, ra, do_perms=True):
    """NeighborListSet.buildPairsAndTriples (NBR:285-321)."""
    ps, ts = [], []
    for m in range(xyzs.shape[0]):
        x = xyzs[m, :nnz[m]]
        ps.append(build_pairs(x, rr, int(nreal[m]), do_perms, m))
        ts.append(build_triples(x, ra, int(nreal[m]), do_perms, None if Zs is None else Zs[m, :nnz[m]], m))
    return np.concatenate(ps, axis=0), np.concatenate(ts, axis=0)


def _slot_index(keys):
    """Running index inside each run of equal consecutive keys (NBR:386-420, 440-465)."""
    n = keys.shape[0]
    if n == 0:
        return np.zeros(0)
    new = np.ones(n, bool)
    new[1:] = np.any(keys[1:] != keys[:-1], axis=1)
    start = np.maximum.accumulate(np.where(new, np.arange(n), 0))
    return (np.arange(n) - start).astype(np.float64)


def build_pairs_and_triples_with_ele_index(xyzs, nnz, nreal, Zs, rr, ra, ele, elep):
    """NeighborListSet.buildPairsAndTriplesWithEleIndex (NBR:344-423).
    Returns trpE_sorted (P,4) [mol,i,j,l], trtE_sorted (T,5) [mol,i,j,k,l], mil_jk (T,4), jk_max.
    dtype float64 like the reference (uint64 (+) int64 concat promotes, Q16)."""
    trp, trt = set_build_pairs_and_triples(xyzs, nnz, nreal, Zs, rr, ra, True)
    ele = np.asarray(ele).reshape(-1)
    elep = np.asarray(elep).reshape(-1, 2)
    trp_i = trp.astype(np.int64)
    trt_i = trt.astype(np.int64)
    Zj = Zs[trp_i[:, 0], trp_i[:, 2]]
    pair_index = np.searchsorted(ele, Zj) if ele.size else np.zeros(0, np.int64)
    # the reference silently DROPS rows whose element is not in `ele` via np.where(...)[1] only if
    # all rows match; we require membership (as the reference effectively does).
    assert np.all(ele[pair_index] == Zj), "neighbour element not in eles"
    Z1 = Zs[trt_i[:, 0], trt_i[:, 2]]
    Z2 = Zs[trt_i[:, 0], trt_i[:, 3]]
    lo = np.minimum(Z1, Z2)
    hi = np.maximum(Z1, Z2)
    trip_index = np.zeros(trt_i.shape[0], np.int64)
    for l, (a, b) in enumerate(elep):
        trip_index[(lo == min(a, b)) & (hi == max(a, b))] = l
    trpE = np.concatenate([trp_i, pair_index.reshape(-1, 1)], axis=-1).astype(np.float64)
    trtE = np.concatenate([trt_i, trip_index.reshape(-1, 1)], axis=-1).astype(np.float64)
    si = np.lexsort((trpE[:, 2], trpE[:, 3], trpE[:, 1], trpE[:, 0]))          # NBR:380
    trpE_sorted = trpE[si]
    si = np.lexsort((trtE[:, 2], trtE[:, 3], trtE[:, 4], trtE[:, 1], trtE[:, 0]))  # NBR:382
    trtE_sorted = trtE[si]
    mil_jk = np.zeros((trt.shape[0], 4))
    if trt.shape[0] == 0:
        return trpE_sorted, trtE_sorted, mil_jk, 0
    mil_jk[:, [0, 1, 2]] = trtE_sorted[:, [0, 1, 4]]
    mil_jk[:, 3] = _slot_index(trtE_sorted[:, [0, 1, 4]])
    return trpE_sorted, trtE_sorted, mil_jk, np.max(mil_jk[:, 3])


def build_pairs_and_triples_with_ele_index_periodic(xyzs, nnz, nreal, Zs, rr, ra, ele, elep):
    """NeighborListSet.buildPairsAndTriplesWithEleIndexPeriodic (NBR:425-467) (also `...Linear`, :469)."""
    trpE_sorted, trtE_sorted, mil_jk, _ = build_pairs_and_triples_with_ele_index(xyzs, nnz, nreal, Zs, rr, ra, ele, elep)
    mil_j = np.zeros((trpE_sorted.shape[0], 4))
    if trpE_sorted.shape[0]:
        mil_j[:, [0, 1, 2]] = trpE_sorted[:, [0, 1, 3]]
        mil_j[:, 3] = _slot_index(trpE_sorted[:, [0, 1, 3]])
    return trpE_sorted, trtE_sorted, mil_j, mil_jk


def build_pairs_with_both_ele_index(xyzs, nnz, nreal, Zs, rng, ele, do_perms):
    """NeighborListSet.buildPairsWithBothEleIndex (NBR:323-342), sort_=False: rows mol,i,j,e_i,e_j."""
    trp = set_build_pairs(xyzs, nnz, nreal, rng, do_perms).astype(np.int64)
    ele = np.asarray(ele).reshape(-1)
    e1 = np.searchsorted(ele, Zs[trp[:, 0], trp[:, 1]])
    e2 = np.searchsorted(ele, Zs[trp[:, 0], trp[:, 2]])
    return np.concatenate([trp, e1.reshape(-1, 1), e2.reshape(-1, 1)], axis=-1)


# --------------------------------------------------------------------------------------
# Lattice   (PER:12-168)
# --------------------------------------------------------------------------------------
def lattice_min_diameter(lattice):
    """Lattice.__init__ (PER:19-25): 2*min distance from the cell centre to the 14 face points."""
    L = np.asarray(lattice, np.float64)
    centre = (L[0] + L[1] + L[2]) / 2.0
    lfp = np.zeros((14, 3))
    lfp[0], lfp[1], lfp[2] = L[0], L[1], L[2]
    lfp[3] = L[0] + L[1]
    lfp[4] = L[0] + L[2]
    lfp[5] = L[1] + L[2]
    lfp[6] = L[0] + L[1] + L[2]
    lfp[7] = 0.0
    lfp[8] = 0.5 * (L[0] + L[1])
    lfp[9] = 0.5 * (L[2] + L[1])
    lfp[10] = 0.5 * (L[0] + L[2])
    lfp[11] = 0.5 * (L[0] + L[1]) + L[2]
    lfp[12] = 0.5 * (L[2] + L[1]) + L[0]
    lfp[13] = 0.5 * (L[0] + L[2]) + L[1]
    return 2.0 * np.min(np.linalg.norm(lfp - centre[None, :], axis=1))


def modulo_lattice(lattice, crds):
    """Lattice.ModuloLattice (PER:87-100) with InLat/FromLat (PER:74-86)."""
    L = np.asarray(lattice, np.float64)
    latmet = np.linalg.inv(np.dot(L, L.T))
    tmp = np.dot(crds, np.dot(L.T, latmet))
    fpart = np.fmod(tmp, 1.0)
    revs = np.where(fpart < 0.0)
    fpart[revs] = 1.0 + fpart[revs]
    return np.dot(fpart, L)


def tess_lattice(lattice, atoms, coords, rng):
    """Lattice.TessLattice (PER:131-168): real atoms first, then (2 ntess+1)^3-1 images in
    i,j,k loop order."""
    L = np.asarray(lattice, np.float64)
    dmin = lattice_min_diameter(L)
    ntess = int(rng / dmin) + 1 if rng > dmin else 1
    natom = atoms.shape[0]
    nimages = (2 * ntess + 1) ** 3
    newAtoms = np.zeros(nimages * natom, dtype=np.uint8)
    newCoords = np.zeros((nimages * natom, 3))
    newAtoms[:natom] = atoms
    newCoords[:natom] = coords
    ind = 1
    for i in range(-ntess, ntess + 1):
        for j in range(-ntess, ntess + 1):
            for k in range(-ntess, ntess + 1):
                if i == 0 and j == 0 and k == 0:
                    continue
                newAtoms[ind * natom:(ind + 1) * natom] = atoms
                newCoords[ind * natom:(ind + 1) * natom] = coords + i * L[0] + j * L[1] + k * L[2]
                ind += 1
    return newAtoms, newCoords
