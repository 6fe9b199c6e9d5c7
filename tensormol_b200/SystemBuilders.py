"""Synthetic systems for benchmarks and tests (the reference's SystemBuilders/MolBuilders.py is a stub;
its water boxes come from files that are not shipped: samples/test_h2o.py:1723-1734)."""
from __future__ import annotations

import numpy as np


def water_box(nx, spacing=3.1044, seed=2, jitter=0.05):
    """nx^3 water molecules on a simple-cubic lattice of spacing `spacing` (3.1044 A -> 0.1002 atoms/A^3)
    with random rigid orientations: O-H 0.9572 A, HOH 104.52 deg, atom order H,H,O per molecule (as in the
    reference's datasets), then Gaussian jitter.  Returns (Z int32 [N], xyz [N,3] A, lattice [3,3])."""
    rng = np.random.default_rng(seed)
    th = np.deg2rad(104.52) / 2.0
    h1 = 0.9572 * np.array([np.sin(th), np.cos(th), 0.0])
    h2 = 0.9572 * np.array([-np.sin(th), np.cos(th), 0.0])
    nmol = nx ** 3
    q = rng.standard_normal((nmol, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((nmol, 3, 3))
    R[:, 0, 0] = a * a + b * b - c * c - d * d
    R[:, 0, 1] = 2 * (b * c - a * d)
    R[:, 0, 2] = 2 * (b * d + a * c)
    R[:, 1, 0] = 2 * (b * c + a * d)
    R[:, 1, 1] = a * a - b * b + c * c - d * d
    R[:, 1, 2] = 2 * (c * d - a * b)
    R[:, 2, 0] = 2 * (b * d - a * c)
    R[:, 2, 1] = 2 * (c * d + a * b)
    R[:, 2, 2] = a * a - b * b - c * c + d * d
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nx), indexing="ij")
    o = (np.stack([ii, jj, kk], axis=-1).reshape(-1, 3) + 0.5) * spacing
    xyz = np.empty((nmol, 3, 3))
    xyz[:, 0] = o + R @ h1
    xyz[:, 1] = o + R @ h2
    xyz[:, 2] = o
    xyz = xyz.reshape(-1, 3) + jitter * rng.standard_normal((3 * nmol, 3))
    Z = np.tile(np.array([1, 1, 8], np.int32), nmol)
    return Z, xyz, np.eye(3) * (nx * spacing)


def wrap_into_cell(xyz, lattice):
    """Lattice.ModuloLattice (Periodic.py:87-100) for callers that have no Lattice object yet."""
    L = np.asarray(lattice, np.float64)
    f = np.fmod(np.asarray(xyz, np.float64) @ np.linalg.inv(L), 1.0)
    f[f < 0.0] += 1.0
    return f @ L


def perturbed_molecule_batch(Z, xyz, n, sigma=0.05, seed=1):
    """n copies of one molecule with i.i.d. N(0, sigma) displacements (config C2)."""
    rng = np.random.default_rng(seed)
    xyzs = np.asarray(xyz, np.float64)[None] + sigma * rng.standard_normal((n,) + np.asarray(xyz).shape)
    Zs = np.tile(np.asarray(Z, np.int32)[None], (n, 1))
    return Zs, xyzs
