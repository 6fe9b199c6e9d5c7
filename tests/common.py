"""Shared helpers for the tests: synthetic water boxes and tolerances.

Tolerances are BASELINE.json's north_star: neighbour pair/triple sets bit-exact after sorting; fp32
descriptors within 1e-5 relative; energies within 1e-5 relative; forces within 1e-4 Hartree/Bohr max-abs.
"""
import numpy as np

BOHRPERA = 1.889725989
DESC_RTOL = 1e-5          # per entry
DESC_ABS_FLOOR_ULPS = 4.0  # + this many fp32 ulps of the row's largest entry (absolute floor)
ENERGY_RTOL = 1e-5
FORCE_ATOL_HA_BOHR = 1e-4


def grad_ha_bohr(g_ha_per_A):
    return np.asarray(g_ha_per_A) / BOHRPERA


def water_box(nx, spacing=3.1044, seed=2, jitter=0.05):
    """nx^3 water molecules on a simple-cubic lattice with random rigid orientations (config C3):
    O-H 0.9572 A, HOH 104.52 deg, atom order H,H,O per molecule, Gaussian jitter."""
    rng = np.random.default_rng(seed)
    th = np.deg2rad(104.52) / 2.0
    h1 = 0.9572 * np.array([np.sin(th), np.cos(th), 0.0])
    h2 = 0.9572 * np.array([-np.sin(th), np.cos(th), 0.0])
    xyz, Z = [], []
    for i in range(nx):
        for j in range(nx):
            for k in range(nx):
                q = rng.standard_normal(4)
                q /= np.linalg.norm(q)
                a, b, c, d = q
                R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                              [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                              [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])
                o = (np.array([i, j, k]) + 0.5) * spacing
                xyz += [o + R @ h1, o + R @ h2, o]
                Z += [1, 1, 8]
    xyz = np.array(xyz) + jitter * rng.standard_normal((len(Z), 3))
    L = nx * spacing
    return np.array(Z, np.int32), xyz, np.eye(3) * L


def sets_equal_csr(off_a, idx_a, off_b, idx_b):
    """Row-wise set equality of two CSR neighbour lists."""
    if not np.array_equal(off_a, off_b):
        return False
    a = idx_a.copy()
    b = idx_b.copy()
    for i in range(len(off_a) - 1):
        a[off_a[i]:off_a[i + 1]].sort()
        b[off_b[i]:off_b[i + 1]].sort()
    return np.array_equal(a, b)


def sort_rows_csr(off, idx):
    out = idx.copy()
    for i in range(len(off) - 1):
        out[off[i]:off[i + 1]].sort()
    return out
