"""General purpose molecule container with the members the drivers use (reference: Containers/Mol.py)."""
from __future__ import annotations

import errno
import os

import numpy as np

from ..Util import *   # noqa: F401,F403


class Mol:
    """atoms: uint8 atomic numbers; coords: (N,3) float64 Angstrom; properties: dict."""

    def __init__(self, atoms_=np.zeros(1, dtype=np.uint8), coords_=np.zeros(shape=(1, 1), dtype=np.float64)):
        self.atoms = np.array(atoms_).copy()
        self.coords = np.array(coords_, dtype=np.float64).copy()
        self.properties = {}
        self.name = None

    # ---- queries -----------------------------------------------------------------------------
    def AtomTypes(self):
        return np.unique(self.atoms)

    def NEles(self):
        return len(self.AtomTypes())

    def NAtoms(self):
        return self.atoms.shape[0]

    def NumOfAtomsE(self, e):
        return int(np.sum(self.atoms == e))

    def Num_of_Heavy_Atom(self):
        return int(np.sum(self.atoms != 1))

    def IsIsomer(self, other):
        return np.array_equal(np.sort(self.atoms), np.sort(other.atoms))

    def Center(self, CenterOf="Atom", MomentOrder=1.):
        if CenterOf == "Mass":
            m = np.array([ATOMICMASSES[z - 1] for z in self.atoms])   # noqa: F405
            return np.einsum("ax,a->x", self.coords, m) / np.sum(m)
        return np.average(self.coords, axis=0)

    def Clean(self):
        pass

    # ---- modifiers ---------------------------------------------------------------------------
    def Distort(self, disp=0.38, movechance=.20):
        """Randomly displace coordinates (reference Mol.py:160-177): each coordinate moves with probability
        `movechance` (Python's `random.uniform`) by normal(0, disp) draws (numpy's generator) that are ADDED one after
        another, up to 100 of them, until the moved atom is farther than 0.35 Angstrom from every other atom -- the
        reference retries on the same array, so rejected trial displacements accumulate."""
        import random
        n = self.NAtoms()
        for i in range(n):
            for j in range(3):
                if random.uniform(0, 1) < movechance:
                    for _ in range(100):
                        self.coords[i, j] += np.random.normal(0.0, disp)
                        d = np.linalg.norm(self.coords - self.coords[i], axis=1)
                        d[i] = 1.0
                        if np.min(d) > 0.35:
                            break

    def Transform(self, ltransf, center=np.array([0.0, 0.0, 0.0])):
        self.coords = np.einsum("ij,kj->ki", ltransf, self.coords - center) + center

    def AlignAtoms(self, m):
        """Reorder the atoms of m to minimise the distance to my atoms, element by element (greedy)."""
        assert self.NAtoms() == m.NAtoms(), "Number of atoms do not match"
        used = np.zeros(m.NAtoms(), bool)
        order = []
        for i in range(self.NAtoms()):
            cand = np.where((m.atoms == self.atoms[i]) & (~used))[0]
            d = np.linalg.norm(m.coords[cand] - self.coords[i], axis=1)
            j = cand[int(np.argmin(d))]
            used[j] = True
            order.append(j)
        m.atoms = m.atoms[order]
        m.coords = m.coords[order]

    # ---- xyz I/O (reference Mol.py:268-359) ----------------------------------------------------
    def ParseProperties(self, s_):
        t = s_.split("Comment:")
        tore = {}
        if len(t) < 2:
            return tore
        for prop in t[1].split(";;;"):
            s = prop.split()
            if len(s) < 2:
                continue
            if s[0] == 'energy':
                tore["energy"] = float(s[1])
            elif s[0] == 'Lattice':
                try:
                    tore["Lattice"] = np.array([float(v) for v in prop.replace("Lattice", "").replace("[", " ").replace("]", " ").split()]).reshape((3, 3))
                except Exception:
                    pass
        return tore

    def PropertyString(self):
        tore = ""
        for prop in self.properties.keys():
            try:
                if prop == "Lattice":
                    tore += ";;;" + prop + " " + " ".join(repr(float(v)) for v in np.asarray(self.properties[prop]).reshape(-1))
                else:
                    tore += ";;;" + prop + " " + str(self.properties[prop])
            except Exception:
                pass
        return tore

    def FromXYZString(self, string):
        lines = string.split("\n")
        natoms = int(lines[0].split()[0])
        if len(lines[1].split()) > 1:
            try:
                self.properties = self.ParseProperties(lines[1])
            except Exception as Ex:
                print("Problem with properties", Ex)
        self.atoms = np.zeros(natoms, dtype=np.uint8)
        self.coords = np.zeros((natoms, 3))
        for i in range(natoms):
            line = lines[i + 2].split()
            if len(line) == 0:
                return
            self.atoms[i] = AtomicNumber(line[0]) if not line[0].isdigit() else int(line[0])   # noqa: F405
            for k in range(3):
                try:
                    self.coords[i, k] = float(line[k + 1])
                except ValueError:
                    self.coords[i, k] = scitodeci(line[k + 1])   # noqa: F405

    def ReadGDB9(self, path, filename=None):
        with open(path) as f:
            self.FromXYZString(f.read())

    def __str__(self, wprop=False):
        natom = self.atoms.shape[0]
        lines = str(natom) + "\nComment: " + (self.PropertyString() if wprop else "") + "\n"
        body = [AtomicSymbol(self.atoms[i]) + "   " + str(self.coords[i][0]) + "  " + str(self.coords[i][1]) + "  " + str(self.coords[i][2])   # noqa: F405
                for i in range(natom)]
        return lines + "\n".join(body)

    def __repr__(self):
        return self.__str__()

    def WriteXYZfile(self, fpath=".", fname="mol", mode="a", wprop=False):
        full = fpath + "/" + fname + ".xyz"
        d = os.path.dirname(full)
        if d and not os.path.exists(d):
            try:
                os.makedirs(d)
            except OSError as exc:
                if exc.errno != errno.EEXIST:
                    raise
        with open(full, mode) as f:
            for line in self.__str__(wprop).split("\n"):
                f.write(line + "\n")
