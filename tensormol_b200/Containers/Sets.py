"""Molecule sets (reference: Containers/Sets.py:17-298): xyz read/write, element filters, pickle save/load."""
from __future__ import annotations

import itertools
import pickle

import numpy as np

from ..Util import *   # noqa: F401,F403
from .Mol import Mol


class MSet:
    def __init__(self, name_="gdb9", path_="./datasets/", center_=True):
        self.mols = []
        self.path = path_
        self.name = name_
        self.suffix = ".pdb"
        self.center = center_

    def Save(self, filename=None):
        filename = self.name if filename is None else filename
        with open(self.path + filename + self.suffix, "wb") as f:
            pickle.dump(self.__dict__, f, protocol=pickle.HIGHEST_PROTOCOL)

    def Load(self, filename=None):
        filename = self.name if filename is None else filename
        with open(self.path + filename + self.suffix, "rb") as f:
            self.__dict__.update(pickle.load(f))

    def CenterSet(self):
        for mol in self.mols:
            mol.coords -= mol.Center()

    def ReadXYZ(self, filename=None, xyz_type='mol'):
        """Reads XYZs concatenated into a single file as a molset (Sets.py:208-233)."""
        filename = self.name if filename is None else filename
        with open(self.path + filename + ".xyz", "r") as f:
            txts = f.readlines()
        line = 0
        while line < len(txts):
            if txts[line].strip() and all(x.isdigit() for x in txts[line].split()):
                nlines = int(txts[line].split()[0])
                m = Mol()
                m.FromXYZString(''.join(txts[line:line + nlines + 2]))
                m.name = str(txts[line + 1])
                m.properties["set_name"] = self.name
                self.mols.append(m)
                line += nlines + 2
            else:
                line += 1
        if self.center:
            self.CenterSet()
        LOGGER.debug("Read " + str(len(self.mols)) + " molecules from XYZ")   # noqa: F405

    def WriteXYZ(self, filename=None):
        filename = self.name if filename is None else filename
        for mol in self.mols:
            mol.WriteXYZfile(self.path, filename)

    def pop(self, ntopop):
        for _ in range(ntopop):
            self.mols.pop()

    def OnlyWithElements(self, allowed_eles):
        self.mols = [m for m in self.mols if set(list(m.atoms)).issubset(allowed_eles)]
        for i in allowed_eles:
            self.name += "_" + str(i)

    def OnlyAtoms(self, allowed_eles):
        for mol in self.mols:
            keep = [i for i, a in enumerate(mol.atoms) if a in allowed_eles]
            mol.atoms = mol.atoms[keep]
            mol.coords = mol.coords[keep]

    def AppendSet(self, b):
        self.mols = self.mols + b.mols

    def NAtoms(self):
        return int(sum(m.NAtoms() for m in self.mols))

    def MaxNAtoms(self):
        return int(max(m.NAtoms() for m in self.mols))

    def AtomTypes(self):
        types = np.array([], dtype=np.uint8)
        for m in self.mols:
            types = np.union1d(types, m.AtomTypes())
        return types

    def BondTypes(self):
        return np.asarray([x for x in itertools.product(self.AtomTypes().tolist(), repeat=2)])   # noqa: F405

    def Clean(self):
        for m in self.mols:
            m.Clean()
