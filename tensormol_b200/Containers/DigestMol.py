"""MolDigester: on the `_Direct` path only its name / output type / element list are consumed
(reference Containers/DigestMol.py:11-24); the descriptors themselves are computed on the GPU."""
from __future__ import annotations


class MolDigester:
    def __init__(self, eles_, name_="Coulomb", OType_="FragEnergy", SensRadius_=6):
        self.name = name_
        self.OType = OType_
        self.lshape = None
        self.eshape = None
        self.egshape = None
        self.SensRadius = SensRadius_
        self.eles = eles_
        self.neles = len(eles_)
        self.ngrid = 5
        self.nsym = self.neles + (self.neles + 1) * self.neles
