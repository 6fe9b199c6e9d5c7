"""Training-data holder of the `_Direct` BP+EE path.  The manager only needs .eles, .MaxNAtoms, .name,
.dig and the batch providers (reference Containers/TensorMolData.py:1222-1245, 1849-1904)."""
from __future__ import annotations

import numpy as np

from ..Util import *   # noqa: F401,F403


class TensorMolData_BP_Direct_EE_WithEle:
    def __init__(self, MSet_=None, Dig_=None, Name_=None, order_=3, num_indis_=1, type_="mol", WithGrad_=False):
        self.set = MSet_
        self.dig = Dig_
        self.order = order_
        self.num_indis = num_indis_
        self.type = type_
        self.HasGrad = WithGrad_
        self.eles = []
        self.MaxNAtoms = None
        self.Nmols = 0
        if MSet_ is not None:
            self.eles = sorted(int(e) for e in MSet_.AtomTypes())
            self.MaxNAtoms = int(np.max([m.NAtoms() for m in self.set.mols]))
            self.Nmols = len(self.set.mols)
        self.name = Name_ if Name_ is not None else (self.set.name if self.set is not None else "")
        self.Rr_cut = PARAMS["AN1_r_Rc"]
        self.Ra_cut = PARAMS["AN1_a_Rc"]
        self.Ree_cut = PARAMS["EECutoffOff"]
        self.ele = None
        self.elep = None
        self.ScratchPointer = 0

    def AtomTypes(self):
        return np.asarray(self.eles)

    def raw_arrays(self):
        """Padded (xyzs, Zs, natom) of the whole set."""
        n = len(self.set.mols)
        xyzs = np.zeros((n, self.MaxNAtoms, 3))
        Zs = np.zeros((n, self.MaxNAtoms), np.int32)
        natom = np.zeros(n, np.int32)
        for i, mol in enumerate(self.set.mols):
            xyzs[i, :mol.NAtoms()] = mol.coords
            Zs[i, :mol.NAtoms()] = mol.atoms
            natom[i] = mol.NAtoms()
        return xyzs, Zs, natom

    def GetTrainBatch(self, ncases):
        """[xyzs, Zs, Elabels, Dlabels, (grads,) rad_p_ele, ang_t_elep, rad_eep, mil_jk, 1/natom] as in the
        reference (TensorMolData.py:1860-1881); labels come from mol.properties when present."""
        from ..ForceModifiers.Neighbors import NeighborListSet
        xyzs, Zs, natom = self.raw_arrays()
        n = xyzs.shape[0]
        if ncases > n:
            raise Exception("Insufficent training data to fill a batch" + str(n) + " vs " + str(ncases))
        if self.ScratchPointer + ncases > n:
            self.ScratchPointer = 0
        sl = slice(self.ScratchPointer, self.ScratchPointer + ncases)
        self.ScratchPointer += ncases
        xyzs, Zs, natom = xyzs[sl], Zs[sl], natom[sl]
        mols = self.set.mols[sl]
        El = np.array([m.properties.get("atomization", m.properties.get("energy", 0.0)) for m in mols], np.float64)
        Dl = np.array([np.asarray(m.properties.get("dipole", np.zeros(3)), np.float64) for m in mols])
        NL = NeighborListSet(xyzs, natom, True, True, Zs, sort_=True)
        rad_p_ele, ang_t_elep, mil_jk, jk_max = NL.buildPairsAndTriplesWithEleIndex(self.Rr_cut, self.Ra_cut, self.ele, self.elep)
        NLEE = NeighborListSet(xyzs, natom, False, False, None)
        rad_eep = NLEE.buildPairs(self.Ree_cut)
        out = [xyzs, Zs, El, Dl]
        if self.HasGrad:
            g = np.zeros_like(xyzs)
            for i, m in enumerate(mols):
                if "gradients" in m.properties:
                    g[i, :m.NAtoms()] = m.properties["gradients"]
            out.append(g)
        return out + [rad_p_ele, ang_t_elep, rad_eep, mil_jk, 1.0 / natom]

    GetTestBatch = GetTrainBatch
