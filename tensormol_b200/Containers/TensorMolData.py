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
        self.ScratchState = None
        self.TestRatio = PARAMS["TestRatio"]
        self.NTrain = 0
        self.NTest = 0

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

    # ---- training-style batches (reference TensorMolData.py:1679-1745, 1860-1904) ---------------------------------------
    def LoadData(self):
        """Shuffles the set's molecules (Python's `random`, as the reference) and returns padded arrays
        xyzs, Zs, Elabels (properties["atomization"]), Dlabels (properties["dipole"] * AUPERDEBYE), natom[, grads]."""
        import random
        if self.dig is not None and getattr(self.dig, "OType", "EnergyAndDipole") != "EnergyAndDipole":
            raise Exception("Output Type is not implemented yet")
        random.shuffle(self.set.mols)
        xyzs, Zs, natom = self.raw_arrays()
        Elabels = np.zeros(self.Nmols, dtype=np.float64)
        Dlabels = np.zeros((self.Nmols, 3), dtype=np.float64)
        grads = np.zeros((self.Nmols, self.MaxNAtoms, 3), dtype=np.float64) if self.HasGrad else None
        for i, mol in enumerate(self.set.mols):
            Elabels[i] = mol.properties["atomization"]
            Dlabels[i] = np.asarray(mol.properties["dipole"]) * AUPERDEBYE   # noqa: F405
            if self.HasGrad:
                grads[i][:mol.NAtoms()] = mol.properties["gradients"]
        if self.HasGrad:
            return xyzs, Zs, Elabels, Dlabels, natom, grads
        return xyzs, Zs, Elabels, Dlabels, natom

    def LoadDataToScratch(self, tformer=None):
        """Loads once, then splits: the last int(TestRatio * N) molecules of the shuffled order are the test cases."""
        if self.ScratchState == 1:
            return
        data = self.LoadData()
        self.xyzs, self.Zs, self.Elabels, self.Dlabels, self.natom = data[:5]
        self.grads = data[5] if self.HasGrad else None
        self.NTestMols = int(self.TestRatio * self.Zs.shape[0])
        self.LastTrainMol = int(self.Zs.shape[0] - self.NTestMols)
        self.NTrain = self.LastTrainMol
        self.NTest = self.NTestMols
        self.test_ScratchPointer = self.LastTrainMol
        self.ScratchPointer = 0
        self.ScratchState = 1

    def _batch_window(self, ncases, test):
        """The reference's batch pointers: training batches wrap to 0 when pointer + ncases >= NTrain (so the last
        partial AND the last exactly-fitting batch are skipped), test batches wrap to LastTrainMol when pointer + ncases
        would pass the end of the data."""
        if self.ScratchState != 1:
            self.LoadDataToScratch()      # the reference leaves this to the network instance (TFMolInstanceDirect.py)
        if not test:
            if ncases > self.NTrain:
                raise Exception("Insufficent training data to fill a batch" + str(self.NTrain) + " vs " + str(ncases))
            if self.ScratchPointer + ncases >= self.NTrain:
                self.ScratchPointer = 0
            self.ScratchPointer += ncases
            return slice(self.ScratchPointer - ncases, self.ScratchPointer)
        if ncases > self.NTest:
            raise Exception("Insufficent training data to fill a batch" + str(self.NTest) + " vs " + str(ncases))
        if self.test_ScratchPointer + ncases > self.Zs.shape[0]:
            self.test_ScratchPointer = self.LastTrainMol
        self.test_ScratchPointer += ncases
        return slice(self.test_ScratchPointer - ncases, self.test_ScratchPointer)

    def _batch(self, ncases, test):
        from ..ForceModifiers.Neighbors import NeighborListSet
        w = self._batch_window(ncases, test)
        xyzs, Zs, natom = self.xyzs[w], self.Zs[w], self.natom[w]
        NL = NeighborListSet(xyzs, natom, True, True, Zs, sort_=True)
        rad_p_ele, ang_t_elep, mil_jk, jk_max = NL.buildPairsAndTriplesWithEleIndex(self.Rr_cut, self.Ra_cut, self.ele, self.elep)
        NLEE = NeighborListSet(xyzs, natom, False, False, None)
        rad_eep = NLEE.buildPairs(self.Ree_cut)
        out = [xyzs, Zs, self.Elabels[w], self.Dlabels[w]]
        if self.HasGrad:
            out.append(self.grads[w])
        return out + [rad_p_ele, ang_t_elep, rad_eep, mil_jk, 1.0 / natom]

    def GetTrainBatch(self, ncases):
        """[xyzs, Zs, Elabels, Dlabels, (grads,) rad_p_ele, ang_t_elep, rad_eep, mil_jk, 1/natom] (TensorMolData.py:1860-1881);
        the neighbour tables are built on the GPU (NeighborListSet -> tm_pairs_triples_ele)."""
        return self._batch(ncases, False)

    def GetTestBatch(self, ncases):
        return self._batch(ncases, True)
