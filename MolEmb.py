"""Top-level `MolEmb` module: the name the reference's CPython extension is imported under (`import MolEmb`,
`from MolEmb import *`; C_API/MolEmb.cpp:2257-2271, imported by Neighbors.py:21, Periodic.py, Mol.py:5 and the sample
scripts). Re-exports the B200 drop-in (tensormol_b200/MolEmb.py: neighbour search on the GPU through libtmolb200,
no CPU fallback)."""
from tensormol_b200.MolEmb import *   # noqa: F401,F403
from tensormol_b200.MolEmb import (CountInRange, GetRDF_Bin, Make_DistMat, Make_DistMat_ForReal, Make_NListLinear,  # noqa: F401
                                   Make_NListNaive, nlist_csr)
