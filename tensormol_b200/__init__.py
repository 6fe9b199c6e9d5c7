"""tensormol_b200: B200-native energy/force path of TensorMol-0.1 behind the reference's Python API surface.

    from tensormol_b200 import *      # what `from TensorMol import *` gives in the reference, for the hot path

Layout: csrc/ (CUDA kernels + C-ABI, libtmolb200.so), engine.py (ctypes handle), and the host-side mirror of the
reference interface (Containers, ForceModifiers, TFNetworks, Simulations, Math, MolEmb).
"""
from .Util import *                         # noqa: F401,F403  PARAMS, LOGGER, constants, TMTiming
from .PhysicalData import *                 # noqa: F401,F403
from .Containers import *                   # noqa: F401,F403  Mol, MSet, MolDigester, TensorMolData_BP_Direct_EE_WithEle
from .Math import *                         # noqa: F401,F403
from .ForceModifiers import *               # noqa: F401,F403  NeighborListSet, Lattice, PeriodicForce
from .TFNetworks import *                   # noqa: F401,F403  TFMolManage
from .Simulations import *                  # noqa: F401,F403  VelocityVerlet, GeomOptimizer, NudgedElasticBand, ...
from . import MolEmb                        # noqa: F401

__version__ = "0.1.0"
