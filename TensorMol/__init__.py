"""Alias package so that driver scripts written for the reference (`from TensorMol import *`,
`from TensorMol.Interfaces.TMIPIinterface import *`, `from TensorMol.ForceModifiers.Neighbors import *`, ...) run
unchanged on the B200 path: everything is re-exported from tensormol_b200, and every submodule of tensormol_b200 is
registered under the matching `TensorMol.` name as THE SAME module object (one PARAMS, one engine)."""
import importlib
import sys

from tensormol_b200 import *          # noqa: F401,F403
from tensormol_b200 import MolEmb     # noqa: F401

for _extra in ("Interfaces", "Interfaces.TMIPIinterface", "Simulations.DeviceMD", "engine", "parallel", "TFNetworks.TFCheckpoint"):
    importlib.import_module("tensormol_b200." + _extra)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("tensormol_b200.") and _mod is not None and ".csrc" not in _name:
        sys.modules.setdefault("TensorMol." + _name[len("tensormol_b200."):], _mod)
        _head = _name[len("tensormol_b200."):]
        if "." not in _head:
            globals().setdefault(_head, _mod)
del _extra, _name, _mod, _head
