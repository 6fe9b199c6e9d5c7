"""Alias package so that driver scripts written for the reference (`from TensorMol import *`) run unchanged on the
B200 path: everything is re-exported from tensormol_b200."""
from tensormol_b200 import *          # noqa: F401,F403
from tensormol_b200 import MolEmb     # noqa: F401
