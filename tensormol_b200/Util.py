"""Globals of the package under the reference's names (TensorMol/Util.py): PARAMS, LOGGER, the TMTiming
decorator and a few scalar helpers.  No TensorFlow."""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np

from .PhysicalData import *        # noqa: F401,F403
from .TMParams import TMBanner, TMLogger, TMParams
from .engine import DSF, DSF_Gradient   # noqa: F401  (Util.py:172-192 in the reference)

PARAMS = TMParams()
LOGGER = TMLogger(PARAMS["log_dir"])
MAX_ATOMIC_NUMBER = 10
HAS_MOLEMB = True      # the MolEmb-compatible module is tensormol_b200.MolEmb (CUDA neighbour search)

TMTIMER = {}
TMSTARTTIME = time.time()


def PrintTMTIMER():
    LOGGER.info("=======    Accumulated Time Information    =======")
    LOGGER.info("Category   |||   Time Per Call   |||   Total Elapsed     ")
    for key in TMTIMER.keys():
        if TMTIMER[key][1] > 0:
            LOGGER.info(key + " ||| %0.5f ||| %0.5f ", TMTIMER[key][0] / (TMTIMER[key][1]), TMTIMER[key][0])


def TMTiming(nm_="Obs"):
    """Accumulates wall time per label (reference Util.py:105-132)."""
    if nm_ not in TMTIMER:
        TMTIMER[nm_] = [0., 0]

    def wrap(f):
        def wf(*args, **kwargs):
            t0 = time.time()
            out = f(*args, **kwargs)
            TMTIMER[nm_][0] += time.time() - t0
            TMTIMER[nm_][1] += 1
            return out
        wf.__name__ = getattr(f, "__name__", "wf")
        wf.__doc__ = getattr(f, "__doc__", None)
        return wf
    return wrap


def scitodeci(sci):
    tmp = str(sci).upper().replace("D", "E").replace("*^", "E")
    return float(tmp)


def AtomicNumber(Symb):
    try:
        return atoi[Symb]          # noqa: F405
    except Exception:
        raise Exception("Unknown Atom")


def AtomicSymbol(number):
    try:
        return itoa[int(number)]   # noqa: F405
    except Exception:
        raise Exception("Unknown Atom")


def LtoS(l):
    return "".join(str(i) + " " for i in l)


def nCr(n, r):
    f = math.factorial
    return int(f(n) / f(r) / f(n - r))


def EluAjust(x, a, x0, shift):
    if x > x0:
        return a * (x - x0) + shift
    return a * (math.exp(x - x0) - 1.0) + shift


def sigmoid_with_param_np(x, alpha=None):
    """numpy form of the reference's activation log(1+exp(alpha x))/alpha (Util.py:200-201), evaluated stably."""
    a = float(PARAMS["sigmoid_alpha"] if alpha is None else alpha)
    t = a * np.asarray(x, np.float64)
    return (np.maximum(t, 0.0) + np.log1p(np.exp(-np.abs(t)))) / a
