"""The global parameter dictionary (reference: TensorMol/TMParams.py:8-165) and the logger
(:188-207), without the TensorFlow import.  Only keys that the hot path and the drivers read are
guaranteed; unknown keys can be added by scripts as in the reference (it is a plain dict)."""
from __future__ import annotations

import logging
import os
import time
from math import pi as Pi

import numpy as np


class TMParams(dict):
    def __init__(self, *args, **kwargs):
        dict.__init__(self, *args, **kwargs)
        self["CheckLevel"] = 1
        self["PrintTMTimer"] = False
        self["MAX_ATOMIC_NUMBER"] = 10
        # ANI-1 symmetry functions (TMParams.py:26-38)
        self["AN1_r_Rc"] = 4.6
        self["AN1_a_Rc"] = 3.1
        self["AN1_eta"] = 4.0
        self["AN1_zeta"] = 8.0
        self["AN1_num_r_Rs"] = 32
        self["AN1_num_a_Rs"] = 8
        self["AN1_num_a_As"] = 8
        self["AN1_r_Rs"] = np.array([self["AN1_r_Rc"] * i / self["AN1_num_r_Rs"] for i in range(self["AN1_num_r_Rs"])])
        self["AN1_a_Rs"] = np.array([self["AN1_a_Rc"] * i / self["AN1_num_a_Rs"] for i in range(self["AN1_num_a_Rs"])])
        self["AN1_a_As"] = np.array([2.0 * Pi * i / self["AN1_num_a_As"] for i in range(self["AN1_num_a_As"])])
        # networks
        self["MonitorSet"] = None
        self["NetNameSuffix"] = ""
        self["NeuronType"] = "relu"
        self["tf_prec"] = "tf.float64"      # kept for script compatibility; the B200 path computes in fp32/3xTF32
        self["HiddenLayers"] = [200, 200, 200]
        self["KeepProb"] = 0.7
        self["sigmoid_alpha"] = 100.0
        self["batch_size"] = 1000
        self["max_checkpoints"] = 1
        self["Profiling"] = False
        self["GradScalar"] = 1.0 / 20.0
        self["EnergyScalar"] = 1.0
        self["DipoleScaler"] = 1.0          # (sic) reference TMParams.py:82; the instance reads "DipoleScalar" (:40)
        self["DipoleScalar"] = 1.0
        self["learning_rate"] = 0.001
        self["learning_rate_dipole"] = 0.0001
        self["learning_rate_energy"] = 0.00001
        self["momentum"] = 0.9
        self["max_steps"] = 1001
        self["test_freq"] = 10
        # read by other network families of the reference; kept so that scripts which touch them run (TMParams.py:67-95)
        self["hidden1"], self["hidden2"], self["hidden3"] = 512, 512, 512
        self["GradWeight"] = 0.01
        self["weight_decay"] = 0.001
        self["InNormRoutine"], self["OutNormRoutine"] = None, None
        self["RandomizeData"] = True
        self["train_gradients"], self["train_dipole"], self["train_quadrupole"], self["train_rotation"] = True, True, False, True
        # optimisation
        self["OptMaxCycles"] = 50
        self["OptThresh"] = 0.0001
        self["OptMaxStep"] = 0.1
        self["OptStepSize"] = 0.1
        self["OptMomentum"] = 0.5
        self["OptMomentumDecay"] = 0.8
        self["OptPrintLvl"] = 1
        self["OptLatticeStep"] = 0.050
        self["GSSearchAlpha"] = 0.05
        self["SDStep"] = 0.05
        self["MaxBFGS"] = 7
        self["TestRatio"] = 0.2          # fraction of the cases withheld for testing (reference TMParams.py:71)
        self["NebSolver"] = "Verlet"
        self["NebNumBeads"] = 18
        self["NebK"] = 0.07
        self["NebKMax"] = 1.0
        self["NebClimbingImage"] = True
        self["DiisSize"] = 20
        self["RemoveInvariant"] = True
        # molecular dynamics
        self["MDMaxStep"] = 20000
        self["MDdt"] = 0.2
        self["MDTemp"] = 300.0
        self["MDV0"] = "Random"
        self["MDThermostat"] = None
        self["MDLogTrajectory"] = True
        self["MDUpdateCharges"] = True
        self["MDIrForceMin"] = False
        self["MDAnnealT0"] = 20.0
        self["MDAnnealTF"] = 300.0
        self["MDAnnealKickBack"] = 1.0
        self["MDAnnealSteps"] = 1000
        self["MDFieldVec"] = np.array([1.0, 0.0, 0.0])
        self["MDFieldAmp"] = 0.0
        self["MDFieldFreq"] = 1.0 / 1.2
        self["MDFieldTau"] = 1.2
        self["MDFieldT0"] = 3.0
        # electrostatic embedding (TMParams.py:150-165)
        self["AddEcc"] = True
        self["OPR12"] = "Poly"
        self["Poly_Width"] = 4.6
        self["Elu_Width"] = 4.6
        self["EEOn"] = True
        self["EESwitchFunc"] = "CosLR"
        self["EEVdw"] = True
        self["EEOrder"] = 2
        self["EEdr"] = 1.0
        self["EECutoff"] = 5.0
        self["EECutoffOn"] = 4.4
        self["EECutoffOff"] = 15.0
        self["Erf_Width"] = 0.2
        self["DSFAlpha"] = 0.18
        # paths
        self["tm_root"] = "."
        self["sets_dir"] = self["tm_root"] + "/datasets/"
        self["networks_directory"] = self["tm_root"] + "/networks/"
        self["output_root"] = "."
        self["results_dir"] = self["output_root"] + "/results/"
        self["dens_dir"] = self["output_root"] + "/densities/"
        self["log_dir"] = self["output_root"] + "/logs/"
        # B200 extensions (not in the reference)
        self["B200Device"] = 0
        self["B200GemmMode"] = 1          # 0 fp32 FFMA, 1 tcgen05 3xTF32

    def __str__(self):
        return "".join(f"{k}:{self[k]}\n" for k in self.keys())


def TMBanner():
    print("--------------------------")
    print("   tensormol_b200  (B200-native BP+EE energy/force path, TensorMol-0.1 API surface)")
    print("--------------------------")


def TMLogger(path_):
    """DEBUG -> logs/<timestamp>.log when the directory can be created, INFO -> stdout."""
    tore = logging.getLogger('TensorMol')
    if tore.handlers:
        return tore
    tore.setLevel(logging.DEBUG)
    try:
        if os.environ.get("TENSORMOL_LOG_FILE", "0") != "1":
            raise OSError("file logging disabled (set TENSORMOL_LOG_FILE=1)")
        if not os.path.exists(path_):
            os.makedirs(path_)
        fh = logging.FileHandler(filename=path_ + time.ctime().replace(" ", "_").replace(":", "_") + '.log')
        fh.setLevel(logging.DEBUG)
        fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        tore.addHandler(fh)
    except OSError:
        pass
    ch = logging.StreamHandler()
    ch.setLevel(logging.INFO)
    ch.setFormatter(logging.Formatter("%(message)s"))
    tore.addHandler(ch)
    return tore
