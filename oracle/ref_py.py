"""TEST INFRASTRUCTURE ONLY.  Runs pieces of the reference's own Python, unmodified, where they lie under
/root/reference, to pin the oracle (oracle/make_golden.py writes the results into tests/golden/ref_python_pins.npz;
nothing here is importable on the GPU box, where /root/reference does not exist).

The reference package cannot be imported (TensorFlow 1.x, Python-2 idioms at import time), but the modules on this
path are plain numpy + the MolEmb extension, so their class / function definitions are taken from the source files by
`ast`, the import statements are dropped, and the definitions are executed in a namespace that provides what those
imports would have provided: numpy, the reference's own compiled MolEmb (oracle/_ref, built from C_API/MolEmb.cpp by
oracle/Makefile), PARAMS, LOGGER and the TMTiming decorator.  No reference source is copied into this repository.

    PhysicalData.py                         -> constants(): BOHRPERA, C6_coff, atomic_vdw_radius, ...
    Util.py:172-192                         -> DSF, DSF_Gradient
    Math/LinearOperations.py MatrixPower    (needed by Lattice.InLat)
    ForceModifiers/Periodic.py:12-180       -> Lattice (ModuloLattice, TessLattice)
    ForceModifiers/Neighbors.py:23-489      -> NeighborList, NeighborListSet, NeighborListSetWithImages
"""
from __future__ import annotations

import ast
import glob
import importlib.util
import logging
import os

import numpy as np

REF = os.environ.get("TM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "TensorMol")) and bool(glob.glob(os.path.join(HERE, "_ref", "MolEmb*.so")))


def _load_molemb():
    so = sorted(glob.glob(os.path.join(HERE, "_ref", "MolEmb*.so")))[0]
    spec = importlib.util.spec_from_file_location("MolEmb", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _defs(path, names=None):
    """Source of the top-level class / function definitions (and plain assignments when names is None) of a reference file,
    imports removed."""
    import warnings
    with open(os.path.join(REF, path)) as fh, warnings.catch_warnings():
        warnings.simplefilter("ignore")          # invalid escape sequences in the reference's docstrings
        tree = ast.parse(fh.read())
    body = []
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            continue
        if names is None:
            if isinstance(node, (ast.Assign, ast.FunctionDef, ast.ClassDef)):
                body.append(node)
        elif isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            body.append(node)
    mod = ast.Module(body=body, type_ignores=[])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return compile(ast.fix_missing_locations(mod), os.path.join(REF, path), "exec")


def _tmtiming(_name):
    def deco(f):
        return f
    return deco


_NS = None


def namespace():
    """The executed reference definitions (cached)."""
    global _NS
    if _NS is not None:
        return _NS
    import math
    import scipy.special
    ns = {"np": np, "math": math, "Pi": math.pi, "scipy": scipy, "time": __import__("time"), "itertools": __import__("itertools"),
          "TMTiming": _tmtiming, "LOGGER": logging.getLogger("reference"), "PARAMS": {"tf_prec": "np.float64"}, "xrange": range,
          "print_function": None}
    exec(_defs("TensorMol/PhysicalData.py"), ns)
    exec(_defs("TensorMol/Util.py", {"DSF", "DSF_Gradient"}), ns)
    exec(_defs("TensorMol/Math/LinearOperations.py", {"MatrixPower"}), ns)
    me = _load_molemb()
    ns["MolEmb"] = me
    for k in dir(me):
        if not k.startswith("_"):
            ns[k] = getattr(me, k)
    exec(_defs("TensorMol/ForceModifiers/Neighbors.py", {"NeighborList", "NeighborListSet", "NeighborListSetWithImages"}), ns)
    exec(_defs("TensorMol/ForceModifiers/Periodic.py", {"Lattice"}), ns)
    _NS = ns
    return ns


def constants():
    ns = namespace()
    keys = ["BOHRPERA", "JOULEPERHARTREE", "KCALPERHARTREE", "KJPERHARTREE", "IDEALGASR", "AVOCONST", "AUPERDEBYE"]
    out = {k: float(ns[k]) for k in keys if k in ns}
    zs = sorted(ns["C6_coff"])
    out["vdw_Z"] = np.asarray(zs, np.int64)
    out["C6_coff"] = np.asarray([ns["C6_coff"][z] for z in zs], np.float64)
    out["atomic_vdw_radius"] = np.asarray([ns["atomic_vdw_radius"][z] for z in zs], np.float64)
    out["ATOMICMASSES"] = np.asarray(ns["ATOMICMASSES"], np.float64)
    return out


def element_tables(Z):
    """eles_np / eles_pairs_np as TFMolInstanceDirect builds them (TFMolInstanceDirect.py:4950-4962): sorted elements,
    upper-triangular pairs."""
    eles = sorted(set(int(z) for z in Z if z > 0))
    eles_np = np.asarray(eles).reshape(len(eles), 1)
    pairs = [[eles[i], eles[j]] for i in range(len(eles)) for j in range(i, len(eles))]
    return eles_np, np.asarray(pairs)


def tables_aperiodic(xyz, Z, Rr, Ra):
    """The manager's call sequence for one molecule (TFMolManage.py:1309-1310)."""
    ns = namespace()
    xyzs = np.asarray(xyz, np.float64)[None].copy()
    Zs = np.asarray(Z, np.int32)[None].copy()
    natom = np.array([len(Z)], np.int32)
    eles_np, elep_np = element_tables(Z)
    NL = ns["NeighborListSet"](xyzs, natom, True, True, Zs, sort_=True)
    rad, ang, mil_jk, jk_max = NL.buildPairsAndTriplesWithEleIndex(Rr, Ra, eles_np, elep_np)
    return np.asarray(rad), np.asarray(ang), np.asarray(mil_jk), int(jk_max)


def tables_periodic(xyz_tess, Z_tess, nreal, Rr, Ra, eles):
    """TFMolManage.py:1343-1344."""
    ns = namespace()
    xyzs = np.asarray(xyz_tess, np.float64)[None].copy()
    Zs = np.asarray(Z_tess, np.int32)[None].copy()
    eles_np = np.asarray(sorted(eles)).reshape(-1, 1)
    e = sorted(eles)
    elep_np = np.asarray([[e[i], e[j]] for i in range(len(e)) for j in range(i, len(e))])
    NL = ns["NeighborListSetWithImages"](xyzs, np.array([len(Z_tess)]), np.array([nreal]), True, True, Zs, sort_=True)
    rad, ang, mil_j, mil_jk = NL.buildPairsAndTriplesWithEleIndexPeriodic(Rr, Ra, eles_np, elep_np)
    return np.asarray(rad), np.asarray(ang), np.asarray(mil_j), np.asarray(mil_jk)


_TFNS = None


def tf_namespace(params):
    """The reference's TensorFlow electrostatics functions (RawSymFunc.py:63-79, 123-134, 1307-1465), executed unmodified on
    the numpy stand-in oracle/tf_shim.py.  `params` supplies PARAMS["EECutoffOff"], PARAMS["Poly_Width"]."""
    global _TFNS
    from oracle import tf_shim
    if _TFNS is None:
        ns = {"tf": tf_shim, "np": np, "BOHRPERA": float(namespace()["BOHRPERA"]), "PARAMS": {}}
        exec(_defs("TensorMol/TFDescriptors/RawSymFunc.py",
                   {"AllDoublesSet", "DifferenceVectorsLinear", "TFCoulombEluSRDSFLR", "TFVdwPolyLR", "TFVdwPolyLRWithEle"}), ns)
        exec(_defs("TensorMol/Util.py", {"sigmoid_with_param"}), ns)             # Util.py:200-201
        _TFNS = ns
    _TFNS["PARAMS"].clear()
    _TFNS["PARAMS"].update(params)
    return _TFNS


def vdw_parameters(eles):
    """self.C6 / self.vdw_R of the instance (TFMolInstanceDirect.py:3763-3767)."""
    ns = namespace()
    B, JPH = float(ns["BOHRPERA"]), float(ns["JOULEPERHARTREE"])
    c6 = np.array([ns["C6_coff"][int(e)] * (B * 10.0) ** 6.0 / JPH for e in eles])
    rv = np.array([ns["atomic_vdw_radius"][int(e)] * B for e in eles])
    return c6, rv


def electrostatics_aperiodic(xyz, Z, q, P):
    """Ecc and Evdw of one molecule exactly as energy_inference / dipole_inference call the TF functions
    (TFMolInstanceDirect.py:5214, 5281; pair list of TFMolManage.py:1311-1312), for given atomic charges q [N]."""
    ns = namespace()
    B = float(ns["BOHRPERA"])
    tfn = tf_namespace({"EECutoffOff": P["EECutoffOff"], "Poly_Width": P["Poly_Width"]})
    N = len(Z)
    X = np.asarray(xyz, np.float64)
    NLEE = ns["NeighborListSet"](X[None].copy(), np.array([N], np.int32), False, False, None)
    ree = np.asarray(NLEE.buildPairs(P["EECutoffOff"])).astype(np.int64)
    elu_a = ns["DSF_Gradient"](P["Elu_Width"] * B, P["EECutoffOff"] * B, P["DSFAlpha"] / B)
    elu_s = ns["DSF"](P["Elu_Width"] * B, P["EECutoffOff"] * B, P["DSFAlpha"] / B)
    eles = sorted(set(int(z) for z in Z))
    c6, rv = vdw_parameters(eles)
    xb = X[None] * B                                                        # xyzsInBohr
    Ecc = tfn["TFCoulombEluSRDSFLR"](xb, np.asarray(q, np.float64)[None], P["Elu_Width"] * B, ree, P["DSFAlpha"], elu_a, elu_s)
    Evdw = tfn["TFVdwPolyLR"](xb, np.asarray(Z, np.int64)[None], np.array(eles, np.int64), c6, rv, P["EECutoffOn"] * B, ree)
    return float(Ecc[0]), float(Evdw[0]), ree


def electrostatics_periodic(xyz_tess, Z_tess, nreal, q_real, eles, P):
    """The periodic forms (TFMolInstanceDirect.py:5824, 5892-5896; pair list of TFMolManage.py:1345-1346): charges tiled
    over the image blocks, both energies halved."""
    ns = namespace()
    B = float(ns["BOHRPERA"])
    tfn = tf_namespace({"EECutoffOff": P["EECutoffOff"], "Poly_Width": P["Poly_Width"]})
    X = np.asarray(xyz_tess, np.float64)
    Zt = np.asarray(Z_tess, np.int32)
    eles = sorted(int(e) for e in eles)
    NLEE = ns["NeighborListSetWithImages"](X[None].copy(), np.array([len(Zt)]), np.array([nreal]), False, True, Zt[None].copy())
    ree = np.asarray(NLEE.buildPairsWithBothEleIndex(P["EECutoffOff"], np.asarray(eles).reshape(-1, 1))).astype(np.int64)
    elu_a = ns["DSF_Gradient"](P["Elu_Width"] * B, P["EECutoffOff"] * B, P["DSFAlpha"] / B)
    elu_s = ns["DSF"](P["Elu_Width"] * B, P["EECutoffOff"] * B, P["DSFAlpha"] / B)
    c6, rv = vdw_parameters(eles)
    xb = X[None] * B
    q_all = np.tile(np.asarray(q_real, np.float64)[None], (1, len(Zt) // nreal))
    Ecc = tfn["TFCoulombEluSRDSFLR"](xb, q_all, P["Elu_Width"] * B, ree[:, :3], P["DSFAlpha"], elu_a, elu_s) / 2.0
    Evdw = tfn["TFVdwPolyLRWithEle"](xb, Zt[None].astype(np.int64), np.array(eles, np.int64), c6, rv, P["EECutoffOn"] * B, ree) / 2.0
    return float(Ecc[0]), float(Evdw[0])


def lattice(latvec):
    return namespace()["Lattice"](np.asarray(latvec, np.float64))
