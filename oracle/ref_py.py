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
    ns["Mol"] = _PinMol                  # Lattice.CenteredInLattice builds one (Periodic.py:71)
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
    import torch
    T = torch.as_tensor
    Ecc = tfn["TFCoulombEluSRDSFLR"](T(xb), T(np.asarray(q, np.float64)[None]), P["Elu_Width"] * B, T(ree), P["DSFAlpha"], elu_a, elu_s)
    Evdw = tfn["TFVdwPolyLR"](T(xb), T(np.asarray(Z, np.int64)[None]), T(np.array(eles, np.int64)), T(c6), T(rv), P["EECutoffOn"] * B, T(ree))
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
    import torch
    T = torch.as_tensor
    Ecc = tfn["TFCoulombEluSRDSFLR"](T(xb), T(q_all), P["Elu_Width"] * B, T(ree[:, :3].copy()), P["DSFAlpha"], elu_a, elu_s) / 2.0
    Evdw = tfn["TFVdwPolyLRWithEle"](T(xb), T(Zt[None].astype(np.int64)), T(np.array(eles, np.int64)), T(c6), T(rv), P["EECutoffOn"] * B, T(ree)) / 2.0
    return float(Ecc[0]), float(Evdw[0])


def _method_defs(path, names, class_name=None):
    """Methods taken out of their class (first class defining them, or `class_name`) as plain functions f(self, ...)."""
    import warnings
    with open(os.path.join(REF, path)) as fh, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = ast.parse(fh.read())
    found = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and (class_name is None or node.name == class_name):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in names and sub.name not in found:
                    found[sub.name] = sub
    mod = ast.Module(body=[found[n] for n in names if n in found], type_ignores=[])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return compile(ast.fix_missing_locations(mod), os.path.join(REF, path), "exec")


INSTANCE_CLASS = "MolInstance_DirectBP_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout"
_GNS = None


def graph_namespace(P):
    """Everything the instance's EvalPrepare wires together (TFMolInstanceDirect.py:5713-5761), executed on the torch
    stand-in: the symmetry-function graph (RawSymFunc.py), the electrostatics, and the methods energy_inference /
    dipole_inference (+ periodic forms) of the instance class, taken out of the class as plain functions."""
    global _GNS
    from oracle import tf_shim
    if _GNS is None:
        import math
        base = namespace()
        ns = {"tf": tf_shim, "np": np, "math": math, "Pi": math.pi, "BOHRPERA": float(base["BOHRPERA"]), "PARAMS": {},
              "TMTiming": _tmtiming, "print": lambda *a, **k: None}
        exec(_defs("TensorMol/TFDescriptors/RawSymFunc.py",
                   {"AllSinglesSet", "AllDoublesSet", "DifferenceVectorsLinear", "TFSymRSet_Linear_WithEle", "TFSymASet_Linear_WithEle",
                    "TFSymSet_Scattered_Linear_WithEle", "TFSymRSet_Linear_WithElePeriodic", "TFSymASet_Linear_WithElePeriodic",
                    "TFSymSet_Scattered_Linear_WithEle_Periodic", "TFCoulombEluSRDSFLR", "TFVdwPolyLR", "TFVdwPolyLRWithEle"}), ns)
        exec(_defs("TensorMol/Util.py", {"sigmoid_with_param"}), ns)
        exec(_method_defs("TensorMol/TFNetworks/TFMolInstanceDirect.py",
                          ["energy_inference", "dipole_inference", "energy_inference_periodic", "dipole_inference_periodic"], INSTANCE_CLASS), ns)
        exec(_method_defs("TensorMol/TFNetworks/TFMolInstanceDirect.py", ["SetANI1Param"]), ns)
        _GNS = ns
    _GNS["PARAMS"].clear()
    _GNS["PARAMS"].update(P)
    return _GNS


class _FakeInstance:
    """The attributes of the reference instance that the extracted methods read (TFMolInstanceDirect.py:4969-4984, 3755-3767)."""

    def __init__(self, ns, eles, hidden, P, nmol, maxnatom, weights):
        from oracle import tf_shim
        base = namespace()
        B = float(base["BOHRPERA"])
        self.eles = list(eles)
        self.eles_np = np.asarray(self.eles).reshape(-1, 1)
        self.eles_pairs = [[self.eles[i], self.eles[j]] for i in range(len(self.eles)) for j in range(i, len(self.eles))]
        self.eles_pairs_np = np.asarray(self.eles_pairs)
        self.HiddenLayers = list(hidden)
        self.batch_size, self.MaxNAtoms = nmol, maxnatom
        self.tf_prec = tf_shim.float64
        self.activation_function = ns["sigmoid_with_param"]
        self.DSFAlpha = P["DSFAlpha"]
        self.elu_shift = base["DSF"](P["Elu_Width"] * B, P["EECutoffOff"] * B, P["DSFAlpha"] / B)          # :4371
        self.elu_alpha = base["DSF_Gradient"](P["Elu_Width"] * B, P["EECutoffOff"] * B, P["DSFAlpha"] / B)
        self.C6, self.vdw_R = vdw_parameters(self.eles)
        ns["SetANI1Param"](self)
        self._weights = weights
        self._queue = []

    def load(self, net):
        """Queue this net's variables in the order the inference method creates them: per element, per hidden layer
        (weights, biases), then the linear output layer."""
        from oracle import tf_shim
        assert not self._queue and not tf_shim._bias_queue, "the previous inference did not consume all its variables"
        self._queue = []
        biases = []
        for z in self.eles:
            for W, b in self._weights[net][int(z)]:
                self._queue.append(np.asarray(W, np.float64))
                biases.append(np.asarray(b, np.float64))
        tf_shim.push_biases(biases)

    def _variable_with_weight_decay(self, var_name, var_shape, var_stddev, var_wd):
        import torch
        W = self._queue.pop(0)
        assert list(W.shape) == [int(v) for v in var_shape], (W.shape, var_shape)
        return torch.as_tensor(W)


def weights_with_biases(weights, seed):
    """random_weights() keeps the reference's zero bias initialisation; the pins use seeded non-zero biases so that the
    bias wiring is exercised as well."""
    rng = np.random.default_rng(seed)
    return {net: {z: [(W, 0.05 * rng.standard_normal(b.shape)) for W, b in layers] for z, layers in d.items()} for net, d in weights.items()}


def full_graph_aperiodic(xyz, Z, hidden, weights, P):
    """One molecule through the reference's own graph code in the order of EvalPrepare / evaluate
    (TFMolInstanceDirect.py:5748-5761, 5684-5711; feed of TFMolManage.py:1300-1317): returns Etotal, Ebp, Ecc, Evdw,
    Ebp_atom, dipole, charge, gradient dE/dx (Hartree / Angstrom)."""
    import torch
    from oracle import tf_shim as tf
    ns = graph_namespace(P)
    X = np.asarray(xyz, np.float64)
    Zs = np.asarray(Z, np.int64)[None]
    N = Zs.shape[1]
    eles = sorted(set(int(z) for z in Z))
    inst = _FakeInstance(ns, eles, hidden, P, 1, N, weights)
    rad, ang, mil_jk, _ = tables_aperiodic(X, Z, P["AN1_r_Rc"], P["AN1_a_Rc"])
    base = namespace()
    NLEE = base["NeighborListSet"](X[None].copy(), np.array([N], np.int32), False, False, None)
    reep = np.asarray(NLEE.buildPairs(P["EECutoffOff"])).astype(np.int64)
    xyzs = torch.tensor(X[None], dtype=torch.float64, requires_grad=True)
    Zt = torch.as_tensor(Zs)
    natom = torch.tensor([1.0 / N], dtype=torch.float64)
    keep = torch.ones(len(hidden) + 1, dtype=torch.float64)
    Ele, Elep = torch.as_tensor(inst.eles_np), torch.as_tensor(inst.eles_pairs_np)
    SFPa2, SFPr2 = torch.as_tensor(inst.SFPa2), torch.as_tensor(inst.SFPr2)
    sym, idx = ns["TFSymSet_Scattered_Linear_WithEle"](xyzs, Zt, Ele, SFPr2, inst.Rr_cut, Elep, SFPa2, inst.zeta, inst.eta, inst.Ra_cut,
                                                       torch.as_tensor(rad.astype(np.int64)), torch.as_tensor(ang.astype(np.int64)),
                                                       torch.as_tensor(mil_jk.astype(np.int64)))
    inst.load("charge")
    Ecc, dipole, charge, _ = ns["dipole_inference"](inst, sym, idx, xyzs, natom, P["Elu_Width"], P["EECutoffOff"], torch.as_tensor(reep), True, keep)
    inst.load("energy")
    Etotal, Ebp, Evdw, _, Ebp_atom = ns["energy_inference"](inst, sym, idx, Ecc, xyzs, Zt, Ele, torch.as_tensor(inst.C6), torch.as_tensor(inst.vdw_R),
                                                            torch.as_tensor(reep), P["EECutoffOn"], P["EECutoffOff"], keep)
    grad = tf.gradients(Etotal, xyzs)[0]
    D = torch.cat([torch.zeros(0)]).numpy()   # placeholder to keep numpy import obvious
    desc = np.zeros((N, int(inst.inshape)))
    for e in range(len(eles)):
        rows = idx[e][:, 1].numpy()
        desc[rows] = sym[e].detach().numpy()
    del D
    return dict(Etotal=Etotal.detach().numpy(), Ebp=Ebp.detach().numpy(), Ecc=Ecc.detach().numpy(), Evdw=Evdw.detach().numpy(),
                Ebp_atom=Ebp_atom.detach().numpy(), dipole=dipole.detach().numpy(), charge=charge.detach().numpy(),
                gradient=grad.detach().numpy(), descriptors=desc)


def full_graph_periodic(xyz_tess, Z_tess, nreal, eles, hidden, weights, P):
    """The periodic form: EvalPrepare_Periodic / evaluate_periodic (TFMolInstanceDirect.py:5919-6003) fed as
    TFMolManage.py:1335-1351 does.  Returns the reference's outputs; gradient rows beyond nreal are the image rows the
    manager discards."""
    import torch
    from oracle import tf_shim as tf
    ns = graph_namespace(P)
    X = np.asarray(xyz_tess, np.float64)
    Zt_np = np.asarray(Z_tess, np.int32)
    NT = len(Zt_np)
    eles = sorted(int(e) for e in eles)
    inst = _FakeInstance(ns, eles, hidden, P, 1, NT, weights)
    inst.nreal = int(nreal)
    rad, ang, mil_j, mil_jk = tables_periodic(X, Zt_np, nreal, P["AN1_r_Rc"], P["AN1_a_Rc"], eles)
    base = namespace()
    NLEE = base["NeighborListSetWithImages"](X[None].copy(), np.array([NT]), np.array([nreal]), False, True, Zt_np[None].copy())
    ree = np.asarray(NLEE.buildPairsWithBothEleIndex(P["EECutoffOff"], np.asarray(eles).reshape(-1, 1))).astype(np.int64)
    T = torch.as_tensor
    xyzs = torch.tensor(X[None], dtype=torch.float64, requires_grad=True)
    Zt = T(Zt_np.astype(np.int64)[None])
    natom = torch.tensor([1.0 / nreal], dtype=torch.float64)
    keep = torch.ones(len(hidden) + 1, dtype=torch.float64)
    Ele, Elep = T(inst.eles_np), T(inst.eles_pairs_np)
    sym, idx = ns["TFSymSet_Scattered_Linear_WithEle_Periodic"](xyzs, Zt, Ele, T(inst.SFPr2), inst.Rr_cut, Elep, T(inst.SFPa2), inst.zeta, inst.eta,
                                                                inst.Ra_cut, T(rad.astype(np.int64)), T(ang.astype(np.int64)), T(mil_j.astype(np.int64)),
                                                                T(mil_jk.astype(np.int64)), int(nreal))
    inst.load("charge")
    Ecc, dipole, charge, _ = ns["dipole_inference_periodic"](inst, sym, idx, xyzs, natom, P["Elu_Width"], P["EECutoffOff"], T(ree[:, :3].copy()), True, keep)
    inst.load("energy")
    Etotal, Ebp, Evdw, _, Ebp_atom = ns["energy_inference_periodic"](inst, sym, idx, Ecc, xyzs, Zt, Ele, T(inst.C6), T(inst.vdw_R), T(ree),
                                                                     P["EECutoffOn"], P["EECutoffOff"], keep)
    grad = tf.gradients(Etotal, xyzs)[0]
    return dict(Etotal=Etotal.detach().numpy(), Ebp=Ebp.detach().numpy(), Ecc=Ecc.detach().numpy(), Evdw=Evdw.detach().numpy(),
                Ebp_atom=Ebp_atom.detach().numpy(), dipole=dipole.detach().numpy(), charge=charge.detach().numpy(),
                gradient=grad.detach().numpy())


def md_namespace(P):
    """The reference's integrator functions (Simulations/SimpleMD.py:14-129, PeriodicMD.py:21-60), plain numpy."""
    base = namespace()
    ns = {"np": np, "PARAMS": dict(P), "IDEALGASR": base["IDEALGASR"], "KineticEnergy": None, "LOGGER": base["LOGGER"]}
    exec(_defs("TensorMol/Simulations/SimpleMD.py", {"VelocityVerletStep", "KineticEnergy", "Thermostat", "NoseThermostat"}), ns)
    exec(_defs("TensorMol/Simulations/PeriodicMD.py", {"PeriodicVelocityVerletStep", "PeriodicNoseThermostat"}), ns)
    return ns


class _ToyPeriodicForce:
    """A PeriodicForce stand-in for the integrator pins: the reference Lattice plus a smooth analytic energy / force
    (harmonic tethers to the initial positions and a soft pair term), returned in the callback's units
    (energy Hartree-like scalar, force in J/mol/Angstrom)."""

    def __init__(self, lat, x0):
        self.lattice = lattice(lat)
        self.x0 = np.array(x0)

    def __call__(self, x, DoForce=True):
        d = x - self.x0
        e = 0.5 * 3.0e5 * float(np.sum(d * d)) + 1.0e4 * float(np.sum(np.sin(x)))
        f = -(3.0e5 * d + 1.0e4 * np.cos(x))
        return e, f


def md_pins(lat, x0, m, v0, dt, nsteps, P):
    """nsteps of PeriodicVelocityVerletStep and of PeriodicNoseThermostat.step on the toy force."""
    ns = md_namespace(P)
    pf = _ToyPeriodicForce(lat, x0)
    out = {}
    x, v, a = np.array(x0), np.array(v0), np.zeros_like(x0)
    traj = []
    for _ in range(nsteps):
        x, v, a, e = ns["PeriodicVelocityVerletStep"](pf, a, x, v, m, dt)
        traj.append(np.concatenate([x.ravel(), v.ravel(), [e]]))
    out["nve"] = np.array(traj)
    vv = np.array(v0)
    th = ns["PeriodicNoseThermostat"](m, vv)          # rescales vv in place (SimpleMD.py:90-100)
    out["nose_v0"] = vv.copy()
    x, v, a = np.array(x0), vv, np.zeros_like(x0)
    traj = []
    for _ in range(nsteps):
        x, v, a, e = th.step(pf, a, x, v, m, dt)
        traj.append(np.concatenate([x.ravel(), v.ravel(), [e, th.eta]]))
    out["nose"] = np.array(traj)
    out["ke"] = np.float64(ns["KineticEnergy"](np.array(v0), m))
    return out


def lattice(latvec):
    return namespace()["Lattice"](np.asarray(latvec, np.float64))


# ---- host drivers off the hot path: optimisers, NEB, xyz wire format -------------------------------

class _NumpyCompat:
    """numpy with the aliases the reference still uses (np.float, removed in NumPy 1.24); everything else is numpy's."""
    float = float

    def __getattr__(self, k):
        return getattr(np, k)


def mol_class():
    """The reference's Mol class (Containers/Mol.py:7-...) executed in place; only the xyz reading / writing members are
    exercised (FromXYZString, ParseProperties, PropertyString, __str__, WriteXYZfile)."""
    base = namespace()
    ns = {"np": _NumpyCompat(), "os": os, "errno": __import__("errno"), "atoi": base["atoi"], "PARAMS": {"GoK": 1.0}, "LOGGER": base["LOGGER"], "random": __import__("random"),
          "MolEmb": base["MolEmb"], "ELEHEATFORM": {}, "re": __import__("re")}
    exec(_defs("TensorMol/Util.py", {"scitodeci", "AtomicNumber", "AtomicSymbol"}), ns)
    exec(_defs("TensorMol/Containers/Mol.py", {"Mol"}), ns)
    ns["Mol"].CalculateAtomization = lambda self: None      # needs the heat-of-formation tables; not part of the format
    return ns["Mol"]


class _QuietMol:
    """Mol stand-in for the optimiser pins: the drivers only construct it, set properties and write trajectories."""

    def __init__(self, atoms_, coords_):
        self.atoms, self.coords, self.properties = np.array(atoms_), np.array(coords_), {}

    def WriteXYZfile(self, *a, **k):
        pass

    def __str__(self):
        return "Mol(%d atoms)" % len(self.atoms)


def opt_namespace(P):
    """ConjGradient / RemoveInvariantForce (Math/QuasiNewtonTools.py:10-25,350-465,551-577), the solvers of Math/BFGS.py
    and Math/DIIS.py, GeomOptimizer (Simulations/Opt.py:17-75) and NudgedElasticBand (Simulations/Neb.py:17-220) executed
    in place with the PARAMS given; prints silenced, trajectory files not written."""
    base = namespace()
    ns = {k: base[k] for k in ("JOULEPERHARTREE", "GOLDENRATIO", "BOHRPERA", "KCALPERHARTREE", "ATOMICMASSES", "ATOMICMASSESAMU",
                               "IDEALGASR", "atoi", "LOGGER")}
    ns.update({"np": np, "PARAMS": dict(P), "Mol": _QuietMol, "print": lambda *a, **k: None, "xrange": range,
               "TMTiming": _tmtiming, "time": __import__("time"), "random": __import__("random")})
    exec(_defs("TensorMol/Math/LinearOperations.py", {"PseudoInverse", "MatrixPower", "Normalize"}), ns)
    exec(_defs("TensorMol/Math/QuasiNewtonTools.py", {"RmsForce", "CenterOfMass", "InertiaTensor", "FdiffGradient", "ConjGradient",
                                                      "RemoveInvariantForce", "LineSearch"}), ns)
    exec(_defs("TensorMol/Math/BFGS.py", {"SteepestDescent", "VerletOptimizer", "BFGS", "BFGS_WithLinesearch"}), ns)
    exec(_defs("TensorMol/Math/DIIS.py", {"DIIS"}), ns)
    exec(_defs("TensorMol/Simulations/Opt.py", {"GeomOptimizer"}), ns)
    exec(_defs("TensorMol/Simulations/Neb.py", {"NudgedElasticBand"}), ns)
    return ns


def toy_surface(x, DoForce=True):
    """Smooth analytic molecule-like surface for the optimiser / NEB pins, in the callback's convention
    (E in Hartree, force in J/mol/Angstrom): harmonic bonds between consecutive atoms (r0 = 1.1), a weak 1-3 term and a
    double well in the first atom's x coordinate."""
    JPH = 2625499.638
    d1 = x[1:] - x[:-1]
    r1 = np.sqrt((d1 * d1).sum(1))
    d2 = x[2:] - x[:-2]
    r2 = np.sqrt((d2 * d2).sum(1))
    w = x[0, 0]
    e = 0.4 * np.sum((r1 - 1.1) ** 2) + 0.05 * np.sum((r2 - 1.9) ** 2) + 0.02 * (w * w - 1.0) ** 2
    if not DoForce:
        return float(e)
    g = np.zeros_like(x)
    c1 = (0.8 * (r1 - 1.1) / r1)[:, None] * d1
    g[1:] += c1
    g[:-1] -= c1
    c2 = (0.1 * (r2 - 1.9) / r2)[:, None] * d2
    g[2:] += c2
    g[:-2] -= c2
    g[0, 0] += 0.08 * w * (w * w - 1.0)
    return float(e), -JPH * g


def opt_pins(atoms, x0, x1, P, nsteps=6):
    """Reference outputs on the toy surface: RemoveInvariantForce, `nsteps` ConjGradient iterations, GeomOptimizer.Opt, and
    `nsteps` NudgedElasticBand solver iterations per solver (bead positions, Es, NEB forces)."""
    out = {}
    ns = opt_namespace(P)
    e0, f0 = toy_surface(x0)
    out["rif"] = ns["RemoveInvariantForce"](x0, f0, np.asarray(atoms, np.float64))
    go = ns["GeomOptimizer"](toy_surface)
    go.m = _QuietMol(atoms, x0)
    cg = ns["ConjGradient"](go.WrappedEForce, x0)
    x, tr = x0.copy(), []
    for _ in range(nsteps):
        x, e, g = cg(x)
        tr.append(np.concatenate([x.ravel(), g.ravel(), [e, cg.alpha]]))
    out["cg"] = np.array(tr)
    m = ns["GeomOptimizer"](toy_surface).Opt(_QuietMol(atoms, x0))
    out["opt_coords"] = m.coords
    out["opt_energy"] = np.float64(m.properties["Energy"])
    out["opt_step"] = np.int64(m.properties["Step"])
    for solver in ("Verlet", "BFGS", "DIIS", "CG"):      # "SD" raises in the reference (Neb.py:51-62: an `if` where an `elif` is meant)
        Ps = dict(P)
        Ps["NebSolver"] = solver
        nss = opt_namespace(Ps)
        neb = nss["NudgedElasticBand"](toy_surface, _QuietMol(atoms, x0), _QuietMol(atoms, x1), nbeads_=7)
        tr = []
        for _ in range(nsteps):
            neb.beads, e, neb.Fs = neb.Solver(neb.beads)
            neb.IntegrateEnergy()
            neb.TSI = np.argmax(neb.Es)
            tr.append(np.concatenate([neb.beads.ravel(), neb.Fs.ravel(), neb.Es, neb.Esi, [e]]))
            neb.step += 1
        out["neb_" + solver] = np.array(tr)
    return out


def simple_md_pins(x0, m, v0, dt, nsteps, P):
    """Aperiodic integrators of the reference (Simulations/SimpleMD.py:14-320): VelocityVerletStep and the step functions
    of Thermostat (Rescaling), NoseThermostat, AndersenThermostat, LangevinThermostat on the toy
    surface; the stochastic thermostats draw from numpy's global generator, seeded per run (np.random.seed(7))."""
    base = namespace()
    ns = {"np": np, "PARAMS": dict(P), "IDEALGASR": base["IDEALGASR"], "LOGGER": base["LOGGER"], "print": lambda *a, **k: None,
          "KCONVERT": base.get("KCONVERT"), "time": __import__("time"), "random": __import__("random")}
    exec(_defs("TensorMol/Simulations/SimpleMD.py", {"VelocityVerletStep", "KineticEnergy", "Thermostat", "NoseThermostat",
                                                     "AndersenThermostat", "LangevinThermostat", "NoseChainThermostat"}), ns)
    force = lambda x: toy_surface(x)[1]                     # noqa: E731  (the integrators take a force routine)
    fande = lambda x: toy_surface(x)                        # noqa: E731
    out = {}
    x, v, a = np.array(x0), np.array(v0), np.zeros_like(x0)
    tr = []
    for _ in range(nsteps):
        x, v, a, e = ns["VelocityVerletStep"](force, a, x, v, m, dt, fande)
        tr.append(np.concatenate([x.ravel(), v.ravel(), a.ravel(), [e]]))
    out["vv"] = np.array(tr)
    # NoseChainThermostat cannot be constructed in the reference (PARAMS["MNHChain"] has no default, and its __init__ calls
    # Rescale before any self.m exists, SimpleMD.py:220,250): nothing to pin. LangevinThermostat ("Not Working",
    # :166) overflows within a few steps with masses in kg/mol: rows are kept while finite.
    for name in ("Thermostat", "NoseThermostat", "AndersenThermostat", "LangevinThermostat"):
        np.random.seed(7)
        vv = np.array(v0)
        th = ns[name](m, vv)
        out[name + "_v0"] = vv.copy()
        x, v, a = np.array(x0), vv, np.zeros_like(x0)
        tr = []
        for _ in range(nsteps):
            x, v, a, e = th.step(force, a, x, v, m, dt, fande)[:4]      # some steps also return the force (frc_ = True)
            if not (np.isfinite(x).all() and np.isfinite(v).all() and np.isfinite(a).all() and np.isfinite(e)):
                break
            tr.append(np.concatenate([x.ravel(), v.ravel(), a.ravel(), [e]]))
        out[name] = np.array(tr)
    return out


class _PinMol(_QuietMol):
    """_QuietMol with the members the periodic wrappers call."""

    def __init__(self, atoms_=None, coords_=None):
        _QuietMol.__init__(self, np.zeros(0, np.uint8) if atoms_ is None else atoms_, np.zeros((0, 3)) if coords_ is None else coords_)

    def NAtoms(self):
        return len(self.atoms)

    def Center(self):
        return np.average(self.coords, axis=0)          # Containers/Mol.py: Center


def toy_local_force(z, x, nreal, DoForce=True):
    """A smooth periodic `local force` for the PeriodicForce pins (callback of Periodic.py:368-400: tessellated atoms in,
    energy [Hartree] and forces [J/mol/Angstrom] on the first nreal rows out): Gaussian pair repulsion between every real
    atom and every other tessellated atom, weighted by the atomic numbers."""
    JPH = 2625499.638
    x = np.asarray(x, np.float64)
    w = np.sqrt(np.asarray(z, np.float64))
    d = x[:nreal, None, :] - x[None, :, :]
    r2 = (d * d).sum(-1)
    g = 0.004 * w[:nreal, None] * w[None, :] * np.exp(-r2 / 2.5)
    g[np.arange(nreal), np.arange(nreal)] = 0.0
    e = 0.5 * g.sum()
    if not DoForce:
        return e
    f = (g[:, :, None] * d * (2.0 / 2.5)).sum(1)             # -dE/dx_i of the real atoms (pairs with images counted once per i)
    return e, JPH * f


def periodic_driver_namespace(P):
    """LocalForce / PeriodicForce (ForceModifiers/Periodic.py:182-473), VelocityVerlet and the periodic drivers
    (Simulations/PeriodicMD.py, PeriodicMC.py, OptPeriodic.py) executed in place."""
    base = namespace()
    ns = opt_namespace(P)
    ns.update({"Mol": _PinMol, "MolEmb": base["MolEmb"], "Lattice": base["Lattice"], "AVOCONST": base["AVOCONST"],
               "KAYBEETEE": base["KAYBEETEE"], "KJPERHARTREE": base["KJPERHARTREE"], "PrintTMTIMER": lambda: None, "os": os,
               "map": lambda f, it: list(map(f, it)),          # Periodic.py:309 relies on Python 2's list-returning map
               "GetRDF_Bin": base["MolEmb"].GetRDF_Bin})
    exec(_defs("TensorMol/Math/LinearOperations.py", {"MovingAverage"}), ns)
    exec(_defs("TensorMol/Math/Statistics.py", {"OnlineEstimator"}), ns)
    exec(_defs("TensorMol/ForceModifiers/Periodic.py", {"LocalForce", "PeriodicForce"}), ns)
    exec(_defs("TensorMol/Simulations/SimpleMD.py", {"VelocityVerletStep", "KineticEnergy", "Thermostat", "NoseThermostat", "VelocityVerlet"}), ns)
    exec(_defs("TensorMol/Simulations/PeriodicMD.py", {"PeriodicVelocityVerletStep", "PeriodicNoseThermostat", "PeriodicVelocityVerlet"}), ns)
    exec(_defs("TensorMol/Simulations/PeriodicMC.py", {"PeriodicMonteCarlo"}), ns)
    exec(_defs("TensorMol/Simulations/OptPeriodic.py", {"PeriodicGeomOptimizer"}), ns)
    ns["np"] = _SaveLess()
    return ns


class _SaveLess(_NumpyCompat):
    """numpy whose savetxt is a no-op (the drivers write logs under ./results/)."""

    def savetxt(self, *a, **k):
        pass


def periodic_driver_pins(atoms, x0, lat, P, nsteps=6):
    """PeriodicForce energy / force / RDF / Density, a seeded Metropolis chain of PeriodicMonteCarlo and a
    PeriodicGeomOptimizer.Opt run of the reference on the toy local force."""
    out = {}
    ns = periodic_driver_namespace(P)
    pf = ns["PeriodicForce"](_PinMol(atoms, x0), np.array(lat))
    pf.BindForce(toy_local_force, 6.0)
    out["mol0"] = pf.mol0.coords.copy()
    e, f = pf(pf.mol0.coords)
    out["e"], out["f"] = np.float64(e), f
    out["e_only"] = np.float64(pf(pf.mol0.coords, DoForce=False)[0])
    out["density"] = np.float64(pf.Density())
    out["rdf"] = pf.RDF(pf.mol0.coords, 8, 1, 7.0, 0.05)
    np.random.seed(11)
    mc = ns["PeriodicMonteCarlo"](pf, "pinMC")
    tr = []
    for _ in range(nsteps):
        mc.MetropolisHastings(mc.x)
        tr.append(np.concatenate([mc.x.ravel(), [mc.eold, mc.Pacc, mc.Eav, mc.dE2]]))
    out["mc"] = np.array(tr)
    out["mc_kbt"] = np.float64(mc.kbt)
    pf2 = ns["PeriodicForce"](_PinMol(atoms, x0), np.array(lat))
    pf2.BindForce(toy_local_force, 6.0)
    pf2.Save = lambda *a, **k: None
    m = ns["PeriodicGeomOptimizer"](pf2).Opt(_PinMol(atoms, pf2.mol0.coords.copy()))
    out["popt_coords"] = m.coords
    pf3 = ns["PeriodicForce"](_PinMol(atoms, x0), np.array(lat))
    pf3.BindForce(toy_local_force, 6.0)
    ns["PARAMS"]["OptLatticeStep"] = 0.05
    out["latstep_x"] = pf3.LatticeStep(pf3.mol0.coords)
    out["latstep_lattice"] = pf3.lattice.lattice.copy()
    out["latstep_step"] = np.float64(ns["PARAMS"]["OptLatticeStep"])
    cube = 6.0
    xc = np.mod(np.array(x0), cube)
    out["rdf_inc"] = pf3.RDF_inC(xc, np.asarray(atoms), cube, 8, 1, 7.0, 0.05)
    return out


def host_pin_inputs():
    """Seeded inputs of the host-driver pins (shared by the generator and tests/test_host_api.py)."""
    atoms = np.array([8, 1, 1, 6, 1], np.uint8)
    rs = np.random.RandomState(5)
    x0 = np.cumsum(np.abs(rs.randn(5, 3)) * 0.7, axis=0)
    x0[0, 0] = -1.05
    x1 = x0.copy()
    x1[0, 0] = 1.02
    x1[1:] += 0.05 * rs.randn(4, 3)
    return atoms, x0, x1


def harmonic_pins(atoms, x0, P):
    """FdiffGradient, FdiffHessian (forward / central / gradient modes) and HarmonicSpectra (Math/QuasiNewtonTools.py:43-156,
    188-293) of the reference on the toy surface."""
    base = namespace()
    ns = opt_namespace(P)
    ns.update({k: base[k] for k in ("ELECTRONPERPROTONMASS", "WAVENUMBERPERHARTREE", "BOHRPERA", "ATOMICMASSESAMU")})
    ns["map"] = lambda f, it: list(map(f, it))
    exec(_defs("TensorMol/Math/LinearOperations.py", {"PairOrthogonalize", "SchmidtStep", "Normalize"}), ns)
    exec(_defs("TensorMol/Math/QuasiNewtonTools.py", {"FdiffHessian", "DirectedFdiffHessian", "InternalCoordinates", "HarmonicSpectra",
                                                      "FourPointHessQuad"}), ns)
    energy = lambda x: np.float64(toy_surface(x, False))             # noqa: E731
    grad = lambda x: -toy_surface(x)[1] / 2625499.638                # noqa: E731
    out = {"fd_gradient": ns["FdiffGradient"](energy, x0), "fd_hess_forward": ns["FdiffHessian"](energy, x0, 0.001),
           "fd_hess_central": ns["FdiffHessian"](energy, x0, 0.001, "central"),
           "fd_hess_gradient": ns["FdiffHessian"](energy, x0, 0.001, "gradient", grad)}
    w, v = ns["HarmonicSpectra"](energy, x0, np.asarray(atoms))
    out["harm_w"], out["harm_v"] = w, v
    return out


def toy_surface_bound(x, DoForce=True):
    """toy_surface shifted down by one Hartree: a negative energy, like a bound molecule's, which the annealers need
    (their best-energy tracker starts from 0.0, SimpleMD.py:472)."""
    if not DoForce:
        return toy_surface(x, False) - 1.0
    e, f = toy_surface(x)
    return e - 1.0, f


def toy_charges(x):
    """Smooth geometry-dependent charges for the IR / annealing driver pins (neutral by construction)."""
    q = 0.3 * np.sin(x[:, 0]) + 0.1 * np.cos(x[:, 1] + x[:, 2])
    return q - q.mean()


def md_driver_pins(atoms, x0, lat, P):
    """Whole Prop() runs of the reference's MD drivers with PARAMS["MDV0"] = None (zero initial velocities; "Random"
    reseeds numpy from the OS): VelocityVerlet (NVE and Nose), IRTrajectory with a field pulse and geometry-dependent
    charges, Annealer (Simulations/SimpleMD.py:322-619), PeriodicVelocityVerlet (NVE and Nose), PeriodicAnnealer and
    PeriodicBoxingDynamics (Simulations/PeriodicMD.py:44-285). Recorded: final x, v and the driver's own log."""
    out = {}

    def nsfor(**kw):
        Q = dict(P)
        Q.update(kw)
        ns = periodic_driver_namespace(Q)
        ns["WeightedCoordAverage"] = None
        exec(_defs("TensorMol/ForceModels/Electrostatics.py", {"Dipole_Naive", "ElectricFieldForce"}), ns)
        ns["Dipole"] = ns["Dipole_Naive"]        # Electrostatics.Dipole = the same sum through WeightedCoordAverage
        exec(_defs("TensorMol/Simulations/SimpleMD.py", {"IRTrajectory", "Annealer"}), ns)
        exec(_defs("TensorMol/Simulations/PeriodicMD.py", {"PeriodicAnnealer", "PeriodicBoxingDynamics"}), ns)
        return ns

    force = lambda x: toy_surface(x)[1]          # noqa: E731
    for tag, thermo in (("nve", None), ("nose", "Nose")):
        ns = nsfor(MDThermostat=thermo)
        d = ns["VelocityVerlet"](force, _PinMol(atoms, x0), "pin", toy_surface)
        d.Prop()
        out["vv_" + tag + "_x"], out["vv_" + tag + "_v"], out["vv_" + tag + "_log"] = d.x, d.v, d.md_log
    ns = nsfor(MDThermostat=None, MDFieldAmp=2.0)
    d = ns["IRTrajectory"](toy_surface_bound, toy_charges, _PinMol(atoms, x0), "pinir")
    d.Prop()
    out["ir_x"], out["ir_v"], out["ir_log"] = d.x, d.v, d.mu_his
    ns = nsfor(MDAnnealSteps=8, MDAnnealT0=40.0, MDAnnealTF=5.0)
    d = ns["Annealer"](toy_surface_bound, toy_charges, _PinMol(atoms, x0), "pinan")
    d.Prop()
    out["an_x"], out["an_v"], out["an_minx"], out["an_mine"] = d.x, d.v, d.Minx, np.float64(d.MinE)
    patoms, px0 = np.array([1, 1, 8] * 3, np.uint8), np.random.RandomState(2).rand(9, 3) * 6.0
    for tag, thermo in (("nve", None), ("nose", "Nose")):
        ns = nsfor(MDThermostat=thermo)
        pf = ns["PeriodicForce"](_PinMol(patoms, px0), np.array(lat))
        pf.BindForce(toy_local_force, 6.0)
        d = ns["PeriodicVelocityVerlet"](pf, "pinp")
        d.Prop()
        out["pvv_" + tag + "_x"], out["pvv_" + tag + "_v"], out["pvv_" + tag + "_log"] = d.x, d.v, d.md_log
    ns = nsfor(MDAnnealSteps=8, MDAnnealT0=40.0, MDAnnealTF=5.0)
    pf = ns["PeriodicForce"](_PinMol(patoms, px0), np.array(lat))
    pf.BindForce(toy_local_force, 6.0)
    d = ns["PeriodicAnnealer"](pf, "pinpa")
    d.Prop()
    out["pan_x"], out["pan_v"], out["pan_minx"], out["pan_mine"] = d.x, d.v, d.Minx, np.float64(d.MinE)
    ns = nsfor(MDThermostat="Nose")
    pf = ns["PeriodicForce"](_PinMol(patoms, px0), np.array(lat))
    pf.BindForce(toy_local_force, 6.0)
    d = ns["PeriodicBoxingDynamics"](pf, np.array(lat) * 0.97, "pinbox", 1.0)
    d.Prop()
    out["box_x"], out["box_v"], out["box_log"], out["box_lattice"] = d.x, d.v, d.md_log, pf.lattice.lattice.copy()
    return out


def train_set_inputs():
    """Seeded little training set of the batch-provider pins (shared by the generator and the tests): 11 molecules of two
    sizes (H2O, CH2O + H) with labels."""
    rs = np.random.RandomState(21)
    water = (np.array([8, 1, 1], np.uint8), np.array([[0.0, 0.0, 0.0], [0.757, 0.586, 0.0], [-0.757, 0.586, 0.0]]))
    form = (np.array([6, 8, 1, 1, 1], np.uint8), np.array([[0.0, 0.0, 0.0], [1.21, 0.0, 0.0], [-0.55, 0.94, 0.0], [-0.55, -0.94, 0.0],
                                                            [0.4, 0.1, 1.7]]))
    mols = []
    for i in range(11):
        z, x = form if i % 3 == 0 else water
        mols.append({"atoms": z.copy(), "coords": x + 0.08 * rs.randn(*x.shape), "atomization": float(-0.3 - 0.01 * i + 0.001 * rs.randn()),
                     "dipole": 0.5 * rs.randn(3), "gradients": 0.02 * rs.randn(*x.shape)})
    return mols


class _TrainSelf:
    """The attributes TensorMolData.__init__ (Containers/TensorMolData.py:24-75, 1222-1245, 1546-1553, 1670-1677) leaves on a
    TensorMolData_BP_Direct_EE_WithEle plus what the network instance assigns (ele / elep, TFMolInstanceDirect.py:4963-4964)."""

    def __init__(self, mols, eles_np, elep_np, P, with_grad):
        self.set = type("S", (), {})()
        self.set.mols = mols
        self.Nmols = len(mols)
        self.MaxNAtoms = max(m.NAtoms() for m in mols)
        self.dig = type("D", (), {"OType": "EnergyAndDipole"})()
        self.HasGrad = with_grad
        self.TestRatio = P["TestRatio"]
        self.ScratchState = 0
        self.ScratchPointer = 0
        self.Rr_cut, self.Ra_cut, self.Ree_cut = P["AN1_r_Rc"], P["AN1_a_Rc"], P["EECutoffOff"]
        self.ele, self.elep = eles_np, elep_np


def train_batch_pins(P, ncases=3, ntrain_calls=5, ncases_test=2, ntest_calls=4, seed=7, with_grad=True):
    """LoadData / LoadDataToScratch of TensorMolData_BP_Direct_EE (Containers/TensorMolData.py:1679-1745) and GetTrainBatch /
    GetTestBatch of TensorMolData_BP_Direct_EE_WithEle (:1860-1904) executed in place on train_set_inputs(), Python's
    `random` seeded.  Returns {name: array}: the shuffled order and every entry of every batch."""
    import random
    base = namespace()
    ns = {"np": np, "random": random, "LOGGER": base["LOGGER"], "NeighborListSet": base["NeighborListSet"], "AUPERDEBYE": base["AUPERDEBYE"],
          "print": lambda *a, **k: None}
    exec(_method_defs("TensorMol/Containers/TensorMolData.py", ["LoadData", "LoadDataToScratch"], "TensorMolData_BP_Direct_EE"), ns)
    exec(_method_defs("TensorMol/Containers/TensorMolData.py", ["GetTrainBatch", "GetTestBatch"], "TensorMolData_BP_Direct_EE_WithEle"), ns)
    mols = []
    for i, d in enumerate(train_set_inputs()):
        m = _PinMol(d["atoms"], d["coords"])
        m.properties = {"atomization": d["atomization"], "dipole": d["dipole"], "gradients": d["gradients"], "serial": i}
        mols.append(m)
    eles_np, elep_np = element_tables(np.concatenate([m.atoms for m in mols]))
    s = _TrainSelf(mols, eles_np, elep_np, P, with_grad)
    for name in ("LoadData", "LoadDataToScratch", "GetTrainBatch", "GetTestBatch"):
        setattr(_TrainSelf, name, ns[name])
    random.seed(seed)
    s.LoadDataToScratch(None)
    out = {"order": np.array([m.properties["serial"] for m in s.set.mols]), "NTrain": np.int64(s.NTrain), "NTest": np.int64(s.NTest)}
    names = ["xyzs", "Zs", "Elabels", "Dlabels"] + (["grads"] if with_grad else []) + ["rad_p_ele", "ang_t_elep", "rad_eep", "mil_jk", "inv_natom"]
    for kind, n, nc, fn in (("train", ntrain_calls, ncases, s.GetTrainBatch), ("test", ntest_calls, ncases_test, s.GetTestBatch)):
        for c in range(n):
            for nm, v in zip(names, fn(nc)):
                out["%s%d_%s" % (kind, c, nm)] = np.asarray(v)
    return out


LOSS_CLASS = "MolInstance_DirectBP_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize"     # defines the loss ops the Dropout class inherits


class _TrainInstance(_FakeInstance):
    """_FakeInstance whose variables are torch leaves and whose `_variable_with_weight_decay` is the reference's own
    (TFInstance.py:277-297, executed in place: it registers the 0.001 * l2 weight-decay terms in the 'losses' collection)."""

    def load(self, net):
        from oracle import tf_shim
        assert not tf_shim._weight_queue and not tf_shim._bias_queue
        ws, bs = [], []
        for z in self.eles:
            for W, b in self._weights[net][int(z)]:
                ws.append(W)
                bs.append(b)
        tf_shim.push_weights(ws)
        tf_shim.push_biases(bs)


def train_graph(batch, eles, hidden, weights, P, scalars, add_ecc=True):
    """One minibatch through TrainPrepare's graph (TFMolInstanceDirect.py:4997-5049) on the torch stand-in: descriptors,
    dipole_inference, energy_inference, tf.gradients, then the reference's loss_op / loss_op_dipole / loss_op_EandG
    (:4860-4901) in that order, so that the 'losses' collection grows as in the reference: total_loss = weight decay + loss,
    total_loss_dipole = that + loss_dipole, total_loss_EandG = that + loss_EandG.  Returns the fetched values and the
    gradients the three train ops hand to Adam: d total_loss / d(all variables), d total_loss_dipole / d(DipoleNet variables),
    d total_loss_EandG / d(EnergyNet variables) (TFMolInstanceDirect.py:2626-2646, 5046-5048).
    batch: [xyzs, Zs, Elabels, Dlabels, grads, rad_p_ele, ang_t_elep, rad_eep, mil_jk, 1/natom] as GetTrainBatch returns.
    weights: {"charge"|"energy": {Z: [(W, b), ...]}} numpy; scalars: EnergyScalar, GradScalar, DipoleScalar."""
    import torch
    from oracle import tf_shim as tf
    ns = graph_namespace(P)
    if "loss_op" not in ns:
        exec(_method_defs("TensorMol/TFNetworks/TFMolInstanceDirect.py", ["loss_op", "loss_op_dipole", "loss_op_EandG"], LOSS_CLASS), ns)
        exec(_method_defs("TensorMol/TFNetworks/TFInstance.py", ["_variable_with_weight_decay"]), ns)
    xyz_np, Zs_np, El, Dl, gl, rad, ang, reep, mil_jk, inv_natom = batch
    nmol, maxn = Zs_np.shape
    T = torch.as_tensor
    leaves = {net: {int(z): [(torch.tensor(np.asarray(W, np.float64), requires_grad=True), torch.tensor(np.asarray(b, np.float64), requires_grad=True))
                             for W, b in weights[net][int(z)]] for z in eles} for net in ("charge", "energy")}
    inst = _TrainInstance(ns, eles, hidden, P, nmol, maxn, leaves)
    _TrainInstance._variable_with_weight_decay = ns["_variable_with_weight_decay"]
    inst.EnergyScalar, inst.GradScalar, inst.DipoleScalar = scalars["EnergyScalar"], scalars["GradScalar"], scalars["DipoleScalar"]
    tf.reset_collections()
    tf.CREATE_GRAPH = True
    try:
        xyzs = torch.tensor(np.asarray(xyz_np, np.float64), requires_grad=True)
        Zt = T(np.asarray(Zs_np, np.int64))
        natom = T(np.asarray(inv_natom, np.float64))
        keep = torch.ones(len(hidden) + 1, dtype=torch.float64)
        Ele, Elep = T(inst.eles_np), T(inst.eles_pairs_np)
        reep_t = T(np.asarray(reep).astype(np.int64))
        sym, idx = ns["TFSymSet_Scattered_Linear_WithEle"](xyzs, Zt, Ele, T(inst.SFPr2), inst.Rr_cut, Elep, T(inst.SFPa2), inst.zeta, inst.eta,
                                                           inst.Ra_cut, T(np.asarray(rad).astype(np.int64)), T(np.asarray(ang).astype(np.int64)),
                                                           T(np.asarray(mil_jk).astype(np.int64)))
        inst.load("charge")
        Ecc, dipole, charge, _ = ns["dipole_inference"](inst, sym, idx, xyzs, natom, P["Elu_Width"], P["EECutoffOff"], reep_t, bool(add_ecc), keep)
        inst.load("energy")
        Etotal, Ebp, Evdw, _, Ebp_atom = ns["energy_inference"](inst, sym, idx, Ecc, xyzs, Zt, Ele, T(inst.C6), T(inst.vdw_R), reep_t,
                                                                P["EECutoffOn"], P["EECutoffOff"], keep)
        gradient = tf.gradients(Etotal, xyzs)
        args = (inst, Etotal, gradient, dipole, T(np.asarray(El, np.float64)), T(np.asarray(gl, np.float64)), T(np.asarray(Dl, np.float64)), natom)
        total, loss, e_loss, g_loss, d_loss = ns["loss_op"](*args)
        total_d, loss_d, _, _, _ = ns["loss_op_dipole"](*args)
        total_eg, loss_eg, _, _, _ = ns["loss_op_EandG"](*args)
        flat = lambda net: [t for z in eles for Wb in leaves[net][int(z)] for t in Wb]          # noqa: E731
        cv, ev = flat("charge"), flat("energy")
        g_all = torch.autograd.grad(total, cv + ev, retain_graph=True, allow_unused=True)
        g_dip = torch.autograd.grad(total_d, cv, retain_graph=True, allow_unused=True)
        g_eg = torch.autograd.grad(total_eg, ev, retain_graph=True, allow_unused=True)
    finally:
        tf.CREATE_GRAPH = False
        tf.reset_collections()

    def unflat(net, gs):
        it = iter(gs)
        out = {}
        for z in eles:
            out[int(z)] = []
            for W, b in leaves[net][int(z)]:
                gW, gb = next(it), next(it)
                out[int(z)].append((np.zeros(W.shape) if gW is None else gW.detach().numpy(), np.zeros(b.shape) if gb is None else gb.detach().numpy()))
        return out

    n = lambda t: t.detach().numpy()       # noqa: E731
    return dict(Etotal=n(Etotal), Ecc=n(Ecc), Evdw=n(Evdw), dipole=n(dipole), charge=n(charge), gradient=n(gradient[0]),
                total_loss=n(total), loss=n(loss), energy_loss=n(e_loss), grads_loss=n(g_loss), dipole_loss=n(d_loss),
                total_loss_dipole=n(total_d), loss_dipole=n(loss_d), total_loss_EandG=n(total_eg), loss_EandG=n(loss_eg),
                grad_train_op={"charge": unflat("charge", g_all[:len(cv)]), "energy": unflat("energy", g_all[len(cv):])},
                grad_train_op_dipole=unflat("charge", g_dip), grad_train_op_EandG=unflat("energy", g_eg))
