"""
TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz.  Run in the build container (needs
/root/reference for the geometries and oracle/_ref/MolEmb*.so, built by `make -C oracle ref`):

    python -m oracle.make_golden

Each fixture holds (a) a geometry taken from the reference's datasets/ directory, (b) pins computed by
the REFERENCE'S OWN native code: neighbour lists (MolEmb.Make_NListNaive, C_API/MolEmb.cpp:1180-1247),
descriptors (MolEmb.Make_ANI1_Sym, :1913-1988) and, for the small cases, descriptor Jacobians
(MolEmb.Make_ANI1_Sym_deri, :1844-1911); (c) outputs of the float64 restatement oracle/oracle_graph.py
with seeded random-init weights (tensormol_b200.engine.random_weights(seed)); plus ref_python_pins.npz = outputs of the
reference's own Python for the table builders, the lattice tessellation, the DSF constants and PhysicalData
(oracle/ref_py.py).
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_graph as og  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from tensormol_b200.engine import descriptor_width, random_weights  # noqa: E402

REF = os.environ.get("TM_REFERENCE", "/root/reference")
ATOI = {"H": 1, "C": 6, "N": 7, "O": 8, "Cl": 17}


def read_xyz_frames(path):
    frames = []
    with open(path) as fh:
        lines = fh.read().split("\n")
    i = 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        n = int(lines[i].split()[0])
        comment = lines[i + 1]
        Z, X = [], []
        for k in range(n):
            p = lines[i + 2 + k].split()
            Z.append(ATOI[p[0]] if p[0] in ATOI else int(p[0]))
            X.append([float(p[1]), float(p[2]), float(p[3])])
        frames.append((np.array(Z, np.int32), np.array(X, np.float64), comment))
        i += 2 + n
    return frames


def symparams(P):
    nAs, nRa, nRr = P["AN1_num_a_As"], P["AN1_num_a_Rs"], P["AN1_num_r_Rs"]
    return {"AN1_r_Rs": np.array([P["AN1_r_Rc"] * i / nRr for i in range(nRr)]),
            "AN1_a_Rs": np.array([P["AN1_a_Rc"] * i / nRa for i in range(nRa)]),
            "AN1_a_As": np.array([2.0 * math.pi * i / nAs for i in range(nAs)]),
            "AN1_num_r_Rs": nRr, "AN1_num_a_Rs": nRa, "AN1_num_a_As": nAs,
            "AN1_r_Rc": P["AN1_r_Rc"], "AN1_a_Rc": P["AN1_a_Rc"], "AN1_eta": P["AN1_eta"], "AN1_zeta": P["AN1_zeta"]}


def csr_from_lists(ll):
    off = np.zeros(len(ll) + 1, np.int64)
    off[1:] = np.cumsum([len(r) for r in ll])
    idx = np.concatenate([np.sort(np.asarray(r, np.int64)) for r in ll]) if off[-1] else np.zeros(0, np.int64)
    return off, idx


def aperiodic_case(name, Z, X, hidden, seed, with_jacobian):
    from oracle.ref_py import _load_molemb
    MolEmb = _load_molemb()          # the reference build, loaded by path (a top-level `MolEmb` is this repo's drop-in)
    P = og.default_params()
    eles = sorted(set(int(z) for z in Z))
    D = descriptor_width(len(eles), P)
    W = random_weights(eles, D, hidden, seed)
    orc = og.Oracle(eles, W, P)
    N = len(Z)
    res = orc.evaluate(X[None], Z[None], np.array([N]))
    out = dict(Z=Z, xyz=X, eles=np.array(eles), hidden=np.array(hidden), seed=seed)
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra"), (P["EECutoffOff"], "ree")):
        off, idx = csr_from_lists(MolEmb.Make_NListNaive(np.ascontiguousarray(X), float(rc), N, 1))
        out[f"ref_nl_{tag}_off"], out[f"ref_nl_{tag}_idx"] = off, idx
    off, idx = csr_from_lists(MolEmb.Make_NListNaive(np.ascontiguousarray(X), float(P["EECutoffOff"]), N, 0))
    out["ref_nl_ree_noperm_off"], out["ref_nl_ree_noperm_idx"] = off, idx
    sp = symparams(P)
    out["ref_sym"] = MolEmb.Make_ANI1_Sym(sp, np.ascontiguousarray(X), Z.astype(np.uint8), np.array(eles, np.uint8), -1)
    if with_jacobian:
        out["ref_sym_deri"] = MolEmb.Make_ANI1_Sym_deri(sp, np.ascontiguousarray(X), Z.astype(np.uint8), np.array(eles, np.uint8), -1)
    for k in ("Etotal", "Ebp", "Ebp_atom", "Ecc", "Evdw", "dipole", "charge", "gradient", "descriptors", "rad_p_ele", "ang_t_elep"):
        out["oracle_" + k] = res[k]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "E", res["Etotal"], "max|F|", np.abs(res["gradient"]).max(),
          "sym pin err", np.abs(out["ref_sym"] - res["descriptors"][0]).max())


def periodic_case(name, Z, X, L, hidden, seed):
    from oracle.ref_py import _load_molemb
    MolEmb = _load_molemb()          # the reference build, loaded by path (a top-level `MolEmb` is this repo's drop-in)
    P = og.default_params()
    eles = sorted(set(int(z) for z in Z))
    D = descriptor_width(len(eles), P)
    W = random_weights(eles, D, hidden, seed)
    orc = og.Oracle(eles, W, P)
    lat = np.eye(3) * L
    Xw = onp.modulo_lattice(lat, X)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), Xw, P["EECutoffOff"])
    nreal = len(Z)
    res = orc.evaluate_periodic(Xt, Zt, nreal)
    ntess = int(round((len(Zt) / nreal) ** (1.0 / 3.0)) - 1) // 2
    out = dict(Z=Z, xyz=Xw, lattice=lat, ntess=ntess, eles=np.array(eles), hidden=np.array(hidden), seed=seed)
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra")):
        off, idx = csr_from_lists(MolEmb.Make_NListNaive(np.ascontiguousarray(Xt), float(rc), nreal, 1))
        out[f"ref_nl_{tag}_off"], out[f"ref_nl_{tag}_idx"] = off, idx
    off, idx = csr_from_lists(MolEmb.Make_NListNaive(np.ascontiguousarray(Xt), float(P["EECutoffOff"]), nreal, 1))
    out["ref_nl_ree_count"] = np.diff(off)
    out["ref_nl_ree_checksum"] = np.array([np.bitwise_xor.reduce(idx[off[i]:off[i + 1]]) if off[i + 1] > off[i] else 0 for i in range(nreal)])
    sp = symparams(P)
    # reference-native descriptor pin on the tessellated coordinates, first nreal rows.
    # Make_ANI1_Sym is O(N^2): restrict it to the atoms that can matter (within Rr of a real atom).
    keep = np.unique(np.concatenate([np.arange(nreal), out["ref_nl_rr_idx"]]))
    sub = MolEmb.Make_ANI1_Sym(sp, np.ascontiguousarray(Xt[keep]), Zt[keep].astype(np.uint8), np.array(eles, np.uint8), -1)
    out["ref_sym"] = sub[:nreal]
    for k in ("Etotal", "Ebp", "Ebp_atom", "Ecc", "Evdw", "dipole", "descriptors", "gradient"):
        out["oracle_" + k] = res[k] if k != "gradient" else res[k][:, :nreal]
    out["oracle_charge"] = res["charge"][:, :nreal]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "ntess", ntess, "E", res["Etotal"], "n_ee", res["n_ee"], "sym pin err", np.abs(out["ref_sym"] - res["descriptors"][0]).max())


def protein_case(name, hidden, seed, nsample=48):
    """Config C5: datasets/2evq.xyz -> OnlyAtoms([1, 6, 7, 8]) (1,568 atoms: the peptide in its explicit
    water), shifted to the positive octant, cell = bounding box, as samples/test_neb.py:91-113 sets it up.  Pins: the
    reference MolEmb's neighbour rows on the 27-image tessellation (4.6 / 3.1 A) and its ANI-1 descriptors for a sample of
    rows (the full matrix is 9.6 MB), plus the float64 oracle's outputs with seeded random-init nets."""
    from oracle.ref_py import _load_molemb
    MolEmb = _load_molemb()          # the reference build, loaded by path (a top-level `MolEmb` is this repo's drop-in)
    Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", "2evq.xyz"))[0]
    keep = np.isin(Z, [1, 6, 7, 8])
    Z, X = Z[keep], X[keep]
    X = X - X.min(0)
    lat = np.diag(X.max(0))
    P = og.default_params()
    eles = sorted(set(int(z) for z in Z))
    W = random_weights(eles, descriptor_width(len(eles), P), hidden, seed)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])      # atoms on the faces stay where they are
    nreal = len(Z)
    assert len(Zt) == 27 * nreal
    res = og.Oracle(eles, W, P).evaluate_periodic(Xt, Zt, nreal)
    out = dict(Z=Z, xyz=X, lattice=lat, ntess=1, eles=np.array(eles), hidden=np.array(hidden), seed=seed)
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra")):
        off, idx = csr_from_lists(MolEmb.Make_NListNaive(np.ascontiguousarray(Xt), float(rc), nreal, 1))
        out[f"ref_nl_{tag}_off"], out[f"ref_nl_{tag}_idx"] = off.astype(np.int32), idx.astype(np.int32)
    rows = np.sort(np.random.default_rng(0).choice(nreal, nsample, replace=False))
    near = np.unique(np.concatenate([np.arange(nreal), out["ref_nl_rr_idx"]]))
    sub = MolEmb.Make_ANI1_Sym(symparams(P), np.ascontiguousarray(Xt[near]), Zt[near].astype(np.uint8), np.array(eles, np.uint8), -1)
    out["sym_rows"] = rows
    out["ref_sym"] = sub[:nreal][rows]
    for k in ("Etotal", "Ebp", "Ebp_atom", "Ecc", "Evdw", "dipole"):
        out["oracle_" + k] = res[k]
    out["oracle_gradient"] = res["gradient"][:, :nreal]
    out["oracle_charge"] = res["charge"][:, :nreal]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
    print(name, "atoms", nreal, "E", res["Etotal"], "n_ee", res["n_ee"], "sym pin err", np.abs(out["ref_sym"] - res["descriptors"][0][rows]).max())


def reference_python_pins():
    """Outputs of the reference's OWN Python (Neighbors.py, Periodic.py Lattice, Util.py DSF, PhysicalData.py) executed
    in place by oracle/ref_py.py on top of the reference's own MolEmb build -> tests/golden/ref_python_pins.npz."""
    import hashlib
    from oracle import ref_py
    out = {}
    for k, v in ref_py.constants().items():
        out["const_" + k] = np.asarray(v)
    ns = ref_py.namespace()
    B = float(ns["BOHRPERA"])
    dsf_in = np.array([[4.6 * B, 15.0 * B, 0.18 / B], [4.4 * B, 15.0 * B, 0.18 / B], [5.0 * B, 12.0 * B, 0.2 / B], [3.0, 25.0, 0.1]])
    out["dsf_in"] = dsf_in
    out["dsf"] = np.array([ns["DSF"](*r) for r in dsf_in])
    out["dsf_gradient"] = np.array([ns["DSF_Gradient"](*r) for r in dsf_in])
    for name, fn in (("h2o_cluster", "H2O_cluster.xyz"), ("morphine", "morphine.xyz")):
        Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", fn))[0]
        rad, ang, mil_jk, jk_max = ref_py.tables_aperiodic(X, Z, 4.6, 3.1)
        out[name + "_rad"] = rad.astype(np.int32)
        out[name + "_ang"] = ang.astype(np.int32)
        out[name + "_mil_jk"] = mil_jk.astype(np.int32)
        out[name + "_jk_max"] = np.int64(jk_max)
    Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", "water_tiny.xyz"))[0]
    lat = 9.3215 * np.eye(3)
    L = ref_py.lattice(lat)
    out["lat"] = lat
    out["lat_min_diameter"] = np.float64(L.latticeMinDiameter)
    Xu = X + 3.0 * np.random.RandomState(0).randn(*X.shape)          # far outside the cell
    out["modulo_in"] = Xu
    out["modulo_out"] = L.ModuloLattice(Xu)
    Xw = L.ModuloLattice(X)
    zt, xt = L.TessLattice(Z.astype(np.uint8), Xw, 15.0)
    out["tess_in"] = Xw
    out["tess_Z"] = np.asarray(zt, np.int32)
    out["tess_xyz_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(xt, np.float64).tobytes()).digest(), np.uint8)
    out["tess_xyz_head"] = np.asarray(xt)[: 3 * len(Z)]
    out["tess_n"] = np.int64(len(zt))
    rad, ang, mil_j, mil_jk = ref_py.tables_periodic(xt, zt, len(Z), 4.6, 3.1, [1, 8])
    out["periodic_rad"] = rad.astype(np.int32)
    out["periodic_ang"] = ang.astype(np.int32)
    out["periodic_mil_j"] = mil_j.astype(np.int32)
    out["periodic_mil_jk"] = mil_jk.astype(np.int32)
    NLEE = ns["NeighborListSetWithImages"](np.asarray(xt)[None].copy(), np.array([len(zt)]), np.array([len(Z)]), False, True,
                                           np.asarray(zt, np.int32)[None].copy())
    ree = np.asarray(NLEE.buildPairsWithBothEleIndex(15.0, np.array([[1], [8]]))).astype(np.int32)
    ree = ree[np.lexsort(ree.T[::-1])]                                # the reference's order is the sweep order: pin the SET
    out["periodic_ee_n"] = np.int64(len(ree))
    out["periodic_ee_sorted_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(ree).tobytes()).digest(), np.uint8)
    # electrostatics: the reference's TensorFlow functions (RawSymFunc.py TFCoulombEluSRDSFLR / TFVdwPolyLR / ...WithEle)
    # executed on the numpy stand-in oracle/tf_shim.py, for seeded neutral charge vectors
    P = og.default_params()
    for name, fn in (("h2o_cluster", "H2O_cluster.xyz"), ("morphine", "morphine.xyz")):
        Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", fn))[0]
        q = 0.3 * np.random.RandomState(3).randn(len(Z))
        q -= q.mean()
        ecc, evdw, ree = ref_py.electrostatics_aperiodic(X, Z, q, P)
        out[name + "_q"] = q
        out[name + "_Ecc"] = np.float64(ecc)
        out[name + "_Evdw"] = np.float64(evdw)
        out[name + "_n_ee"] = np.int64(len(ree))
    q = 0.3 * np.random.RandomState(4).randn(len(Z0 := read_xyz_frames(os.path.join(REF, "datasets", "water_tiny.xyz"))[0][0]))
    q -= q.mean()
    ecc, evdw = ref_py.electrostatics_periodic(xt, zt, len(Z0), q, [1, 8], P)
    out["periodic_q"] = q
    out["periodic_Ecc"] = np.float64(ecc)
    out["periodic_Evdw"] = np.float64(evdw)
    # the whole evaluation graph of the reference instance (symmetry functions, dipole_inference, energy_inference,
    # tf.gradients) executed on the torch stand-in with the same seeded weights as the oracle fixtures
    for name, fn, hidden, seed in (("h2o_cluster", "H2O_cluster.xyz", [64, 48, 32], 0), ("morphine", "morphine.xyz", [96, 64, 64], 1)):
        Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", fn))[0]
        eles = sorted(set(int(z) for z in Z))
        W = ref_py.weights_with_biases(random_weights(eles, descriptor_width(len(eles), P), hidden, seed), 100 + seed)
        r = ref_py.full_graph_aperiodic(X, Z, hidden, W, P)
        for k in ("Etotal", "Ebp", "Ecc", "Evdw", "Ebp_atom", "dipole", "charge", "gradient", "descriptors"):
            out["graph_" + name + "_" + k] = np.asarray(r[k])
    W = ref_py.weights_with_biases(random_weights([1, 8], descriptor_width(2, P), [64, 48, 32], 2), 102)
    r = ref_py.full_graph_periodic(xt, zt, len(Z0), [1, 8], [64, 48, 32], W, P)
    for k in ("Etotal", "Ebp", "Ecc", "Evdw", "dipole"):
        out["graph_periodic_" + k] = np.asarray(r[k])
    out["graph_periodic_Ebp_atom"] = np.asarray(r["Ebp_atom"])[:, : len(Z0)]
    out["graph_periodic_charge"] = np.asarray(r["charge"])[:, : len(Z0)]
    out["graph_periodic_gradient"] = np.asarray(r["gradient"])[:, : len(Z0)]       # the rows the manager keeps (TFMolManage.py:1353)
    # integrators: PeriodicVelocityVerletStep / PeriodicNoseThermostat.step (Simulations/PeriodicMD.py:21-60) on a toy force
    rs = np.random.RandomState(0)
    md_lat = np.array([[6.0, 0, 0], [0.5, 6.5, 0], [0, 0.3, 7.0]])
    md_x0, md_m, md_v0 = rs.rand(9, 3) * 5.0, np.array([0.016, 0.001, 0.001] * 3), 1e-3 * rs.randn(9, 3)
    md = ref_py.md_pins(md_lat, md_x0, md_m, md_v0, 0.2, 8, {"MDTemp": 300.0, "MDdt": 0.2, "MDThermostat": "Nose"})
    out.update({"md_lat": md_lat, "md_x0": md_x0, "md_m": md_m, "md_v0": md_v0, "md_nve": md["nve"], "md_nose": md["nose"],
                "md_nose_v0": md["nose_v0"], "md_ke": md["ke"]})
    tfn = ref_py.tf_namespace({"sigmoid_alpha": P["sigmoid_alpha"], "EECutoffOff": P["EECutoffOff"], "Poly_Width": P["Poly_Width"]})
    out["act_in"] = np.linspace(-2.0, 2.0, 81)
    out["act_out"] = tfn["sigmoid_with_param"](out["act_in"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_python_pins.npz"), **out)
    print("ref_python_pins:", {k: np.asarray(v).shape for k, v in out.items() if np.asarray(v).size > 8})


from oracle.ref_py import host_pin_inputs  # noqa: E402,F401


HOST_PARAM_PREFIXES = ("Opt", "Neb", "GS", "Remove", "MaxBFGS", "Diis", "SDStep")


def host_pins():
    """Outputs of the reference's host drivers off the hot path, executed in place by oracle/ref_py.py:
    ConjGradient, RemoveInvariantForce, GeomOptimizer.Opt, NudgedElasticBand with each solver (default windows and
    windows of 3 so that the history roll-over is exercised), the aperiodic integrator / thermostat steps of
    Simulations/SimpleMD.py, PeriodicForce (energy, force, RDF, RDF_inC, Density, LatticeStep), a seeded PeriodicMonteCarlo
    chain and PeriodicGeomOptimizer.Opt on a toy local force, whole Prop() runs of the eight MD drivers, MolEmb.CountInRange / GetRDF_Bin of the reference build, and the
    xyz text the reference's Mol writes -> tests/golden/ref_host_pins.npz."""
    from oracle import ref_py
    from tensormol_b200 import PARAMS
    out = {}
    P = {k: PARAMS[k] for k in PARAMS if k.startswith(HOST_PARAM_PREFIXES)}
    atoms, x0, x1 = host_pin_inputs()
    for k, v in ref_py.opt_pins(atoms, x0, x1, P, 6).items():
        out["opt_" + k] = v
    P3 = dict(P)
    P3["MaxBFGS"], P3["DiisSize"] = 3, 3
    for k, v in ref_py.opt_pins(atoms, x0, x1, P3, 9).items():
        if k in ("neb_BFGS", "neb_DIIS"):
            out["opt3_" + k] = v
    from tensormol_b200.PhysicalData import ATOMICMASSES
    md_m = np.array([ATOMICMASSES[z - 1] for z in atoms])
    md_v0 = 1e-3 * np.random.RandomState(3).randn(5, 3)
    PM = {k: PARAMS[k] for k in PARAMS if k.startswith("MD")}
    PM["MDTemp"], PM["MDdt"] = 300.0, 0.2
    out["smd_m"], out["smd_v0"] = md_m, md_v0
    for k, v in ref_py.simple_md_pins(x0, md_m, md_v0, 0.2, 8, PM).items():
        out["smd_" + k] = v
    out.update(ref_py.harmonic_pins(atoms, x0, P))
    pd_atoms = np.array([1, 1, 8] * 3, np.uint8)
    pd_x0 = np.random.RandomState(2).rand(9, 3) * 6.0
    pd_lat = np.array([[6.0, 0, 0], [0.4, 6.2, 0], [0, 0.3, 6.5]])
    PD = {k: PARAMS[k] for k in PARAMS if k.startswith(HOST_PARAM_PREFIXES + ("MD", "PrintTM"))}
    PD["OptMaxCycles"], PD["MDV0"], PD["MDTemp"] = 12, None, 300.0     # MDV0 "Random" reseeds numpy from the OS (SimpleMD.py:367)
    out["pd_atoms"], out["pd_x0"], out["pd_lat"] = pd_atoms, pd_x0, pd_lat
    for k, v in ref_py.periodic_driver_pins(pd_atoms, pd_x0, pd_lat, PD, 6).items():
        out["pd_" + k] = v
    PMD = dict(PD)
    PMD.update(MDV0=None, MDMaxStep=10, MDdt=0.2, MDTemp=300.0, MDLogTrajectory=False, OptMaxCycles=PARAMS["OptMaxCycles"])
    for k, v in ref_py.md_driver_pins(atoms, x0, pd_lat, PMD).items():
        out["mdd_" + k] = v
    M = ref_py.namespace()["MolEmb"]
    rs = np.random.RandomState(11)
    x = rs.uniform(0.0, 6.0, (30, 3))
    z = np.array([1, 1, 8] * 10, np.uint8)
    out["rdf_x"], out["rdf_z"] = x, z
    out["rdf_bins_8_1"] = np.asarray(M.GetRDF_Bin(x, z, 7.0, 0.1, 6.0, 8, 1), np.int64)
    out["rdf_bins_8_8"] = np.asarray(M.GetRDF_Bin(x, z, 5.0, 0.05, 6.0, 8, 8), np.int64)
    xt = np.concatenate([x, x + 6.0, x - 6.0])
    zt = np.concatenate([z, z, z])
    out["count_8_8"] = M.CountInRange(zt, xt, 30, 8, 8, 5.0, 0.02)
    out["count_8_1"] = M.CountInRange(zt, xt, 30, 8, 1, 6.0, 0.05)
    RMol = ref_py.mol_class()
    m = RMol(np.array([8, 1, 1], np.uint8), np.array([[0.1, 0.2, 0.3], [1.0, -0.25, 1e-7], [-0.75, 0.5, 2.5e-5]]))
    m.properties = {"energy": -76.4, "Step": 3}
    out["xyz_with_properties"] = np.array(m.__str__(True))
    out["xyz_plain"] = np.array(str(m))
    r = RMol()
    r.FromXYZString("3\nComment: ;;;energy -1.5;;;foo bar\nO 0 0 0\nH 1.5*^-3 0 0\nH 0 12.25*^2 0\n")
    out["xyz_parsed_coords"], out["xyz_parsed_atoms"] = r.coords, r.atoms
    out["xyz_parsed_energy"] = np.float64(r.properties["energy"])
    import random
    dm = RMol(np.array([8, 1, 1, 1], np.uint8), np.array([[0., 0, 0], [0.757, 0.586, 0], [-0.757, 0.586, 0], [0.3, 0.3, 0.2]]))
    out["distort_in"] = dm.coords.copy()
    np.random.seed(3)
    random.seed(3)
    dm.Distort(0.3, 0.9)
    out["distort_out"] = dm.coords.copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_host_pins.npz"), **out)
    print("ref_host_pins:", {k: np.asarray(v).shape for k, v in out.items()})


TRAIN_PIN_PARAMS = {"TestRatio": 0.4, "AN1_r_Rc": 4.6, "AN1_a_Rc": 3.1, "EECutoffOff": 15.0}
TRAIN_BATCH_NAMES = ["xyzs", "Zs", "Elabels", "Dlabels", "grads", "rad_p_ele", "ang_t_elep", "rad_eep", "mil_jk", "inv_natom"]
TRAIN_PIN_ELES, TRAIN_PIN_HIDDEN, TRAIN_PIN_SEED = (1, 6, 8), (8, 8, 6), 4
TRAIN_PIN_SCALARS = {"EnergyScalar": 1.0, "GradScalar": 1.0 / 20.0, "DipoleScalar": 1.0}
TRAIN_PIN_GRAPHS = (("train0", True), ("test1", True), ("test0", False))


def train_pins():
    """Batches of the reference's training-style provider (TensorMolData_BP_Direct_EE_WithEle.GetTrainBatch / GetTestBatch
    on top of LoadData / LoadDataToScratch, Containers/TensorMolData.py:1679-1745, 1860-1904), executed in place on the
    seeded set of oracle/ref_py.py:train_set_inputs -> tests/golden/ref_train_pins.npz (data side)."""
    from oracle import ref_py
    out = ref_py.train_batch_pins(dict(TRAIN_PIN_PARAMS))
    for i, d in enumerate(ref_py.train_set_inputs()):
        for k, v in d.items():
            out["in%d_%s" % (i, k)] = np.asarray(v)
    # the training graph on three of those batches: TrainPrepare's losses and the gradients its three train ops hand to Adam
    # (loss_op / loss_op_dipole / loss_op_EandG and _variable_with_weight_decay executed in place on the torch stand-in)
    from oracle.oracle_graph import default_params
    from tensormol_b200.engine import descriptor_width, random_weights
    P = default_params()
    eles, hidden = list(TRAIN_PIN_ELES), list(TRAIN_PIN_HIDDEN)
    W = ref_py.weights_with_biases(random_weights(eles, descriptor_width(len(eles), P), hidden, TRAIN_PIN_SEED), 100 + TRAIN_PIN_SEED)
    for tag, ecc in TRAIN_PIN_GRAPHS:
        batch = [out["%s_%s" % (tag, n)] for n in TRAIN_BATCH_NAMES]
        Q = dict(P)
        Q["AddEcc"] = ecc          # what the step feeds AddEcc_pl: PARAMS["AddEcc"], or False in train_step_dipole / test_dipole
        r = ref_py.train_graph(batch, eles, hidden, W, Q, TRAIN_PIN_SCALARS, add_ecc=ecc)
        pre = "tq_%s_ecc%d_" % (tag, int(ecc))
        for k, v in r.items():
            if not k.startswith("grad_train_op"):
                out[pre + k] = np.asarray(v)
        for op, d in (("all_charge", r["grad_train_op"]["charge"]), ("all_energy", r["grad_train_op"]["energy"]),
                      ("dipole", r["grad_train_op_dipole"]), ("EandG", r["grad_train_op_EandG"])):
            for z in eles:
                for l, (gW, gb) in enumerate(d[z]):
                    if tag == TRAIN_PIN_GRAPHS[0][0]:      # whole gradients for the first graph, Frobenius norms for the others (file size)
                        out["%sg_%s_%d_%d_W" % (pre, op, z, l)] = gW
                        out["%sg_%s_%d_%d_b" % (pre, op, z, l)] = gb
                    else:
                        out["%sgnorm_%s_%d_%d" % (pre, op, z, l)] = np.array([np.linalg.norm(gW), np.linalg.norm(gb)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_train_pins.npz"), **out)
    print("ref_train_pins:", len(out), "arrays; shuffled order", out["order"])


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    if "--only-train" in sys.argv:
        return train_pins()
    if "--only-host" in sys.argv:
        return host_pins()
    if "--only-protein" in sys.argv:
        return protein_case("evq2_periodic", [200, 200, 200], 5)
    reference_python_pins()
    host_pins()
    train_pins()
    Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", "H2O_cluster.xyz"))[0]
    aperiodic_case("h2o_cluster", Z, X, [64, 48, 32], 0, True)
    Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", "morphine.xyz"))[0]
    aperiodic_case("morphine", Z, X, [96, 64, 64], 1, False)
    Z, X, _ = read_xyz_frames(os.path.join(REF, "datasets", "water_tiny.xyz"))[0]
    periodic_case("water_tiny_periodic", Z, X, 9.3215, [64, 48, 32], 2)
    protein_case("evq2_periodic", [200, 200, 200], 5)


if __name__ == "__main__":
    main()
