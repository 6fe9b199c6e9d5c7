"""CPU tests that pin the ORACLE (oracle/) before it is trusted as the checker:
  * neighbour lists against lists produced by the reference's own MolEmb.Make_NListNaive (golden fixtures,
    and live against oracle/_ref when that build is present),
  * descriptors and descriptor Jacobians against MolEmb.Make_ANI1_Sym / Make_ANI1_Sym_deri (golden),
  * the index assembly against a literal loop transcription of Neighbors.py on a tiny case,
  * closed-form constants and finite differences of the oracle energy.
"""
import glob
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import oracle_graph as og
from oracle import oracle_np as onp
from tensormol_b200.engine import descriptor_width, random_weights


def _ref_molemb():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "MolEmb*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("MolEmb", so[0])
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:
        return None
    return mod


@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_oracle_nlist_vs_reference_golden(name):
    g = load_golden(name)
    P = og.default_params()
    N = len(g["Z"])
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra"), (P["EECutoffOff"], "ree")):
        off, idx = onp.nlist_csr(g["xyz"], rc, N, 1)
        assert np.array_equal(off, g[f"ref_nl_{tag}_off"]) and np.array_equal(idx, g[f"ref_nl_{tag}_idx"])
    off, idx = onp.nlist_csr(g["xyz"], P["EECutoffOff"], N, 0)
    assert np.array_equal(off, g["ref_nl_ree_noperm_off"]) and np.array_equal(idx, g["ref_nl_ree_noperm_idx"])


def test_oracle_nlist_periodic_vs_reference_golden():
    g = load_golden("water_tiny_periodic")
    P = og.default_params()
    Zt, Xt = onp.tess_lattice(g["lattice"], g["Z"].astype(np.uint8), g["xyz"], P["EECutoffOff"])
    assert len(Zt) == (2 * int(g["ntess"]) + 1) ** 3 * len(g["Z"])
    nreal = len(g["Z"])
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra")):
        off, idx = onp.nlist_csr(Xt, rc, nreal, 1)
        assert np.array_equal(off, g[f"ref_nl_{tag}_off"]) and np.array_equal(idx, g[f"ref_nl_{tag}_idx"])
    off, idx = onp.nlist_csr(Xt, P["EECutoffOff"], nreal, 1)
    assert np.array_equal(np.diff(off), g["ref_nl_ree_count"])


def test_oracle_protein_box_vs_reference_golden():
    """Config C5 (2evq peptide in water, C/H/N/O, bounding-box cell): neighbour rows and sampled ANI-1 descriptor rows of the
    reference's MolEmb, and the stored oracle outputs (a regression pin for the fixture the GPU test compares with)."""
    from tensormol_b200.engine import descriptor_width, random_weights
    g = load_golden("evq2_periodic")
    P = og.default_params()
    nreal = len(g["Z"])
    assert nreal == 1568 and list(g["eles"]) == [1, 6, 7, 8]
    Zt, Xt = onp.tess_lattice(g["lattice"], g["Z"].astype(np.uint8), g["xyz"], P["EECutoffOff"])
    assert len(Zt) == 27 * nreal
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra")):
        off, idx = onp.nlist_csr(Xt, rc, nreal, 1)
        assert np.array_equal(off, g[f"ref_nl_{tag}_off"]) and np.array_equal(idx, g[f"ref_nl_{tag}_idx"])
    eles = [int(e) for e in g["eles"]]
    W = random_weights(eles, descriptor_width(4, P), list(g["hidden"]), int(g["seed"]))
    o = og.Oracle(eles, W, P).evaluate_periodic(Xt, Zt, nreal)
    assert np.abs(o["descriptors"][0][g["sym_rows"]] - g["ref_sym"]).max() < 1e-11
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        assert abs(o[k][0] - g["oracle_" + k][0]) <= 1e-10 * max(1.0, abs(o[k][0]))
    assert np.abs(o["gradient"][:, :nreal] - g["oracle_gradient"]).max() < 1e-10


def test_oracle_nlist_vs_live_reference_build():
    M = _ref_molemb()
    if M is None:
        pytest.skip("oracle/_ref/MolEmb not built here (make -C oracle ref)")
    rng = np.random.default_rng(5)
    for n, nreal, rc, perms in [(300, 300, 4.6, 1), (300, 100, 3.1, 1), (300, 100, 3.1, 0), (900, 900, 4.6, 0), (64, 64, 15.0, 1)]:
        x = rng.uniform(0, (n / 0.1) ** (1 / 3), (n, 3))
        x[1] = x[0] + np.array([rc, 0, 0])
        x[2] = x[0] + np.array([0, np.nextafter(rc, 0), 0])
        ref = M.Make_NListNaive(np.ascontiguousarray(x), float(rc), int(nreal), int(perms))
        mine = onp.make_nlist_naive(x, rc, nreal, perms)
        assert [sorted(r) for r in ref] == mine


@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_oracle_descriptors_vs_reference_molemb(name):
    g = load_golden(name)
    P = og.default_params()
    W = random_weights(list(g["eles"]), descriptor_width(len(g["eles"]), P), list(g["hidden"]), int(g["seed"]))
    o = og.Oracle(g["eles"], W, P).evaluate(g["xyz"][None], g["Z"][None], np.array([len(g["Z"])]))
    assert np.abs(o["descriptors"][0] - g["ref_sym"]).max() < 1e-11      # reference-native pin
    # regression against the stored oracle outputs
    for k in ("Etotal", "Ebp", "Ecc", "Evdw", "charge", "gradient"):
        assert np.allclose(o[k], g["oracle_" + k], rtol=1e-10, atol=1e-13), k


def test_oracle_descriptor_jacobian_vs_reference_molemb():
    g = load_golden("h2o_cluster")
    P = og.default_params()
    N = len(g["Z"])
    eles_np, elep_np = og.elements_and_pairs(g["eles"])
    x = g["xyz"][None]
    Zs = g["Z"][None].astype(np.int64)
    nat = np.array([N])
    rp, tt, _, _ = onp.build_pairs_and_triples_with_ele_index(x, nat, nat, Zs, P["AN1_r_Rc"], P["AN1_a_Rc"], eles_np, elep_np)
    R = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    GM = og.descriptors(R, torch.as_tensor(rp.astype(np.int64)), torch.as_tensor(tt.astype(np.int64)), P, len(g["eles"]), elep_np.shape[0], N)
    ref = g["ref_sym_deri"]                      # (N, D, 3N)
    rng = np.random.default_rng(0)
    for _ in range(40):
        i, d = int(rng.integers(N)), int(rng.integers(GM.shape[2]))
        (gr,) = torch.autograd.grad(GM[0, i, d], R, retain_graph=True)
        assert np.abs(gr.numpy().reshape(-1) - ref[i, d]).max() < 1e-8


def test_oracle_forces_are_the_gradient_of_its_energy():
    g = load_golden("h2o_cluster")
    P = og.default_params()
    W = random_weights(list(g["eles"]), descriptor_width(len(g["eles"]), P), [24, 16], 3)
    orc = og.Oracle(g["eles"], W, P)
    N = len(g["Z"])
    x0 = g["xyz"][None].copy()
    o = orc.evaluate(x0, g["Z"][None], np.array([N]))
    rng = np.random.default_rng(1)
    h = 1e-5
    for _ in range(6):
        a, d = int(rng.integers(N)), int(rng.integers(3))
        xp, xm = x0.copy(), x0.copy()
        xp[0, a, d] += h
        xm[0, a, d] -= h
        fd = (orc.energy_only(xp, g["Z"][None], np.array([N]))[0] - orc.energy_only(xm, g["Z"][None], np.array([N]))[0]) / (2 * h)
        assert abs(fd - o["gradient"][0, a, d]) < 1e-8


def test_known_answer_constants():
    orc = og.Oracle([1, 8], random_weights([1, 8], 256, [4], 0))
    assert abs(orc.elu_shift - 0.027735738454600072) < 1e-15
    assert abs(orc.elu_alpha - (-0.009423816979327905)) < 1e-15
    C6, Rv = og.vdw_constants([1, 6, 7, 8])
    assert np.allclose(C6, [2.42833773, 30.35422166, 21.33468151, 12.14168866], rtol=1e-8)
    assert np.allclose(Rv, [1.89161571, 2.74388214, 2.63994721, 2.53601228], rtol=1e-8)
    eles, pairs = og.elements_and_pairs([8, 1, 6])
    assert eles.reshape(-1).tolist() == [1, 6, 8]
    assert pairs.tolist() == [[1, 1], [1, 6], [1, 8], [6, 6], [6, 8], [8, 8]]


def _loop_tables(x, Z, rr, ra, eles, elep):
    """Literal loop transcription of Neighbors.py:117-201 + 344-423 + 440-465 for ONE molecule."""
    N = len(Z)
    pair = onp.make_nlist_naive(x, rr, N, 1)
    tpair = onp.make_nlist_naive(x, ra, N, 1)
    p, t = [], []
    for i in range(N):
        for j in pair[i]:
            p.append([0, i, j])
        for j in tpair[i]:
            for k in tpair[i]:
                if k > j:
                    if Z[j] > Z[k]:
                        t.append([0, i, k, j])
                    else:
                        t.append([0, i, j, k])
    p, t = np.array(p).reshape(-1, 3), np.array(t).reshape(-1, 4)
    el = [int(e) for e in eles]
    pl = [el.index(int(Z[r[2]])) for r in p]
    tl = []
    for r in t:
        for l, (a, b) in enumerate(elep):
            if sorted([int(Z[r[2]]), int(Z[r[3]])]) == sorted([int(a), int(b)]):
                tl.append(l)
    trpE = np.concatenate([p, np.array(pl).reshape(-1, 1)], axis=1)
    trtE = np.concatenate([t, np.array(tl).reshape(-1, 1)], axis=1)
    trpE = trpE[np.lexsort((trpE[:, 2], trpE[:, 3], trpE[:, 1], trpE[:, 0]))]
    trtE = trtE[np.lexsort((trtE[:, 2], trtE[:, 3], trtE[:, 4], trtE[:, 1], trtE[:, 0]))]

    def slots(keys):
        out, prev, c = [], None, 0
        for k in keys:
            k = tuple(k)
            c = c + 1 if k == prev else 0
            prev = k
            out.append(c)
        return np.array(out)

    mil_jk = np.concatenate([trtE[:, [0, 1, 4]], slots(trtE[:, [0, 1, 4]]).reshape(-1, 1)], axis=1)
    mil_j = np.concatenate([trpE[:, [0, 1, 3]], slots(trpE[:, [0, 1, 3]]).reshape(-1, 1)], axis=1)
    return trpE, trtE, mil_j, mil_jk


def test_index_assembly_vs_loop_transcription():
    g = load_golden("h2o_cluster")
    P = og.default_params()
    eles_np, elep_np = og.elements_and_pairs(g["eles"])
    N = len(g["Z"])
    x, Z = g["xyz"], g["Z"].astype(np.int64)
    a = onp.build_pairs_and_triples_with_ele_index_periodic(x[None], np.array([N]), np.array([N]), Z[None], P["AN1_r_Rc"], P["AN1_a_Rc"], eles_np, elep_np)
    b = _loop_tables(x, Z, P["AN1_r_Rc"], P["AN1_a_Rc"], eles_np.reshape(-1), elep_np)
    for u, v in zip(a, b):
        assert np.array_equal(u.astype(np.int64), v)
    assert a[0].dtype == np.float64          # quirk Q16


def test_tessellation_order_and_modulo():
    lat = np.array([[9.0, 0, 0], [0.5, 8.0, 0], [0, 0.3, 7.0]])
    x = np.array([[-1.0, 2.0, 3.0], [10.0, 9.0, 8.0], [4.0, 4.0, 4.0]])
    w = onp.modulo_lattice(lat, x)
    f = w @ np.linalg.inv(lat)
    assert np.all(f >= -1e-12) and np.all(f < 1 + 1e-12)
    Zt, Xt = onp.tess_lattice(lat, np.array([1, 1, 8], np.uint8), w, 3.0)
    assert len(Zt) == 27 * 3 and np.array_equal(Xt[:3], w)
    # first image block is (i,j,k) = (-1,-1,-1)
    assert np.allclose(Xt[3:6], w - lat[0] - lat[1] - lat[2])
    assert np.array_equal(Zt[3:6], [1, 1, 8])


# ---- pins produced by the reference's OWN Python (oracle/ref_py.py executes Neighbors.py, Periodic.py Lattice, Util.py DSF
# and PhysicalData.py where they lie, on the reference's own MolEmb build; oracle/make_golden.py stores the outputs) ----
def _pins():
    return load_golden("ref_python_pins")


@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_oracle_tables_equal_reference_python_aperiodic(name):
    """NeighborListSet(...).buildPairsAndTriplesWithEleIndex (Neighbors.py:344-391, TFMolManage.py:1309-1310): bit-exact."""
    p, g = _pins(), load_golden(name)
    Z, X = g["Z"].astype(np.int32), g["xyz"]
    eles = sorted(set(int(z) for z in Z))
    eles_np = np.asarray(eles).reshape(-1, 1)
    elep_np = np.asarray([[eles[i], eles[j]] for i in range(len(eles)) for j in range(i, len(eles))])
    rad, ang, mil_jk, jk_max = onp.build_pairs_and_triples_with_ele_index(X[None], np.array([len(Z)]), np.array([len(Z)]), Z[None], 4.6, 3.1,
                                                                          eles_np, elep_np)
    assert np.array_equal(np.asarray(rad), p[name + "_rad"])
    assert np.array_equal(np.asarray(ang), p[name + "_ang"])
    assert np.array_equal(np.asarray(mil_jk), p[name + "_mil_jk"])
    assert int(jk_max) == int(p[name + "_jk_max"])


def test_oracle_lattice_and_periodic_tables_equal_reference_python():
    """Lattice.ModuloLattice / TessLattice (Periodic.py:87-168) and buildPairsAndTriplesWithEleIndexPeriodic /
    buildPairsWithBothEleIndex (Neighbors.py:393-470): coordinates bit-exact (sha256 of the float64 bytes), tables bit-exact,
    the 15 A pair list as a set (the reference keeps MolEmb's sweep order)."""
    import hashlib
    p, g = _pins(), load_golden("water_tiny_periodic")
    lat = p["lat"]
    assert np.array_equal(onp.modulo_lattice(lat, p["modulo_in"]), p["modulo_out"])
    assert abs(onp.lattice_min_diameter(lat) - float(p["lat_min_diameter"])) < 1e-9
    Z = g["Z"]
    zt, xt = onp.tess_lattice(lat, Z.astype(np.uint8), p["tess_in"], 15.0)
    assert len(zt) == int(p["tess_n"]) and np.array_equal(np.asarray(zt, np.int32), p["tess_Z"])
    assert np.array_equal(xt[: len(p["tess_xyz_head"])], p["tess_xyz_head"])
    assert hashlib.sha256(np.ascontiguousarray(xt, np.float64).tobytes()).digest() == p["tess_xyz_sha256"].tobytes()
    nreal = len(Z)
    eles_np, elep_np = np.array([[1], [8]]), np.array([[1, 1], [1, 8], [8, 8]])
    o = onp.build_pairs_and_triples_with_ele_index_periodic(xt[None], np.array([len(zt)]), np.array([nreal]), zt.astype(np.int32)[None], 4.6, 3.1,
                                                            eles_np, elep_np)
    for got, key in zip(o, ("periodic_rad", "periodic_ang", "periodic_mil_j", "periodic_mil_jk")):
        assert np.array_equal(np.asarray(got), p[key]), key
    ee = np.asarray(onp.build_pairs_with_both_ele_index(xt[None], np.array([len(zt)]), np.array([nreal]), zt.astype(np.int32)[None], 15.0, eles_np,
                                                        True)).astype(np.int32)
    ee = ee[np.lexsort(ee.T[::-1])]
    assert len(ee) == int(p["periodic_ee_n"])
    assert hashlib.sha256(np.ascontiguousarray(ee).tobytes()).digest() == p["periodic_ee_sorted_sha256"].tobytes()


def test_constants_and_dsf_equal_reference_python():
    """PhysicalData.py tables and Util.py DSF / DSF_Gradient (the ELU shift / slope of the Coulomb kernel) as the host
    layer and the oracle use them."""
    from tensormol_b200 import PhysicalData as PD
    from tensormol_b200.engine import DSF, DSF_Gradient
    p = _pins()
    for k in ("BOHRPERA", "JOULEPERHARTREE", "KCALPERHARTREE", "KJPERHARTREE", "IDEALGASR", "AVOCONST", "AUPERDEBYE"):
        assert getattr(PD, k) == float(p["const_" + k]), k
    for z, c6, rv in zip(p["const_vdw_Z"], p["const_C6_coff"], p["const_atomic_vdw_radius"]):
        assert PD.C6_coff[int(z)] == c6 and PD.atomic_vdw_radius[int(z)] == rv
    mine = np.asarray(PD.ATOMICMASSES, np.float64)      # H..Kr here, the first 100 elements in the reference
    assert len(mine) >= 36 and np.array_equal(mine, p["const_ATOMICMASSES"][: len(mine)])
    for (R, Rc, a), v, dv in zip(p["dsf_in"], p["dsf"], p["dsf_gradient"]):
        assert abs(DSF(R, Rc, a) - v) <= 1e-15 * max(1.0, abs(v))
        assert abs(DSF_Gradient(R, Rc, a) - dv) <= 1e-15 * max(1.0, abs(dv))
    P = og.default_params()
    B = float(p["const_BOHRPERA"])
    sh, al = og.elu_constants(P) if hasattr(og, "elu_constants") else (None, None)
    if sh is not None:
        assert abs(sh - p["dsf"][0]) < 1e-14 and abs(al - p["dsf_gradient"][0]) < 1e-14 and P["Elu_Width"] * B == p["dsf_in"][0][0]


@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_oracle_electrostatics_equal_reference_tf_functions(name):
    """TFCoulombEluSRDSFLR and TFVdwPolyLR (RawSymFunc.py:1307-1414), executed unmodified on the numpy TF stand-in for a
    seeded neutral charge vector and the manager's i<j pair list: the oracle's restatements agree to rounding."""
    from oracle.oracle_graph import BOHRPERA, coulomb_elu_sr_dsf_lr, vdw_poly_lr
    from tensormol_b200.engine import DSF, DSF_Gradient
    p, g = _pins(), load_golden(name)
    P = og.default_params()
    Z, X = g["Z"].astype(np.int32), g["xyz"]
    N = len(Z)
    q = p[name + "_q"]
    ree = np.asarray(onp.build_pairs(X, P["EECutoffOff"], N, False))
    ree = np.concatenate([np.zeros((len(ree), 1), np.int64), ree.astype(np.int64)], axis=1) if ree.shape[1] == 2 else ree.astype(np.int64)
    assert len(ree) == int(p[name + "_n_ee"])
    elu_a = DSF_Gradient(P["Elu_Width"] * BOHRPERA, P["EECutoffOff"] * BOHRPERA, P["DSFAlpha"] / BOHRPERA)
    elu_s = DSF(P["Elu_Width"] * BOHRPERA, P["EECutoffOff"] * BOHRPERA, P["DSFAlpha"] / BOHRPERA)
    Rb = torch.tensor(X[None] * BOHRPERA)
    pr = torch.tensor(ree)
    Ecc = coulomb_elu_sr_dsf_lr(Rb, torch.tensor(q[None]), P["Elu_Width"] * BOHRPERA, pr, P["DSFAlpha"], elu_a, elu_s, P)
    assert abs(float(Ecc[0]) - float(p[name + "_Ecc"])) <= 1e-13 * abs(float(p[name + "_Ecc"]))
    eles = sorted(set(Z.tolist()))
    C6, vdw_R = og.vdw_constants(eles)
    ei = np.array([eles.index(z) for z in Z[ree[:, 1]]])
    ej = np.array([eles.index(z) for z in Z[ree[:, 2]]])
    Evdw = vdw_poly_lr(Rb, torch.as_tensor(C6[ei]), torch.as_tensor(C6[ej]), torch.as_tensor(vdw_R[ei]), torch.as_tensor(vdw_R[ej]),
                       P["EECutoffOn"] * BOHRPERA, pr, P)
    assert abs(float(Evdw[0]) - float(p[name + "_Evdw"])) <= 1e-13 * abs(float(p[name + "_Evdw"]))


def test_oracle_periodic_electrostatics_equal_reference_tf_functions():
    """The periodic forms (TFMolInstanceDirect.py:5824, 5892-5896): charges tiled over the image blocks, TFVdwPolyLRWithEle
    and TFCoulombEluSRDSFLR over the real-centre pair list, both halved."""
    from oracle.oracle_graph import BOHRPERA, coulomb_elu_sr_dsf_lr, vdw_poly_lr
    from tensormol_b200.engine import DSF, DSF_Gradient
    p, g = _pins(), load_golden("water_tiny_periodic")
    P = og.default_params()
    Z = g["Z"]
    nreal = len(Z)
    zt, xt = onp.tess_lattice(p["lat"], Z.astype(np.uint8), p["tess_in"], P["EECutoffOff"])
    eles_np = np.array([[1], [8]])
    ree = np.asarray(onp.build_pairs_with_both_ele_index(xt[None], np.array([len(zt)]), np.array([nreal]), zt.astype(np.int32)[None],
                                                         P["EECutoffOff"], eles_np, True)).astype(np.int64)
    q_all = np.tile(p["periodic_q"][None], (1, len(zt) // nreal))
    elu_a = DSF_Gradient(P["Elu_Width"] * BOHRPERA, P["EECutoffOff"] * BOHRPERA, P["DSFAlpha"] / BOHRPERA)
    elu_s = DSF(P["Elu_Width"] * BOHRPERA, P["EECutoffOff"] * BOHRPERA, P["DSFAlpha"] / BOHRPERA)
    Rb = torch.tensor(xt[None] * BOHRPERA)
    pr = torch.tensor(ree[:, :3])
    Ecc = coulomb_elu_sr_dsf_lr(Rb, torch.tensor(q_all), P["Elu_Width"] * BOHRPERA, pr, P["DSFAlpha"], elu_a, elu_s, P) / 2.0
    assert abs(float(Ecc[0]) - float(p["periodic_Ecc"])) <= 1e-11 * abs(float(p["periodic_Ecc"]))
    C6, vdw_R = og.vdw_constants([1, 8])
    ei, ej = ree[:, 3], ree[:, 4]
    Evdw = vdw_poly_lr(Rb, torch.as_tensor(C6[ei]), torch.as_tensor(C6[ej]), torch.as_tensor(vdw_R[ei]), torch.as_tensor(vdw_R[ej]),
                       P["EECutoffOn"] * BOHRPERA, pr, P) / 2.0
    assert abs(float(Evdw[0]) - float(p["periodic_Evdw"])) <= 1e-11 * abs(float(p["periodic_Evdw"]))


def test_oracle_activation_equals_reference_sigmoid_with_param():
    """Util.py:200-201 executed on the numpy TF stand-in (alpha = 100, |x| <= 2 stays inside float64 range)."""
    p = _pins()
    P = og.default_params()
    got = og.activation(torch.tensor(p["act_in"]), P).numpy()
    assert np.abs(got - p["act_out"]).max() <= 1e-15 + 1e-14 * np.abs(p["act_out"]).max()


# ---- the whole evaluation graph of the reference (its TF symmetry functions, dipole_inference, energy_inference and
# tf.gradients, executed by oracle/ref_py.py on the torch stand-in with seeded weights AND non-zero biases) ----
_GRAPH_CASES = {"h2o_cluster": ([64, 48, 32], 0), "morphine": ([96, 64, 64], 1)}


@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_oracle_equals_reference_graph_aperiodic(name):
    from oracle.ref_py import weights_with_biases
    p, g = _pins(), load_golden(name)
    hidden, seed = _GRAPH_CASES[name]
    P = og.default_params()
    Z, X = g["Z"], g["xyz"]
    eles = sorted(set(int(z) for z in Z))
    W = weights_with_biases(random_weights(eles, descriptor_width(len(eles), P), hidden, seed), 100 + seed)
    o = og.Oracle(eles, W, P).evaluate(X[None], Z[None].astype(np.int32), np.array([len(Z)]))
    pre = "graph_" + name + "_"
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        assert abs(o[k][0] - p[pre + k][0]) <= 1e-12 * max(abs(p[pre + k][0]), 1e-6), k
    assert np.abs(o["gradient"] - p[pre + "gradient"]).max() <= 1e-13
    assert np.abs(o["charge"] - p[pre + "charge"]).max() <= 1e-14
    assert np.abs(o["Ebp_atom"] - p[pre + "Ebp_atom"]).max() <= 1e-13
    assert np.abs(o["dipole"] - p[pre + "dipole"]).max() <= 1e-12
    assert np.abs(o["descriptors"][0] - p[pre + "descriptors"]).max() <= 1e-12


def test_oracle_equals_reference_graph_periodic():
    from oracle.ref_py import weights_with_biases
    p, g = _pins(), load_golden("water_tiny_periodic")
    P = og.default_params()
    Z = g["Z"]
    nreal = len(Z)
    W = weights_with_biases(random_weights([1, 8], descriptor_width(2, P), [64, 48, 32], 2), 102)
    zt, xt = onp.tess_lattice(p["lat"], Z.astype(np.uint8), p["tess_in"], P["EECutoffOff"])
    o = og.Oracle([1, 8], W, P).evaluate_periodic(xt, zt, nreal)
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        assert abs(o[k][0] - p["graph_periodic_" + k][0]) <= 1e-11 * max(abs(p["graph_periodic_" + k][0]), 1e-6), k
    assert np.abs(o["gradient"][0][:nreal] - p["graph_periodic_gradient"][0]).max() <= 1e-13
    assert np.abs(np.asarray(o["charge"])[0][:nreal] - p["graph_periodic_charge"][0]).max() <= 1e-14


# ---- training quantities: the oracle's restatement of TrainPrepare's losses and of the gradients its three train
# ops hand to Adam, against the reference's loss_op / loss_op_dipole / loss_op_EandG / _variable_with_weight_decay executed in
# place on the torch stand-in (oracle/ref_py.py:train_graph -> tests/golden/ref_train_pins.npz) ----
TRAIN_PIN_ELES, TRAIN_PIN_HIDDEN, TRAIN_PIN_SEED = [1, 6, 8], [8, 8, 6], 4       # oracle/make_golden.py
TRAIN_PIN_GRAPHS = (("train0", True), ("test1", True), ("test0", False))
LOSS_KEYS = ("energy_loss", "grads_loss", "dipole_loss", "loss", "loss_dipole", "loss_EandG", "total_loss", "total_loss_dipole", "total_loss_EandG")


def train_pin_weights():
    from oracle.ref_py import weights_with_biases
    P = og.default_params()
    return weights_with_biases(random_weights(TRAIN_PIN_ELES, descriptor_width(len(TRAIN_PIN_ELES), P), TRAIN_PIN_HIDDEN, TRAIN_PIN_SEED),
                               100 + TRAIN_PIN_SEED)


def _train_quantities(g, tag, ecc, W):
    P = og.default_params()
    P["AddEcc"] = ecc
    b = {n: g["%s_%s" % (tag, n)] for n in ("xyzs", "Zs", "Elabels", "Dlabels", "grads", "inv_natom")}
    natom = np.rint(1.0 / b["inv_natom"]).astype(np.int64)
    return og.Oracle(TRAIN_PIN_ELES, W, P).train_quantities(b["xyzs"], b["Zs"], natom, b["Elabels"], b["Dlabels"], b["grads"],
                                                            EnergyScalar=1.0, GradScalar=1.0 / 20.0, DipoleScalar=1.0)


@pytest.mark.parametrize("tag,ecc", TRAIN_PIN_GRAPHS)
def test_oracle_training_losses_and_weight_gradients_equal_reference_graph(tag, ecc):
    g = load_golden("ref_train_pins")
    r = _train_quantities(g, tag, ecc, train_pin_weights())
    pre = "tq_%s_ecc%d_" % (tag, int(ecc))
    for k in LOSS_KEYS + ("Etotal", "dipole", "gradient"):
        ref = g[pre + k]
        assert np.abs(r[k] - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-30), k
    # the 'losses' collection quirk: every total is the sum of all losses registered so far, weight decay included
    assert abs(g[pre + "total_loss_dipole"] - g[pre + "total_loss"] - g[pre + "loss_dipole"]) <= 1e-13
    assert abs(g[pre + "total_loss_EandG"] - g[pre + "total_loss_dipole"] - g[pre + "loss_EandG"]) <= 1e-13
    assert g[pre + "total_loss"] > g[pre + "loss"]
    sets = (("all_charge", r["grad_train_op"]["charge"]), ("all_energy", r["grad_train_op"]["energy"]),
            ("dipole", r["grad_train_op_dipole"]), ("EandG", r["grad_train_op_EandG"]))
    for op, d in sets:
        for z in TRAIN_PIN_ELES:
            for l, (gW, gb) in enumerate(d[z]):
                if pre + "g_%s_%d_%d_W" % (op, z, l) in g:
                    rW, rb = g[pre + "g_%s_%d_%d_W" % (op, z, l)], g[pre + "g_%s_%d_%d_b" % (op, z, l)]
                    assert np.abs(gW - rW).max() <= 1e-11 * max(np.abs(rW).max(), 1e-6), (op, z, l)
                    assert np.abs(gb - rb).max() <= 1e-11 * max(np.abs(rb).max(), 1e-6), (op, z, l)
                else:
                    nW, nb = g[pre + "gnorm_%s_%d_%d" % (op, z, l)]
                    assert abs(np.linalg.norm(gW) - nW) <= 1e-11 * max(nW, 1e-6) and abs(np.linalg.norm(gb) - nb) <= 1e-11 * max(nb, 1e-6)


def test_oracle_weight_gradient_equals_finite_difference_of_total_loss():
    """An independent check of the double back-pass (the force term of the loss differentiates dE/dx w.r.t. the weights):
    central differences of total_loss in single entries of an energy-net and a charge-net weight matrix."""
    g = load_golden("ref_train_pins")
    W = train_pin_weights()
    r = _train_quantities(g, "train0", True, W)
    for net, z, l, ij in (("energy", 8, 1, (2, 3)), ("charge", 1, 0, (40, 1)), ("energy", 6, 3, (4, 0))):
        h = 1e-5
        vals = []
        for s in (+1, -1):
            Wp = {n: {e: [(a.copy(), b.copy()) for a, b in layers] for e, layers in d.items()} for n, d in W.items()}
            Wp[net][z][l][0][ij] += s * h
            vals.append(float(_train_quantities(g, "train0", True, Wp)["total_loss"]))
        fd = (vals[0] - vals[1]) / (2 * h)
        an = r["grad_train_op"][net][z][l][0][ij]
        assert abs(fd - an) <= 1e-6 * max(abs(an), 1e-3), (net, z, l, fd, an)


def test_adam_known_answers_and_dipole_stage_descends():
    """TensorFlow's Adam (un-vendored third party; restated from its documentation): the first step of any component is
    lr sqrt(1-b2) g / (sqrt(1-b2) |g| + eps) ~ lr sign(g); a constant gradient keeps that step size; and ten train_op_dipole
    steps on one pinned minibatch lower its dipole loss while leaving the EnergyNet untouched."""
    gvec = np.array([0.5, -2.0, 1e-3])
    step, m, v = og.adam_update(gvec, np.zeros(3), np.zeros(3), 1, 0.01)
    want = 0.01 * np.sqrt(1 - 0.999) * gvec / (np.sqrt(1 - 0.999) * np.abs(gvec) + 1e-8)
    assert np.allclose(step, want, rtol=1e-14, atol=0) and np.allclose(step, 0.01 * np.sign(gvec), rtol=1e-3)
    assert np.allclose(m, 0.1 * gvec) and np.allclose(v, 0.001 * gvec ** 2)
    for t in range(2, 6):
        step, m, v = og.adam_update(gvec, m, v, t, 0.01)
        assert np.allclose(step, 0.01 * np.sign(gvec), rtol=1e-3)
    g = load_golden("ref_train_pins")
    P = og.default_params()
    o = og.Oracle(TRAIN_PIN_ELES, train_pin_weights(), P)
    b = {n: g["train0_" + n] for n in ("xyzs", "Zs", "Elabels", "Dlabels", "grads", "inv_natom")}
    natom = np.rint(1.0 / b["inv_natom"]).astype(np.int64)
    e0 = [W.clone() for z in o.eles for W, _ in o.w["energy"][z]]
    state, hist = {}, []
    for _ in range(10):
        r = o.train_step("dipole", state, b["xyzs"], b["Zs"], natom, b["Elabels"], b["Dlabels"], b["grads"], 1e-3)
        hist.append(float(r["dipole_loss"]))
    assert abs(hist[0] - float(g["tq_train0_ecc1_dipole_loss"])) <= 1e-12 * hist[0]
    assert hist[-1] < hist[0] and state["t"] == 10
    assert all(torch.equal(a, W) for a, W in zip(e0, [W for z in o.eles for W, _ in o.w["energy"][z]]))


def test_tangent_pass_weight_gradient_equals_double_backward():
    """The algorithm planned for the device training step: the force term of the weight gradient from ONE
    forward-mode pass of the energy graph along v = w^2 (dE/dx - g_label) followed by an ordinary back-pass, instead of
    differentiating the force back-pass.  Equal to the pinned double-backward gradients of train_op."""
    g = load_golden("ref_train_pins")
    P = og.default_params()
    o = og.Oracle(TRAIN_PIN_ELES, train_pin_weights(), P)
    b = {n: g["train0_" + n] for n in ("xyzs", "Zs", "Elabels", "Dlabels", "grads", "inv_natom")}
    natom = np.rint(1.0 / b["inv_natom"]).astype(np.int64)
    t = o.total_loss_gradient_by_tangent_pass(b["xyzs"], b["Zs"], natom, b["Elabels"], b["Dlabels"], b["grads"])
    for net, op in (("charge", "all_charge"), ("energy", "all_energy")):
        for z in TRAIN_PIN_ELES:
            for l, (gW, gb) in enumerate(t[net][z]):
                rW, rb = g["tq_train0_ecc1_g_%s_%d_%d_W" % (op, z, l)], g["tq_train0_ecc1_g_%s_%d_%d_b" % (op, z, l)]
                assert np.abs(gW - rW).max() <= 1e-10 * max(np.abs(rW).max(), 1e-6), (net, z, l)
                assert np.abs(gb - rb).max() <= 1e-10 * max(np.abs(rb).max(), 1e-6), (net, z, l)


def test_oracle_elu_and_selu_formulas():
    """tf.nn.elu and TFInstance.selu (TFInstance.py:365-369) written out."""
    x = torch.tensor([-3.0, -0.5, 0.0, 0.25, 2.0], dtype=torch.float64)
    P = og.default_params()
    P["NeuronType"] = "elu"
    assert torch.allclose(og.activation(x, P), torch.where(x > 0, x, torch.exp(x) - 1), rtol=1e-15, atol=0)
    P["NeuronType"] = "selu"
    a, s = 1.6732632423543772848170429916717, 1.0507009873554804934193349852946
    assert torch.allclose(og.activation(x, P), s * torch.where(x >= 0, x, a * (torch.exp(x) - 1)), rtol=1e-15, atol=1e-300)


def test_tangent_stage_quantities_are_directional_derivatives():
    """The per-stage tangents the device training kernels will be checked against: dEtotal equals <v, dE/dx> of the pinned
    gradient, and dG / dq equal central differences of the descriptors / charges along v."""
    g = load_golden("ref_train_pins")
    P = og.default_params()
    o = og.Oracle(TRAIN_PIN_ELES, train_pin_weights(), P)
    b = {n: g["train0_" + n] for n in ("xyzs", "Zs", "inv_natom")}
    natom = np.rint(1.0 / b["inv_natom"]).astype(np.int64)
    mask = (np.arange(b["Zs"].shape[1])[None, :] < natom[:, None])[:, :, None]
    v = np.random.default_rng(0).normal(size=b["xyzs"].shape) * mask
    t = o.tangent_stage_quantities(b["xyzs"], b["Zs"], natom, v)
    want = (g["tq_train0_ecc1_gradient"] * v).sum((1, 2))
    assert np.abs(t["dEtotal"] - want).max() <= 1e-12 * np.abs(want).max()
    assert np.abs(t["dEtotal"] - (t["dEbp_atom"].sum(1) + t["dEcc"] + t["dEvdw"])).max() <= 1e-14
    h = 1e-6
    p, m = (o.evaluate(b["xyzs"] + s * h * v, b["Zs"], natom) for s in (1, -1))
    for key, ref in (("dG", "descriptors"), ("dq", "charge")):
        fd = (p[ref] - m[ref]) / (2 * h)
        assert np.abs(t[key] - fd).max() <= 1e-7 * max(np.abs(fd).max(), 1e-3), key
