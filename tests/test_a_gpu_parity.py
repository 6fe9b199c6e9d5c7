"""GPU parity tests: the CUDA path (through the C-ABI, via tensormol_b200.engine.Engine) against the
float64 oracle (oracle/) and the committed golden fixtures (tests/golden/, which carry pins computed by
the reference's own MolEmb).  Tolerances: tests/common.py (= BASELINE.json north_star)."""
import numpy as np
import pytest

from common import DESC_ABS_FLOOR_ULPS, DESC_RTOL, ENERGY_RTOL, FORCE_ATOL_HA_BOHR, grad_ha_bohr, sort_rows_csr, water_box
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _engine(eles, hidden, seed, gemm_mode=None, params=None):
    from oracle import oracle_graph as og
    from tensormol_b200.engine import Engine, random_weights
    P = og.default_params()
    if params:
        P.update(params)
    eng = Engine(eles, hidden, P)
    W = random_weights(eng.eles, eng.D, hidden, seed)
    eng.set_weights(W)
    if gemm_mode is not None:
        eng.set_gemm_mode(gemm_mode)
    return eng, W, P


def _check_desc(got, want):
    """north_star: fp32 descriptors within 1e-5 relative -- per ENTRY, with an absolute floor of one fp32 ulp of the
    row's largest entry (an entry far below that is below what an fp32 row can resolve next to its neighbours)."""
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64)
    floor = DESC_ABS_FLOOR_ULPS * 2.0 ** -23 * np.abs(want).max(axis=-1, keepdims=True)
    tol = DESC_RTOL * np.abs(want) + floor
    err = np.abs(got - want)
    worst = np.unravel_index(np.argmax(err - tol), err.shape)
    assert np.all(err <= tol), f"descriptor entry {worst}: got {got[worst]:.9e} want {want[worst]:.9e} err {err[worst]:.3e} tol {tol[worst]:.3e}"


def _check_energy(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    tol = ENERGY_RTOL * np.maximum(np.abs(want), 1e-3)   # absolute floor 1e-8 Ha for components that vanish
    assert np.all(np.abs(got - want) <= tol), f"{what}: got {got} want {want}"


def _check_grad(got, want):
    err = np.abs(grad_ha_bohr(got) - grad_ha_bohr(want)).max()
    assert err <= FORCE_ATOL_HA_BOHR, f"force error {err:.3e} Ha/Bohr"
    # and much tighter relative to the force scale, so a sign / factor bug cannot hide under the absolute tolerance
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-4 * scale + 1e-7, f"relative force error {np.abs(got - want).max() / scale:.3e}"


# ---------------------------------------------------------------------------------------- neighbour lists
@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_nlist_matches_reference_molemb(name):
    g = load_golden(name)
    eng, _, P = _engine(g["eles"], list(g["hidden"]), int(g["seed"]))
    N = len(g["Z"])
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra"), (P["EECutoffOff"], "ree")):
        off, idx = eng.nlist(g["xyz"], rc, N, 1)
        assert np.array_equal(off, g[f"ref_nl_{tag}_off"])
        assert np.array_equal(sort_rows_csr(off, idx), g[f"ref_nl_{tag}_idx"])
    off, idx = eng.nlist(g["xyz"], P["EECutoffOff"], N, 0)
    assert np.array_equal(off, g["ref_nl_ree_noperm_off"])
    assert np.array_equal(sort_rows_csr(off, idx), g["ref_nl_ree_noperm_idx"])


def test_nlist_periodic_images_matches_reference_molemb():
    from oracle import oracle_np as onp
    g = load_golden("water_tiny_periodic")
    eng, _, P = _engine(g["eles"], list(g["hidden"]), int(g["seed"]))
    Zt, Xt = onp.tess_lattice(g["lattice"], g["Z"].astype(np.uint8), g["xyz"], P["EECutoffOff"])
    nreal = len(g["Z"])
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra")):
        off, idx = eng.nlist(Xt, rc, nreal, 1)
        assert np.array_equal(off, g[f"ref_nl_{tag}_off"])
        assert np.array_equal(sort_rows_csr(off, idx), g[f"ref_nl_{tag}_idx"])
    off, idx = eng.nlist(Xt, P["EECutoffOff"], nreal, 1)
    assert np.array_equal(np.diff(off), g["ref_nl_ree_count"])
    chk = np.array([np.bitwise_xor.reduce(idx[off[i]:off[i + 1]]) for i in range(nreal)])
    assert np.array_equal(chk, g["ref_nl_ree_checksum"])


@pytest.mark.parametrize("n,nreal,rc,perms", [(0, 0, 4.6, 1), (1, 1, 4.6, 1), (2, 2, 4.6, 0), (700, 700, 4.6, 1), (700, 200, 3.1, 1),
                                               (700, 200, 3.1, 0), (3000, 3000, 4.6, 1), (500, 500, 15.0, 0)])
def test_nlist_random_vs_oracle(n, nreal, rc, perms):
    from oracle import oracle_np as onp
    eng, _, _ = _engine([1, 8], [16], 0)
    rng = np.random.default_rng(n + nreal)
    L = max(3.0, (n / 0.1) ** (1 / 3))
    x = rng.uniform(0, L, (n, 3))
    if n >= 4:   # borderline pairs: exactly rc apart along an axis, and one ulp inside / outside
        x[1] = x[0] + np.array([rc, 0, 0])
        x[2] = x[0] + np.array([0, np.nextafter(rc, 0), 0])
        x[3] = x[0] + np.array([0, 0, rc - 2e-13])
    off, idx = eng.nlist(x, rc, nreal, perms)
    o_off, o_idx = onp.nlist_csr(x, rc, nreal, perms)
    assert np.array_equal(off, o_off)
    assert np.array_equal(sort_rows_csr(off, idx), o_idx)


@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_pairs_triples_ele_tables(name):
    from oracle import oracle_np as onp
    g = load_golden(name)
    eng, _, P = _engine(g["eles"], list(g["hidden"]), int(g["seed"]))
    N = len(g["Z"])
    # a padded set of two molecules: the fixture and its first 2/3 of atoms displaced
    n2 = (2 * N) // 3
    xyzs = np.zeros((2, N, 3))
    Zs = np.zeros((2, N), np.int32)
    xyzs[0], Zs[0] = g["xyz"], g["Z"]
    xyzs[1, :n2], Zs[1, :n2] = g["xyz"][:n2] * 1.03, g["Z"][:n2]
    nat = np.array([N, n2])
    rad, ang, mil_j, mil_jk = eng.pairs_triples_ele(xyzs, Zs, nat, nat, P["AN1_r_Rc"], P["AN1_a_Rc"])
    from oracle import oracle_graph as og
    eles_np, elep_np = og.elements_and_pairs(g["eles"])
    o_rad, o_ang, o_mil_j, o_mil_jk = onp.build_pairs_and_triples_with_ele_index_periodic(xyzs, nat, nat, Zs, P["AN1_r_Rc"], P["AN1_a_Rc"], eles_np, elep_np)
    assert np.array_equal(rad, o_rad.astype(np.int64))
    assert np.array_equal(ang, o_ang.astype(np.int64))
    assert np.array_equal(mil_j, o_mil_j.astype(np.int64))
    assert np.array_equal(mil_jk, o_mil_jk.astype(np.int64))
    # first molecule against the golden oracle tables
    assert np.array_equal(rad[rad[:, 0] == 0], g["oracle_rad_p_ele"])
    assert np.array_equal(ang[ang[:, 0] == 0], g["oracle_ang_t_elep"])
    # ... and against the tables the reference's own Python builds for it (Neighbors.py executed in place, ref_python_pins)
    pins = load_golden("ref_python_pins")
    assert np.array_equal(rad[rad[:, 0] == 0], pins[name + "_rad"])
    assert np.array_equal(ang[ang[:, 0] == 0], pins[name + "_ang"])
    assert np.array_equal(mil_jk[mil_jk[:, 0] == 0], pins[name + "_mil_jk"])


# ---------------------------------------------------------------------------------------- fused evaluation
GEMM_MODES = [0, 1, 2, 3, 4]    # 0 = fp32 FFMA tiles, 1 = tcgen05 split-fp16 (library default: tile width by size), 2 = the same on CTA pairs (cta_group::2), 3 / 4 = 64- / 128-column tiles forced


@pytest.mark.parametrize("mode", GEMM_MODES)
@pytest.mark.parametrize("name", ["h2o_cluster", "morphine"])
def test_eval_aperiodic_golden(name, mode):
    g = load_golden(name)
    eng, _, _ = _engine(g["eles"], list(g["hidden"]), int(g["seed"]), gemm_mode=mode)
    N = len(g["Z"])
    r = eng.evaluate(g["xyz"][None], g["Z"][None], np.array([N]), descriptors=True)
    _check_desc(r["descriptors"][0], g["ref_sym"])             # reference-native pin (MolEmb.Make_ANI1_Sym)
    _check_desc(r["descriptors"][0], g["oracle_descriptors"][0])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], g["oracle_" + k], k)
    assert np.abs(r["charge"] - g["oracle_charge"]).max() <= 1e-5 * max(np.abs(g["oracle_charge"]).max(), 1e-3)
    assert np.abs(r["Ebp_atom"] - g["oracle_Ebp_atom"]).max() <= 1e-5 * np.abs(g["oracle_Ebp_atom"]).max()
    assert np.abs(r["dipole"] - g["oracle_dipole"]).max() <= 1e-5 * max(np.abs(g["oracle_dipole"]).max(), 1e-2)
    _check_grad(r["gradient"], g["oracle_gradient"])


@pytest.mark.parametrize("mode", GEMM_MODES)
def test_eval_periodic_golden_images_and_lattice(mode):
    from oracle import oracle_np as onp
    g = load_golden("water_tiny_periodic")
    eng, _, P = _engine(g["eles"], list(g["hidden"]), int(g["seed"]), gemm_mode=mode)
    nreal = len(g["Z"])
    Zt, Xt = onp.tess_lattice(g["lattice"], g["Z"].astype(np.uint8), g["xyz"], P["EECutoffOff"])
    r1 = eng.evaluate_images(Xt, Zt.astype(np.int32), nreal, descriptors=True)
    r2 = eng.evaluate_lattice(g["xyz"], g["Z"], g["lattice"], int(g["ntess"]), descriptors=True)
    for r in (r1, r2):
        _check_desc(r["descriptors"][0], g["ref_sym"])
        _check_desc(r["descriptors"][0], g["oracle_descriptors"][0])
        for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
            _check_energy(r[k], g["oracle_" + k], k)
        assert np.abs(r["charge"][:, :nreal] - g["oracle_charge"]).max() <= 1e-5 * max(np.abs(g["oracle_charge"]).max(), 1e-3)
        _check_grad(r["gradient"], g["oracle_gradient"])
    assert np.array_equal(r1["charge"][0, nreal:2 * nreal], r1["charge"][0, :nreal])      # tiling (TFMolInstanceDirect.py:5892)
    assert r2["charge"].shape == (1, nreal)


def test_protein_box_periodic_golden_and_wide_nets():
    """Config C5: the 1,568-atom 2evq peptide + water box (C/H/N/O, D = 768, bounding-box cell, atoms on the faces) against the
    fixture (reference MolEmb neighbour rows and descriptor rows, oracle energies / charges / gradient, nets 200^3), then with
    the 2000^3 nets BASELINE.json names for C/H/N/O against the oracle evaluated here."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    g = load_golden("evq2_periodic")
    eles = [int(e) for e in g["eles"]]
    nreal = len(g["Z"])
    eng, _, P = _engine(eles, list(g["hidden"]), int(g["seed"]))
    Zt, Xt = onp.tess_lattice(g["lattice"], g["Z"].astype(np.uint8), g["xyz"], P["EECutoffOff"])
    for rc, tag in ((P["AN1_r_Rc"], "rr"), (P["AN1_a_Rc"], "ra")):
        off, idx = eng.nlist(Xt, rc, nreal, 1)
        assert np.array_equal(off, g[f"ref_nl_{tag}_off"])
        assert np.array_equal(sort_rows_csr(off, idx), g[f"ref_nl_{tag}_idx"])
    r = eng.evaluate_lattice(g["xyz"], g["Z"], g["lattice"], 1, descriptors=True)
    _check_desc(r["descriptors"][0][g["sym_rows"]], g["ref_sym"])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], g["oracle_" + k], k)
    assert np.abs(r["charge"] - g["oracle_charge"]).max() <= 1e-5 * np.abs(g["oracle_charge"]).max()
    _check_grad(r["gradient"], g["oracle_gradient"])
    eng.close()
    hidden = [2000, 2000, 2000]
    eng, W, P = _engine(eles, hidden, 6)
    o = og.Oracle(eles, W, P).evaluate_periodic(Xt, Zt, nreal)
    r = eng.evaluate_lattice(g["xyz"], g["Z"], g["lattice"], 1)
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    _check_grad(r["gradient"], o["gradient"][:, :nreal])


@pytest.mark.parametrize("mode", [1, 3])
def test_wide_nets_long_k_loops(mode):
    """Hidden layers of 1024: K loops of 16 k-blocks = 4 accumulator chunks per tile.  Regression test for a dead-lock of the
    64-column GEMM tile variant on K > 768 (the launcher now keeps that variant to short K loops; mode 3 asks for it)."""
    from oracle import oracle_graph as og
    g = load_golden("morphine")
    hidden = [1024, 1024, 1024]
    eng, W, P = _engine(g["eles"], hidden, 9, gemm_mode=mode)
    N = len(g["Z"])
    r = eng.evaluate(g["xyz"][None], g["Z"][None], np.array([N]))
    o = og.Oracle(g["eles"], W, P).evaluate(g["xyz"][None], g["Z"][None], np.array([N]))
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    _check_grad(r["gradient"], o["gradient"])


def test_eval_set_of_molecules_vs_oracle():
    """EvalBPDirectEEUpdateSet contract: molecules of different size padded to MaxNAtoms."""
    from oracle import oracle_graph as og
    g = load_golden("morphine")
    hidden = [64, 64]
    eng, W, P = _engine(g["eles"], hidden, 5, gemm_mode=1)
    N = len(g["Z"])
    rng = np.random.default_rng(7)
    nmol = 5
    xyzs = np.zeros((nmol, N, 3))
    Zs = np.zeros((nmol, N), np.int32)
    nat = np.array([N, N - 7, N, 3, N - 1])
    for m in range(nmol):
        xyzs[m, :nat[m]] = g["xyz"][:nat[m]] + 0.05 * rng.standard_normal((nat[m], 3))
        Zs[m, :nat[m]] = g["Z"][:nat[m]]
    r = eng.evaluate(xyzs, Zs, nat, descriptors=True)
    o = og.Oracle(g["eles"], W, P).evaluate(xyzs, Zs, nat)
    _check_desc(r["descriptors"], o["descriptors"])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    assert np.abs(r["charge"] - o["charge"]).max() <= 1e-5 * max(np.abs(o["charge"]).max(), 1e-3)   # includes the padded slots (Q11)
    _check_grad(r["gradient"], o["gradient"])


@pytest.mark.parametrize("mode,hidden", [(0, [128, 96, 64]), (1, [128, 96, 64]), (1, [500, 500, 500]), (3, [500, 500, 500]), (4, [500, 500, 500])])
def test_eval_water_box_periodic_vs_oracle(mode, hidden):
    """216-water periodic box (648 atoms, L=18.6 A > 15 A so ntess=1, 27 images)."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    Z, X, lat = water_box(6)
    eng, W, P = _engine([1, 8], hidden, 11, gemm_mode=mode)
    Xw = onp.modulo_lattice(lat, X)
    r = eng.evaluate_lattice(Xw, Z, lat, 1, descriptors=True)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), Xw, P["EECutoffOff"])
    assert len(Zt) == 27 * len(Z)
    o = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, len(Z))
    _check_desc(r["descriptors"][0], o["descriptors"][0])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    _check_grad(r["gradient"], o["gradient"][:, :len(Z)])


def test_fold_images_gradient_is_the_oracles_gradient_summed_over_image_rows():
    """TM_F_FOLD_IMAGES (non-reference option): the gradient rows the reference drops (image rows) are folded
    onto slot % nreal -- descriptor terms by the force kernel, pair terms by full weight on image partners; the result is
    the oracle's autograd gradient summed over each atom's 27 rows, i.e. the derivative of the periodic energy."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    Z, X, lat = water_box(6)
    eng, W, P = _engine([1, 8], [128, 96, 64], 11)
    Xw = onp.modulo_lattice(lat, X)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), Xw, P["EECutoffOff"])
    o = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, len(Z))
    g_fold = o["gradient"][0].reshape(27, len(Z), 3).sum(0)
    assert np.abs(g_fold.sum(0)).max() <= 1e-12                      # translation invariance of the oracle energy
    for r in (eng.evaluate_lattice(Xw, Z, lat, 1, fold=True), eng.evaluate_images(Xt, Zt.astype(np.int32), len(Z), fold=True)):
        _check_energy(r["Etotal"], o["Etotal"], "Etotal")
        _check_grad(r["gradient"][0], g_fold)
    assert np.abs(g_fold - o["gradient"][0, :len(Z)]).max() > 1e-3   # and it is not the reference-convention gradient


def test_energy_only_and_no_ecc_flags():
    from oracle import oracle_graph as og
    g = load_golden("h2o_cluster")
    eng, W, P = _engine(g["eles"], list(g["hidden"]), int(g["seed"]), params={"AddEcc": False})
    N = len(g["Z"])
    r = eng.evaluate(g["xyz"][None], g["Z"][None], np.array([N]), do_force=False)
    o = og.Oracle(g["eles"], W, P).evaluate(g["xyz"][None], g["Z"][None], np.array([N]))
    _check_energy(r["Etotal"], o["Etotal"], "Etotal(AddEcc=False)")
    assert r["Ecc"][0] == 0.0
    assert np.all(r["gradient"] == 0.0)
    r = eng.evaluate(g["xyz"][None], g["Z"][None], np.array([N]))
    _check_grad(r["gradient"], o["gradient"])


def test_unknown_element_and_missing_weights_fail_loudly():
    from oracle import oracle_graph as og
    from tensormol_b200._lib import TMolB200Error
    from tensormol_b200.engine import Engine
    eng = Engine([1, 8], [16], og.default_params())
    x = np.zeros((1, 2, 3))
    x[0, 1, 0] = 1.0
    with pytest.raises(TMolB200Error):
        eng.evaluate(x, np.array([[1, 8]], np.int32), np.array([2]))          # weights not set
    eng2, _, _ = _engine([1, 8], [16], 0)
    with pytest.raises(TMolB200Error):
        eng2.evaluate(x, np.array([[1, 6]], np.int32), np.array([2]))         # carbon is not in eles


def test_eval_lattice_unwrapped_input_falls_back_to_bbox_grid():
    """tm_eval_lattice lays the cell grid out from the lattice (wrapped input); coordinates outside the cell must still give
    the images-mode answer for the same (unwrapped) tessellation, through the bounding-box grid."""
    from oracle import oracle_np as onp
    g = load_golden("water_tiny_periodic")
    eng, _, P = _engine(g["eles"], list(g["hidden"]), int(g["seed"]))
    X = g["xyz"].copy()
    X[::3] += g["lattice"][1]            # every third atom one cell over
    X[1::7] -= 0.4 * g["lattice"][0]
    nreal = len(g["Z"])
    Zt, Xt = onp.tess_lattice(g["lattice"], g["Z"].astype(np.uint8), X, P["EECutoffOff"])
    r1 = eng.evaluate_images(Xt, Zt.astype(np.int32), nreal)
    r2 = eng.evaluate_lattice(X, g["Z"], g["lattice"], int(g["ntess"]))
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r2[k], r1[k], k)
    assert np.abs(r2["gradient"] - r1["gradient"]).max() <= 1e-5 * np.abs(r1["gradient"]).max()
    # and the wrapped call still works on the same context afterwards
    r3 = eng.evaluate_lattice(g["xyz"], g["Z"], g["lattice"], int(g["ntess"]))
    _check_energy(r3["Etotal"], g["oracle_Etotal"], "Etotal")


@pytest.mark.parametrize("nx,shear", [(3, 0.35), (6, 0.2)])
def test_eval_lattice_triclinic_windowed_binning(nx, shear):
    """The lattice path bins only the real atoms and the images inside the cell + halo window (fractional coordinates of a
    general lattice).  A sheared water box must give what the caller-supplied full tessellation gives (same slot ids, same
    image arithmetic), and what the oracle gives on that tessellation: nx = 3 needs two shells of images (125 blocks),
    nx = 6 one shell with most of it outside the window."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
    Z, X, lat = water_box(nx)
    lat = lat.copy()
    lat[1, 0] = shear * lat[0, 0]          # b leans along a, c leans along a and b
    lat[2, 0] = -0.5 * shear * lat[0, 0]
    lat[2, 1] = 0.7 * shear * lat[1, 1]
    X = wrap_into_cell(X, lat)
    hidden = [32, 32]
    eng, W, P = _engine([1, 8], hidden, 5)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])
    nimg = len(Zt) // len(Z)
    ntess = (round(nimg ** (1 / 3)) - 1) // 2
    nreal = len(Z)
    r1 = eng.evaluate_images(Xt, Zt.astype(np.int32), nreal, descriptors=True)
    r2 = eng.evaluate_lattice(X, Z, lat, ntess, descriptors=True)
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r2[k], r1[k], k)
    _check_desc(r2["descriptors"][0], r1["descriptors"][0].astype(np.float64))
    assert np.abs(r2["charge"][0] - r1["charge"][0, :nreal]).max() <= 1e-6 * max(np.abs(r1["charge"]).max(), 1e-3)
    assert np.abs(r2["gradient"] - r1["gradient"]).max() <= 1e-5 * np.abs(r1["gradient"]).max()
    if nx == 3:
        o = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, nreal)
        for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
            _check_energy(r2[k], o[k], k)
        _check_grad(r2["gradient"][0], o["gradient"][0, :nreal])


@pytest.mark.parametrize("delta", [-1e-4, -2e-6, 2e-6, 1e-4])
def test_ee_cutoff_boundary_pair_set_has_no_energy_effect(delta):
    """The 15 A electrostatics pair SET is decided in fp32 on the device (reference: float64, strict, MolEmb.cpp:1217-1218).
    A pair that sits on the boundary may therefore be kept or dropped differently -- the DSF kernel and its slope vanish at
    EECutoffOff and the C6 tail is 5e-10 Hartree there, so this must not be visible: both sides of the boundary against the
    oracle, energies at the usual tolerance with a 1e-9 Hartree floor, forces at 1e-7 Hartree/Bohr."""
    from oracle import oracle_graph as og
    eng, W, P = _engine([1, 8], [32, 32], 2)
    d = P["EECutoffOff"] + delta
    X = np.array([[[0.0, 0.0, 0.0], [0.7, 0.6, 0.0], [-0.7, 0.6, 0.0],
                   [d, 0.0, 0.0], [d + 0.7, 0.6, 0.0], [d - 0.25, -0.9, 0.0]]])
    Z = np.array([[8, 1, 1, 8, 1, 1]], np.int32)
    r = eng.evaluate(X, Z, np.array([6]))
    o = og.Oracle([1, 8], W, P).evaluate(X, Z, np.array([6]))
    for k in ("Etotal", "Ecc", "Evdw"):
        assert abs(r[k][0] - o[k][0]) <= ENERGY_RTOL * abs(o[k][0]) + 1e-9, (k, r[k], o[k])
    assert np.abs(grad_ha_bohr(r["gradient"]) - grad_ha_bohr(o["gradient"])).max() <= 1e-6


@pytest.mark.parametrize("case", ["one_water_small_cell", "single_atom", "long_thin_cell"])
def test_eval_lattice_edge_cells_vs_oracle(case):
    """Windowed binning at its corners: a cell far smaller than the cutoffs (ntess = 3: every atom has 343 images inside the
    window), a single atom (no neighbours inside the descriptor cutoffs, images only at Coulomb range), and a long thin
    triclinic cell (different image counts per axis)."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from tensormol_b200.SystemBuilders import wrap_into_cell
    eng, W, P = _engine([1, 8], [16, 16], 11)
    if case == "one_water_small_cell":
        Z = np.array([1, 1, 8], np.int32)
        X = np.array([[0.76, 0.59, 0.1], [-0.76, 0.59, 0.0], [0.0, 0.0, 0.05]]) + 2.9
        lat = np.eye(3) * 6.0
    elif case == "single_atom":
        Z = np.array([8], np.int32)
        X = np.array([[1.0, 2.0, 3.0]])
        lat = np.eye(3) * 9.0
    else:
        rs = np.random.RandomState(4)
        lat = np.array([[5.2, 0.0, 0.0], [1.1, 7.3, 0.0], [0.4, -0.9, 31.0]])
        nw = 9
        f = rs.rand(nw, 3)
        O = f @ lat
        X = np.concatenate([np.stack([o + [0.76, 0.59, 0.0], o + [-0.76, 0.59, 0.0], o]) for o in O])
        Z = np.tile(np.array([1, 1, 8], np.int32), nw)
    X = wrap_into_cell(X, lat)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])
    n = len(Z)
    ntess = (round((len(Zt) / n) ** (1 / 3)) - 1) // 2
    r = eng.evaluate_lattice(X, Z, lat, ntess, descriptors=True)
    o = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, n)
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    _check_desc(r["descriptors"][0], o["descriptors"][0])
    _check_grad(r["gradient"][0], o["gradient"][0, :n])
    r1 = eng.evaluate_images(Xt, Zt.astype(np.int32), n)
    assert np.abs(r["gradient"] - r1["gradient"]).max() <= 1e-5 * max(np.abs(r1["gradient"]).max(), 1e-6)


def test_graph_replay_equals_eager_device_call():
    """engine.GraphedCall: the captured tm_eval_lattice_dev step, replayed after the positions were changed in place,
    gives the numbers of an eager call on the new positions."""
    import ctypes as C
    import torch
    from tensormol_b200.engine import GraphedCall
    Z, X, lat = water_box(4)
    from oracle import oracle_np as onp
    X = onp.modulo_lattice(lat, X)
    eng, _, _ = _engine([1, 8], [128, 128], 3)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(C.c_void_p(stream.cuda_stream))
    n = len(Z)
    with torch.cuda.stream(stream):
        x = torch.tensor(X, dtype=torch.float64, device=dev)
        z = torch.tensor(Z, dtype=torch.int32, device=dev)
        e = torch.zeros(6, dtype=torch.float64, device=dev)
        g = torch.zeros(n, 3, dtype=torch.float64, device=dev)

        def step():
            eng.evaluate_lattice_dev(C.c_void_p(x.data_ptr()), C.c_void_p(z.data_ptr()), n, lat, 1, C.c_void_p(e.data_ptr()), C.c_void_p(g.data_ptr()))

        gc = GraphedCall(step, stream)
        X2 = onp.modulo_lattice(lat, X + 0.05 * np.random.RandomState(2).randn(*X.shape))
        x.copy_(torch.tensor(X2, dtype=torch.float64))
        gc()
        stream.synchronize()
        e_graph, g_graph = e.cpu().numpy().copy(), g.cpu().numpy().copy()
        step()
        stream.synchronize()
    eng.sync()   # surfaces device flags of the pointer API
    assert np.array_equal(e_graph[:4], e.cpu().numpy()[:4]) or np.allclose(e_graph[:4], e.cpu().numpy()[:4], rtol=1e-7, atol=1e-9)
    assert np.abs(g_graph - g.cpu().numpy()).max() <= 1e-6 * np.abs(g_graph).max()
    r = eng.evaluate_lattice(X2, Z, lat, 1)
    _check_energy(e_graph[0:1], r["Etotal"], "Etotal")


def test_molecule_batch_members_equal_single_evaluations():
    """Config C2 shape (many geometries of one C,H,N,O molecule in one padded set): a size-independent property at a size
    the oracle cannot reach -- a molecule evaluated inside the batch equals the same molecule evaluated alone, and the
    batch is invariant under a permutation of its members."""
    from tensormol_b200.SystemBuilders import perturbed_molecule_batch
    g = load_golden("morphine")
    eng, _, _ = _engine(g["eles"], [128, 128, 128], 4)
    nmol = 3000
    Zs, xyzs = perturbed_molecule_batch(g["Z"], g["xyz"], nmol, sigma=0.05, seed=3)
    nat = np.full(nmol, Zs.shape[1], np.int64)
    r = eng.evaluate(xyzs, Zs, nat)
    assert np.all(np.isfinite(r["Etotal"])) and np.all(np.isfinite(r["gradient"]))
    for m in (0, 1234, nmol - 1):
        r1 = eng.evaluate(xyzs[m:m + 1], Zs[m:m + 1], nat[m:m + 1])
        assert abs(r1["Etotal"][0] - r["Etotal"][m]) <= 2e-6 * abs(r["Etotal"][m])
        assert np.abs(r1["gradient"][0] - r["gradient"][m]).max() <= 1e-6
    perm = np.random.default_rng(0).permutation(nmol)
    rp = eng.evaluate(xyzs[perm], Zs[perm], nat[perm])
    assert np.abs(rp["Etotal"] - r["Etotal"][perm]).max() <= 2e-6 * np.abs(r["Etotal"]).max()
    assert np.abs(rp["gradient"] - r["gradient"][perm]).max() <= 1e-6


def test_host_call_graph_replay_equals_eager(monkeypatch):
    """tm_eval_lattice replays its device work from a CUDA graph from the third same-shaped call on: the replayed
    results on new positions equal those of a context that never uses the graph (TM_NO_GRAPH=1), and a change of the
    call shape (other lattice) drops back to the eager path."""
    from oracle import oracle_np as onp
    Z, X, lat = water_box(4)
    X = onp.modulo_lattice(lat, X)
    eng_g, _, _ = _engine([1, 8], [128, 128], 3)
    monkeypatch.setenv("TM_NO_GRAPH", "1")
    eng_e, _, _ = _engine([1, 8], [128, 128], 3)
    monkeypatch.delenv("TM_NO_GRAPH")
    rs = np.random.RandomState(9)
    for it in range(6):
        Xi = onp.modulo_lattice(lat, X + 0.03 * rs.randn(*X.shape))
        rg = eng_g.evaluate_lattice(Xi, Z, lat, 1)
        re = eng_e.evaluate_lattice(Xi, Z, lat, 1)
        for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
            assert abs(rg[k][0] - re[k][0]) <= 1e-7 * abs(re[k][0]) + 1e-10, (it, k)
        assert np.abs(rg["gradient"] - re["gradient"]).max() <= 1e-6 * np.abs(re["gradient"]).max()
        assert np.abs(rg["charge"] - re["charge"]).max() <= 1e-6
    assert eng_g.timings()["launches"] > 10 and eng_g.timings()["nlist"] == 0.0     # last call came from the graph
    lat2 = lat * 1.01
    X2 = onp.modulo_lattice(lat2, X)
    rg = eng_g.evaluate_lattice(X2, Z, lat2, 1)
    re = eng_e.evaluate_lattice(X2, Z, lat2, 1)
    assert abs(rg["Etotal"][0] - re["Etotal"][0]) <= 1e-7 * abs(re["Etotal"][0])
    assert eng_g.timings()["nlist"] > 0.0                                            # eager again


# ---- the CUDA path against outputs of the REFERENCE'S OWN evaluation graph (RawSymFunc.py symmetry functions,
# dipole_inference, energy_inference, tf.gradients executed unmodified on the torch TF stand-in by oracle/ref_py.py;
# seeded weights with non-zero biases; tests/golden/ref_python_pins.npz) ----
@pytest.mark.parametrize("mode", GEMM_MODES)
@pytest.mark.parametrize("name,hidden,seed", [("h2o_cluster", [64, 48, 32], 0), ("morphine", [96, 64, 64], 1)])
def test_eval_vs_reference_graph_aperiodic(name, hidden, seed, mode):
    from oracle import oracle_graph as og
    from oracle.ref_py import weights_with_biases
    from tensormol_b200.engine import Engine, descriptor_width, random_weights
    p, g = load_golden("ref_python_pins"), load_golden(name)
    P = og.default_params()
    Z, X = g["Z"], g["xyz"]
    eles = sorted(set(int(z) for z in Z))
    W = weights_with_biases(random_weights(eles, descriptor_width(len(eles), P), hidden, seed), 100 + seed)
    eng = Engine(eles, hidden, P)
    eng.set_weights(W)
    eng.set_gemm_mode(mode)
    r = eng.evaluate(X[None], Z[None].astype(np.int32), np.array([len(Z)]), descriptors=True)
    pre = "graph_" + name + "_"
    _check_desc(r["descriptors"][0], p[pre + "descriptors"])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], p[pre + k], k)
    _check_grad(r["gradient"], p[pre + "gradient"])
    assert np.abs(r["charge"] - p[pre + "charge"]).max() <= 1e-5 * max(np.abs(p[pre + "charge"]).max(), 1e-3)
    assert np.abs(r["Ebp_atom"] - p[pre + "Ebp_atom"]).max() <= 1e-5 * np.abs(p[pre + "Ebp_atom"]).max()
    assert np.abs(r["dipole"] - p[pre + "dipole"]).max() <= 1e-5 * max(np.abs(p[pre + "dipole"]).max(), 1e-3)


@pytest.mark.parametrize("mode", GEMM_MODES)
def test_eval_vs_reference_graph_periodic(mode):
    from oracle import oracle_graph as og
    from oracle.ref_py import weights_with_biases
    from tensormol_b200.engine import Engine, descriptor_width, random_weights
    p, g = load_golden("ref_python_pins"), load_golden("water_tiny_periodic")
    P = og.default_params()
    Z = g["Z"]
    nreal = len(Z)
    hidden = [64, 48, 32]
    W = weights_with_biases(random_weights([1, 8], descriptor_width(2, P), hidden, 2), 102)
    eng = Engine([1, 8], hidden, P)
    eng.set_weights(W)
    eng.set_gemm_mode(mode)
    r = eng.evaluate_lattice(p["tess_in"], Z, p["lat"], int(g["ntess"]))
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], p["graph_periodic_" + k], k)
    _check_grad(r["gradient"], p["graph_periodic_gradient"])
    assert np.abs(r["charge"][:, :nreal] - p["graph_periodic_charge"]).max() <= 1e-5 * max(np.abs(p["graph_periodic_charge"]).max(), 1e-3)
    assert np.abs(r["Ebp_atom"] - p["graph_periodic_Ebp_atom"]).max() <= 1e-5 * np.abs(p["graph_periodic_Ebp_atom"]).max()


def test_edge_cases_empty_single_atom_and_capacity():
    """Ragged / degenerate input and the capacity errors: an empty molecule and a lone atom inside a padded set evaluate
    (lone atom = net(0 descriptor), zero force) and agree with the oracle; more than TM_ANG_CAP = 64 atoms inside the
    angular cutoff of one centre fails loudly with TM_ECAP instead of truncating."""
    from oracle import oracle_graph as og
    from tensormol_b200._lib import TMolB200Error
    eng, W, P = _engine([1, 8], [32, 32], 2)
    xyzs = np.zeros((3, 4, 3))
    Zs = np.zeros((3, 4), np.int32)
    nat = np.array([0, 1, 3])
    Zs[1, 0] = 8
    xyzs[2, :3] = [[0.0, 0.0, 0.0], [0.96, 0.0, 0.0], [-0.24, 0.93, 0.0]]
    Zs[2, :3] = [8, 1, 1]
    r = eng.evaluate(xyzs, Zs, nat)
    o = og.Oracle([1, 8], W, P).evaluate(xyzs, Zs, nat)
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    assert r["Etotal"][0] == 0.0 and np.all(r["gradient"][0] == 0.0) and np.all(r["gradient"][1] == 0.0)
    _check_grad(r["gradient"], o["gradient"])
    # 80 hydrogens inside a 1.2 A ball around an oxygen: > 64 angular neighbours
    rs = np.random.RandomState(0)
    pts = rs.randn(80, 3)
    pts *= (1.2 * rs.rand(80, 1) ** (1 / 3)) / np.linalg.norm(pts, axis=1, keepdims=True)
    X = np.concatenate([[[0.0, 0.0, 0.0]], pts])[None]
    Z = np.array([[8] + [1] * 80], np.int32)
    with pytest.raises(TMolB200Error, match="angular"):
        eng.evaluate(X, Z, np.array([81]))
    # the context stays usable afterwards
    r2 = eng.evaluate(xyzs, Zs, nat)
    assert np.allclose(r2["Etotal"], r["Etotal"], rtol=1e-9, atol=0)


GRIDS = [dict(AN1_num_r_Rs=16, AN1_num_a_Rs=4, AN1_num_a_As=4, AN1_eta=3.5, AN1_zeta=6.0),      # small grid, zeta != 8
         dict(AN1_num_r_Rs=24, AN1_num_a_Rs=6, AN1_num_a_As=10, AN1_eta=4.5, AN1_zeta=8.0),     # neither 8 x 8 nor 32
         dict(AN1_r_Rc=5.2, AN1_a_Rc=3.5, AN1_num_r_Rs=32, AN1_num_a_Rs=8, AN1_num_a_As=8)]     # default counts, other cutoffs (fast kernels)


@pytest.mark.gpu
@pytest.mark.parametrize("grid", GRIDS, ids=["r16_a4x4_zeta6", "r24_a6x10", "cutoffs_5.2_3.5"])
def test_other_symmetry_function_grids_vs_oracle(grid):
    """SURVEY a17 (SetANI1Param, TFMolInstanceDirect.py:1262-1267, 1293-1328): grids other than the released 32 / 8 x 8 take the
    general descriptor and force kernels (k_desc / k_force); descriptors, energies, charges and gradient against the
    oracle built with the same PARAMS, for a molecule and for a small periodic box."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from tensormol_b200.SystemBuilders import wrap_into_cell
    g = load_golden("h2o_cluster")
    hidden = [48, 32]
    eng, W, P = _engine(g["eles"], hidden, 11, params=grid)
    N = len(g["Z"])
    r = eng.evaluate(g["xyz"][None], g["Z"][None], np.array([N]), descriptors=True)
    o = og.Oracle(g["eles"], W, P).evaluate(g["xyz"][None], g["Z"][None], np.array([N]))
    assert r["descriptors"].shape[-1] == o["descriptors"].shape[-1] == eng.D
    _check_desc(r["descriptors"], o["descriptors"])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(r[k], o[k], k)
    assert np.abs(r["charge"] - o["charge"]).max() <= 1e-5 * max(np.abs(o["charge"]).max(), 1e-3)
    _check_grad(r["gradient"], o["gradient"])
    Z, X, lat = water_box(3, jitter=0.04)
    Xw = wrap_into_cell(X, lat)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), Xw, P["EECutoffOff"])
    rp = eng.evaluate_lattice(Xw, Z, lat, int(round((len(Zt) / len(Z)) ** (1 / 3.0)) - 1) // 2, descriptors=True)
    op = og.Oracle(g["eles"], W, P).evaluate_periodic(Xt, Zt, len(Z))
    _check_desc(rp["descriptors"][0], op["descriptors"][0])
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        _check_energy(rp[k], op[k], k)
    _check_grad(rp["gradient"], op["gradient"][:, :len(Z)])
