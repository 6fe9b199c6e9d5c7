"""GPU parity at BASELINE.json's full sizes (config C4: 24,000-atom periodic water box, nets 500^3; config C2: 10,000
molecules), where the float64 oracle cannot evaluate the system directly.  The checks are size-independent properties:

* a 24,000-atom box that is an exact 5x5x5 replication of a 192-atom cell has 125 x the cell's energy, and its
  gradient rows follow from the cell's gradient rows -- and the 192-atom cell with its 125 images (24,000 positions,
  Lattice.TessLattice, Periodic.py:131-168) IS within the oracle's reach, so the full-size result is checked against
  the oracle atom by atom, in the reference's force convention (image rows dropped) and in the folded one;
* translation of the wrapped box, permutation of the atom order, checksums (sum of atom energies, sum of charges);
* neighbour rows of the 648,000-position tessellation against a float64 brute-force scan of sampled centres, bit-exact;
* the folded gradient is the derivative of the periodic energy: zero net force and a central finite difference.

Everything goes through the C-ABI (tensormol_b200.engine.Engine -> libtmolb200.so)."""
import numpy as np
import pytest

from common import ENERGY_RTOL, FORCE_ATOL_HA_BOHR, grad_ha_bohr

pytestmark = pytest.mark.gpu

HIDDEN = [500, 500, 500]
NREP = 5


def _cell_and_box():
    """The 192-atom cell of config C4 and its exact 5^3 replication (24,000 atoms, L = 62.14 A);
    atom (c, a) of the box sits at row c*192 + a with c = (ci*5 + cj)*5 + ck."""
    from oracle import oracle_np as onp
    from tensormol_b200.SystemBuilders import water_box
    Z, X, lat = water_box(4, spacing=3.1072, seed=3, jitter=0.05)
    Xw = onp.modulo_lattice(lat, X)
    offs = np.array([(i, j, k) for i in range(NREP) for j in range(NREP) for k in range(NREP)], np.float64)
    Xb = (Xw[None, :, :] + (offs @ lat)[:, None, :]).reshape(-1, 3)
    Zb = np.tile(Z, NREP ** 3)
    return Z, Xw, lat, Zb, Xb, NREP * lat, offs.astype(np.int64)


def _engine(seed=11):
    from oracle import oracle_graph as og
    from tensormol_b200.engine import Engine, random_weights
    P = og.default_params()
    eng = Engine([1, 8], HIDDEN, P)
    W = random_weights(eng.eles, eng.D, HIDDEN, seed)
    eng.set_weights(W)
    return eng, W, P


@pytest.fixture(scope="module")
def fullsize():
    """One oracle evaluation of the cell and the GPU evaluations of the box that the tests below share."""
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    Z, Xw, lat, Zb, Xb, latb, offs = _cell_and_box()
    eng, W, P = _engine()
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), Xw, P["EECutoffOff"])
    nimg = len(Zt) // len(Z)
    assert nimg == 125                                     # L = 12.43 A < 15 A: two shells of images
    o = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, len(Z))
    # image index -> cell offset, in TessLattice's order (real cell first, then i, j, k loops without (0,0,0))
    img_off = [(0, 0, 0)] + [(i, j, k) for i in range(-2, 3) for j in range(-2, 3) for k in range(-2, 3) if (i, j, k) != (0, 0, 0)]
    g_img = {d: o["gradient"][0].reshape(nimg, len(Z), 3)[n] for n, d in enumerate(img_off)}
    r_ref = eng.evaluate_lattice(Xb, Zb, latb, 1)
    r_fold = eng.evaluate_lattice(Xb, Zb, latb, 1, fold=True)
    return dict(Z=Z, Xw=Xw, lat=lat, Zb=Zb, Xb=Xb, latb=latb, offs=offs, eng=eng, P=P, o=o, g_img=g_img, r_ref=r_ref, r_fold=r_fold)


def test_24k_box_energy_is_125_cell_energies_of_the_oracle(fullsize):
    o, r = fullsize["o"], fullsize["r_ref"]
    n = NREP ** 3
    etot = n * o["Etotal"][0]
    for k in ("Etotal", "Ebp", "Ecc", "Evdw"):
        want = n * o[k][0]
        tol = ENERGY_RTOL * max(abs(want), 1e-3 * abs(etot))
        assert abs(r[k][0] - want) <= tol, f"{k}: got {r[k][0]!r} want {want!r}"
    # atom energies and charges repeat the cell's, copy by copy
    ea = r["Ebp_atom"][0].reshape(n, -1)
    assert np.abs(ea - o["Ebp_atom"][0][None]).max() <= 2 * ENERGY_RTOL * np.abs(o["Ebp_atom"]).max()
    q = r["charge"][0].reshape(n, -1)
    nat = len(fullsize["Z"])
    assert np.abs(q - o["charge"][0][:nat][None]).max() <= 1e-5 * np.abs(o["charge"]).max()
    # checksums: the parts add up, the box is neutral
    assert abs(ea.sum() - r["Ebp"][0]) <= 1e-9 * np.abs(ea).sum()
    assert abs(r["Etotal"][0] - (r["Ebp"][0] + r["Ecc"][0] + r["Evdw"][0])) <= 1e-12 * abs(r["Etotal"][0]) + 1e-12
    assert abs(q.sum()) <= 1e-6 * np.abs(q).sum()
    # the fold flag only changes the gradient
    assert abs(fullsize["r_fold"]["Etotal"][0] - r["Etotal"][0]) <= 1e-7 * abs(etot)


def test_24k_box_folded_gradient_equals_the_oracles_periodic_gradient(fullsize):
    """TM_F_FOLD_IMAGES: every copy of atom a carries the cell's gradient summed over a's image rows."""
    g_fold = sum(fullsize["g_img"].values())
    got = fullsize["r_fold"]["gradient"][0].reshape(NREP ** 3, -1, 3)
    err = np.abs(grad_ha_bohr(got) - grad_ha_bohr(g_fold)[None]).max()
    assert err <= FORCE_ATOL_HA_BOHR, f"folded force error {err:.3e} Ha/Bohr"
    assert np.abs(got - g_fold[None]).max() <= 2e-4 * np.abs(g_fold).max()
    # derivative of a translation-invariant energy: no net force
    assert np.abs(got.reshape(-1, 3).sum(0)).max() <= 1e-4 * np.abs(got).sum() / got.size ** 0.5


def test_24k_box_reference_convention_gradient_atom_by_atom(fullsize):
    """Reference convention: only rows of real atoms receive gradient.  Atom (c, a) of the box is a real row
    for every centre cell c' inside the box, where it plays the role of image c - c' of the cell calculation; so its
    gradient is the sum of the cell's image rows d with c - d inside the box."""
    g_img, offs = fullsize["g_img"], fullsize["offs"]
    nat = len(fullsize["Z"])
    want = np.zeros((NREP ** 3, nat, 3))
    for n, c in enumerate(offs):
        for d, g in g_img.items():
            cp = c - np.array(d)
            if np.all(cp >= 0) and np.all(cp < NREP):
                want[n] += g
    got = fullsize["r_ref"]["gradient"][0].reshape(NREP ** 3, nat, 3)
    err = np.abs(grad_ha_bohr(got) - grad_ha_bohr(want)).max()
    assert err <= FORCE_ATOL_HA_BOHR, f"force error {err:.3e} Ha/Bohr"
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()
    # the two conventions differ where a cell touches the boundary of the box, and only there
    centre = (2 * NREP + 2) * NREP + 2
    assert np.abs(want[centre] - sum(g_img.values())).max() <= 1e-12
    assert np.abs(want[0] - sum(g_img.values())).max() > 1e-3


def test_24k_box_translation_and_permutation(fullsize):
    from oracle import oracle_np as onp
    eng, r = fullsize["eng"], fullsize["r_ref"]
    Zb, Xb, latb = fullsize["Zb"], fullsize["Xb"], fullsize["latb"]
    # jittered copy (no exact replication: every atom in its own environment), evaluated as is, shifted + wrapped, permuted
    rng = np.random.default_rng(5)
    X0 = onp.modulo_lattice(latb, Xb + 0.02 * rng.standard_normal(Xb.shape))
    r0 = eng.evaluate_lattice(X0, Zb, latb, 1)
    assert abs(r0["Etotal"][0] - r["Etotal"][0]) > 1e-4 * abs(r["Etotal"][0])        # it is a different system
    scale_f = np.abs(r0["gradient"]).max()
    X1 = onp.modulo_lattice(latb, X0 + np.array([7.3, -11.9, 29.4]))
    r1 = eng.evaluate_lattice(X1, Zb, latb, 1)
    # the reference convention ties a row's gradient to which partners are images, so a shift changes gradients near the
    # faces; the energy is invariant
    assert abs(r1["Etotal"][0] - r0["Etotal"][0]) <= 5e-6 * abs(r0["Etotal"][0])
    assert np.abs(r1["Ebp_atom"] - r0["Ebp_atom"]).max() <= 1e-5 * np.abs(r0["Ebp_atom"]).max()
    r1f = eng.evaluate_lattice(X1, Zb, latb, 1, fold=True)
    r0f = eng.evaluate_lattice(X0, Zb, latb, 1, fold=True)
    assert np.abs(r1f["gradient"] - r0f["gradient"]).max() <= 2e-5 * scale_f       # the folded gradient is invariant too
    # permutation of the rows (keeps H,H,O counts per element but not the order)
    perm = rng.permutation(len(Zb))
    rp = eng.evaluate_lattice(X0[perm], Zb[perm], latb, 1)
    assert abs(rp["Etotal"][0] - r0["Etotal"][0]) <= 2e-6 * abs(r0["Etotal"][0])
    assert np.abs(rp["gradient"][0] - r0["gradient"][0][perm]).max() <= 2e-5 * scale_f
    assert np.abs(rp["charge"][0] - r0["charge"][0][perm]).max() <= 2e-6 * np.abs(r0["charge"]).max()


def test_24k_box_folded_gradient_is_the_energy_derivative(fullsize):
    """Central finite difference of the periodic energy along a random direction against the folded gradient."""
    from oracle import oracle_np as onp
    eng = fullsize["eng"]
    Zb, Xb, latb = fullsize["Zb"], fullsize["Xb"], fullsize["latb"]
    rng = np.random.default_rng(6)
    X0 = onp.modulo_lattice(latb, Xb + 0.02 * rng.standard_normal(Xb.shape))
    g = eng.evaluate_lattice(X0, Zb, latb, 1, fold=True)["gradient"][0]
    u = g / np.linalg.norm(g)                 # steepest direction: the largest signal for a given step
    h = 0.05                                  # |u| = 1 over 72,000 components: no atom moves more than ~1e-3 A
    assert np.abs(h * u).max() < 5e-3
    ep = eng.evaluate_lattice(onp.modulo_lattice(latb, X0 + h * u), Zb, latb, 1, do_force=False)["Etotal"][0]
    em = eng.evaluate_lattice(onp.modulo_lattice(latb, X0 - h * u), Zb, latb, 1, do_force=False)["Etotal"][0]
    fd = (ep - em) / (2 * h)
    an = float((g * u).sum())
    assert abs(fd - an) <= 2e-3 * abs(an), f"finite difference {fd!r} vs analytic {an!r}"


@pytest.mark.parametrize("rc", [4.6, 3.1])
def test_648k_position_neighbour_rows_vs_bruteforce_sample(fullsize, rc):
    """tm_nlist on the full 27-image tessellation of the 24,000-atom box (648,000 positions): sampled rows equal a float64
    brute-force scan with the reference's accept test sqrt(dx^2+dy^2+dz^2)+1e-13 < rc (MolEmb.cpp:1213-1218), bit-exact."""
    from oracle import oracle_np as onp
    eng, P = fullsize["eng"], fullsize["P"]
    Zb, Xb, latb = fullsize["Zb"], fullsize["Xb"], fullsize["latb"]
    Zt, Xt = onp.tess_lattice(latb, Zb.astype(np.uint8), Xb, P["EECutoffOff"])
    nreal = len(Zb)
    assert len(Zt) == 27 * nreal
    off, idx = eng.nlist(Xt, rc, nreal, 1)
    assert off[0] == 0 and len(off) == nreal + 1 and off[-1] == len(idx)
    assert idx.min() >= 0 and idx.max() < len(Zt)
    rng = np.random.default_rng(7)
    sample = np.concatenate([[0, nreal - 1], rng.choice(nreal, 96, replace=False)])
    for i in sample:
        d = Xt - Xt[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        want = np.nonzero(np.sqrt(d2) + 1e-13 < rc)[0]
        want = want[want != i]
        got = np.sort(idx[off[i]:off[i + 1]])
        assert np.array_equal(got, want), f"row {i}"
    # every real-real pair appears in both rows (DoPerms): the multiset of (i, j) equals that of (j, i)
    rows = np.repeat(np.arange(nreal), np.diff(off))
    rr = idx < nreal
    a = rows[rr] * nreal + idx[rr]
    b = idx[rr] * nreal + rows[rr]
    assert np.array_equal(np.sort(a), np.sort(b))
    # the density of liquid water: SURVEY.md section 8 quotes 40.3 / 10.4 neighbours per atom
    per_atom = len(idx) / nreal
    assert (37.0 < per_atom < 44.0) if rc == 4.6 else (8.0 < per_atom < 15.0)


def test_10k_molecule_batch_members_equal_single_evaluations():
    """Config C2 at its full size: 10,000 geometries of a 40-atom C,H,N,O molecule in one padded call."""
    from conftest import load_golden
    from oracle import oracle_graph as og
    from tensormol_b200.engine import Engine, random_weights
    from tensormol_b200.SystemBuilders import perturbed_molecule_batch
    g = load_golden("morphine")
    P = og.default_params()
    eng = Engine(g["eles"], HIDDEN, P)
    W = random_weights(eng.eles, eng.D, HIDDEN, 4)
    eng.set_weights(W)
    nmol = 10000
    Zs, xyzs = perturbed_molecule_batch(g["Z"], g["xyz"], nmol, sigma=0.05, seed=1)
    nat = np.full(nmol, Zs.shape[1], np.int64)
    r = eng.evaluate(xyzs, Zs, nat)
    assert np.all(np.isfinite(r["Etotal"])) and np.all(np.isfinite(r["gradient"]))
    # three members against the float64 oracle, and against themselves evaluated alone
    pick = [0, 4321, nmol - 1]
    o = og.Oracle(g["eles"], W, P).evaluate(xyzs[pick], Zs[pick], nat[pick])
    for n, m in enumerate(pick):
        assert abs(r["Etotal"][m] - o["Etotal"][n]) <= ENERGY_RTOL * abs(o["Etotal"][n])
        assert np.abs(grad_ha_bohr(r["gradient"][m]) - grad_ha_bohr(o["gradient"][n])).max() <= FORCE_ATOL_HA_BOHR
        r1 = eng.evaluate(xyzs[m:m + 1], Zs[m:m + 1], nat[m:m + 1])
        assert abs(r1["Etotal"][0] - r["Etotal"][m]) <= 2e-6 * abs(r["Etotal"][m])
        assert np.abs(r1["gradient"][0] - r["gradient"][m]).max() <= 1e-6
