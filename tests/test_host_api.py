"""CPU tests of the host-side mirror of the reference interface (containers, lattice, drivers, math) on
analytic potentials -- no GPU, no oracle needed except for the lattice cross-check."""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp
from tensormol_b200 import (PARAMS, ConjGradient, GeomOptimizer, Lattice, Mol, MSet, NudgedElasticBand, PeriodicForce,
                            PeriodicVelocityVerlet, RemoveInvariantForce, VelocityVerlet, HarmonicSpectra, FdiffGradient, KineticEnergy)
from tensormol_b200.PhysicalData import ATOMICMASSES, IDEALGASR, JOULEPERHARTREE


@pytest.fixture(autouse=True)
def _results_dir(tmp_path):
    old = PARAMS["results_dir"]
    PARAMS["results_dir"] = str(tmp_path) + "/"
    PARAMS["MDLogTrajectory"] = False
    yield
    PARAMS["results_dir"] = old


def _water():
    return Mol(np.array([1, 1, 8], np.uint8), np.array([[0.757, 0.586, 0.0], [-0.757, 0.586, 0.0], [0.0, 0.0, 0.0]]))


def _spring_potential(pairs, r0, k=0.5):
    """E (Hartree) = sum k (r-r0)^2 ; force returned in J/mol/A like the manager's callbacks."""
    def f(x, DoForce=True):
        E, F = 0.0, np.zeros_like(x)
        for (i, j), r in zip(pairs, r0):
            d = x[i] - x[j]
            n = np.linalg.norm(d)
            E += k * (n - r) ** 2
            g = 2 * k * (n - r) * d / n
            F[i] -= g
            F[j] += g
        if DoForce:
            return E, F * JOULEPERHARTREE
        return E
    return f


def test_mol_xyz_round_trip(tmp_path):
    m = _water()
    m.properties["energy"] = -76.4
    m.properties["Lattice"] = np.eye(3) * 9.3215
    m.WriteXYZfile(str(tmp_path), "w", "w", True)
    m.WriteXYZfile(str(tmp_path), "w", "a", True)
    s = MSet("w", str(tmp_path) + "/", center_=False)
    s.ReadXYZ()
    assert len(s.mols) == 2 and s.MaxNAtoms() == 3 and s.NAtoms() == 6
    assert np.array_equal(s.mols[1].atoms, m.atoms) and np.allclose(s.mols[1].coords, m.coords)
    assert abs(s.mols[0].properties["energy"] + 76.4) < 1e-12
    assert np.allclose(s.mols[0].properties["Lattice"], np.eye(3) * 9.3215)
    assert s.AtomTypes().tolist() == [1, 8]
    s.OnlyAtoms([8])
    assert s.mols[0].NAtoms() == 1


def test_lattice_matches_oracle_restatement():
    lat = np.array([[9.0, 0, 0], [0.5, 8.0, 0], [0, 0.3, 7.0]])
    L = Lattice(lat)
    rng = np.random.default_rng(0)
    x = rng.uniform(-12, 20, (17, 3))
    assert np.allclose(L.ModuloLattice(x), onp.modulo_lattice(lat, x), atol=1e-12)
    assert abs(L.latticeMinDiameter - onp.lattice_min_diameter(lat)) < 1e-12
    z = rng.integers(1, 9, 17).astype(np.uint8)
    w = L.ModuloLattice(x)
    for r in (3.0, 15.0):
        Za, Xa = L.TessLattice(z, w, r)
        Zb, Xb = onp.tess_lattice(lat, z, w, r)
        assert np.array_equal(Za, Zb) and np.array_equal(Xa, Xb)
    assert L.NTess(15.0) == int(15.0 / L.latticeMinDiameter) + 1
    Zt, Xt = L.TessNTimes(z, w, 2)
    assert len(Zt) == 8 * 17 and np.allclose(Xt[17:34], w + lat[2])


def test_remove_invariant_force():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(6, 3))
    f = rng.normal(size=(6, 3))
    m = np.array([1, 1, 8, 6, 7, 1.0])
    x = x - np.einsum("a,ax->x", m, x) / m.sum()          # centre of mass at the origin
    g = RemoveInvariantForce(x, f, m)
    assert np.abs(g.sum(0)).max() < 1e-12                 # no net force
    assert np.abs(np.cross(x, g).sum(0)).max() < 1e-10    # no net torque
    # angular part: mass-weighted torque removed
    PARAMS["RemoveInvariant"] = False
    assert RemoveInvariantForce(x, f, m) is f
    PARAMS["RemoveInvariant"] = True


def test_geom_optimizer_on_springs():
    m = _water()
    pairs, r0 = [(0, 2), (1, 2), (0, 1)], [0.96, 0.96, 1.52]
    f = _spring_potential(pairs, r0)
    PARAMS["OptMaxCycles"] = 200
    PARAMS["OptThresh"] = 1e-5
    out = GeomOptimizer(f).Opt(m, "springs")
    d = [np.linalg.norm(out.coords[i] - out.coords[j]) for i, j in pairs]
    assert np.allclose(d, r0, atol=2e-3)
    assert f(out.coords, False) < 1e-5


def test_velocity_verlet_conserves_energy():
    m = _water()
    f = _spring_potential([(0, 2), (1, 2), (0, 1)], [0.96, 0.96, 1.52], k=0.3)
    PARAMS["MDMaxStep"] = 400
    PARAMS["MDdt"] = 0.2
    PARAMS["MDV0"] = None
    PARAMS["MDThermostat"] = None
    md = VelocityVerlet(None, m, "nve", f)
    md.Prop()
    etot = md.md_log[5:, 4] * 3 + (md.md_log[5:, 5]) * JOULEPERHARTREE      # KE is per atom
    assert np.ptp(etot) < 2e-3 * np.abs(md.md_log[5:, 5] * JOULEPERHARTREE).max() + 50.0
    assert md.md_log[-1, 0] == pytest.approx(399 * 0.2)


def test_nose_thermostat_reaches_temperature():
    np.random.seed(3)
    n = 24
    atoms = np.array([8] * n, np.uint8)
    x = np.random.normal(size=(n, 3)) * 3
    f = lambda y, DoForce=True: (0.0, -0.002 * JOULEPERHARTREE * y) if DoForce else 0.0   # noqa: E731
    PARAMS["MDMaxStep"] = 300
    PARAMS["MDdt"] = 0.5
    PARAMS["MDV0"] = "Random"
    PARAMS["MDTemp"] = 300.0
    PARAMS["MDThermostat"] = "Nose"
    md = VelocityVerlet(None, Mol(atoms, x), "nvt", f)
    T0 = (2. / 3.) * KineticEnergy(md.v, md.m) / IDEALGASR
    assert T0 == pytest.approx(300.0, rel=1e-6)          # per-atom rescale at start
    md.Prop()
    assert md.Tstat.name == "Nose" and np.isfinite(md.KE)
    PARAMS["MDThermostat"] = None
    PARAMS["MDV0"] = None


def test_neb_finds_barrier_on_double_well():
    # one "atom" moving in a 2D double well E = (x^2-1)^2 + 2 y^2, plus a spectator far away
    def f(x, DoForce=True):
        a, b = x[0, 0], x[0, 1]
        E = (a * a - 1) ** 2 + 2 * b * b
        F = np.zeros_like(x)
        F[0, 0] = -4 * a * (a * a - 1)
        F[0, 1] = -4 * b
        return (E, F * JOULEPERHARTREE) if DoForce else E
    PARAMS["NebSolver"] = "Verlet"
    PARAMS["NebNumBeads"] = 9
    PARAMS["OptMaxCycles"] = 200
    PARAMS["RemoveInvariant"] = False
    g0 = Mol(np.array([1], np.uint8), np.array([[-1.0, 0.3, 0.0]]))
    g1 = Mol(np.array([1], np.uint8), np.array([[1.0, 0.3, 0.0]]))
    neb = NudgedElasticBand(f, g0, g1, "dw", thresh_=1e-3)
    neb.Opt("dw")
    PARAMS["RemoveInvariant"] = True
    assert np.max(neb.Es) == pytest.approx(1.0, abs=0.15)            # barrier height of the double well
    assert abs(neb.beads[4, 0, 0]) < 0.15                            # middle bead sits near the saddle


def test_conj_gradient_and_fdiff():
    f = lambda x, DoForce=True: ((np.sum((x - 1.0) ** 2)), -2 * (x - 1.0)) if DoForce else np.sum((x - 1.0) ** 2)   # noqa: E731
    x = np.zeros((2, 3))
    CG = ConjGradient(f, x)
    for _ in range(30):
        x, e, g = CG(x)
    assert np.allclose(x, 1.0, atol=1e-3)
    g = FdiffGradient(lambda y: np.sum(y ** 3), np.ones((2, 3)))
    assert np.allclose(g, 3.0 + 3.0e-4, atol=1e-6)      # forward difference, eps = 1e-4: 3 + 3 eps + eps^2, like the reference's


def test_harmonic_spectra_diatomic():
    k = 0.5    # Hartree/A^2 on the bond
    at = np.array([1, 1], np.uint8)
    x0 = np.array([[0.0, 0, 0], [0.74, 0, 0]])
    E = lambda x: k * (np.linalg.norm(x[0] - x[1]) - 0.74) ** 2   # noqa: E731
    w, v = HarmonicSpectra(E, x0, at)
    assert np.sum(np.abs(w) > 100.0) == 1      # one stretch, five zero modes
    assert w[-1] > 1000.0


def test_periodic_force_wrapper_and_md():
    # LJ-like soft pair force on the tessellated images, computed in numpy
    def lf(z, x, nreal, DoForce=True):
        E, F = 0.0, np.zeros((nreal, 3))
        for i in range(nreal):
            d = x[i] - x
            r = np.linalg.norm(d, axis=1)
            m = (r > 1e-9) & (r < 4.0)
            w = np.where(np.arange(len(x))[m] < nreal, 1.0, 0.5)
            E += 0.5 * np.sum(0.01 * (4.0 - r[m]) ** 2)
            F[i] += np.sum((w * 0.02 * (4.0 - r[m]) / r[m])[:, None] * d[m], axis=0)
        return (E, F * JOULEPERHARTREE) if DoForce else E
    rng = np.random.default_rng(2)
    atoms = np.array([8] * 8, np.uint8)
    x = rng.uniform(0, 6, (8, 3))
    pf = PeriodicForce(Mol(atoms, x), np.eye(3) * 6.0)
    pf.BindForce(lf, 4.0)
    e, f = pf(pf.mol0.coords)
    assert f.shape == (8, 3) and np.isfinite(e)
    e2, _ = pf(pf.mol0.coords + np.array([6.0, 0, 0]))        # invariance under a lattice translation
    assert e2 == pytest.approx(e, rel=1e-10)
    assert 0.9 < pf.Density() / (8 * 15.9994 / 6.02214086e23 / (216e-24)) < 1.1
    PARAMS["MDMaxStep"] = 5
    PARAMS["MDThermostat"] = None
    PARAMS["MDV0"] = None
    md = PeriodicVelocityVerlet(pf, "pmd")
    md.Prop()
    assert md.md_log.shape == (5, 7) and np.all(np.isfinite(md.x))
    assert np.all(pf.lattice.InLat(md.x) >= -1e-9) and np.all(pf.lattice.InLat(md.x) < 1 + 1e-9)


def _soft_periodic_force(n=8, L=6.0, seed=2):
    def lf(z, x, nreal, DoForce=True):
        E, F = 0.0, np.zeros((nreal, 3))
        for i in range(nreal):
            d = x[i] - x
            r = np.linalg.norm(d, axis=1)
            m = (r > 1e-9) & (r < 4.0)
            E += 0.5 * np.sum(0.01 * (4.0 - r[m]) ** 2)
            F[i] += np.sum((0.02 * (4.0 - r[m]) / r[m])[:, None] * d[m], axis=0)
        return (E, F * JOULEPERHARTREE) if DoForce else E
    rng = np.random.default_rng(seed)
    pf = PeriodicForce(Mol(np.array([8] * n, np.uint8), rng.uniform(0, L, (n, 3))), np.eye(3) * L)
    pf.BindForce(lf, 4.0)
    return pf


def test_periodic_boxing_dynamics_reaches_the_target_cell():
    """PeriodicBoxingDynamics (Simulations/PeriodicMD.py:147-215): the cell moves linearly to BoxingLatp_ over BoxingT_ fs,
    fractional coordinates are carried along, the density ends at the target's."""
    from tensormol_b200 import PeriodicBoxingDynamics
    pf = _soft_periodic_force()
    rho0 = pf.Density()
    PARAMS["MDMaxStep"] = 12
    PARAMS["MDdt"] = 0.5
    PARAMS["MDThermostat"] = None
    PARAMS["MDV0"] = None
    target = np.eye(3) * 5.0
    md = PeriodicBoxingDynamics(pf, target, "box", BoxingT_=5.0)
    frac0 = pf.lattice.InLat(md.x)
    md._deform()                                    # t = 0: the cell is still the initial one
    assert np.allclose(pf.lattice.lattice, np.eye(3) * 6.0) and np.allclose(pf.lattice.InLat(md.x), frac0)
    md.Prop()
    assert np.allclose(pf.lattice.lattice, target)  # reached at t = 5 fs (step 10), kept afterwards
    assert pf.Density() == pytest.approx(rho0 * (6.0 / 5.0) ** 3, rel=1e-9)
    assert np.all(np.isfinite(md.x)) and md.md_log.shape == (12, 7)
    f = pf.lattice.InLat(md.x)
    assert np.all(f >= -1e-9) and np.all(f < 1 + 1e-9)
    PARAMS["MDdt"] = 0.1


def test_periodic_annealer_keeps_the_lowest_energy_geometry():
    """PeriodicAnnealer (Simulations/PeriodicMD.py:218-285): starts at rest with dt = 0.1 fs, Nose target temperature follows the
    MDAnnealT0 -> MDAnnealTF schedule, the best geometry so far is kept and written with its lattice."""
    from tensormol_b200 import PeriodicAnnealer
    pf = _soft_periodic_force()
    e0, _ = pf(pf.mol0.coords)
    PARAMS["MDMaxStep"] = 40
    PARAMS["MDAnnealSteps"] = 40
    PARAMS["MDAnnealT0"], PARAMS["MDAnnealTF"], PARAMS["MDAnnealKickBack"] = 20.0, 5.0, 1.0
    PARAMS["MDThermostat"] = "Nose"
    PARAMS["MDV0"] = None
    an = PeriodicAnnealer(pf, "anneal", AnnealThresh_=1e-7)
    assert an.dt == 0.1 and np.all(an.v == 0.0)
    an.Prop()
    assert an.Minx is not None and an.MinE < e0       # the soft repulsion relaxes from the random start
    assert pf(an.Minx)[0] == pytest.approx(an.MinE, rel=1e-9)
    assert abs(an.Tstat.T - (20.0 / 40 * 1 + 5.0 * 39 / 40)) < 25.0    # the schedule ended near MDAnnealTF (kick-backs allowed)
    assert os.path.exists(os.path.join(PARAMS["results_dir"], "PAnnealMin.xyz"))
    PARAMS["MDThermostat"] = None
    PARAMS["MDAnnealSteps"] = 1000
    PARAMS["MDAnnealT0"], PARAMS["MDAnnealTF"] = 20.0, 300.0


def test_integrators_equal_reference_python():
    """PeriodicVelocityVerletStep, PeriodicNoseThermostat.step and KineticEnergy of this package against the reference's own
    functions (Simulations/PeriodicMD.py:21-60, SimpleMD.py:42-129) executed by oracle/ref_py.py on a toy analytic force
    (pins in tests/golden/ref_python_pins.npz)."""
    from conftest import load_golden
    from tensormol_b200 import PARAMS
    from tensormol_b200.ForceModifiers.Periodic import Lattice
    from tensormol_b200.Simulations.PeriodicMD import PeriodicNoseThermostat, PeriodicVelocityVerletStep
    from tensormol_b200.Simulations.SimpleMD import KineticEnergy
    p = load_golden("ref_python_pins")
    lat, x0, m, v0 = p["md_lat"], p["md_x0"], p["md_m"], p["md_v0"]

    class Toy:
        def __init__(self):
            self.lattice = Lattice(lat)

        def __call__(self, x, DoForce=True):
            d = x - x0
            return 0.5 * 3.0e5 * float(np.sum(d * d)) + 1.0e4 * float(np.sum(np.sin(x))), -(3.0e5 * d + 1.0e4 * np.cos(x))

    pf = Toy()
    assert abs(KineticEnergy(v0, m) - float(p["md_ke"])) <= 1e-12 * float(p["md_ke"])
    x, v, a = x0.copy(), v0.copy(), np.zeros_like(x0)
    for row in p["md_nve"]:
        x, v, a, e = PeriodicVelocityVerletStep(pf, a, x, v, m, 0.2)
        assert np.abs(x.ravel() - row[:27]).max() <= 1e-12 and np.abs(v.ravel() - row[27:54]).max() <= 1e-14 and abs(e - row[54]) <= 1e-9 * abs(row[54])
    PARAMS["MDTemp"], PARAMS["MDdt"] = 300.0, 0.2
    vv = v0.copy()
    th = PeriodicNoseThermostat(m, vv)
    assert np.abs(vv - p["md_nose_v0"]).max() <= 1e-15
    x, v, a = x0.copy(), vv, np.zeros_like(x0)
    for row in p["md_nose"]:
        x, v, a, e = th.step(pf, a, x, v, m, 0.2)
        assert np.abs(x.ravel() - row[:27]).max() <= 1e-12 and np.abs(v.ravel() - row[27:54]).max() <= 1e-14
        assert abs(th.eta - row[55]) <= 1e-12 * max(1.0, abs(row[55]))


# ---- host drivers against the reference's own Python / MolEmb build executed in place (pins: tests/golden/ref_host_pins.npz,
# ---- written by oracle/make_golden.py:host_pins through oracle/ref_py.py) ----------------------------------------------

def _host_pin_setup():
    from conftest import load_golden
    from oracle.ref_py import host_pin_inputs
    from oracle.ref_py import toy_surface
    return load_golden("ref_host_pins"), toy_surface, host_pin_inputs()


def test_conj_gradient_and_geom_optimizer_equal_reference_python():
    """RemoveInvariantForce, six ConjGradient iterations (point, force, energy, line-search alpha) and a whole
    GeomOptimizer.Opt run against Math/QuasiNewtonTools.py:350-465,551-577 and Simulations/Opt.py:17-75."""
    p, ts, (atoms, x0, x1) = _host_pin_setup()
    e0, f0 = ts(x0)
    assert np.abs(RemoveInvariantForce(x0, f0, atoms.astype(np.float64)) - p["opt_rif"]).max() <= 1e-12 * np.abs(p["opt_rif"]).max()
    go = GeomOptimizer(ts)
    go.m = Mol(atoms, x0)
    cg = ConjGradient(go.WrappedEForce, x0)
    x = x0.copy()
    for row in p["opt_cg"]:
        x, e, g = cg(x)
        assert np.abs(x.ravel() - row[:15]).max() <= 1e-12 and np.abs(g.ravel() - row[15:30]).max() <= 1e-12
        assert abs(e - row[30]) <= 1e-13 and cg.alpha == row[31]
    old = {k: PARAMS[k] for k in ("OptMaxCycles", "OptThresh")}
    PARAMS["OptMaxCycles"], PARAMS["OptThresh"] = 50, 0.0001       # the package defaults the pins were written with
    try:
        m = GeomOptimizer(ts).Opt(Mol(atoms, x0))
    finally:
        PARAMS.update(old)
    assert m.properties["Step"] == int(p["opt_opt_step"])
    assert np.abs(m.coords - p["opt_opt_coords"]).max() <= 1e-6 and abs(m.properties["Energy"] - float(p["opt_opt_energy"])) <= 1e-9


@pytest.mark.parametrize("solver,key,window", [("Verlet", "opt_neb_Verlet", None), ("CG", "opt_neb_CG", None), ("BFGS", "opt_neb_BFGS", None),
                                               ("DIIS", "opt_neb_DIIS", None), ("BFGS", "opt3_neb_BFGS", 3), ("DIIS", "opt3_neb_DIIS", 3)])
def test_neb_solvers_equal_reference_python(solver, key, window):
    """NudgedElasticBand driven step by step with each of the reference's solvers (Simulations/Neb.py:17-220, Math/BFGS.py,
    Math/DIIS.py): bead positions, NEB forces, bead energies and the integrated profile after every solver call. With
    windows of 3 the quasi-Newton / DIIS histories roll over, including the reference's lagging DIIS overlap table."""
    p, ts, (atoms, x0, x1) = _host_pin_setup()
    old = {k: PARAMS[k] for k in ("NebSolver", "MaxBFGS", "DiisSize")}
    try:
        PARAMS["NebSolver"] = solver
        if window:
            PARAMS["MaxBFGS"] = PARAMS["DiisSize"] = window
        neb = NudgedElasticBand(ts, Mol(atoms, x0), Mol(atoms, x1), nbeads_=7)
        n = 7 * 15
        for row in p[key]:
            neb.beads, e, neb.Fs = neb.Solver(neb.beads)
            neb.IntegrateEnergy()
            neb.TSI = int(np.argmax(neb.Es))
            neb.step += 1
            assert np.abs(neb.beads.ravel() - row[:n]).max() <= 1e-9, solver
            assert np.abs(neb.Fs.ravel() - row[n:2 * n]).max() <= 1e-9
            assert np.abs(neb.Es - row[2 * n:2 * n + 7]).max() <= 1e-10 and np.abs(neb.Esi - row[2 * n + 7:2 * n + 14]).max() <= 1e-10
            assert abs(e - row[-1]) <= 1e-10
    finally:
        PARAMS.update(old)


def test_rdf_helpers_equal_reference_molemb():
    """MolEmb.GetRDF_Bin / CountInRange (C_API/MolEmb.cpp:1082-1178) against the reference build's outputs, list order
    included."""
    from conftest import load_golden
    from tensormol_b200 import MolEmb
    p = load_golden("ref_host_pins")
    x, z = p["rdf_x"], p["rdf_z"]
    assert MolEmb.GetRDF_Bin(x, z, 7.0, 0.1, 6.0, 8, 1) == p["rdf_bins_8_1"].tolist()
    assert MolEmb.GetRDF_Bin(x, z, 5.0, 0.05, 6.0, 8, 8) == p["rdf_bins_8_8"].tolist()
    xt, zt = np.concatenate([x, x + 6.0, x - 6.0]), np.concatenate([z, z, z])
    assert np.array_equal(MolEmb.CountInRange(zt, xt, 30, 8, 8, 5.0, 0.02), p["count_8_8"])
    assert np.array_equal(MolEmb.CountInRange(zt, xt, 30, 8, 1, 6.0, 0.05), p["count_8_1"])


def test_xyz_text_equals_reference_mol():
    """The xyz text of Mol.__str__ / WriteXYZfile (';;;'-separated properties on the comment line) and FromXYZString with
    Mathematica-style exponents against the reference's Mol class (Containers/Mol.py:268-359)."""
    from conftest import load_golden
    p = load_golden("ref_host_pins")
    m = Mol(np.array([8, 1, 1], np.uint8), np.array([[0.1, 0.2, 0.3], [1.0, -0.25, 1e-7], [-0.75, 0.5, 2.5e-5]]))
    m.properties = {"energy": -76.4, "Step": 3}
    assert m.__str__(True) == str(p["xyz_with_properties"]) and str(m) == str(p["xyz_plain"])
    r = Mol()
    r.FromXYZString("3\nComment: ;;;energy -1.5;;;foo bar\nO 0 0 0\nH 1.5*^-3 0 0\nH 0 12.25*^2 0\n")
    assert np.array_equal(r.coords, p["xyz_parsed_coords"]) and np.array_equal(r.atoms, p["xyz_parsed_atoms"])
    assert r.properties["energy"] == float(p["xyz_parsed_energy"])
    # Mol.Distort with both generators seeded (collision-avoiding retries included: Mol.py:160-177)
    import random
    d = Mol(np.array([8, 1, 1, 1], np.uint8), p["distort_in"].copy())
    np.random.seed(3)
    random.seed(3)
    d.Distort(0.3, 0.9)
    assert np.array_equal(d.coords, p["distort_out"]) and np.abs(d.coords - p["distort_in"]).max() > 0.1


def test_aperiodic_integrators_equal_reference_python():
    """VelocityVerletStep and the step functions of the Rescaling, Nose, Andersen and Langevin thermostats against
    Simulations/SimpleMD.py:14-204 executed in place on the toy surface (positions, velocities, accelerations, energy after
    every step; the stochastic ones with numpy's global generator seeded as the generator did). The reference's Langevin
    integrator overflows after a few steps (it is marked "Not Working"); its finite rows are compared."""
    import tensormol_b200.Simulations.SimpleMD as S
    p, ts, (atoms, x0, x1) = _host_pin_setup()
    m, v0 = p["smd_m"], p["smd_v0"]
    force = lambda x: ts(x)[1]      # noqa: E731
    old = {k: PARAMS[k] for k in ("MDTemp", "MDdt")}
    PARAMS["MDTemp"], PARAMS["MDdt"] = 300.0, 0.2
    try:
        x, v, a = x0.copy(), v0.copy(), np.zeros_like(x0)
        for row in p["smd_vv"]:
            x, v, a, e = S.VelocityVerletStep(force, a, x, v, m, 0.2, ts)
            assert np.array_equal(np.concatenate([x.ravel(), v.ravel(), a.ravel(), [e]]), row)
        for name in ("Thermostat", "NoseThermostat", "AndersenThermostat", "LangevinThermostat"):
            np.random.seed(7)
            vv = v0.copy()
            th = getattr(S, name)(m, vv)
            assert np.abs(vv - p["smd_" + name + "_v0"]).max() <= 1e-18, name
            x, v, a = x0.copy(), vv, np.zeros_like(x0)
            assert len(p["smd_" + name]) >= 3
            for row in p["smd_" + name]:
                with np.errstate(all="ignore"):
                    x, v, a, e = th.step(force, a, x, v, m, 0.2, ts)[:4]
                got = np.concatenate([x.ravel(), v.ravel(), a.ravel(), [e]])
                assert np.abs(got - row).max() <= 1e-12 * max(1.0, np.abs(row).max()), name
    finally:
        PARAMS.update(old)


@pytest.mark.parametrize("solver", ["Verlet", "CG", "BFGS"])
def test_neb_batched_beads_equal_per_bead_path(solver):
    """NudgedElasticBand(fb_=...) evaluates the whole band with one batched callback per solver iteration; with a
    batched callback that returns what the per-geometry callback returns, every iterate equals the per-bead path's
    (climbing image included: 14 iterations)."""
    from oracle.ref_py import host_pin_inputs
    from oracle.ref_py import toy_surface as ts
    atoms, x0, x1 = host_pin_inputs()
    calls = []

    def fb(xs, DoForce=True):
        calls.append(len(xs))
        if DoForce:
            out = [ts(x) for x in xs]
            return np.array([o[0] for o in out]), np.array([o[1] for o in out])
        return np.array([ts(x, False) for x in xs])

    old = PARAMS["NebSolver"]
    PARAMS["NebSolver"] = solver
    try:
        a = NudgedElasticBand(ts, Mol(atoms, x0), Mol(atoms, x1), nbeads_=6)
        b = NudgedElasticBand(None, Mol(atoms, x0), Mol(atoms, x1), nbeads_=6, fb_=fb)
        for it in range(14):
            for neb in (a, b):
                neb.beads, e, neb.Fs = neb.Solver(neb.beads)
                neb.IntegrateEnergy()
                neb.TSI = int(np.argmax(neb.Es))
                neb.step += 1
            assert np.array_equal(a.beads, b.beads) and np.array_equal(a.Es, b.Es) and np.array_equal(a.Fs, b.Fs), (solver, it)
        assert set(calls) == {6}
    finally:
        PARAMS["NebSolver"] = old


def test_ipi_client_serves_forces_over_a_socket():
    """TMIPIManger (reference Interfaces/TMIPIinterface.py:7-61) against an in-process stand-in for the i-PI server
    that speaks the driver protocol the reference's client implements (12-byte headers, POSDATA in atomic units, FORCEREADY
    reply) and delivers its messages in small TCP fragments. Units: Bohr -> Angstrom into the callback, J/mol/Angstrom ->
    Hartree/Bohr out."""
    import socket
    import threading
    from tensormol_b200.PhysicalData import BOHRPERA
    from TensorMol.Interfaces.TMIPIinterface import TMIPIManger        # the reference's import path

    natom = 4
    rs = np.random.RandomState(0)
    pos_bohr = [rs.rand(natom, 3) * 5.0 for _ in range(3)]
    seen, replies = [], []

    def field(x):
        seen.append(x.copy())
        return -1.5 - 0.1 * len(seen), JOULEPERHARTREE * (x * 2.0 - 1.0)

    srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    port = srv.getsockname()[1]

    def hdr(m):
        return m.encode().ljust(12)

    def recvn(c, n):
        b = b""
        while len(b) < n:
            chunk = c.recv(n - len(b))
            assert chunk
            b += chunk
        return b

    def server():
        c, _ = srv.accept()
        c.sendall(hdr("STATUS"))
        assert recvn(c, 12).strip() == b"READY"
        c.sendall(hdr("INIT") + np.int32(0).tobytes() + np.int32(5).tobytes() + b"hello")
        for p in pos_bohr:
            cell = (np.eye(3) * 20.0).tobytes()
            msg = hdr("POSDATA") + cell + cell + np.int32(natom).tobytes() + p.tobytes()
            for k in range(0, len(msg), 37):        # fragments: the client must reassemble
                c.sendall(msg[k:k + 37])
            c.sendall(hdr("STATUS"))
            assert recvn(c, 12).strip() == b"HAVEDATA"
            c.sendall(hdr("GETFORCE"))
            assert recvn(c, 12).strip() == b"FORCEREADY"
            e = np.frombuffer(recvn(c, 8), np.float64)[0]
            n = np.frombuffer(recvn(c, 4), np.int32)[0]
            f = np.frombuffer(recvn(c, 3 * n * 8), np.float64).reshape(n, 3)
            vir = np.frombuffer(recvn(c, 72), np.float64)
            nx = np.frombuffer(recvn(c, 4), np.int32)[0]
            extra = recvn(c, nx)
            replies.append((e, n, f.copy(), vir.copy(), extra))
            c.sendall(hdr("STATUS"))
            assert recvn(c, 12).strip() == b"READY"
        c.sendall(hdr("EXIT"))
        c.close()

    t = threading.Thread(target=server, daemon=True)
    t.start()
    client = TMIPIManger(field, "127.0.0.1", port)
    assert client.md_run() == 3
    t.join(10)
    srv.close()
    assert len(replies) == 3
    for k, (e, n, f, vir, extra) in enumerate(replies):
        x_ang = pos_bohr[k] / BOHRPERA
        assert np.allclose(seen[k], x_ang, rtol=0, atol=1e-15)
        assert e == -1.5 - 0.1 * (k + 1) and n == natom and extra == b"nothing" and not vir.any()
        assert np.allclose(f, (x_ang * 2.0 - 1.0) / BOHRPERA, rtol=1e-14, atol=0)
    with pytest.raises(OSError):
        TMIPIManger(field, "127.0.0.1", port)      # nothing listens any more: a failed connection raises


def test_reference_import_paths_resolve_to_the_same_modules():
    """`from TensorMol.<sub>.<module> import *` and `import MolEmb`, as the reference's sample scripts write them, give the
    modules of this package (one PARAMS, one engine)."""
    import MolEmb as top
    import TensorMol
    import TensorMol.ForceModifiers.Neighbors as n1
    import TensorMol.Simulations.SimpleMD as s1
    import tensormol_b200
    import tensormol_b200.ForceModifiers.Neighbors as n2
    import tensormol_b200.Simulations.SimpleMD as s2
    assert n1 is n2 and s1 is s2 and TensorMol.PARAMS is tensormol_b200.PARAMS
    assert top.Make_NListNaive is tensormol_b200.MolEmb.Make_NListNaive
    for f in ("Make_NListNaive", "Make_NListLinear", "Make_DistMat", "Make_DistMat_ForReal", "CountInRange", "GetRDF_Bin"):
        assert callable(getattr(top, f))


def test_periodic_force_and_drivers_equal_reference_python(tmp_path):
    """PeriodicForce (centred molecule, energy / force / energy-only call, Density, RDF, RDF_inC, LatticeStep), a seeded
    Metropolis chain of PeriodicMonteCarlo (accepted and rejected moves) and PeriodicGeomOptimizer.Opt against
    ForceModifiers/Periodic.py:277-473, Simulations/PeriodicMC.py:17-123 and OptPeriodic.py:8-60 executed in place on a toy
    local force."""
    from conftest import load_golden
    from oracle.ref_py import toy_local_force as tl
    from tensormol_b200 import PeriodicGeomOptimizer, PeriodicMonteCarlo
    p = load_golden("ref_host_pins")
    atoms, x0, lat = p["pd_atoms"], p["pd_x0"], p["pd_lat"]
    keys = ("OptMaxCycles", "MDV0", "MDTemp", "OptLatticeStep", "OptThresh")
    old = {k: PARAMS[k] for k in keys}
    PARAMS["OptMaxCycles"], PARAMS["MDV0"], PARAMS["MDTemp"], PARAMS["OptThresh"] = 12, None, 300.0, 0.0001
    try:
        pf = PeriodicForce(Mol(atoms, x0), lat)
        pf.BindForce(tl, 6.0)
        assert np.array_equal(pf.mol0.coords, p["pd_mol0"])
        e, f = pf(pf.mol0.coords)
        assert e == float(p["pd_e"]) and np.array_equal(f, p["pd_f"])
        assert pf(pf.mol0.coords, DoForce=False)[0] == float(p["pd_e_only"])
        assert abs(pf.Density() - float(p["pd_density"])) <= 1e-15
        assert np.array_equal(pf.RDF(pf.mol0.coords, 8, 1, 7.0, 0.05), p["pd_rdf"])
        np.random.seed(11)
        mc = PeriodicMonteCarlo(pf, "pinMC")
        assert mc.kbt == float(p["pd_mc_kbt"])
        accepted = []
        for row in p["pd_mc"]:
            mc.MetropolisHastings(mc.x)
            accepted.append(mc.Pacc)
            assert np.abs(np.concatenate([mc.x.ravel(), [mc.eold, mc.Pacc, mc.Eav, mc.dE2]]) - row).max() <= 1e-10
        assert min(accepted) < 1.0            # the pinned chain contains a rejected move
        pf2 = PeriodicForce(Mol(atoms, x0), lat)
        pf2.BindForce(tl, 6.0)
        m = PeriodicGeomOptimizer(pf2).Opt(Mol(atoms, pf2.mol0.coords.copy()))
        assert np.abs(m.coords - p["pd_popt_coords"]).max() <= 1e-10
        pf3 = PeriodicForce(Mol(atoms, x0), lat)
        pf3.BindForce(tl, 6.0)
        PARAMS["OptLatticeStep"] = 0.05
        xx = pf3.LatticeStep(pf3.mol0.coords)
        assert np.abs(xx - p["pd_latstep_x"]).max() <= 1e-12 and np.abs(pf3.lattice.lattice - p["pd_latstep_lattice"]).max() <= 1e-12
        assert PARAMS["OptLatticeStep"] == float(p["pd_latstep_step"])
        assert np.array_equal(pf3.RDF_inC(np.mod(x0, 6.0), atoms, 6.0, 8, 1, 7.0, 0.05), p["pd_rdf_inc"])
    finally:
        PARAMS.update(old)


def test_finite_difference_tools_and_harmonic_spectra_equal_reference_python():
    """FdiffGradient, FdiffHessian (forward, central, gradient modes) and HarmonicSpectra (wavenumbers and modes) against
    Math/QuasiNewtonTools.py:43-156,228-293 executed in place on the toy surface."""
    from tensormol_b200.Math.QuasiNewtonTools import FdiffHessian
    p, ts, (atoms, x0, x1) = _host_pin_setup()
    energy = lambda x: np.float64(ts(x, False))           # noqa: E731
    grad = lambda x: -ts(x)[1] / 2625499.638               # noqa: E731
    assert np.abs(FdiffGradient(energy, x0) - p["fd_gradient"]).max() <= 1e-12
    assert np.abs(FdiffHessian(energy, x0, 0.001) - p["fd_hess_forward"]).max() <= 1e-9
    assert np.abs(FdiffHessian(energy, x0, 0.001, "central") - p["fd_hess_central"]).max() <= 1e-9
    assert np.abs(FdiffHessian(energy, x0, 0.001, "gradient", grad) - p["fd_hess_gradient"]).max() <= 1e-9
    w, v = HarmonicSpectra(energy, x0, atoms)
    assert np.abs(w - p["harm_w"]).max() <= 1e-6 * np.abs(p["harm_w"]).max()
    assert np.abs(np.abs(v) - np.abs(p["harm_v"])).max() <= 1e-6


def test_md_drivers_whole_runs_equal_reference_python():
    """Ten-step Prop() runs (zero initial velocities) of VelocityVerlet (NVE, Nose), IRTrajectory with a field pulse and
    geometry-dependent charges, Annealer, PeriodicVelocityVerlet (NVE, Nose), PeriodicAnnealer and PeriodicBoxingDynamics
    against Simulations/SimpleMD.py:322-619 and PeriodicMD.py:44-285 executed in place: final positions, velocities, the
    drivers' logs (time, dipole, KE, EPot, total), best annealed geometry, final boxed lattice."""
    from conftest import load_golden
    from oracle import ref_py
    from tensormol_b200 import Annealer, IRTrajectory, PeriodicAnnealer, PeriodicBoxingDynamics
    p = load_golden("ref_host_pins")
    atoms, x0, _ = ref_py.host_pin_inputs()
    patoms, px0, lat = p["pd_atoms"], p["pd_x0"], p["pd_lat"]
    ts, tsb, tq, tl = ref_py.toy_surface, ref_py.toy_surface_bound, ref_py.toy_charges, ref_py.toy_local_force
    keys = ("MDV0", "MDMaxStep", "MDdt", "MDTemp", "MDLogTrajectory", "MDThermostat", "MDFieldAmp", "MDAnnealSteps", "MDAnnealT0", "MDAnnealTF")
    old = {k: PARAMS[k] for k in keys}

    def setp(**kw):
        PARAMS.update(old)
        PARAMS.update(dict(MDV0=None, MDMaxStep=10, MDdt=0.2, MDTemp=300.0, MDLogTrajectory=False))
        PARAMS.update(kw)

    def same(tag, **got):
        for k, a in got.items():
            ref = p["mdd_" + tag + "_" + k]
            assert np.abs(np.asarray(a, np.float64) - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (tag, k)

    def pforce():
        pf = PeriodicForce(Mol(patoms, px0), lat)
        pf.BindForce(tl, 6.0)
        return pf

    try:
        for tag, th in (("nve", None), ("nose", "Nose")):
            setp(MDThermostat=th)
            d = VelocityVerlet(lambda x: ts(x)[1], Mol(atoms, x0), "pin", ts)
            d.Prop()
            same("vv_" + tag, x=d.x, v=d.v, log=d.md_log)
            d = PeriodicVelocityVerlet(pforce(), "pinp")
            d.Prop()
            same("pvv_" + tag, x=d.x, v=d.v, log=d.md_log)
        setp(MDThermostat=None, MDFieldAmp=2.0)
        d = IRTrajectory(tsb, tq, Mol(atoms, x0), "pinir")
        d.Prop()
        same("ir", x=d.x, v=d.v, log=d.mu_his)
        assert np.abs(d.mu_his[:, 1:4]).max() > 0
        setp(MDAnnealSteps=8, MDAnnealT0=40.0, MDAnnealTF=5.0)
        d = Annealer(tsb, tq, Mol(atoms, x0), "pinan")
        d.Prop()
        same("an", x=d.x, v=d.v, minx=d.Minx, mine=d.MinE)
        d = PeriodicAnnealer(pforce(), "pinpa")
        d.Prop()
        same("pan", x=d.x, v=d.v, minx=d.Minx, mine=d.MinE)
        setp(MDThermostat="Nose")
        pf = pforce()
        d = PeriodicBoxingDynamics(pf, lat * 0.97, "pinbox", 1.0)
        d.Prop()
        same("box", x=d.x, v=d.v, log=d.md_log, lattice=pf.lattice.lattice)
    finally:
        PARAMS.update(old)


def test_names_used_by_the_reference_sample_scripts_resolve():
    """Static check of the drop-in surface: every global name the reference's hot-path sample scripts
    use (samples/test_tensormol01.py, test_h2o.py: BoxAndDensity / TestNeb / Eval, test_neb.py: GetChemSpider12 / Eval /
    TestBetaHairpin) and every TFMolManage method they call exists in `from TensorMol import *` plus the two explicit
    imports those scripts make. Needs the reference tree (skipped on the GPU box)."""
    import ast
    import builtins
    import TensorMol as tm
    import TensorMol.Interfaces.TMIPIinterface as ipi
    ref = os.environ.get("TM_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "samples")):
        pytest.skip("reference tree not present")
    wanted = {"test_tensormol01.py": None, "test_h2o.py": {"BoxAndDensity", "TestNeb", "Eval"},
              "test_neb.py": {"GetChemSpider12", "Eval", "TestBetaHairpin"}}
    off_path = {"WriteDerDipoleCorrelationFunction"}           # IR post-processing (Simulations/InfraredMD.py): out of scope
    problems = []
    for fname, funcs in wanted.items():
        tree = ast.parse(open(os.path.join(ref, "samples", fname)).read())
        top = {n.name for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
        top |= {t.id for n in ast.walk(tree) if isinstance(n, ast.Assign) for t in n.targets if isinstance(t, ast.Name)}   # script globals
        for node in tree.body:
            if not isinstance(node, ast.FunctionDef) or (funcs is not None and node.name not in funcs):
                continue
            bound, used, mgr = set(), set(), set()
            for n in ast.walk(node):
                if isinstance(n, ast.Name):
                    (bound if isinstance(n.ctx, (ast.Store, ast.Del)) else used).add(n.id)
                elif isinstance(n, ast.arg):
                    bound.add(n.arg)
                elif isinstance(n, ast.FunctionDef):
                    bound.add(n.name)
                elif isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "manager" and n.attr != "Instances":
                    mgr.add(n.attr)
            for u in sorted(used - bound - top - off_path):
                if not (hasattr(builtins, u) or hasattr(tm, u) or hasattr(ipi, u)):
                    problems.append((fname, node.name, u))
            problems += [(fname, node.name, "TFMolManage." + a) for a in sorted(mgr) if not hasattr(tm.TFMolManage, a)]
    assert not problems, problems


TRAIN_BATCH_NAMES = ["xyzs", "Zs", "Elabels", "Dlabels", "grads", "rad_p_ele", "ang_t_elep", "rad_eep", "mil_jk", "inv_natom"]


def train_pin_set(g):
    """The molecule set of tests/golden/ref_train_pins.npz as this package's containers."""
    from tensormol_b200 import MolDigester, TensorMolData_BP_Direct_EE_WithEle
    a = MSet("train_pins")
    i = 0
    while "in%d_atoms" % i in g:
        m = Mol(g["in%d_atoms" % i], g["in%d_coords" % i])
        m.properties = {"atomization": float(g["in%d_atomization" % i]), "dipole": g["in%d_dipole" % i],
                        "gradients": g["in%d_gradients" % i], "serial": i}
        a.mols.append(m)
        i += 1
    old = PARAMS["TestRatio"]
    PARAMS["TestRatio"] = 0.4            # oracle/make_golden.py:TRAIN_PIN_PARAMS
    try:
        t = TensorMolData_BP_Direct_EE_WithEle(a, MolDigester(a.AtomTypes(), name_="ANI1_Sym_Direct", OType_="EnergyAndDipole"),
                                               order_=1, num_indis_=1, type_="mol", WithGrad_=True)
    finally:
        PARAMS["TestRatio"] = old
    eles = sorted(int(e) for e in a.AtomTypes())
    t.ele = np.asarray(eles).reshape(-1, 1)
    t.elep = np.asarray([[eles[i], eles[j]] for i in range(len(eles)) for j in range(i, len(eles))])
    return t


def check_train_batches(t, g):
    import random
    random.seed(7)
    t.LoadDataToScratch(None)
    assert [m.properties["serial"] for m in t.set.mols] == list(g["order"])
    assert (t.NTrain, t.NTest) == (int(g["NTrain"]), int(g["NTest"]))
    for kind, n, nc, fn in (("train", 5, 3, t.GetTrainBatch), ("test", 4, 2, t.GetTestBatch)):
        for c in range(n):
            batch = fn(nc)
            assert len(batch) == len(TRAIN_BATCH_NAMES)
            for nm, v in zip(TRAIN_BATCH_NAMES, batch):
                ref = g["%s%d_%s" % (kind, c, nm)]
                v = np.asarray(v)
                assert v.shape == ref.shape, (kind, c, nm)
                if nm == "rad_eep":      # the unsorted list keeps MolEmb's x-sweep order inside a row: equal as sets of rows
                    rows = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]                   # noqa: E731
                    assert v.dtype == ref.dtype and np.array_equal(rows(v), rows(ref)), (kind, c, nm)
                elif nm in ("Zs", "rad_p_ele", "ang_t_elep", "mil_jk"):
                    assert v.dtype == ref.dtype and np.array_equal(v, ref), (kind, c, nm)        # index work: bit-exact
                else:
                    assert np.array_equal(v, ref), (kind, c, nm)                                  # copies of the inputs / 1/natom
    with pytest.raises(Exception, match="Insufficent training data"):
        t.GetTrainBatch(t.NTrain + 1)
    with pytest.raises(Exception, match="Insufficent training data"):
        t.GetTestBatch(t.NTest + 1)


def test_train_batches_equal_reference_python(monkeypatch):
    """GetTrainBatch / GetTestBatch (data side): shuffle, train / test split, the wrapping batch pointers, labels
    and gradient blocks equal the reference's Python executed in place (tests/golden/ref_train_pins.npz).  No GPU here, so the
    neighbour tables of this CPU test come from the oracle through a stand-in for NeighborListSet -- which pins the oracle's
    multi-molecule tables against the reference at the same time; tests/test_zz_gpu_widen.py runs the same check on the CUDA tables."""
    from conftest import load_golden
    import tensormol_b200.ForceModifiers.Neighbors as NB

    class OracleNeighborListSet:
        def __init__(self, x_, nnz_, DoTriples_=False, DoPerms_=False, ele_=None, alg_=None, sort_=False):
            self.x, self.nnz, self.perms, self.Zs = x_, np.asarray(nnz_), DoPerms_, ele_

        def buildPairsAndTriplesWithEleIndex(self, rr, ra, ele, elep):
            return onp.build_pairs_and_triples_with_ele_index(self.x, self.nnz, self.nnz, self.Zs, rr, ra, ele, elep)

        def buildPairs(self, rng):
            return onp.set_build_pairs(self.x, self.nnz, self.nnz, rng, self.perms)

    monkeypatch.setattr(NB, "NeighborListSet", OracleNeighborListSet)
    g = load_golden("ref_train_pins")
    check_train_batches(train_pin_set(g), g)
