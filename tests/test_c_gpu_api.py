"""GPU tests of the reference-facing Python surface (TFMolManage Eval* methods, NeighborListSet, MolEmb shim,
PeriodicForce callbacks, drivers) against the oracle."""
import numpy as np
import pytest

from common import ENERGY_RTOL, FORCE_ATOL_HA_BOHR, water_box
from conftest import load_golden

pytestmark = pytest.mark.gpu
from tensormol_b200._lib import TMolB200Error  # noqa: E402

NET = "fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout"


def _setup_params(hidden):
    from tensormol_b200 import PARAMS
    PARAMS["NeuronType"] = "sigmoid_with_param"
    PARAMS["sigmoid_alpha"] = 100.0
    PARAMS["HiddenLayers"] = list(hidden)
    PARAMS["EECutoffOn"] = 0
    PARAMS["Elu_Width"] = 4.6
    PARAMS["EECutoffOff"] = 15.0
    PARAMS["DSFAlpha"] = 0.18
    PARAMS["AddEcc"] = True
    PARAMS["KeepProb"] = [1.0, 1.0, 1.0, 1.0]
    PARAMS["MDLogTrajectory"] = False
    return PARAMS


def _manager(mols, hidden, seed):
    from tensormol_b200 import MolDigester, MSet, TensorMolData_BP_Direct_EE_WithEle, TFMolManage
    _setup_params(hidden)
    a = MSet("t", center_=False)
    a.mols = list(mols)
    d = MolDigester(a.AtomTypes(), name_="ANI1_Sym_Direct", OType_="EnergyAndDipole")
    tset = TensorMolData_BP_Direct_EE_WithEle(a, d, order_=1, num_indis_=1, type_="mol", WithGrad_=True)
    manager = TFMolManage("", tset, False, NET, False, False)
    W = manager.InitRandom(seed)
    return manager, W


def _oracle(eles, W):
    from oracle import oracle_graph as og
    return og.Oracle(eles, W, og.default_params())


def test_manager_single_matches_oracle_and_units():
    from tensormol_b200 import PARAMS, Mol
    from tensormol_b200.PhysicalData import BOHRPERA, JOULEPERHARTREE
    g = load_golden("h2o_cluster")
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], list(g["hidden"]), int(g["seed"]))
    out = manager.EvalBPDirectEEUpdateSingle(m, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)
    Etotal, Ebp, Ebp_atom, Ecc, Evdw, mol_dipole, atom_charge, force = out
    assert Etotal.shape == (1,) and force.shape == (1, len(g["Z"]), 3) and atom_charge.shape == (1, len(g["Z"]))
    assert abs(Etotal[0] - g["oracle_Etotal"][0]) <= ENERGY_RTOL * abs(g["oracle_Etotal"][0])
    assert abs(Ecc[0] - g["oracle_Ecc"][0]) <= ENERGY_RTOL * max(abs(g["oracle_Ecc"][0]), 1e-3)
    # force = -JOULEPERHARTREE * dE/dx (J/mol/A)
    grad = -force / JOULEPERHARTREE
    assert np.abs(grad - g["oracle_gradient"]).max() / BOHRPERA <= FORCE_ATOL_HA_BOHR
    assert np.abs(grad - g["oracle_gradient"]).max() <= 2e-4 * np.abs(g["oracle_gradient"]).max()
    six = manager.EvalBPDirectEEUpdateSingle(m, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], False)
    assert len(six) == 6 and six[0][0] == Etotal[0]
    dip, q = manager.EvalBPDirectChargeSingle(m, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"])
    assert np.allclose(q, atom_charge) and np.abs(dip - g["oracle_dipole"]).max() < 1e-5
    assert manager.Instances.eles_np.reshape(-1).tolist() == [1, 8]
    assert manager.Instances.eles_pairs_np.tolist() == [[1, 1], [1, 8], [8, 8]]


def test_manager_weights_save_load_round_trip(tmp_path):
    from tensormol_b200 import PARAMS, Mol
    g = load_golden("h2o_cluster")
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], [32, 16], 4)
    e1 = manager.EvalBPDirectEEUpdateSingle(m, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)[0][0]
    p = manager.SaveWeights(str(tmp_path / "w.npz"))
    manager2, _ = _manager([m], [32, 16], 99)
    manager2.LoadWeights(p)
    e2 = manager2.EvalBPDirectEEUpdateSingle(m, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)[0][0]
    assert abs(e1 - e2) <= 1e-9 * abs(e1)        # same weights; reductions use atomics, so not bit-equal by contract


def test_manager_restores_a_reference_style_network_by_name(tmp_path):
    """The reference's load chain (TFMolManage.py:1468-1470, TFMolInstance.py:104-113): <name>.tfm -> TrainedNetworks[0] ->
    <instance>.tfn -> chk_file -> TensorFlow checkpoint, read without TensorFlow (TFNetworks/TFCheckpoint.py)."""
    import pickle
    from tensormol_b200 import PARAMS, Mol, MolDigester, MSet, TensorMolData_BP_Direct_EE_WithEle, TFMolManage
    g = load_golden("h2o_cluster")
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], [32, 16, 24], 4)
    for net in W:                                   # trained networks have biases
        for z in W[net]:
            W[net][z] = [(w, 0.01 * np.arange(len(b)) - 0.02) for w, b in W[net][z]]
    manager.SetWeights(W)
    args = (m, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)
    e1, f1 = manager.EvalBPDirectEEUpdateSingle(*args)[0][0], manager.EvalBPDirectEEUpdateSingle(*args)[7]
    nets = str(tmp_path) + "/"
    inst = "Mol_t_ANI1_Sym_Direct_" + NET
    manager.SaveCheckpoint(nets + inst + "/" + inst + "-chk-7", np.float64)
    with open(nets + inst + ".tfn", "wb") as fh:
        pickle.dump(dict(name=inst, HiddenLayers=[32, 16, 24], eles=[1, 8], chk_file="./networks/" + inst + "/" + inst + "-chk-7"), fh, protocol=2)
    with open(nets + "water_network.tfm", "wb") as fh:
        pickle.dump(dict(name="water_network", NetType=NET, TrainedNetworks=[inst]), fh, protocol=2)
    old = PARAMS["networks_directory"]
    PARAMS["networks_directory"] = nets
    try:
        a = MSet("t", center_=False)
        a.mols = [m]
        tset = TensorMolData_BP_Direct_EE_WithEle(a, MolDigester(a.AtomTypes(), name_="ANI1_Sym_Direct", OType_="EnergyAndDipole"), order_=1,
                                                  num_indis_=1, type_="mol", WithGrad_=True)
        manager2 = TFMolManage("water_network", tset, False, NET, False, False)
        out = manager2.EvalBPDirectEEUpdateSingle(*args)
        # same weights through the checkpoint: equal up to the summation order of the atomically reduced forces
        assert abs(out[0][0] - e1) <= 1e-9 * abs(e1)
        assert np.abs(out[7] - f1).max() <= 1e-5 * np.abs(f1).max()
        PARAMS["HiddenLayers"] = [32, 16, 25]      # a network of another shape is refused, not silently mis-read
        with pytest.raises(Exception, match="HiddenLayers"):
            TFMolManage("water_network", tset, False, NET, False, False)
    finally:
        PARAMS["networks_directory"] = old


def test_manager_periodic_callbacks_match_oracle():
    from oracle import oracle_np as onp
    from tensormol_b200 import PARAMS, Mol, PeriodicForce
    from tensormol_b200.PhysicalData import BOHRPERA, JOULEPERHARTREE
    g = load_golden("water_tiny_periodic")
    nreal = len(g["Z"])
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], list(g["hidden"]), int(g["seed"]))

    def EnAndForce(z_, x_, nreal_, DoForce=True):      # the closure of samples/test_h2o.py:1661-1671
        mtmp = Mol(z_, x_)
        if DoForce:
            en, f = manager.EvalBPDirectEEUpdateSinglePeriodic(mtmp, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], nreal_, True)
            return en[0], f[0]
        return manager.EvalBPDirectEEUpdateSinglePeriodic(mtmp, PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], nreal_, True, DoForce)[0]

    pf = PeriodicForce(m, g["lattice"])
    pf.BindForce(EnAndForce, 15.0)
    x0 = g["xyz"]          # already wrapped
    e, f = pf(x0)
    assert abs(e - g["oracle_Etotal"][0]) <= ENERGY_RTOL * abs(g["oracle_Etotal"][0])
    grad = -f / JOULEPERHARTREE
    assert np.abs(grad - g["oracle_gradient"][0]).max() / BOHRPERA <= FORCE_ATOL_HA_BOHR
    assert np.abs(grad - g["oracle_gradient"][0]).max() <= 2e-4 * np.abs(g["oracle_gradient"]).max()
    e_only, _ = pf(x0, False)
    assert abs(e_only - e) < 1e-12 * abs(e) + 1e-9
    # B200 extension: images generated on the device give the same numbers
    pf2 = PeriodicForce(m, g["lattice"])
    pf2.BindLatticeForce(manager.LatticeForce(), 15.0)
    e2, f2 = pf2(x0)
    # (a different cell grid, hence a different fp32 summation order: equal to rounding, not bit for bit)
    assert abs(e2 - e) <= 1e-6 * abs(e) and np.abs(f2 - f).max() <= 1e-5 * np.abs(f).max()
    out = manager.EvalBPDirectEEUpdateSinglePeriodic(Mol(*pf.lattice.TessLattice(pf.atoms, x0, 15.0)), PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"],
                                                     PARAMS["EECutoffOff"], nreal, True, True, True)
    assert out[2].shape == (1, nreal) and np.abs(out[2] - g["oracle_charge"]).max() < 1e-5


def test_neighbor_list_set_tables_match_oracle():
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from tensormol_b200 import MolEmb, NeighborListSet, NeighborListSetWithImages
    g = load_golden("morphine")
    N = len(g["Z"])
    xyzs = np.zeros((2, N, 3))
    Zs = np.zeros((2, N), np.int32)
    nat = np.array([N, N - 5])
    xyzs[0], Zs[0] = g["xyz"], g["Z"]
    xyzs[1, :N - 5], Zs[1, :N - 5] = g["xyz"][:N - 5] + 0.01, g["Z"][:N - 5]
    eles_np, elep_np = og.elements_and_pairs(g["eles"])
    NL = NeighborListSet(xyzs, nat, True, True, Zs, sort_=True)
    rad, ang, mil_jk, jk_max = NL.buildPairsAndTriplesWithEleIndex(4.6, 3.1, eles_np, elep_np)
    o = onp.build_pairs_and_triples_with_ele_index(xyzs, nat, nat, Zs, 4.6, 3.1, eles_np, elep_np)
    assert rad.dtype == np.float64
    assert np.array_equal(rad, o[0]) and np.array_equal(ang, o[1]) and np.array_equal(mil_jk, o[2]) and jk_max == o[3]
    rad2, ang2, mil_j, mil_jk2 = NL.buildPairsAndTriplesWithEleIndexLinear(4.6, 3.1, eles_np, elep_np)
    o2 = onp.build_pairs_and_triples_with_ele_index_periodic(xyzs, nat, nat, Zs, 4.6, 3.1, eles_np, elep_np)
    assert np.array_equal(mil_j, o2[2])
    NLEE = NeighborListSet(xyzs, nat, False, False, None)
    ree = NLEE.buildPairs(15.0)
    o_ree = onp.set_build_pairs(xyzs, nat, nat, 15.0, False)
    assert ree.dtype == np.uint64
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]   # noqa: E731
    assert np.array_equal(key(ree.astype(np.int64)), key(o_ree.astype(np.int64)))
    p, t = NL.buildPairsAndTriples(4.6, 3.1)
    op, ot = onp.set_build_pairs_and_triples(xyzs, nat, nat, Zs, 4.6, 3.1, True)
    key4 = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]   # noqa: E731
    assert np.array_equal(key(p.astype(np.int64)), key(op.astype(np.int64)))
    assert np.array_equal(key4(t.astype(np.int64)), key4(ot.astype(np.int64)))
    # images form
    NLI = NeighborListSetWithImages(xyzs[:1], np.array([N]), np.array([10]), False, True, Zs[:1])
    both = NLI.buildPairsWithBothEleIndex(15.0, eles_np)
    ob = onp.build_pairs_with_both_ele_index(xyzs[:1], np.array([N]), np.array([10]), Zs[:1], 15.0, eles_np, True)
    key5 = lambda a: a[np.lexsort((a[:, 2], a[:, 1]))]   # noqa: E731
    assert np.array_equal(key5(both), key5(ob))
    # MolEmb shim returns list-of-lists like the C extension
    ll = MolEmb.Make_NListNaive(g["xyz"], 4.6, N, 1)
    off, idx = g["ref_nl_rr_off"], g["ref_nl_rr_idx"]
    assert [sorted(r) for r in ll] == [idx[off[i]:off[i + 1]].tolist() for i in range(N)]
    assert [sorted(r) for r in MolEmb.Make_NListLinear(g["xyz"], 4.6, N, 1)] == [sorted(r) for r in ll]


def test_nve_md_on_gpu_forces_conserves_energy():
    """3x3x3 water box, 40 velocity-Verlet steps through PeriodicVelocityVerlet with the device-tessellated force."""
    from tensormol_b200 import PARAMS, Mol, PeriodicForce, PeriodicVelocityVerlet
    from tensormol_b200.PhysicalData import JOULEPERHARTREE
    Z, X, lat = water_box(3, jitter=0.0)
    m = Mol(Z.astype(np.uint8), X)
    manager, W = _manager([m], [64, 64], 7)
    pf = PeriodicForce(m, lat)
    pf.BindLatticeForce(manager.LatticeForce(), 15.0)
    PARAMS["MDMaxStep"] = 40
    PARAMS["MDdt"] = 0.2
    PARAMS["MDV0"] = None
    PARAMS["MDThermostat"] = None
    md = PeriodicVelocityVerlet(pf, "gpu_nve")
    md.Prop()
    ke = md.md_log[:, 4] * len(Z)                       # J/mol total (KE is per atom), logged one step behind EPot
    etot = ke[1:] + md.md_log[:-1, 5] * JOULEPERHARTREE
    drift = np.ptp(etot[2:])
    scale = np.ptp(md.md_log[:, 5] * JOULEPERHARTREE) + 1.0
    assert np.all(np.isfinite(md.md_log)) and drift < 0.05 * scale + 1e-3 * abs(etot[2])


@pytest.mark.gpu
@pytest.mark.parametrize("thermostat,graph", [(None, True), (None, False), ("Nose", True)])
def test_device_md_matches_host_driver(thermostat, graph):
    """the on-device integrator (state on the GPU, CUDA-graph replay) follows the host driver
    (PeriodicVelocityVerlet / PeriodicNoseThermostat on numpy arrays, same device forces) step for step."""
    from tensormol_b200 import PARAMS, Mol, PeriodicForce, PeriodicVelocityVerlet
    from tensormol_b200.Simulations.DeviceMD import DevicePeriodicVelocityVerlet
    Z, X, lat = water_box(3, jitter=0.02)
    m = Mol(Z.astype(np.uint8), X)
    manager, W = _manager([m], [64, 64], 7)
    nstep = 12
    PARAMS["MDMaxStep"] = nstep
    PARAMS["MDdt"] = 0.2
    PARAMS["MDV0"] = None
    PARAMS["MDTemp"] = 300.0
    PARAMS["MDThermostat"] = thermostat
    rs = np.random.RandomState(5)
    v0 = 1e-3 * rs.randn(len(Z), 3)
    pf = PeriodicForce(m, lat)
    pf.BindLatticeForce(manager.LatticeForce(), 15.0)
    host = PeriodicVelocityVerlet(pf, "host_md", v0_=v0.copy())
    v_start = host.v.copy()                       # the Nose constructor rescales v0 to MDTemp
    host.Prop()
    dev = DevicePeriodicVelocityVerlet(manager, m, lat, "dev_md", v0_=v0.copy(), graph_=graph, sync_every_=5)
    log = dev.Prop()
    if thermostat is None:
        assert np.allclose(v_start, v0)
    xh = pf.lattice.ModuloLattice(host.x)
    d = dev.x - xh
    d -= np.round(d @ np.linalg.inv(lat)) @ lat   # same point modulo the lattice
    assert np.abs(d).max() < 1e-7
    assert np.abs(dev.v - host.v).max() < 1e-7 * max(1.0, np.abs(host.v).max() / 1e-3)
    assert abs(dev.EPot - host.EPot) < 1e-6 * abs(host.EPot)
    assert np.all(np.isfinite(log)) and abs(log[nstep - 1, 5] - dev.EPot) <= 1e-9 * abs(dev.EPot)


def test_geometry_optimizer_lowers_energy_on_gpu_potential():
    from tensormol_b200 import PARAMS, GeomOptimizer, Mol
    g = load_golden("h2o_cluster")
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], [32, 32], 9)

    def EnAndForce(x_, DoForce=True):
        out = manager.EvalBPDirectEEUpdateSingle(Mol(m.atoms, x_), PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)
        return (out[0][0], out[-1][0]) if DoForce else out[0][0]
    PARAMS["OptMaxCycles"] = 8
    e0 = EnAndForce(m.coords, False)
    out = GeomOptimizer(EnAndForce).Opt(m, "gpuopt")
    assert EnAndForce(out.coords, False) < e0


@pytest.fixture(autouse=True)
def _tmp_results(tmp_path):
    from tensormol_b200 import PARAMS
    old = PARAMS["results_dir"]
    PARAMS["results_dir"] = str(tmp_path) + "/"
    yield
    PARAMS["results_dir"] = old


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_phases_emulated_ranks_match_single_call(world):
    """The three slab phases of `world` ranks, run one after the other on ONE GPU with the all-reduces emulated by
    sums, must reproduce tm_eval_lattice (energies to fp32 summation-order noise, gradients to fp32 atomics noise)."""
    import torch
    from oracle import oracle_graph as og
    from tensormol_b200.engine import Engine, random_weights
    from tensormol_b200.parallel import EngineSlabBackend
    from tensormol_b200._lib import TM_F_FORCE, TM_F_VDW
    from tensormol_b200.SystemBuilders import wrap_into_cell
    Z, X, lat = water_box(5, jitter=0.03)
    X = wrap_into_cell(X, lat)
    n = len(Z)
    P = og.default_params()
    hidden = [64, 48]
    W = random_weights([1, 8], 256, hidden, 3)
    ref_eng = Engine([1, 8], hidden, P)
    ref_eng.set_weights(W)
    ref = ref_eng.evaluate_lattice(X, Z, lat, 1)
    dev = torch.device("cuda", 0)
    xt = torch.tensor(X, dtype=torch.float64, device=dev)
    zt = torch.tensor(Z, dtype=torch.int32, device=dev)
    engines = []
    for r in range(world):
        e = Engine([1, 8], hidden, P)
        e.set_weights(W)
        engines.append(EngineSlabBackend(e))
    qraw = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(world)]
    for r, b in enumerate(engines):
        b.slab_phase_a(xt, zt, n, lat, 1, r, world, qraw[r])
        b.eng.sync()
    qsum = torch.stack(qraw).sum(0)
    es = [torch.zeros(6, dtype=torch.float64, device=dev) for _ in range(world)]
    for r, b in enumerate(engines):
        b.slab_phase_b(qsum, es[r])
        b.eng.sync()
    esum = torch.stack(es).sum(0)
    grads = [torch.zeros(n, 3, dtype=torch.float64, device=dev) for _ in range(world)]
    for r, b in enumerate(engines):
        b.slab_phase_c(esum, TM_F_FORCE | TM_F_VDW, grads[r])
        b.eng.sync()
    g = torch.stack(grads).sum(0).cpu().numpy()
    e = esum.cpu().numpy()
    etot = e[1] + e[2] + e[3]
    # each rank bins only its slab + halo, so cell order (and with it the fp32 summation order) differs from the single call
    assert abs(etot - ref["Etotal"][0]) <= 5e-7 * abs(ref["Etotal"][0])
    assert abs(e[2] - ref["Ecc"][0]) <= 1e-6 * max(abs(ref["Ecc"][0]), 1e-3)
    assert np.abs(g - ref["gradient"][0]).max() <= 2e-6 * np.abs(ref["gradient"]).max() + 1e-9


@pytest.mark.parametrize("world", [2, 3])
def test_slab_peer_memory_exchange_on_one_gpu(world):
    """The peer-memory form of the slab phases (tm_slab_p2p_setup): `world` contexts on ONE GPU, each on its own stream,
    exchange q_raw / energy partials / force partials through each other's "symmetric" buffers with device-side
    signal/wait flags, no host collective.  Two consecutive steps (the flag epochs advance) reproduce tm_eval_lattice."""
    import ctypes as C
    import torch
    from oracle import oracle_graph as og
    from tensormol_b200.engine import Engine, random_weights
    from tensormol_b200.parallel import EngineSlabBackend
    from tensormol_b200._lib import TM_F_FORCE, TM_F_VDW
    from tensormol_b200.SystemBuilders import wrap_into_cell
    Z, X, lat = water_box(5, jitter=0.03)
    n = len(Z)
    P = og.default_params()
    hidden = [64, 48]
    W = random_weights([1, 8], 256, hidden, 3)
    ref_eng = Engine([1, 8], hidden, P)
    ref_eng.set_weights(W)
    dev = torch.device("cuda", 0)
    backs, streams, bufs = [], [], []
    for r in range(world):
        e = Engine([1, 8], hidden, P)
        e.set_weights(W)
        st = torch.cuda.Stream(device=dev)
        e.set_stream(C.c_void_p(st.cuda_stream))
        backs.append(EngineSlabBackend(e))
        streams.append(st)
    zt = torch.tensor(Z, dtype=torch.int32, device=dev)
    es = [torch.zeros(6, dtype=torch.float64, device=dev) for _ in range(world)]
    gs = [torch.zeros(n, 3, dtype=torch.float64, device=dev) for _ in range(world)]
    # one ordinary (host-collective form) pass per context first, so that every workspace buffer exists: an allocation
    # may wait for the device, and in this ONE-host-thread emulation that would wait for a peer whose work is not
    # enqueued yet.  Real ranks are separate processes on separate GPUs.
    xt = torch.tensor(wrap_into_cell(X, lat), dtype=torch.float64, device=dev)
    qtmp = torch.zeros(n, dtype=torch.float64, device=dev)
    for r, b in enumerate(backs):
        b.slab_phase_a(xt, zt, n, lat, 1, r, world, qtmp)
        b.slab_phase_b(qtmp, es[r])
        b.slab_phase_c(es[r], TM_F_FORCE | TM_F_VDW, gs[r])
        b.eng.sync()
    nbytes = backs[0].p2p_bytes(world, n)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    for r, b in enumerate(backs):
        b.p2p_setup(world, r, n, [t.data_ptr() for t in bufs])
    rs = np.random.RandomState(1)
    for step in range(3):
        Xs = wrap_into_cell(X + 0.02 * step * rs.randn(*X.shape), lat)
        ref = ref_eng.evaluate_lattice(Xs, Z, lat, 1)
        xt = torch.tensor(Xs, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        # launch order: phase by phase over the ranks, everything asynchronous; the waits resolve on the device
        for r, b in enumerate(backs):
            b.slab_phase_a(xt, zt, n, lat, 1, r, world, es[r])        # qraw argument unused in this mode
        for r, b in enumerate(backs):
            b.slab_phase_b(es[r], es[r])
        for r, b in enumerate(backs):
            b.slab_phase_c(es[r], TM_F_FORCE | TM_F_VDW, gs[r])
        for r, b in enumerate(backs):
            try:
                b.eng.sync()                                          # also surfaces a wait time-out (flag 32)
            except Exception as ex:
                raise AssertionError(f"step {step} rank {r}: {ex}")
        for r in range(world):
            e = es[r].cpu().numpy()
            assert abs(e[0] - ref["Etotal"][0]) <= 5e-7 * abs(ref["Etotal"][0]), (step, r)
            assert abs(e[1] - ref["Ebp"][0]) <= 5e-7 * abs(ref["Ebp"][0])
            assert abs(e[2] - ref["Ecc"][0]) <= 1e-6 * abs(ref["Ecc"][0]) + 1e-9
            assert abs(e[3] - ref["Evdw"][0]) <= 5e-7 * abs(ref["Evdw"][0])
            assert np.abs(gs[r].cpu().numpy() - ref["gradient"][0]).max() <= 2e-6 * np.abs(ref["gradient"]).max()


# ---------------------------------------------------------------------------------------- Verlet skin
def _skin_system(nx, margin=0.75):
    """Water box whose atoms keep `margin` A clear of the cell faces (cell = box + 2 margin), so that displaced coordinates
    are still inside the cell and can be evaluated from scratch without re-wrapping."""
    from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
    Z, X, lat = water_box(nx)
    X = wrap_into_cell(X, lat) + margin
    return Z, X, lat + 2.0 * margin * np.eye(3)


@pytest.mark.parametrize("nx", [3, 6])
def test_verlet_skin_reuse_equals_rebuild_and_oracle(nx):
    """tm_set_skin + TM_F_REUSE_NLIST: rows built once out to cutoff + skin, then evaluations at displaced positions that only
    refresh the coordinates must equal a from-scratch evaluation of the same coordinates (engine without skin), and the
    oracle on the reference's tessellation of them (nx = 3: two image shells)."""
    import ctypes as C
    import torch
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from test_a_gpu_parity import _check_energy, _check_grad, _engine
    Z, X0, lat = _skin_system(nx)
    n = len(Z)
    eng, W, P = _engine([1, 8], [32, 32], 7)
    ref, _, _ = _engine([1, 8], [32, 32], 7)
    eng.set_skin(0.6)
    dev = torch.device("cuda", 0)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X0, P["EECutoffOff"])
    ntess = (round((len(Zt) / n) ** (1 / 3)) - 1) // 2
    xd = torch.tensor(X0, dtype=torch.float64, device=dev)
    zd = torch.tensor(Z, dtype=torch.int32, device=dev)
    e = torch.zeros(6, dtype=torch.float64, device=dev)
    g = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    args = lambda: (C.c_void_p(xd.data_ptr()), C.c_void_p(zd.data_ptr()), n, lat, ntess, C.c_void_p(e.data_ptr()), C.c_void_p(g.data_ptr()))
    eng.evaluate_lattice_dev(*args())                      # builds the rows (with skin)
    eng.sync()
    r0 = ref.evaluate_lattice(X0, Z, lat, ntess)
    _check_energy(e.cpu().numpy()[0:1], r0["Etotal"], "Etotal (build step)")
    rs = np.random.RandomState(3)
    X = X0.copy()
    for step in range(3):                                   # a short walk, every atom stays within skin / 2 of X0
        d = rs.randn(n, 3)
        d *= (0.09 * rs.rand(n, 1)) / np.linalg.norm(d, axis=1, keepdims=True)
        X = X + d
        xd.copy_(torch.tensor(X, dtype=torch.float64))
        eng.evaluate_lattice_dev(*args(), reuse_nlist=True)
        eng.sync()
        r = ref.evaluate_lattice(X, Z, lat, ntess)
        ee, gg = e.cpu().numpy(), g.cpu().numpy()
        scale = abs(r["Ebp"][0]) + abs(r["Ecc"][0]) + abs(r["Evdw"][0])     # (the parts cancel in Etotal for a small box)
        for k, name in enumerate(("Etotal", "Ebp", "Ecc", "Evdw")):
            assert abs(ee[k] - r[name][0]) <= 1e-6 * scale, (step, name, ee[k], r[name][0])
        assert np.abs(gg - r["gradient"][0]).max() <= 2e-6 * np.abs(r["gradient"]).max()
    if nx == 3:
        Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])
        o = og.Oracle([1, 8], W, P).evaluate_periodic(Xt, Zt, n)
        _check_energy(ee[0:1], o["Etotal"], "Etotal vs oracle")
        _check_grad(gg, o["gradient"][0, :n])
    # an atom that outruns the skin is reported at the next synchronisation, and a rebuild clears the condition
    X[0] += 0.5
    xd.copy_(torch.tensor(X, dtype=torch.float64))
    eng.evaluate_lattice_dev(*args(), reuse_nlist=True)
    with pytest.raises(TMolB200Error, match="skin"):
        eng.sync()
    eng.evaluate_lattice_dev(*args())
    eng.sync()
    r = ref.evaluate_lattice(X, Z, lat, ntess)
    assert abs(e.cpu().numpy()[0] - r["Etotal"][0]) <= 1e-6 * (abs(r["Ebp"][0]) + abs(r["Ecc"][0]) + abs(r["Evdw"][0]))


def test_device_md_with_verlet_skin_follows_the_every_step_rebuild():
    """DevicePeriodicVelocityVerlet(skin_, nl_every_): 40 NVE steps with the rows rebuilt every 8 steps stay on the trajectory
    of the run that rebuilds every step (same forces up to summation order)."""
    from tensormol_b200 import PARAMS, Mol
    from tensormol_b200.Simulations.DeviceMD import DevicePeriodicVelocityVerlet
    Z, X, lat = _skin_system(4)
    m = Mol(Z.astype(np.uint8), X)
    manager, _ = _manager([m], [32, 32], 1)
    PARAMS["MDMaxStep"] = 40; PARAMS["MDdt"] = 0.2; PARAMS["MDV0"] = None; PARAMS["MDThermostat"] = None; PARAMS["MDTemp"] = 300.0
    PARAMS["MDLogTrajectory"] = False
    v0 = 2e-3 * np.random.RandomState(5).randn(len(Z), 3)
    a = DevicePeriodicVelocityVerlet(manager, m, lat, "md_every", v0_=v0.copy(), sync_every_=20)
    la = a.Prop()
    xa, va = a.x.copy(), a.v.copy()
    b = DevicePeriodicVelocityVerlet(manager, m, lat, "md_skin", v0_=v0.copy(), sync_every_=20, skin_=0.5, nl_every_=8)
    lb = b.Prop()
    d = b.x - xa
    d -= np.round(d @ np.linalg.inv(lat)) @ lat          # same point modulo the lattice (b wraps only on building steps)
    assert np.abs(d).max() < 1e-6
    assert np.abs(b.v - va).max() < 1e-6 * max(1.0, np.abs(va).max() / 1e-3)
    assert np.abs(lb[:40, 5] - la[:40, 5]).max() <= 1e-6 * np.abs(la[:40, 5]).max()
    manager.Instances.engine.set_skin(0.0)


@pytest.mark.gpu
def test_page_locked_caller_arrays_and_bound_call_match_the_plain_call():
    """tm_eval_lattice wires page-locked caller arrays straight into the replayed graph's copy nodes: results equal
    the staged path, the arrays are re-read every call, and a change of arrays re-captures."""
    from oracle import oracle_graph as og
    from tensormol_b200.SystemBuilders import wrap_into_cell
    from tensormol_b200.engine import Engine, random_weights
    Z, X, lat = water_box(5, jitter=0.03)
    X = wrap_into_cell(X, lat)
    n = len(Z)
    eng = Engine([1, 8], [64, 64, 64], og.default_params())
    eng.set_weights(random_weights([1, 8], eng.D, [64, 64, 64], 0))
    ref = eng.evaluate_lattice(X, Z, lat, 1)
    Xp, Zp = eng.pinned(X.shape), eng.pinned(Z.shape, np.int32)
    Xp[:] = X
    Zp[:] = Z
    into = {"gradient": eng.pinned((1, n, 3)), "charge": np.zeros((1, n))}   # one page-locked, one ordinary
    call = eng.bind_lattice(Xp, Zp, lat, 1, into=into)
    for it in range(5):          # eager calls, capture, replays
        into["gradient"][:] = 7.0
        into["charge"][:] = 7.0
        r = call()
        assert r["gradient"] is into["gradient"]
        np.testing.assert_allclose(r["gradient"], ref["gradient"], rtol=0, atol=2e-8)
        np.testing.assert_allclose(r["charge"], ref["charge"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["Etotal"], ref["Etotal"], rtol=1e-12)
    X2 = wrap_into_cell(X + 0.05 * np.random.default_rng(1).normal(size=X.shape), lat)
    Xp[:] = X2                    # same arrays, new numbers: the replay must read them
    r2 = {k: v.copy() for k, v in call().items()}
    ref2 = eng.evaluate_lattice(X2, Z, lat, 1)
    assert abs(ref2["Etotal"][0] - ref["Etotal"][0]) > 1e-6
    np.testing.assert_allclose(r2["gradient"], ref2["gradient"], rtol=0, atol=2e-8)
    np.testing.assert_allclose(r2["Etotal"], ref2["Etotal"], rtol=1e-12)
    other = {"gradient": eng.pinned((1, n, 3))}     # different page-locked array: must not write the old one
    into["gradient"][:] = -3.0
    r3 = eng.evaluate_lattice(Xp, Zp, lat, 1, into=other)
    np.testing.assert_allclose(r3["gradient"], ref2["gradient"], rtol=0, atol=2e-8)
    assert np.all(into["gradient"] == -3.0)
    with pytest.raises(ValueError):
        eng.evaluate_lattice(Xp, Zp, lat, 1, into={"gradient": np.zeros((n, 2))})


@pytest.mark.gpu
def test_c3_nve_1000_steps_with_skin_sets_at_rebuild_steps_and_conservation():
    """Config C3 (3,000-atom water box), 1,000 NVE steps on the device driver with a 0.5 A Verlet skin and a rebuild every
    5 steps, run as four legs of 250 steps.  Every leg ends on a rebuild step: there the neighbour SETS of the current
    wrapped positions (tm_nlist through the MolEmb drop-in, images included) are bit-exact against the reference's own
    compiled Make_NListNaive (C_API/MolEmb.cpp:1180-1247, oracle/_ref), the evaluation that reuses the skin rows equals
    one that rebuilds them, and the total energy is conserved over the whole run."""
    import glob, importlib.util, os
    import MolEmb
    from conftest import ROOT
    from oracle import oracle_np as onp
    from tensormol_b200 import PARAMS, Mol
    from tensormol_b200.PhysicalData import JOULEPERHARTREE
    from tensormol_b200.Simulations.DeviceMD import DevicePeriodicVelocityVerlet
    from tensormol_b200.SystemBuilders import water_box as big_water_box, wrap_into_cell
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "MolEmb*.so"))
    if not so:
        pytest.skip("oracle/_ref/MolEmb not built (make -C oracle ref)")
    spec = importlib.util.spec_from_file_location("MolEmb", so[0])   # (not registered in sys.modules: `MolEmb` stays the drop-in)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    Z, X, lat = big_water_box(10, spacing=3.1072, seed=3, jitter=0.02)
    n = len(Z)
    m = Mol(Z.astype(np.uint8), wrap_into_cell(X, lat))
    manager, _ = _manager([m], [64, 64, 64], 0)
    eng = manager.Instances.engine
    PARAMS["MDMaxStep"] = 250; PARAMS["MDdt"] = 0.2; PARAMS["MDV0"] = None; PARAMS["MDThermostat"] = None; PARAMS["MDTemp"] = 300.0
    PARAMS["MDLogTrajectory"] = False
    v = 1e-3 * np.random.RandomState(1).randn(n, 3)
    etot, ekin = [], []
    for leg in range(4):
        md = DevicePeriodicVelocityVerlet(manager, m, lat, "c3_leg%d" % leg, v0_=v.copy(), sync_every_=250, skin_=0.5, nl_every_=5)
        log = md.Prop()
        etot.append(log[:250, 4] * n + log[:250, 5] * JOULEPERHARTREE)   # J/mol: kinetic per atom * n + potential
        ekin.append(log[:250, 4] * n)
        x = wrap_into_cell(md.x, lat)
        v = md.v.copy()
        m = Mol(Z.astype(np.uint8), x)
        # neighbour sets of the wrapped positions with their images, radial and angular cutoffs
        Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), x, PARAMS["EECutoffOff"])
        Xt = np.ascontiguousarray(Xt)
        for rc in (PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"]):
            want = ref.Make_NListNaive(Xt, float(rc), n, 1)
            got = MolEmb.Make_NListNaive(Xt, float(rc), n, 1)
            assert [sorted(r) for r in got] == [sorted(r) for r in want], "leg %d rc %g" % (leg, rc)
        # skin rows reused vs rebuilt, same positions (one more build with the skin on, then a reuse call through the driver's engine)
        eng.set_skin(0.5)
        want_out = ("Etotal", "Ebp", "Ecc", "Evdw", "gradient")
        a = eng.evaluate_lattice(x, Z, lat, md.ntess, outputs=want_out)
        eng.set_skin(0.0)
        b = eng.evaluate_lattice(x, Z, lat, md.ntess, outputs=want_out)
        scale = abs(b["Ebp"][0]) + abs(b["Ecc"][0]) + abs(b["Evdw"][0])      # the parts cancel in Etotal
        assert abs(a["Etotal"][0] - b["Etotal"][0]) <= 1e-7 * scale           # same sets, other summation order (fp32)
        assert np.abs(a["gradient"] - b["gradient"]).max() < 2e-7
    # the first leg again with a rebuild on every step: the skin run follows it energy for energy (same forces up to the
    # fp32 summation order; 50 fs is far below the time over which two such trajectories part)
    PARAMS["MDMaxStep"] = 250
    md0 = DevicePeriodicVelocityVerlet(manager, Mol(Z.astype(np.uint8), wrap_into_cell(X, lat)), lat, "c3_every",
                                       v0_=1e-3 * np.random.RandomState(1).randn(n, 3), sync_every_=250)
    log0 = md0.Prop()
    etot0 = log0[:250, 4] * n + log0[:250, 5] * JOULEPERHARTREE
    assert np.abs(etot[0] - etot0).max() < 1e-4 * np.ptp(ekin[0]), (np.abs(etot[0] - etot0).max(), np.ptp(ekin[0]))
    etot, ekin = np.concatenate(etot), np.concatenate(ekin)
    # kinetic and potential energy trade an amount ptp(ekin) (the random-weight potential is far from its minimum and, with
    # softplus(100 x) neurons, nearly piecewise linear: the box heats from 6 K to ~900 K); what velocity Verlet at 0.2 fs
    # loses or gains on that surface stays a small part of it
    assert np.ptp(etot[5:]) < 0.10 * np.ptp(ekin), (np.ptp(etot[5:]), np.ptp(ekin))
    eng.set_skin(0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("thermostat", [None, "Nose"])
def test_device_md_of_an_isolated_molecule_matches_the_host_driver(thermostat):
    """SURVEY 8f N1 remainder: DeviceVelocityVerlet (state on the GPU, tm_eval_dev, one CUDA graph per step) follows the
    host VelocityVerlet / NoseThermostat (numpy arrays, SimpleMD.py:14-129, 322-425) driven by the same device forces."""
    from tensormol_b200 import PARAMS, Mol, VelocityVerlet
    from tensormol_b200.Simulations.DeviceMD import DeviceVelocityVerlet
    g = load_golden("h2o_cluster")
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], [64, 64], 3)
    nstep = 15
    PARAMS["MDMaxStep"] = nstep; PARAMS["MDdt"] = 0.2; PARAMS["MDV0"] = None; PARAMS["MDTemp"] = 300.0
    PARAMS["MDThermostat"] = thermostat
    PARAMS["MDLogTrajectory"] = False

    def EnAndForce(x_, DoForce=True):
        out = manager.EvalBPDirectEEUpdateSingle(Mol(m.atoms, x_), PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)
        return (out[0][0], out[-1][0]) if DoForce else out[0][0]
    v0 = 1e-3 * np.random.RandomState(8).randn(len(g["Z"]), 3)
    host = VelocityVerlet(None, m, "host_mol_md", EnAndForce)
    host.v = v0.copy()
    if thermostat == "Nose":
        from tensormol_b200.Simulations.SimpleMD import NoseThermostat
        host.Tstat = NoseThermostat(host.m, host.v)     # rescales host.v to MDTemp, like the device driver's constructor
    host.Prop()
    dev = DeviceVelocityVerlet(manager, m, "dev_mol_md", v0_=v0.copy(), sync_every_=5)
    log = dev.Prop()
    assert np.abs(dev.x - host.x).max() < 1e-7
    assert np.abs(dev.v - host.v).max() < 1e-7 * max(1.0, np.abs(host.v).max() / 1e-3)
    assert abs(dev.EPot - host.EPot) < 1e-6 * abs(host.EPot)
    assert np.all(np.isfinite(log)) and abs(log[nstep - 1, 5] - dev.EPot) <= 1e-9 * abs(dev.EPot)
    # the library call itself: device pointers in and out, a padded two-molecule set against the host-buffer call
    import ctypes as C
    import torch
    eng = manager.Instances.engine
    n = len(g["Z"])
    xyz = np.zeros((2, n + 3, 3)); Z = np.zeros((2, n + 3), np.int32)
    xyz[0, :n] = g["xyz"]; Z[0, :n] = g["Z"]
    xyz[1, :n - 3] = g["xyz"][:n - 3] + 0.01; Z[1, :n - 3] = g["Z"][:n - 3]
    want = eng.evaluate(xyz, Z, np.array([n, n - 3]))
    dv = torch.device("cuda", 0)
    xt, zt = torch.tensor(xyz, device=dv), torch.tensor(Z, device=dv)
    e = torch.zeros(8, dtype=torch.float64, device=dv); gr = torch.zeros(2, n + 3, 3, dtype=torch.float64, device=dv)
    q = torch.zeros(2, n + 3, dtype=torch.float64, device=dv)
    eng.evaluate_dev(C.c_void_p(xt.data_ptr()), C.c_void_p(zt.data_ptr()), 2, n + 3, C.c_void_p(e.data_ptr()), C.c_void_p(gr.data_ptr()),
                     C.c_void_p(q.data_ptr()))
    eng.sync()
    e = e.cpu().numpy().reshape(4, 2)
    np.testing.assert_allclose(e[0], want["Etotal"], rtol=1e-9)
    np.testing.assert_allclose(e[2], want["Ecc"], rtol=1e-7)
    np.testing.assert_allclose(gr.cpu().numpy(), want["gradient"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(q.cpu().numpy(), want["charge"], rtol=0, atol=1e-8)
