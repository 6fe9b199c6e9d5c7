"""GPU tests of the widening rows added after the round's GPU budget was spent: they were written
against paths that ARE covered on the GPU elsewhere (the molecule-batch call of tests/test_a_gpu_parity.py), but were
themselves never run on a B200 before the round end -- hence a file that sorts last, so that under `-x` a surprise
here cannot hide the results of the established parity tests."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _tmp_results(tmp_path):
    from tensormol_b200 import PARAMS
    old = PARAMS["results_dir"]
    PARAMS["results_dir"] = str(tmp_path) + "/"
    yield
    PARAMS["results_dir"] = old


def _manager_and_mol():
    from test_c_gpu_api import _manager
    from tensormol_b200 import Mol
    g = load_golden("h2o_cluster")
    m = Mol(g["Z"].astype(np.uint8), g["xyz"])
    manager, W = _manager([m], [32, 32], 9)
    return manager, m


def test_batch_force_members_equal_single_evaluations():
    """TFMolManage.BatchForce: every member of one batched call equals EvalBPDirectEEUpdateSingle on that geometry
    (energy 1e-6 relative: the fp32 reductions of a batch member and of a lone molecule take the same order; force
    1e-4 Hartree/Bohr, the north-star tolerance)."""
    from tensormol_b200 import PARAMS, Mol
    from tensormol_b200.PhysicalData import BOHRPERA, JOULEPERHARTREE
    manager, m = _manager_and_mol()
    rs = np.random.RandomState(0)
    xs = m.coords[None] + 0.03 * rs.randn(5, *m.coords.shape)
    fb = manager.BatchForce(m.atoms)
    E, F = fb(xs)
    assert E.shape == (5,) and F.shape == xs.shape
    assert np.allclose(fb(xs, False), E, rtol=1e-7, atol=0)
    for i in range(5):
        out = manager.EvalBPDirectEEUpdateSingle(Mol(m.atoms, xs[i]), PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)
        assert abs(out[0][0] - E[i]) <= 1e-6 * abs(E[i])
        assert np.abs(out[-1][0] - F[i]).max() / JOULEPERHARTREE / BOHRPERA <= 1e-4
    with pytest.raises(ValueError):
        fb(xs[:, :-1])


def test_neb_batched_beads_on_gpu_potential():
    """A short nudged-elastic-band run between two distorted water clusters with all beads evaluated per iteration by
    one molecule-batch call follows the per-bead run (same solver, same callbacks otherwise)."""
    from tensormol_b200 import PARAMS, Mol, NudgedElasticBand
    manager, m = _manager_and_mol()
    rs = np.random.RandomState(1)
    m1 = Mol(m.atoms, m.coords + 0.05 * rs.randn(*m.coords.shape))

    def f(x_, DoForce=True):
        out = manager.EvalBPDirectEEUpdateSingle(Mol(m.atoms, x_), PARAMS["AN1_r_Rc"], PARAMS["AN1_a_Rc"], PARAMS["EECutoffOff"], True)
        return (out[0][0], out[-1][0]) if DoForce else out[0][0]

    old = PARAMS["NebSolver"]
    PARAMS["NebSolver"] = "Verlet"
    try:
        a = NudgedElasticBand(f, m, m1, nbeads_=5)
        b = NudgedElasticBand(None, m, m1, nbeads_=5, fb_=manager.BatchForce(m.atoms))
        for it in range(4):
            for neb in (a, b):
                neb.beads, e, neb.Fs = neb.Solver(neb.beads)
                neb.step += 1
            assert np.abs(a.beads - b.beads).max() <= 1e-6
            assert np.abs(a.Es - b.Es).max() <= 1e-6 * np.abs(a.Es).max()
    finally:
        PARAMS["NebSolver"] = old


def test_train_batches_equal_reference_python_on_cuda_tables():
    """The training-style batch provider (TensorMolData_BP_Direct_EE_WithEle.GetTrainBatch / GetTestBatch) with its neighbour
    tables built by the CUDA path equals the reference's Python executed in place, entry by entry (index tables bit-exact)."""
    from test_host_api import check_train_batches, train_pin_set
    g = load_golden("ref_train_pins")
    check_train_batches(train_pin_set(g), g)


def test_batch_losses_and_test_steps_equal_reference_graph():
    """The evaluation half of the reference's training loop on the CUDA path: BPInstance.batch_losses on the
    reference's own minibatches against the reference's loss ops executed in place (tests/golden/ref_train_pins.npz), and
    test() / test_dipole() / test_EandG() over the test cases against the oracle (itself pinned to those loss ops).
    Tolerances: the losses are sums of squared differences between fp32-path outputs and O(0.1) labels, so the north-star
    tolerances (energy 1e-5 relative, force 1e-4 Hartree/Bohr) map to 1e-4 relative on energy_loss / dipole_loss and 1e-3
    relative on grads_loss."""
    import random
    from test_c_gpu_api import NET, _setup_params
    from test_host_api import train_pin_set
    from test_oracle import TRAIN_PIN_ELES, TRAIN_PIN_HIDDEN, _train_quantities, train_pin_weights
    from tensormol_b200 import PARAMS, TFMolManage
    g = load_golden("ref_train_pins")
    _setup_params(TRAIN_PIN_HIDDEN)
    PARAMS["EnergyScalar"], PARAMS["GradScalar"], PARAMS["DipoleScalar"] = 1.0, 1.0 / 20.0, 1.0
    t = train_pin_set(g)
    assert [int(e) for e in t.eles] == TRAIN_PIN_ELES
    manager = TFMolManage("", t, False, NET, False, False)
    W = train_pin_weights()
    manager.SetWeights(W)
    I = manager.Instances
    random.seed(7)
    t.LoadDataToScratch(None)
    rtol = dict(energy_loss=1e-4, dipole_loss=1e-4, grads_loss=1e-3, loss=1e-4, loss_dipole=1e-4, loss_EandG=1e-4)

    def check(L, pre):
        for k, tol in rtol.items():
            assert abs(L[k] - float(g[pre + k])) <= tol * abs(float(g[pre + k])), (pre, k, L[k], float(g[pre + k]))
        assert np.abs(L["Etotal"] - g[pre + "Etotal"]).max() <= 1e-5 * np.abs(g[pre + "Etotal"]).max()

    check(I.batch_losses(t.GetTrainBatch(3), True), "tq_train0_ecc1_")
    check(I.batch_losses(t.GetTestBatch(2), False), "tq_test0_ecc0_")
    check(I.batch_losses(t.GetTestBatch(2), True), "tq_test1_ecc1_")
    assert PARAMS["AddEcc"] is True
    # the sharded form on one rank is the same arithmetic (tensormol_b200.parallel.sharded_batch_losses)
    from tensormol_b200.parallel import sharded_batch_losses
    b = t.GetTrainBatch(3)
    one, sh = I.batch_losses(b, True), sharded_batch_losses(I.engine, b, 0, 1, GradScalar=PARAMS["GradScalar"])
    for k in sh:
        assert abs(one[k] - sh[k]) <= 1e-5 * abs(one[k]), k
    # whole test steps: two test batches of two molecules (NTest = 4), pointer back at the first test case
    old = PARAMS["batch_size"]
    PARAMS["batch_size"] = 2
    try:
        for fn, key, ecc in ((I.test_EandG, "loss_EandG", True), (I.test_dipole, "loss_dipole", False), (I.test, "loss", True)):
            t.test_ScratchPointer = t.LastTrainMol
            want = sum(float(_train_quantities(g, tag, ecc, W)[key]) for tag in ("test0", "test1"))
            got = fn(0)
            assert abs(got - want) <= 1e-4 * abs(want), (key, got, want)
    finally:
        PARAMS["batch_size"] = old


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("kind", ["relu", "softplus", "tanh", "sigmoid", "elu", "selu"])
def test_other_neuron_types_against_oracle(kind, mode):
    """PARAMS["NeuronType"] other than the released nets' sigmoid_with_param (TFInstance.AssignActivation, TFInstance.py:108-140;
    SURVEY 8a a15): energy, charges and forces of the H2O cluster against the float64 oracle at the north-star tolerances, both
    GEMM paths.  (The back-pass forms a'(z) from the stored activation, which is why only the monotonic options are offered.)"""
    from oracle import oracle_graph as og
    from test_a_gpu_parity import _check_energy, _check_grad, _engine
    g = load_golden("h2o_cluster")
    eng, W, P = _engine(g["eles"], [64, 48, 32], 3, gemm_mode=mode, params={"NeuronType": kind})
    X, Z = g["xyz"][None], g["Z"][None].astype(np.int32)
    natom = np.array([Z.shape[1]])
    r = eng.evaluate(X, Z, natom)
    o = og.Oracle(g["eles"], W, P).evaluate(X, Z, natom)
    # signed activations let the atomic energies cancel in the sum: the 1e-5 is taken relative to sum |E_atom| (the conditioning
    # scale of Etotal) where that is larger than |Etotal|
    scale = max(abs(o["Etotal"][0]), np.abs(o["Ebp_atom"]).sum())
    assert abs(r["Etotal"][0] - o["Etotal"][0]) <= 1e-5 * scale, (kind, r["Etotal"], o["Etotal"])
    _check_energy(r["Ecc"], o["Ecc"], kind + " Ecc")
    assert np.abs(r["charge"] - o["charge"]).max() <= 1e-5 * max(np.abs(o["charge"]).max(), 1e-3)
    _check_grad(r["gradient"], o["gradient"])


def test_batch_shards_on_the_engine_equal_the_whole_batch():
    """parallel.BatchShardEvaluator's blocks evaluated one after another on ONE GPU (emulated ranks, no process group) and
    stitched together equal the unsharded tm_eval call member by member (molecules are independent units; fp32 sums inside a
    molecule do not depend on its neighbours in the batch: 1e-6 relative on energies, 1e-6 Ha/A on gradients)."""
    from test_a_gpu_parity import _engine
    from tensormol_b200.parallel import BatchShardEvaluator, batch_shard_bounds
    g = load_golden("h2o_cluster")
    eng, W, P = _engine(g["eles"], [32, 32], 5)
    rs = np.random.RandomState(3)
    nmol, N = 7, len(g["Z"])
    natom = np.array([N, 6, N, 3, 9, N, 12])
    xyzs = np.zeros((nmol, N, 3))
    Zs = np.zeros((nmol, N), np.int32)
    for m in range(nmol):
        xyzs[m, :natom[m]] = g["xyz"][:natom[m]] + 0.03 * rs.randn(natom[m], 3)
        Zs[m, :natom[m]] = g["Z"][:natom[m]]
    whole = eng.evaluate(xyzs, Zs, natom)
    for world in (2, 3, 8):
        b = batch_shard_bounds(natom, world)
        got = {k: [] for k in ("Etotal", "gradient", "charge", "dipole")}
        for rank in range(world):
            lo, hi, r = BatchShardEvaluator(eng, rank, world).evaluate_local(xyzs, Zs, natom)
            assert (lo, hi) == (b[rank], b[rank + 1])
            for k in got:
                got[k].append(np.asarray(r[k]).reshape((hi - lo,) + whole[k].shape[1:]))
        for k in got:
            a = np.concatenate(got[k], axis=0)
            assert a.shape == whole[k].shape
            assert np.abs(a - whole[k]).max() <= 1e-6 * max(np.abs(whole[k]).max(), 1.0), (world, k)
