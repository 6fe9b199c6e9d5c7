"""The manager the drivers' closures call (reference: TFNetworks/TFMolManage.py:33-63, 1260-1465).

`TFMolManage(Name_, TData_, Train_, NetType_, RandomTData_, Trainable_)` keeps the reference's constructor and the
`EvalBPDirectEE*` methods with their return tuples; the TensorFlow instance behind them is replaced by one
tensormol_b200.engine.Engine (CUDA).  Pair/triple tables are NOT built on the host for these calls -- the fused
library call does neighbour search, descriptors, nets, electrostatics and forces on the device.

Weights: the reference restores a TensorFlow checkpoint (TFMolInstanceDirect.py:5765).  `LoadCheckpoint(prefix)` reads
such a checkpoint without TensorFlow (TFCheckpoint.py: the V2 tensor-bundle format and the reference's variable names),
and a manager built with a Name_ follows the reference's `<Name_>.tfm` -> `<instance>.tfn` -> `chk_file` chain
(TFMolManage.py:1468-1470, TFMolInstance.py:104-113).  Extensions: `InitRandom(seed)`, `SetWeights(dict)`,
`SaveWeights(path)` / `LoadWeights(path)` (npz, looked for first as PARAMS["networks_directory"]/<Name_>.npz),
`SaveCheckpoint(prefix)`.
"""
from __future__ import annotations

import os

import numpy as np

from ..Containers.Mol import Mol
from ..Containers.Sets import MSet
from ..Util import *   # noqa: F401,F403
from ..engine import Engine, element_pairs, random_weights

SUPPORTED_NETTYPES = ("fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize_Dropout",
                      "fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu_Normalize",
                      "fc_sqdiff_BP_Direct_EE_ChargeEncode_Update_vdw_DSF_elu",
                      "fc_sqdiff_BP_Direct_EE_SymFunction",
                      "fc_sqdiff_BP_Direct_EandG_SymFunction")


class BPInstance:
    """What `manager.Instances` exposes to scripts (eles_np, eles_pairs_np, ...) plus the CUDA engine."""

    def __init__(self, TData_, NetType_):
        self.TData = TData_
        self.NetType = NetType_
        self.eles, pairs = element_pairs(TData_.eles)
        self.n_eles = len(self.eles)
        self.eles_np = np.asarray(self.eles).reshape((self.n_eles, 1))
        self.eles_pairs = pairs
        self.eles_pairs_np = np.asarray(pairs)
        self.HiddenLayers = list(PARAMS["HiddenLayers"])
        self.Rr_cut = PARAMS["AN1_r_Rc"]
        self.Ra_cut = PARAMS["AN1_a_Rc"]
        self.Ree_on = PARAMS["EECutoffOn"]
        self.Ree_off = PARAMS["EECutoffOff"]
        self.DSFAlpha = PARAMS["DSFAlpha"]
        self.elu_width = PARAMS["Elu_Width"]
        self.inshape = int(self.n_eles * PARAMS["AN1_num_r_Rs"] + len(pairs) * PARAMS["AN1_num_a_Rs"] * PARAMS["AN1_num_a_As"])
        if self.Ree_on != 0.0:
            raise Exception("EECutoffOn should equal to zero in DSF_elu")
        self.engine = Engine(self.eles, self.HiddenLayers, self._params(), device=int(PARAMS.get("B200Device", 0)))
        self.engine.set_gemm_mode(int(PARAMS.get("B200GemmMode", 1)))
        self.elu_shift, self.elu_alpha = self.engine.elu_shift, self.engine.elu_alpha
        self.weights = None
        self.name = "Mol_" + str(TData_.name) + "_" + str(getattr(TData_.dig, "name", "")) + "_" + NetType_

    @staticmethod
    def _params():
        keys = ("AN1_r_Rc", "AN1_a_Rc", "AN1_eta", "AN1_zeta", "AN1_num_r_Rs", "AN1_num_a_Rs", "AN1_num_a_As", "EECutoffOn", "EECutoffOff",
                "Elu_Width", "Poly_Width", "DSFAlpha", "AddEcc", "sigmoid_alpha", "NeuronType")
        return {k: PARAMS[k] for k in keys}

    def refresh(self):
        """The reference re-reads NeuronType / AddEcc / EECutoffOff / Poly_Width from PARAMS at evaluate time."""
        self.engine.update_params({k: PARAMS[k] for k in ("NeuronType", "AddEcc", "EECutoffOff", "Poly_Width", "sigmoid_alpha")})

    def set_weights(self, w):
        self.engine.set_weights(w)
        self.weights = w

    # ---- the evaluation half of the training loop (reference TFMolInstanceDirect.py:4860-4901, 5420-5494) ----
    def batch_losses(self, batch_data, AddEcc=None):
        """The reference's loss values of one minibatch, from ONE forward + force evaluation on the CUDA path.
        batch_data = [xyzs, Zs, Elabels, Dlabels, grads, rad_p_ele, ang_t_elep, rad_eep, mil_jk, 1/natom] as
        TData.GetTrainBatch / GetTestBatch return it (WithGrad_=True); the neighbour tables in it are not needed, the
        library builds its own.  With w_m = MaxNAtoms / natom_m (the graph is fed 1/natom and multiplies by the padded
        width, :4861): energy_loss = 1/2 sum ((E - Elabel) w)^2, grads_loss = 1/2 sum ((dE/dx - grads) w)^2,
        dipole_loss = 1/2 sum ((dipole - Dlabel) w)^2; loss = EnergyScalar energy_loss + GradScalar grads_loss +
        DipoleScalar dipole_loss, loss_dipole = dipole_loss, loss_EandG = the first two terms.  Sums in float64 on the host
        (a few numbers per molecule).  Returns a dict with those six plus Etotal, Ecc, Evdw, dipole, charge."""
        if len(batch_data) < 10:
            raise Exception("batch_losses needs batches with gradient labels (TensorMolData built with WithGrad_=True)")
        xyzs, Zs, Elabels, Dlabels, grads = batch_data[:5]
        inv_natom = np.asarray(batch_data[9], np.float64)
        if not np.all(np.isfinite(Elabels)):
            raise Exception("DontEatShit")                                   # fill_feed_dict, :5157-5159
        natom = np.rint(1.0 / inv_natom).astype(np.int32)
        old = PARAMS["AddEcc"]
        if AddEcc is not None:
            PARAMS["AddEcc"] = bool(AddEcc)
        try:
            self.refresh()
            r = self.engine.evaluate(xyzs, Zs, natom, do_force=True, has_vdw=True)
        finally:
            PARAMS["AddEcc"] = old
            self.refresh()                      # the engine follows PARAMS again (a no-op when nothing changed)
        w = float(np.asarray(Zs).shape[1]) * inv_natom
        e_loss = 0.5 * np.sum(((r["Etotal"] - Elabels) * w) ** 2)
        g_loss = 0.5 * np.sum(((r["gradient"] - grads) * w[:, None, None]) ** 2)
        d_loss = 0.5 * np.sum(((r["dipole"] - Dlabels) * w[:, None]) ** 2)
        loss_eg = e_loss * PARAMS["EnergyScalar"] + g_loss * PARAMS["GradScalar"]
        return dict(loss=loss_eg + d_loss * PARAMS["DipoleScalar"], loss_dipole=d_loss, loss_EandG=loss_eg, energy_loss=e_loss,
                    grads_loss=g_loss, dipole_loss=d_loss, Etotal=r["Etotal"], Ecc=r["Ecc"], Evdw=r["Evdw"], dipole=r["dipole"], charge=r["charge"])

    def _test(self, step, which, AddEcc):
        import time
        bs = int(PARAMS["batch_size"])
        start = time.time()
        tot = dict(loss=0.0, energy_loss=0.0, grads_loss=0.0, dipole_loss=0.0)
        nmols = 0
        for _ in range(int(self.TData.NTest / bs)):
            L = self.batch_losses(self.TData.GetTestBatch(bs), AddEcc)
            tot["loss"] += L[which]
            for k in ("energy_loss", "grads_loss", "dipole_loss"):
                tot[k] += L[k]
            nmols += bs
        LOGGER.info("testing...")
        self.print_training(step, tot["loss"], tot["energy_loss"], tot["grads_loss"], tot["dipole_loss"], nmols, time.time() - start, False)
        return tot["loss"]

    def test(self, step):
        """Sum of `loss` over the test batches (:5600-5640); test_dipole feeds AddEcc = False, the others PARAMS["AddEcc"]."""
        return self._test(step, "loss", PARAMS["AddEcc"])

    def test_dipole(self, step):
        return self._test(step, "loss_dipole", False)

    def test_EandG(self, step):
        return self._test(step, "loss_EandG", PARAMS["AddEcc"])

    def print_training(self, step, loss, energy_loss, grads_loss, dipole_loss, Ncase, duration, Train=True):
        LOGGER.info("step: %7d  duration: %.5f  %s loss: %.10f  energy_loss: %.10f  grad_loss: %.10f, dipole_loss: %.10f", step, duration,
                    "train" if Train else "test", float(loss) / Ncase, float(energy_loss) / Ncase, float(grads_loss) / Ncase, float(dipole_loss) / Ncase)


class TFMolManage:
    def __init__(self, Name_="", TData_=None, Train_=True, NetType_="fc_sqdiff", RandomTData_=True, Trainable_=True):
        self.path = PARAMS["networks_directory"]
        self.TData = TData_
        self.NetType = NetType_
        self.n_train = PARAMS.get("max_steps", 0)
        self.Instances = None
        self.Trainable = Trainable_
        self.name = Name_
        if NetType_ not in SUPPORTED_NETTYPES:
            raise Exception("Unknown / unsupported Network Type on the B200 path: " + str(NetType_) + "  (supported: " + ", ".join(SUPPORTED_NETTYPES) + ")")
        if TData_ is None:
            raise Exception("TFMolManage needs a TensorMolData (element list, MaxNAtoms)")
        if Train_:
            raise NotImplementedError("training is outside the B200 hot path (SURVEY.md section 8f, N2); build the manager with Train_=False")
        self.Instances = BPInstance(TData_, NetType_)
        self.energy_only = NetType_.endswith("EandG_SymFunction")
        if Name_:
            f = os.path.join(self.path, Name_ + ".npz")
            if os.path.exists(f):
                self.LoadWeights(f)
            else:
                from .TFCheckpoint import find_reference_network
                chk, inst = find_reference_network(Name_, self.path)
                if chk:
                    self.TrainedNetworks = [os.path.basename(os.path.dirname(chk))]
                    self.LoadCheckpoint(chk, inst)
                else:
                    LOGGER.info("No %s and no %s.tfm with a checkpoint; call manager.InitRandom(seed) / SetWeights(...) / "
                                "LoadWeights(path) / LoadCheckpoint(prefix) before evaluating", f, Name_)

    # ---- weights (extension) -------------------------------------------------------------------
    def InitRandom(self, seed=0):
        """Random-init weights with the reference's initialiser statistics (TFInstance.py:277-297)."""
        I = self.Instances
        w = random_weights(I.eles, I.inshape, I.HiddenLayers, seed)
        I.set_weights(w)
        return w

    def SetWeights(self, weights):
        """weights = {"charge": {Z: [(W,b), ...]}, "energy": {Z: [(W,b), ...]}}; y = a(xW+b), last layer linear."""
        self.Instances.set_weights(weights)

    def LoadCheckpoint(self, prefix, instance_state=None):
        """Weights from a TensorFlow checkpoint of the reference (`<prefix>.index` + `.data-00000-of-00001`), read without
        TensorFlow.  `instance_state` (the unpickled .tfn, optional) is checked against this manager's network shape."""
        from .TFCheckpoint import CheckpointError, read_bundle, variable_names, weights_from_variables
        I = self.Instances
        if instance_state:
            hl = instance_state.get("HiddenLayers")
            if hl is not None and list(hl) != list(I.HiddenLayers):
                raise CheckpointError(f"the stored network has HiddenLayers {list(hl)}, PARAMS asks for {list(I.HiddenLayers)}")
            el = instance_state.get("eles")
            if el is not None and sorted(int(e) for e in el) != sorted(I.eles):
                raise CheckpointError(f"the stored network was trained for elements {sorted(int(e) for e in el)}, the set has {sorted(I.eles)}")
        wanted = [n for pair in variable_names(I.eles, len(I.HiddenLayers)).values() for n in pair]
        try:
            variables = read_bundle(prefix, names=wanted)
        except CheckpointError:
            variables = read_bundle(prefix)          # names under an outer scope: let weights_from_variables resolve them
        w = weights_from_variables(variables, I.eles, I.HiddenLayers, I.inshape)
        I.set_weights(w)
        return w

    def SaveCheckpoint(self, prefix, dtype=np.float64):
        """The current weights as a TensorFlow V2 checkpoint under the reference's variable names."""
        from .TFCheckpoint import variables_from_weights, write_bundle
        write_bundle(prefix, variables_from_weights(self.Instances.weights, dtype))
        return prefix

    def SaveWeights(self, path=None):
        path = os.path.join(self.path, self.name + ".npz") if path is None else path
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        out = {}
        for net, d in self.Instances.weights.items():
            for z, layers in d.items():
                for l, (W, b) in enumerate(layers):
                    out[f"{net}/{z}/{l}/W"] = W
                    out[f"{net}/{z}/{l}/b"] = b
        np.savez_compressed(path, **out)
        return path

    def LoadWeights(self, path):
        d = np.load(path)
        w = {"charge": {}, "energy": {}}
        for key in d.files:
            net, z, l, kind = key.split("/")
            w[net].setdefault(int(z), {}).setdefault(int(l), {})[kind] = d[key]
        w = {net: {z: [(ls[l]["W"], ls[l]["b"]) for l in sorted(ls)] for z, ls in dz.items()} for net, dz in w.items()}
        self.SetWeights(w)

    # ---- marshalling -----------------------------------------------------------------------------
    def _pad_set(self, mols):
        nmols = len(mols)
        maxn = max(m.NAtoms() for m in mols)
        self.TData.MaxNAtoms = maxn
        xyzs = np.zeros((nmols, maxn, 3), dtype=np.float64)
        Zs = np.zeros((nmols, maxn), dtype=np.int32)
        natom = np.zeros((nmols), dtype=np.int32)
        for i, mol in enumerate(mols):
            xyzs[i][:mol.NAtoms()] = mol.coords
            Zs[i][:mol.NAtoms()] = mol.atoms
            natom[i] = mol.NAtoms()
        return xyzs, Zs, natom

    def _check_cutoffs(self, Rr_cut, Ra_cut, Ree_cut=None):
        I = self.Instances
        if abs(Rr_cut - I.Rr_cut) > 1e-12 or abs(Ra_cut - I.Ra_cut) > 1e-12:
            raise Exception("cutoffs passed to Eval* must equal the instance's AN1_r_Rc / AN1_a_Rc")
        if Ree_cut is not None and abs(Ree_cut - PARAMS["EECutoffOff"]) > 1e-12:
            raise Exception("Ree_cut must equal PARAMS['EECutoffOff']")

    def _eval_set(self, mols, Rr_cut, Ra_cut, Ree_cut, do_force=True, has_vdw=True):
        self._check_cutoffs(Rr_cut, Ra_cut, Ree_cut)
        self.Instances.refresh()
        xyzs, Zs, natom = self._pad_set(mols)
        return self.Instances.engine.evaluate(xyzs, Zs, natom, do_force=do_force, has_vdw=has_vdw)

    # ---- aperiodic (TFMolManage.py:1260-1321, 1410-1439) ---------------------------------------------
    def EvalBPDirectEEUpdateSet(self, mol_set, Rr_cut, Ra_cut, Ree_cut, HasVdw=False):
        """Returns Etotal, Ebp, [Ebp_atom,] Ecc, [Evdw,] mol_dipole, atom_charge, force (J/mol/A, = -JOULEPERHARTREE * dE/dx)."""
        r = self._eval_set(mol_set.mols, Rr_cut, Ra_cut, Ree_cut)
        F = -JOULEPERHARTREE * r["gradient"]   # noqa: F405
        if not HasVdw:
            return r["Etotal"], r["Ebp"], r["Ecc"], r["dipole"], r["charge"], F
        return r["Etotal"], r["Ebp"], r["Ebp_atom"], r["Ecc"], r["Evdw"], r["dipole"], r["charge"], F

    @TMTiming("EvalBPDirectEEUpdateSingle")
    def EvalBPDirectEEUpdateSingle(self, mol, Rr_cut, Ra_cut, Ree_cut, HasVdw=False):
        s = MSet()
        s.mols.append(mol)
        return self.EvalBPDirectEEUpdateSet(s, Rr_cut, Ra_cut, Ree_cut, HasVdw)

    def EvalBPDirectEELinearSingle(self, mol, Rr_cut, Ra_cut, Ree_cut, HasVdw=False):
        return self.EvalBPDirectEEUpdateSingle(mol, Rr_cut, Ra_cut, Ree_cut, HasVdw)

    def BatchForce(self, atoms, HasVdw=True):
        """B200 extension: an energy / force callback over MANY geometries of one molecule,
        fb(xs[B, N, 3], DoForce=True) -> (E[B] Hartree, F[B, N, 3] J/mol/A) or E[B], evaluated as ONE molecule-batch call
        of the C-ABI (tm_eval) -- the marshalling of EvalBPDirectEEUpdateSet without building Mol objects. Each member
        equals EvalBPDirectEEUpdateSingle on that geometry. Used by NudgedElasticBand(fb_=...) to evaluate all beads of
        a band per solver iteration in one launch sequence."""
        Z = np.asarray(atoms, np.int32)
        I = self.Instances

        def fb(xs, DoForce=True):
            xs = np.ascontiguousarray(xs, np.float64)
            if xs.ndim != 3 or xs.shape[1] != len(Z) or xs.shape[2] != 3:
                raise ValueError("BatchForce expects coordinates of shape (B, %d, 3)" % len(Z))
            I.refresh()
            B = xs.shape[0]
            self.TData.MaxNAtoms = len(Z)
            r = I.engine.evaluate(xs, np.tile(Z, (B, 1)), np.full(B, len(Z), np.int32), do_force=DoForce, has_vdw=HasVdw)
            if not DoForce:
                return r["Etotal"]
            return r["Etotal"], -JOULEPERHARTREE * r["gradient"]   # noqa: F405
        return fb

    def EvalBPDirectEESingle(self, mol, Rr_cut, Ra_cut, Ree_cut):
        return self.EvalBPDirectEEUpdateSingle(mol, Rr_cut, Ra_cut, Ree_cut, False)

    def EvalBPDirectEESet(self, mol_set, Rr_cut=None, Ra_cut=None, Ree_cut=None):
        Rr_cut = PARAMS["AN1_r_Rc"] if Rr_cut is None else Rr_cut
        Ra_cut = PARAMS["AN1_a_Rc"] if Ra_cut is None else Ra_cut
        Ree_cut = PARAMS["EECutoffOff"] if Ree_cut is None else Ree_cut
        return self.EvalBPDirectEEUpdateSet(mol_set, Rr_cut, Ra_cut, Ree_cut, False)

    def EvalBPDirectChargeSingle(self, mol, Rr_cut, Ra_cut, Ree_cut, HasVdw=False):
        r = self._eval_set([mol], Rr_cut, Ra_cut, Ree_cut, do_force=False)
        return r["dipole"], r["charge"]

    def EvalBPDirectEandGLinearSingle(self, mol, Rr_cut, Ra_cut):
        """BP energy and gradient only (no electrostatics): Etotal, Ebp, Ebp_atom, force."""
        self._check_cutoffs(Rr_cut, Ra_cut)
        old = PARAMS["AddEcc"]
        PARAMS["AddEcc"] = False
        try:
            r = self._eval_set([mol], Rr_cut, Ra_cut, None, has_vdw=False)
        finally:
            PARAMS["AddEcc"] = old
        return r["Etotal"], r["Ebp"], r["Ebp_atom"], -JOULEPERHARTREE * r["gradient"]   # noqa: F405

    # ---- periodic (TFMolManage.py:1323-1358, 1441-1442) ---------------------------------------------
    @TMTiming("EvalBPDirectEEUpdateSinglePeriodic")
    def EvalBPDirectEEUpdateSinglePeriodic(self, mol, Rr_cut, Ra_cut, Ree_cut, nreal, HasVdw=True, DoForce=True, DoCharge=False):
        """mol holds the real atoms first and then their periodic images (PeriodicForce tessellation)."""
        self._check_cutoffs(Rr_cut, Ra_cut, Ree_cut)
        self.Instances.refresh()
        self.TData.MaxNAtoms = mol.NAtoms()
        r = self.Instances.engine.evaluate_images(mol.coords, np.asarray(mol.atoms, np.int32), int(nreal), do_force=DoForce, has_vdw=True)
        if not DoForce:
            return r["Etotal"]
        F = -JOULEPERHARTREE * r["gradient"][0][:nreal].reshape(1, nreal, 3)   # noqa: F405
        if not DoCharge:
            return r["Etotal"], F
        return r["Etotal"], F, r["charge"][0][:nreal].reshape(1, nreal)

    def EvalBPDirectEELinearSinglePeriodic(self, mol, Rr_cut, Ra_cut, Ree_cut, nreal, HasVdw=True, DoForce=True, DoCharge=False):
        return self.EvalBPDirectEEUpdateSinglePeriodic(mol, Rr_cut, Ra_cut, Ree_cut, nreal, HasVdw, DoForce, DoCharge)

    def EvalBPDirectEandGLinearSinglePeriodic(self, mol, Rr_cut, Ra_cut, nreal, DoForce=True):
        self._check_cutoffs(Rr_cut, Ra_cut)
        old = PARAMS["AddEcc"]
        PARAMS["AddEcc"] = False
        try:
            self.Instances.refresh()
            r = self.Instances.engine.evaluate_images(mol.coords, np.asarray(mol.atoms, np.int32), int(nreal), do_force=DoForce, has_vdw=False)
        finally:
            PARAMS["AddEcc"] = old
        if not DoForce:
            return r["Etotal"]
        return r["Etotal"], r["Ebp"], r["Ebp_atom"], -JOULEPERHARTREE * r["gradient"]   # noqa: F405

    # ---- periodic, images made on the device (B200 extension) ------------------------------------------
    def EvalBPDirectEEUpdateSingleLattice(self, atoms, coords_wrapped, lattice, ntess, DoForce=True, DoCharge=False):
        """Same results as EvalBPDirectEEUpdateSinglePeriodic on Lattice.TessLattice(atoms, coords, rng) but the
        (2 ntess+1)^3 image blocks are generated on the GPU (tm_eval_lattice)."""
        self.Instances.refresh()
        nreal = len(atoms)
        # only what this call returns crosses the bus (the reference returns Etotal, force[, charge]: TFMolManage.py:1353-1358)
        want = ("Etotal",) + (("gradient",) if DoForce else ()) + (("charge",) if DoCharge else ())
        r = self.Instances.engine.evaluate_lattice(coords_wrapped, np.asarray(atoms, np.int32), lattice, int(ntess), do_force=DoForce, has_vdw=True,
                                                   outputs=want)
        if not DoForce:
            return r["Etotal"]
        F = -JOULEPERHARTREE * r["gradient"][0].reshape(1, nreal, 3)   # noqa: F405
        if not DoCharge:
            return r["Etotal"], F
        return r["Etotal"], F, r["charge"][0][:nreal].reshape(1, nreal)

    def LatticeForce(self):
        """A callback for PeriodicForce.BindLatticeForce: f(z, x, lattice, ntess, DoForce) -> (E, F[nreal,3])."""
        def f(z, x, lattice, ntess, DoForce=True):
            if DoForce:
                e, frc = self.EvalBPDirectEEUpdateSingleLattice(z, x, lattice, ntess, True)
                return e[0], frc[0]
            return self.EvalBPDirectEEUpdateSingleLattice(z, x, lattice, ntess, False)[0]
        return f
