"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, float64, autograd) of the TensorFlow graph
the reference builds for the BP+EE energy/force evaluation.

Nothing in the product package may import this module (see oracle/oracle_np.py header).

Parity pinning: PINNED by executing the reference's own code in the build container (the reference's tests hold no
golden vectors for this path, SURVEY.md §4, and TensorFlow is not installable):
  * oracle/ref_py.py runs the reference's evaluation graph itself -- RSF symmetry functions and electrostatics, the
    instance methods dipole_inference / energy_inference (+ periodic forms), tf.gradients -- unmodified on the eager
    torch stand-in oracle/tf_shim.py, with seeded weights and non-zero biases; the outputs are stored in
    tests/golden/ref_python_pins.npz (oracle/make_golden.py) and this module reproduces them to 1e-12 relative
    (tests/test_oracle.py: test_oracle_equals_reference_graph_*);
  * descriptor values and Jacobians against the reference's C implementation compiled from /root/reference
    (oracle/_ref/MolEmb: Make_ANI1_Sym, C_API/MolEmb.cpp:1913-1988; Make_ANI1_Sym_deri, :1844-1911);
  * central finite differences of its own energy and the closed-form constants of SURVEY.md §8 a11/a13.

Abbreviations: RSF = TensorMol/TFDescriptors/RawSymFunc.py, TMD =
TensorMol/TFNetworks/TFMolInstanceDirect.py, MGR = TensorMol/TFNetworks/TFMolManage.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import oracle_np as onp

BOHRPERA = 1.889725989                      # PhysicalData.py:30
KJPERHARTREE = 2625.499638                  # PhysicalData.py:34
JOULEPERHARTREE = KJPERHARTREE * 1000.0     # PhysicalData.py:35
C6_coff = {1: 0.14, 2: 0.08, 3: 1.16, 4: 1.61, 5: 3.13, 6: 1.75, 7: 1.23, 8: 0.70, 9: 0.75, 10: 0.63}   # PhysicalData.py:26
atomic_vdw_radius = {1: 1.001, 2: 1.012, 3: 0.825, 4: 1.408, 5: 1.485, 6: 1.452, 7: 1.397, 8: 1.342, 9: 1.287, 10: 1.243}  # :25

F64 = torch.float64


def default_params():
    """Hyper-parameters of the hot path with the reference defaults (TMParams.py:26-38,
    150-165) as set by the water/chemspider scripts (EECutoffOn=0)."""
    return dict(AN1_r_Rc=4.6, AN1_a_Rc=3.1, AN1_eta=4.0, AN1_zeta=8.0,
                AN1_num_r_Rs=32, AN1_num_a_Rs=8, AN1_num_a_As=8,
                EECutoffOn=0.0, EECutoffOff=15.0, Elu_Width=4.6, Poly_Width=4.6, DSFAlpha=0.18,
                AddEcc=True, sigmoid_alpha=100.0, NeuronType="sigmoid_with_param")


def dsf(R, R_c, alpha):
    """Util.py:172-181."""
    if R > R_c:
        return 0.0
    from scipy.special import erfc
    XX = alpha * R_c
    ZZ = erfc(XX) / R_c
    YY = 1.1283791671 * alpha * math.exp(-XX * XX) / R_c
    return erfc(alpha * R) / R - ZZ + (R - R_c) * (ZZ / R_c + YY)


def dsf_gradient(R, R_c, alpha):
    """Util.py:183-192."""
    if R > R_c:
        return 0.0
    from scipy.special import erfc
    XX = alpha * R_c
    ZZ = erfc(XX) / R_c
    YY = 1.1283791671 * alpha * math.exp(-XX * XX) / R_c
    return -((erfc(alpha * R) / R / R + 1.1283791671 * alpha * math.exp(-alpha * R * alpha * R) / R) - (ZZ / R_c + YY))


def elements_and_pairs(eles):
    """TMD:1262-1267: eles sorted ascending, pairs upper-triangular row-major."""
    eles = sorted(int(e) for e in eles)
    pairs = [[eles[i], eles[j]] for i in range(len(eles)) for j in range(i, len(eles))]
    return np.asarray(eles).reshape(-1, 1), np.asarray(pairs)


def ani1_params(P):
    """SetANI1Param (TMD:1293-1328): SFPr2 (1,nRs_r); SFPa2 (2,nAs,nRs_a) = (theta_a, Rs_s)."""
    nAs, nRa, nRr = P["AN1_num_a_As"], P["AN1_num_a_Rs"], P["AN1_num_r_Rs"]
    thetas = np.array([2.0 * math.pi * i / nAs for i in range(nAs)])
    rs = np.array([P["AN1_a_Rc"] * i / nRa for i in range(nRa)])
    rs_R = np.array([P["AN1_r_Rc"] * i / nRr for i in range(nRr)])
    SFPa2 = np.stack([np.tile(thetas[:, None], (1, nRa)), np.tile(rs[None, :], (nAs, 1))], axis=0)
    SFPr2 = rs_R.reshape(1, nRr)
    return SFPr2, SFPa2


def vdw_constants(eles):
    """TMD:3763-3767."""
    C6 = np.array([C6_coff[int(e)] * (BOHRPERA * 10.0) ** 6.0 / JOULEPERHARTREE for e in eles])
    Rv = np.array([atomic_vdw_radius[int(e)] * BOHRPERA for e in eles])
    return C6, Rv


def activation(x, P):
    """sigmoid_with_param (Util.py:200-201): log(1+exp(a x))/a, evaluated as a stable softplus
    (identical wherever the reference's expression does not overflow float64)."""
    kind = P.get("NeuronType", "sigmoid_with_param")
    if kind == "sigmoid_with_param":
        a = P["sigmoid_alpha"]
        return torch.nn.functional.softplus(x, beta=a, threshold=1e9)
    if kind == "relu":
        return torch.relu(x)
    if kind == "softplus":
        return torch.nn.functional.softplus(x, threshold=1e9)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "sigmoid":
        return torch.sigmoid(x)
    if kind == "elu":
        return torch.nn.functional.elu(x)
    if kind == "selu":                       # TFInstance.py:365-369: scale * where(x >= 0, x, alpha * elu(x))
        return 1.0507009873554804934193349852946 * torch.where(x >= 0, x, 1.6732632423543772848170429916717 * torch.nn.functional.elu(x))
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# Weights  (TFInstance.py:277-297; layer shapes TMD:5188-5202, 5244-5264)
# --------------------------------------------------------------------------------------
def mlp(x, layers, P):
    """layers = [(W1,b1),...,(W4,b4)]; y = a(xW+b) for hidden layers, last layer linear."""
    h = x
    for W, b in layers[:-1]:
        h = activation(h @ W + b, P)
    W, b = layers[-1]
    return (h @ W + b)[:, 0]


# --------------------------------------------------------------------------------------
# Descriptors (RSF:1696-1863 radial; RSF:868-1136 angular; RSF:2223-2398 element partition)
# --------------------------------------------------------------------------------------
def sym_radial(R, pairs_ele, SFPr2, eta, R_cut, nrows, nele):
    """TFSymRSet_Linear_WithEle(/Periodic): pairs_ele rows [mol,i,j,l]; returns (nmol,nrows,nele*nR).
    scatter_nd into slots followed by reduce_sum over slots == index_add over (mol,i,l)."""
    nmol = R.shape[0]
    nr = SFPr2.shape[1]
    out = torch.zeros(nmol * nrows * nele, nr, dtype=F64)
    if pairs_ele.shape[0] == 0:
        return out.reshape(nmol, nrows, nele * nr)
    m, i, j, l = (pairs_ele[:, c] for c in range(4))
    Rij = R[m, i] - R[m, j]                                            # DifferenceVectorsLinear RSF:123-133
    r = torch.sqrt((Rij * Rij).sum(1) + 1e-27)                         # RSF:1731
    tet = r[:, None] - torch.as_tensor(SFPr2[0], dtype=F64)[None, :]
    fac1 = torch.exp(-eta * tet * tet)                                 # RSF:1736
    fac2 = 0.5 * (torch.cos(3.14159265359 * r / R_cut) + 1.0)          # RSF:1738 (truncated pi, Q3)
    Gm = fac1 * fac2[:, None]
    out.index_add_(0, (m * nrows + i) * nele + l, Gm)
    return out.reshape(nmol, nrows, nele * nr)


def sym_angular(R, trip_elep, SFPa2, zeta, eta, R_cut, nrows, nelep):
    """TFSymASet_Linear_WithEle(/Periodic): trip_elep rows [mol,i,j,k,l]; layout theta-major."""
    nmol = R.shape[0]
    ntheta, nr = SFPa2.shape[1], SFPa2.shape[2]
    nsym = ntheta * nr
    out = torch.zeros(nmol * nrows * nelep, nsym, dtype=F64)
    if trip_elep.shape[0] == 0:
        return out.reshape(nmol, nrows, nelep * nsym)
    m, i, j, k, l = (trip_elep[:, c] for c in range(5))
    Rij = R[m, i] - R[m, j]
    Rik = R[m, i] - R[m, k]
    rij = torch.sqrt((Rij * Rij).sum(1) + 1e-27)                       # RSF:911
    rik = torch.sqrt((Rik * Rik).sum(1) + 1e-27)                       # RSF:913
    ToACos = (Rij * Rik).sum(1) / (rij * rik)
    onescalar = 1.0 - 0.0000000000000001
    ToACos = torch.where(ToACos >= 1.0, torch.full_like(ToACos, onescalar), ToACos)    # RSF:918
    ToACos = torch.where(ToACos <= -1.0, torch.full_like(ToACos, -onescalar), ToACos)  # RSF:919
    theta = torch.acos(ToACos)
    thetas = torch.as_tensor(SFPa2[0], dtype=F64)[None]               # (1,ntheta,nr)
    rs = torch.as_tensor(SFPa2[1], dtype=F64)[None]
    Tijk = torch.cos(theta[:, None, None] - thetas)
    fac1 = (2.0 ** (1.0 - zeta)) * torch.pow(1.0 + Tijk, zeta)         # RSF:929
    tet = ((rij + rik) / 2.0)[:, None, None] - rs
    fac2 = torch.exp(-eta * tet * tet)                                 # RSF:933
    fac3 = 0.5 * (torch.cos(3.14159265359 * rij / R_cut) + 1.0)        # RSF:935
    fac4 = 0.5 * (torch.cos(3.14159265359 * rik / R_cut) + 1.0)        # RSF:936
    Gm = (fac1 * fac2 * (fac3 * fac4)[:, None, None]).reshape(-1, nsym)
    out.index_add_(0, (m * nrows + i) * nelep + l, Gm)
    return out.reshape(nmol, nrows, nelep * nsym)


def descriptors(R, pairs_ele, trip_elep, P, nele, nelep, nrows):
    """GM = concat([GMR, GMA]) (RSF:2248)."""
    SFPr2, SFPa2 = ani1_params(P)
    GMR = sym_radial(R, pairs_ele, SFPr2, P["AN1_eta"], P["AN1_r_Rc"], nrows, nele)
    GMA = sym_angular(R, trip_elep, SFPa2, P["AN1_zeta"], P["AN1_eta"], P["AN1_a_Rc"], nrows, nelep)
    return torch.cat([GMR, GMA], dim=2)


# --------------------------------------------------------------------------------------
# Electrostatics / vdW   (RSF:1307-1465)
# --------------------------------------------------------------------------------------
def coulomb_elu_sr_dsf_lr(Rb, Qs, R_cut, pairs, alpha, elu_a, elu_shift, P):
    """TFCoulombEluSRDSFLR (RSF:1307-1359).  Rb in Bohr, pairs rows [mol,i,j]."""
    nmol = Rb.shape[0]
    out = torch.zeros(nmol, dtype=F64)
    if pairs.shape[0] == 0:
        return out
    alpha = alpha / BOHRPERA
    R_lrcut = P["EECutoffOff"] * BOHRPERA
    m, i, j = pairs[:, 0], pairs[:, 1], pairs[:, 2]
    Rij = Rb[m, i] - Rb[m, j]
    r = torch.sqrt((Rij * Rij).sum(1) + 1e-27)
    SR_sub = torch.where(r > R_cut, elu_a * (r - R_cut) + elu_shift, elu_a * (torch.exp(torch.clamp(r - R_cut, max=0.0)) - 1.0) + elu_shift)
    Qij = Qs[m, i] * Qs[m, j]
    XX = alpha * R_lrcut
    ZZ = math.erfc(XX) / R_lrcut
    YY = 1.1283791671 * alpha * math.exp(-XX * XX) / R_lrcut
    LR = Qij * (torch.erfc(alpha * r) / r - ZZ + (r - R_lrcut) * (ZZ / R_lrcut + YY))
    LR = torch.where(torch.isnan(LR), torch.zeros_like(LR), LR)
    LR = torch.where(r > R_lrcut, torch.zeros_like(LR), LR)
    SR = Qij * SR_sub
    K = torch.where(r > R_cut, LR, SR)
    out.index_add_(0, m, K)
    return out


def vdw_poly_lr(Rb, c6_i, c6_j, Rv_i, Rv_j, R_cut, pairs, P):
    """TFVdwPolyLR / TFVdwPolyLRWithEle (RSF:1361-1465).  Rb already in Bohr and scaled by
    BOHRPERA AGAIN (RSF:1377/1433, quirk Q6)."""
    nmol = Rb.shape[0]
    out = torch.zeros(nmol, dtype=F64)
    if pairs.shape[0] == 0:
        return out
    R = Rb * BOHRPERA
    R_width = P["Poly_Width"] * BOHRPERA
    m, i, j = pairs[:, 0], pairs[:, 1], pairs[:, 2]
    Rij = R[m, i] - R[m, j]
    r = torch.sqrt((Rij * Rij).sum(1) + 1e-27)
    t = (r - R_cut) / R_width
    Cut1 = torch.where(t > 0.0, -t * t * (2.0 * t - 3.0), torch.zeros_like(t))
    Cut = torch.where(t > 1.0, torch.ones_like(t), Cut1)
    Kern = -Cut * torch.sqrt(c6_i * c6_j) / torch.pow(r, 6.0) * 1.0 / (1.0 + 6.0 * torch.pow(r / (Rv_i + Rv_j), -12.0))
    out.index_add_(0, m, Kern)
    return out


# --------------------------------------------------------------------------------------
# Adam as TensorFlow 1.x applies it (tf.train.AdamOptimizer, training(), TMD:2639).  TensorFlow is an un-vendored,
# un-pinned dependency of the reference (README.md:42): restated from its published documentation -- parity unpinned.
#     lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);  m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;
#     variable -= lr_t m / (sqrt(v) + epsilon)          (epsilon OUTSIDE the bias correction: TF's "epsilon hat")
# --------------------------------------------------------------------------------------
def adam_update(g, m, v, t, lr, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """Returns (step to SUBTRACT from the variable, new m, new v) for the t-th application (t = 1, 2, ...)."""
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    return lr_t * m / (np.sqrt(v) + epsilon), m, v


# --------------------------------------------------------------------------------------
# Whole evaluation (TMD:5164-5285 aperiodic, 5774-5898 periodic; MGR:1260-1358)
# --------------------------------------------------------------------------------------
class Oracle:
    """weights = {"charge": {Z: [(W,b)...]}, "energy": {Z: [(W,b)...]}} of numpy float64 arrays."""

    def __init__(self, eles, weights, params=None):
        self.P = default_params()
        if params:
            self.P.update(params)
        self.eles_np, self.eles_pairs_np = elements_and_pairs(eles)
        self.eles = [int(e) for e in self.eles_np.reshape(-1)]
        self.nele = len(self.eles)
        self.nelep = self.eles_pairs_np.shape[0]
        self.C6, self.vdw_R = vdw_constants(self.eles)
        P = self.P
        if P["EECutoffOn"] != 0.0:
            raise Exception("EECutoffOn should equal to zero in DSF_elu")      # TMD:4366
        self.elu_shift = dsf(P["Elu_Width"] * BOHRPERA, P["EECutoffOff"] * BOHRPERA, P["DSFAlpha"] / BOHRPERA)      # TMD:4371
        self.elu_alpha = dsf_gradient(P["Elu_Width"] * BOHRPERA, P["EECutoffOff"] * BOHRPERA, P["DSFAlpha"] / BOHRPERA)
        self.w = {net: {int(z): [(torch.tensor(np.asarray(W, np.float64)), torch.tensor(np.asarray(b, np.float64)))     # copies: train_step updates in place
                                 for W, b in layers] for z, layers in d.items()} for net, d in weights.items()}

    # ---- shared pieces -------------------------------------------------------------
    def _nets(self, GM, Zs_rows, net):
        """per-element MLP + scatter back to [nmol, nrows] (TMD:5180-5211 / 5236-5272)."""
        nmol, nrows, D = GM.shape
        out = torch.zeros(nmol * nrows, dtype=F64)
        flatZ = torch.as_tensor(Zs_rows.reshape(-1).astype(np.int64))
        G2 = GM.reshape(nmol * nrows, D)
        for z in self.eles:
            idx = torch.nonzero(flatZ == z)[:, 0]
            if idx.numel() == 0:
                continue
            y = mlp(G2[idx], self.w[net][z], self.P)
            out = out.index_add(0, idx, y)
        return out.reshape(nmol, nrows)

    # ---- aperiodic -----------------------------------------------------------------
    def _graph(self, xyzs, Zs, natom, R=None):
        """The aperiodic graph up to Etotal as torch tensors (coordinates are the differentiable leaf R, or the tensor handed in)."""
        P = self.P
        xyzs = np.ascontiguousarray(xyzs, np.float64)
        Zs = np.asarray(Zs)
        natom = np.asarray(natom, np.int64)
        nmol, N, _ = xyzs.shape
        rp, tt, mil_jk, _ = onp.build_pairs_and_triples_with_ele_index(xyzs, natom, natom, Zs, P["AN1_r_Rc"], P["AN1_a_Rc"], self.eles_np, self.eles_pairs_np)
        ree = onp.set_build_pairs(xyzs, natom, natom, P["EECutoffOff"], False).astype(np.int64)
        rp = torch.as_tensor(rp.astype(np.int64))
        tt = torch.as_tensor(tt.astype(np.int64))
        ree = torch.as_tensor(ree)
        if R is None:
            R = torch.tensor(xyzs, dtype=F64, requires_grad=True)
        GM = descriptors(R, rp, tt, P, self.nele, self.nelep, N)
        # charges
        q_raw = self._nets(GM, Zs, "charge")
        inv_n = torch.as_tensor(1.0 / natom.astype(np.float64))
        q = q_raw - (q_raw.sum(1) * inv_n)[:, None]                          # TMD:5274-5277 (padded atoms also get -mean, Q11)
        Rb = R * BOHRPERA
        dipole = (Rb * q[:, :, None]).sum(1)                                 # TMD:5278-5279
        if P["AddEcc"]:
            Ecc = coulomb_elu_sr_dsf_lr(Rb, q, P["Elu_Width"] * BOHRPERA, ree, P["DSFAlpha"], self.elu_alpha, self.elu_shift, P)
        else:
            Ecc = torch.zeros(nmol, dtype=F64)
        Ebp_atom = self._nets(GM, Zs, "energy")
        Ebp = Ebp_atom.sum(1)
        if ree.shape[0]:
            zi = Zs[ree[:, 0].numpy(), ree[:, 1].numpy()]
            zj = Zs[ree[:, 0].numpy(), ree[:, 2].numpy()]
            ei = np.searchsorted(np.asarray(self.eles), zi)
            ej = np.searchsorted(np.asarray(self.eles), zj)
            Evdw = vdw_poly_lr(Rb, torch.as_tensor(self.C6[ei]), torch.as_tensor(self.C6[ej]), torch.as_tensor(self.vdw_R[ei]), torch.as_tensor(self.vdw_R[ej]),
                               P["EECutoffOn"] * BOHRPERA, ree, P)
        else:
            Evdw = torch.zeros(nmol, dtype=F64)
        Etotal = Ebp + Ecc + Evdw                                            # TMD:5213-5215
        return dict(R=R, GM=GM, q=q, dipole=dipole, Ecc=Ecc, Ebp_atom=Ebp_atom, Ebp=Ebp, Evdw=Evdw, Etotal=Etotal, rp=rp, tt=tt, ree=ree)

    def evaluate(self, xyzs, Zs, natom, has_vdw=True, want=("all",)):
        """EvalBPDirectEEUpdateSet/Single (MGR:1260-1321) + evaluate (TMD:5684-5711).
        xyzs (nmol,N,3) f64 padded with zeros, Zs (nmol,N) int padded with 0, natom (nmol,)."""
        g = self._graph(xyzs, Zs, natom)
        (grad,) = torch.autograd.grad(g["Etotal"].sum(), g["R"])             # TMD:5761
        n = lambda k: g[k].detach().numpy()                                  # noqa: E731
        return dict(Etotal=n("Etotal"), Ebp=n("Ebp"), Ebp_atom=n("Ebp_atom"), Ecc=n("Ecc"), Evdw=n("Evdw"), dipole=n("dipole"),
                    charge=n("q"), gradient=grad.numpy(), force=-JOULEPERHARTREE * grad.numpy(),
                    descriptors=n("GM"), rad_p_ele=g["rp"].numpy(), ang_t_elep=g["tt"].numpy(), ree=g["ree"].numpy())

    # ---- training quantities ------------------------------------------
    def train_quantities(self, xyzs, Zs, natom, Elabels, Dlabels, grads, EnergyScalar=1.0, GradScalar=1.0 / 20.0, DipoleScalar=1.0,
                         weight_decay=0.001):
        """Losses of one minibatch and the gradients the three train ops hand to Adam (TrainPrepare, TMD:4997-5049).
        With w = maxatom / natom per molecule (natom_pl is fed 1/natom, TMD:5146-5160; maxatom = the padded width, TMD:4861):
            energy_loss = 1/2 sum_m ((E_m - Elabel_m) w_m)^2,   grads_loss = 1/2 sum ((dE/dx - grads) w_m)^2,
            dipole_loss = 1/2 sum ((dipole - Dlabel) w_m)^2                               (tf.nn.l2_loss = sum(t^2)/2; TMD:4860-4868)
            loss = EnergyScalar energy_loss + GradScalar grads_loss + DipoleScalar dipole_loss; loss_dipole = dipole_loss;
            loss_EandG = EnergyScalar energy_loss + GradScalar grads_loss                   (TMD:4870-4901)
        The ops add to ONE 'losses' collection which already holds weight_decay * l2_loss(W) of every hidden-layer weight matrix
        of both nets (TFInstance.py:292-295, var_wd=0.001 at TMD:5188-5196, 5244-5256; the linear output layers have none),
        and every total is the sum of the collection AS IT STANDS when the op is built, so
            total_loss = decay + loss;  total_loss_dipole = total_loss + loss_dipole;  total_loss_EandG = total_loss_dipole + loss_EandG.
        train_op minimises total_loss over all variables, train_op_dipole total_loss_dipole over the DipoleNet variables,
        train_op_EandG total_loss_EandG over the EnergyNet variables (TMD:5046-5048, 2626-2646)."""
        frozen = self.w
        self.w = {net: {z: [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers] for z, layers in d.items()}
                  for net, d in frozen.items()}
        try:
            g = self._graph(xyzs, Zs, natom)
            (dEdx,) = torch.autograd.grad(g["Etotal"].sum(), g["R"], create_graph=True)
            w = torch.as_tensor(float(np.asarray(Zs).shape[1]) / np.asarray(natom, np.float64))
            e_loss = (((g["Etotal"] - torch.as_tensor(np.asarray(Elabels, np.float64))) * w) ** 2).sum() / 2
            g_loss = (((dEdx - torch.as_tensor(np.asarray(grads, np.float64))) * w[:, None, None]) ** 2).sum() / 2
            d_loss = (((g["dipole"] - torch.as_tensor(np.asarray(Dlabels, np.float64))) * w[:, None]) ** 2).sum() / 2
            loss_eg = e_loss * EnergyScalar + g_loss * GradScalar
            loss = loss_eg + d_loss * DipoleScalar
            decay = sum((W ** 2).sum() / 2 * weight_decay for d in self.w.values() for layers in d.values() for W, _ in layers[:-1])
            total = decay + loss
            total_d = total + d_loss
            total_eg = total_d + loss_eg
            flat = lambda net: [t for z in self.eles for Wb in self.w[net][z] for t in Wb]      # noqa: E731
            cv, ev = flat("charge"), flat("energy")
            zero = lambda gs, vs: [torch.zeros_like(v) if a is None else a for a, v in zip(gs, vs)]   # noqa: E731
            g_all = zero(torch.autograd.grad(total, cv + ev, retain_graph=True, allow_unused=True), cv + ev)
            g_dip = zero(torch.autograd.grad(total_d, cv, retain_graph=True, allow_unused=True), cv)
            g_eg = zero(torch.autograd.grad(total_eg, ev, allow_unused=True), ev)
        finally:
            live, self.w = self.w, frozen

        def unflat(net, gs):
            it = iter(gs)
            return {z: [(next(it).numpy(), next(it).numpy()) for _ in live[net][z]] for z in self.eles}

        n = lambda t: t.detach().numpy()                                     # noqa: E731
        return dict(Etotal=n(g["Etotal"]), dipole=n(g["dipole"]), gradient=n(dEdx), energy_loss=n(e_loss), grads_loss=n(g_loss),
                    dipole_loss=n(d_loss), loss=n(loss), loss_dipole=n(d_loss), loss_EandG=n(loss_eg), total_loss=n(total),
                    total_loss_dipole=n(total_d), total_loss_EandG=n(total_eg),
                    grad_train_op={"charge": unflat("charge", g_all[:len(cv)]), "energy": unflat("energy", g_all[len(cv):])},
                    grad_train_op_dipole=unflat("charge", g_dip), grad_train_op_EandG=unflat("energy", g_eg))

    def total_loss_gradient_by_tangent_pass(self, xyzs, Zs, natom, Elabels, Dlabels, grads, EnergyScalar=1.0, GradScalar=1.0 / 20.0,
                                            DipoleScalar=1.0, weight_decay=0.001):
        """d total_loss / d(all variables) WITHOUT differentiating the force back-pass -- the algorithm planned for the device
        (DESIGN.md section 6b).  With v = w^2 (dE/dx - g_label) held constant,
            d grads_loss / dW = d/dW <v, dE/dx> = d/dW [ d/d(eps) E(x + eps v; W) at eps = 0 ],
        so ONE forward-mode (tangent) pass of the energy graph along v gives a scalar whose ordinary back-pass is the force
        term of the weight gradient; the energy / dipole / decay terms are ordinary back-passes.  Returns the same nested dict
        as train_quantities()["grad_train_op"]; tests/test_oracle.py checks it equals that (double-backward) result."""
        import torch.autograd.forward_ad as fwAD
        r0 = self.evaluate(xyzs, Zs, natom)
        w = float(np.asarray(Zs).shape[1]) / np.asarray(natom, np.float64)
        v = (r0["gradient"] - np.asarray(grads, np.float64)) * (w ** 2)[:, None, None]
        frozen = self.w
        self.w = {net: {z: [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers] for z, layers in d.items()}
                  for net, d in frozen.items()}
        try:
            wt = torch.as_tensor(w)
            with fwAD.dual_level():
                R = fwAD.make_dual(torch.tensor(np.ascontiguousarray(xyzs, np.float64)), torch.as_tensor(v))
                g = self._graph(xyzs, Zs, natom, R=R)
                E, dE = fwAD.unpack_dual(g["Etotal"])
                dip = fwAD.unpack_dual(g["dipole"]).primal
                e_loss = (((E - torch.as_tensor(np.asarray(Elabels, np.float64))) * wt) ** 2).sum() / 2
                d_loss = (((dip - torch.as_tensor(np.asarray(Dlabels, np.float64))) * wt[:, None]) ** 2).sum() / 2
                decay = sum((W ** 2).sum() / 2 * weight_decay for d in self.w.values() for layers in d.values() for W, _ in layers[:-1])
                total = decay + EnergyScalar * e_loss + DipoleScalar * d_loss + GradScalar * dE.sum()
                flat = [t for net in ("charge", "energy") for z in self.eles for Wb in self.w[net][z] for t in Wb]
                gs = torch.autograd.grad(total, flat, allow_unused=True)
        finally:
            live, self.w = self.w, frozen
        it = iter(torch.zeros_like(t) if a is None else a for a, t in zip(gs, flat))
        return {net: {z: [(next(it).numpy(), next(it).numpy()) for _ in live[net][z]] for z in self.eles} for net in ("charge", "energy")}

    def tangent_stage_quantities(self, xyzs, Zs, natom, v):
        """Stage outputs of the tangent pass along the coordinate direction v (same shape as xyzs), as parity targets for the
        device kernels of the training step: descriptor tangents dG = (dG/dx) v, tangents of the atomic charges and atomic
        energies, and dE = <v, dE/dx> per molecule."""
        import torch.autograd.forward_ad as fwAD
        with torch.no_grad(), fwAD.dual_level():
            R = fwAD.make_dual(torch.tensor(np.ascontiguousarray(xyzs, np.float64)), torch.as_tensor(np.ascontiguousarray(v, np.float64)))
            g = self._graph(xyzs, Zs, natom, R=R)
            t = lambda k: fwAD.unpack_dual(g[k]).tangent                     # noqa: E731
            return dict(dG=t("GM").numpy().copy(), dq=t("q").numpy().copy(), dEbp_atom=t("Ebp_atom").numpy().copy(),
                        dEcc=t("Ecc").numpy().copy(), dEvdw=t("Evdw").numpy().copy(), dEtotal=t("Etotal").numpy().copy())

    def train_step(self, op, state, xyzs, Zs, natom, Elabels, Dlabels, grads, learning_rate, **scalars):
        """One minibatch of train_op ("all"), train_op_dipole ("dipole") or train_op_EandG ("EandG"): the gradients of
        train_quantities applied by Adam to that op's variable list (TMD:2626-2646: tf.train.AdamOptimizer(learning_rate) with
        TensorFlow's defaults, one optimizer -- its own moments and step count -- per op; `state` is that optimizer's dict,
        {} before its first step).  Updates self.w in place; returns train_quantities' dict (values BEFORE the update, as
        sess.run fetches them together with the train op)."""
        r = self.train_quantities(xyzs, Zs, natom, Elabels, Dlabels, grads, **scalars)
        sets = {"all": (("charge", r["grad_train_op"]["charge"]), ("energy", r["grad_train_op"]["energy"])),
                "dipole": (("charge", r["grad_train_op_dipole"]),), "EandG": (("energy", r["grad_train_op_EandG"]),)}[op]
        state["t"] = state.get("t", 0) + 1
        for net, gs in sets:
            for z in self.eles:
                for l, (gW, gb) in enumerate(gs[z]):
                    W, b = self.w[net][z][l]
                    for name, var, gr in (("W", W, gW), ("b", b, gb)):
                        m, v = state.setdefault((net, z, l, name), (np.zeros(gr.shape), np.zeros(gr.shape)))
                        step, m, v = adam_update(gr, m, v, state["t"], learning_rate)
                        state[(net, z, l, name)] = (m, v)
                        var -= torch.as_tensor(step)
        return r

    # ---- periodic (images supplied by the caller) -----------------------------------
    def evaluate_periodic(self, xyz_tess, Z_tess, nreal, do_force=True):
        """EvalBPDirectEEUpdateSinglePeriodic (MGR:1323-1358) + evaluate_periodic (TMD:5918-5947):
        xyz_tess (Ntess,3) with the nreal real atoms first, then images."""
        P = self.P
        x = np.ascontiguousarray(xyz_tess, np.float64)[None]
        Zs = np.asarray(Z_tess)[None].astype(np.int64)
        Nt = x.shape[1]
        nnz = np.array([Nt])
        nr = np.array([nreal])
        rp, tt, mil_j, mil_jk = onp.build_pairs_and_triples_with_ele_index_periodic(x, nnz, nr, Zs, P["AN1_r_Rc"], P["AN1_a_Rc"], self.eles_np, self.eles_pairs_np)
        ree5 = onp.build_pairs_with_both_ele_index(x, nnz, nr, Zs, P["EECutoffOff"], self.eles_np, True)
        rp = torch.as_tensor(rp.astype(np.int64))
        tt = torch.as_tensor(tt.astype(np.int64))
        ree = torch.as_tensor(ree5[:, :3].astype(np.int64))
        R = torch.tensor(x, dtype=F64, requires_grad=True)
        GM = descriptors(R, rp, tt, P, self.nele, self.nelep, nreal)
        q_raw = self._nets(GM, Zs[:, :nreal], "charge")
        q = q_raw - (q_raw.sum(1) * (1.0 / nreal))[:, None]                  # natom fed as nreal (MGR:1342)
        Rb = R * BOHRPERA
        dipole = (Rb[:, :nreal] * q[:, :, None]).sum(1)
        ntess = Nt // nreal
        q_all = q.repeat(1, ntess)                                           # TMD:5892-5893
        if P["AddEcc"]:
            Ecc = coulomb_elu_sr_dsf_lr(Rb, q_all, P["Elu_Width"] * BOHRPERA, ree, P["DSFAlpha"], self.elu_alpha, self.elu_shift, P) / 2.0   # TMD:5896
        else:
            Ecc = torch.zeros(1, dtype=F64)
        Ebp_atom = self._nets(GM, Zs[:, :nreal], "energy")
        Ebp = Ebp_atom.sum(1)
        ei, ej = ree5[:, 3].astype(np.int64), ree5[:, 4].astype(np.int64)
        Evdw = vdw_poly_lr(Rb, torch.as_tensor(self.C6[ei]), torch.as_tensor(self.C6[ej]), torch.as_tensor(self.vdw_R[ei]), torch.as_tensor(self.vdw_R[ej]),
                           P["EECutoffOn"] * BOHRPERA, ree, P) / 2.0         # TMD:5824
        Etotal = Ebp + Ecc + Evdw
        res = dict(Etotal=Etotal.detach().numpy(), Ebp=Ebp.detach().numpy(), Ebp_atom=Ebp_atom.detach().numpy(),
                   Ecc=Ecc.detach().numpy(), Evdw=Evdw.detach().numpy(), dipole=dipole.detach().numpy(),
                   charge=q_all.detach().numpy(), descriptors=GM.detach().numpy(),
                   rad_p_ele=rp.numpy(), ang_t_elep=tt.numpy(), n_ee=int(ree.shape[0]))
        if do_force:
            (grad,) = torch.autograd.grad(Etotal.sum(), R)                   # TMD:5999
            res["gradient"] = grad.numpy()
            res["force"] = -JOULEPERHARTREE * grad.numpy()[0, :nreal].reshape(1, nreal, 3)   # MGR:1353
        return res

    def energy_only(self, xyzs, Zs, natom):
        return self.evaluate(xyzs, Zs, natom)["Etotal"]
