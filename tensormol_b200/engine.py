"""Host-side handle on one libtmolb200 context: what `Instances.evaluate*` is in the reference
(TFMolInstanceDirect.py:5684-5711, 5918-5947) minus the TensorFlow session.

All numerics happen in the CUDA library; this file only marshals numpy arrays to the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import (TM_ACT, TM_F_DESCRIPTORS, TM_F_FOLD_IMAGES, TM_F_FORCE, TM_F_REUSE_NLIST, TM_F_VDW, TM_GEMM_FP32, TM_GEMM_TC_3XTF32, TM_GEMM_TC_SPLIT,
                   TM_NET_CHARGE, TM_NET_ENERGY, check, tm_model_desc, tm_outputs, tm_params, tm_timings)

BOHRPERA = 1.889725989
KJPERHARTREE = 2625.499638
JOULEPERHARTREE = KJPERHARTREE * 1000.0
_C6_coff = {1: 0.14, 2: 0.08, 3: 1.16, 4: 1.61, 5: 3.13, 6: 1.75, 7: 1.23, 8: 0.70, 9: 0.75, 10: 0.63}
_vdw_radius = {1: 1.001, 2: 1.012, 3: 0.825, 4: 1.408, 5: 1.485, 6: 1.452, 7: 1.397, 8: 1.342, 9: 1.287, 10: 1.243}


def DSF(R, R_c, alpha):
    """Damped shifted force kernel value (reference: Util.py:172-181)."""
    if R > R_c:
        return 0.0
    XX = alpha * R_c
    ZZ = math.erfc(XX) / R_c
    YY = 1.1283791671 * alpha * math.exp(-XX * XX) / R_c
    return math.erfc(alpha * R) / R - ZZ + (R - R_c) * (ZZ / R_c + YY)


def DSF_Gradient(R, R_c, alpha):
    """d/dR of DSF (reference: Util.py:183-192)."""
    if R > R_c:
        return 0.0
    XX = alpha * R_c
    ZZ = math.erfc(XX) / R_c
    YY = 1.1283791671 * alpha * math.exp(-XX * XX) / R_c
    return -((math.erfc(alpha * R) / R / R + 1.1283791671 * alpha * math.exp(-alpha * R * alpha * R) / R) - (ZZ / R_c + YY))


def default_params():
    """Hyper-parameters of the hot path with the reference's defaults (TMParams.py:26-38, 150-165) as set by its water /
    chemspider scripts (EECutoffOn = 0, sigmoid_with_param)."""
    return dict(AN1_r_Rc=4.6, AN1_a_Rc=3.1, AN1_eta=4.0, AN1_zeta=8.0, AN1_num_r_Rs=32, AN1_num_a_Rs=8, AN1_num_a_As=8,
                EECutoffOn=0.0, EECutoffOff=15.0, Elu_Width=4.6, Poly_Width=4.6, DSFAlpha=0.18, AddEcc=True,
                sigmoid_alpha=100.0, NeuronType="sigmoid_with_param")


def element_pairs(eles):
    """eles ascending; pairs upper-triangular row-major (TFMolInstanceDirect.py:1262-1267)."""
    eles = sorted(int(e) for e in eles)
    pairs = [[eles[i], eles[j]] for i in range(len(eles)) for j in range(i, len(eles))]
    return eles, pairs


def descriptor_width(n_ele, P):
    """inshape (TFMolInstanceDirect.py:1317)."""
    return n_ele * int(P["AN1_num_r_Rs"]) + (n_ele * (n_ele + 1) // 2) * int(P["AN1_num_a_Rs"]) * int(P["AN1_num_a_As"])


def random_weights(eles, D, hidden, seed=0):
    """Seeded stand-in for TF's truncated_normal initialiser with the reference's conventions
    (TFInstance.py:277-297; shapes TFMolInstanceDirect.py:5188-5202): W ~ N(0, sigma) resampled beyond
    2 sigma with sigma = 1/(10+sqrt(fan_in)), biases 0; creation order DipoleNet (all elements) then
    EnergyNet.  Returns {"charge": {Z: [(W,b),...]}, "energy": {...}} of float64 arrays."""
    rng = np.random.default_rng(seed)

    def tn(shape, sigma):
        w = rng.standard_normal(shape)
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        return w * sigma

    out = {}
    for net in ("charge", "energy"):
        out[net] = {}
        for z in eles:
            layers = []
            fan = D
            for h in hidden:
                layers.append((tn((fan, h), 1.0 / (10.0 + math.sqrt(float(fan)))), np.zeros(h)))
                fan = h
            layers.append((tn((fan, 1), 1.0 / (10.0 + math.sqrt(float(fan)))), np.zeros(1)))
            out[net][int(z)] = layers
    return out


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class GraphedCall:
    """Captures `fn()` -- device-pointer calls of this library (tm_eval_lattice_dev, the slab phases) and torch ops,
    all issued on `stream` -- into a CUDA graph and replays it: one launch per step instead of ~40, no host work
    between kernels.  The buffers `fn` touches must stay allocated and keep their addresses; new positions are
    written into the captured input tensor in place.  `fn` is run `warmup` times first so that every library buffer
    has its final size (allocation is not capturable)."""

    def __init__(self, fn, stream, warmup=3):
        import torch
        self.stream = stream
        with torch.cuda.stream(stream):
            for _ in range(max(1, warmup)):
                fn()
        stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=stream):
            fn()

    def __call__(self):
        self.graph.replay()


class Engine:
    """One CUDA context of the BP+EE evaluator."""

    def __init__(self, eles, hidden, params, device=0):
        self.lib = _lib.load()
        self.eles, self.eles_pairs = element_pairs(eles)
        self.hidden = [int(h) for h in hidden]
        self.P = dict(params)
        self.D = descriptor_width(len(self.eles), self.P)
        self._desc = tm_model_desc()
        self._desc.n_ele = len(self.eles)
        for i, z in enumerate(self.eles):
            self._desc.eles[i] = z
        self._desc.n_hidden = len(self.hidden)
        for i, h in enumerate(self.hidden):
            self._desc.hidden[i] = h
        self._params = self._make_params(self.P)
        self.ctx = self.lib.tm_create(int(device), C.byref(self._desc), C.byref(self._params))
        if not self.ctx:
            raise _lib.TMolB200Error("tm_create failed: " + self.lib.tm_last_error().decode())
        self.device = int(device)
        self.weights_set = False

    # ---- parameters ----------------------------------------------------------------------
    def _make_params(self, P):
        p = tm_params()
        p.r_Rc, p.a_Rc, p.eta, p.zeta = float(P["AN1_r_Rc"]), float(P["AN1_a_Rc"]), float(P["AN1_eta"]), float(P["AN1_zeta"])
        p.num_r_Rs, p.num_a_Rs, p.num_a_As = int(P["AN1_num_r_Rs"]), int(P["AN1_num_a_Rs"]), int(P["AN1_num_a_As"])
        p.ee_cutoff_on = float(P["EECutoffOn"])
        p.ee_cutoff_off = float(P["EECutoffOff"])
        p.elu_width = float(P["Elu_Width"])
        p.poly_width = float(P["Poly_Width"])
        p.dsf_alpha = float(P["DSFAlpha"])
        # TFMolInstanceDirect.py:4371-4372
        p.elu_shift = DSF(p.elu_width * BOHRPERA, p.ee_cutoff_off * BOHRPERA, p.dsf_alpha / BOHRPERA)
        p.elu_alpha = DSF_Gradient(p.elu_width * BOHRPERA, p.ee_cutoff_off * BOHRPERA, p.dsf_alpha / BOHRPERA)
        p.add_ecc = 1 if P["AddEcc"] else 0
        act = P.get("NeuronType", "sigmoid_with_param")
        if act not in TM_ACT:
            raise ValueError(f"NeuronType {act!r} is not supported on the B200 path (have {sorted(TM_ACT)})")
        p.activation = TM_ACT[act]
        p.sigmoid_alpha = float(P["sigmoid_alpha"])
        for i, z in enumerate(self.eles):   # TFMolInstanceDirect.py:3763-3767
            p.C6[i] = _C6_coff[z] * (BOHRPERA * 10.0) ** 6.0 / JOULEPERHARTREE
            p.Rvdw[i] = _vdw_radius[z] * BOHRPERA
        self.elu_shift, self.elu_alpha = p.elu_shift, p.elu_alpha
        return p

    def update_params(self, P):
        """Re-read the PARAMS the reference re-reads at evaluate time (NeuronType, AddEcc, EECutoffOff,
        Poly_Width: TFMolInstanceDirect.py:5691,5706; RawSymFunc.py:1324,1378)."""
        newP = dict(self.P)
        newP.update({k: P[k] for k in P if k in newP})
        if newP != self.P:
            self.P = newP
            self._params = self._make_params(newP)
            check(self.lib.tm_set_params(self.ctx, C.byref(self._params)), "tm_set_params")

    def set_weights(self, weights):
        for net_name, net_id in (("charge", TM_NET_CHARGE), ("energy", TM_NET_ENERGY)):
            for ei, z in enumerate(self.eles):
                layers = weights[net_name][z]
                if len(layers) != len(self.hidden) + 1:
                    raise ValueError("wrong number of layers")
                for li, (W, b) in enumerate(layers):
                    W = np.ascontiguousarray(W, np.float64)
                    b = np.ascontiguousarray(b, np.float64).reshape(-1)
                    check(self.lib.tm_set_weights(self.ctx, net_id, ei, li, _ptr(W), _ptr(b), W.shape[0], W.shape[1]), "tm_set_weights")
        self.weights_set = True

    def set_gemm_mode(self, mode):
        check(self.lib.tm_set_gemm_mode(self.ctx, int(mode)), "tm_set_gemm_mode")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.tm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- neighbour tables ------------------------------------------------------------------
    def nlist(self, xyz, rc, nreal, do_perms):
        xyz = np.ascontiguousarray(xyz, np.float64)
        n = xyz.shape[0]
        off, idx = C.c_void_p(), C.c_void_p()
        check(self.lib.tm_nlist(self.ctx, _ptr(xyz), n, int(nreal), float(rc), int(do_perms), C.byref(off), C.byref(idx)), "tm_nlist")
        offsets = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int64)), shape=(int(nreal) + 1,)).copy()
        total = int(offsets[-1])
        if total == 0:
            return offsets, np.zeros(0, np.int64)
        ind = np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_int64)), shape=(total,)).copy()
        return offsets, ind

    def pairs_triples_ele(self, xyzs, Zs, nnz, nreal, rr, ra):
        xyzs = np.ascontiguousarray(xyzs, np.float64)
        Zs = np.ascontiguousarray(Zs, np.int32)
        nmol, maxn = Zs.shape
        nnz = np.ascontiguousarray(nnz, np.int64)
        nreal = np.ascontiguousarray(nreal, np.int64)
        Pn, Tn = C.c_int64(), C.c_int64()
        ptrs = [C.c_void_p() for _ in range(4)]
        check(self.lib.tm_pairs_triples_ele(self.ctx, _ptr(xyzs), _ptr(Zs), nmol, maxn, _ptr(nnz), _ptr(nreal), float(rr), float(ra),
                                            C.byref(Pn), C.byref(Tn), *[C.byref(p) for p in ptrs]), "tm_pairs_triples_ele")

        def arr(p, rows, cols):
            if rows == 0:
                return np.zeros((0, cols), np.int64)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(rows, cols)).copy()

        return arr(ptrs[0], Pn.value, 4), arr(ptrs[1], Tn.value, 5), arr(ptrs[2], Pn.value, 4), arr(ptrs[3], Tn.value, 4)

    # ---- evaluation --------------------------------------------------------------------------
    @staticmethod
    def _flags(do_force, has_vdw, descriptors, fold=False):
        return (TM_F_FORCE if do_force else 0) | (TM_F_VDW if has_vdw else 0) | (TM_F_DESCRIPTORS if descriptors else 0) | (TM_F_FOLD_IMAGES if fold else 0)

    def evaluate(self, xyzs, Zs, natom, do_force=True, has_vdw=True, descriptors=False):
        """Aperiodic padded set: returns dict like Instances.evaluate (gradient = dE/dx in Hartree/A)."""
        xyzs = np.ascontiguousarray(xyzs, np.float64)
        Zs = np.ascontiguousarray(Zs, np.int32)
        natom = np.ascontiguousarray(natom, np.int64)
        nmol, maxn = Zs.shape
        res = dict(Etotal=np.zeros(nmol), Ebp=np.zeros(nmol), Ebp_atom=np.zeros((nmol, maxn)), Ecc=np.zeros(nmol), Evdw=np.zeros(nmol),
                   dipole=np.zeros((nmol, 3)), charge=np.zeros((nmol, maxn)), gradient=np.zeros((nmol, maxn, 3)))
        if descriptors:
            res["descriptors"] = np.zeros((nmol, maxn, self.D), np.float32)
        out = tm_outputs()
        for k in res:
            setattr(out, k, _ptr(res[k]))
        check(self.lib.tm_eval(self.ctx, _ptr(xyzs), _ptr(Zs), nmol, maxn, _ptr(natom), self._flags(do_force, has_vdw, descriptors), C.byref(out)), "tm_eval")
        return res

    def _periodic_result(self, nreal, ntot, descriptors, outputs=None, into=None):
        """Result arrays + the tm_outputs that points at them.  `outputs` = names to compute AND transfer (None = all):
        a pointer left NULL is skipped by the library, and its block never crosses the bus.  `into`: caller-owned float64
        arrays by name, written in place (page-locked ones — Engine.pinned — receive the device copy directly)."""
        shapes = dict(Etotal=(1,), Ebp=(1,), Ebp_atom=(1, nreal), Ecc=(1,), Evdw=(1,), dipole=(1, 3), charge=(1, ntot), gradient=(1, nreal, 3))
        names = list(shapes) if outputs is None else [k for k in shapes if k in outputs]
        res = {}
        for k in names:
            a = into.get(k) if into else None
            if a is not None and not (a.dtype == np.float64 and a.flags.c_contiguous and a.size == int(np.prod(shapes[k]))):
                raise ValueError("evaluate: into[%r] must be C-contiguous float64 with %d elements" % (k, int(np.prod(shapes[k]))))
            res[k] = a if a is not None else np.zeros(shapes[k])
        if descriptors:
            res["descriptors"] = np.zeros((1, nreal, self.D), np.float32)
        out = tm_outputs()
        for k in res:
            setattr(out, k, _ptr(res[k]))
        return res, out

    def evaluate_images(self, xyz_tess, Z_tess, nreal, do_force=True, has_vdw=True, descriptors=False, fold=False):
        xyz_tess = np.ascontiguousarray(xyz_tess, np.float64)
        Z_tess = np.ascontiguousarray(Z_tess, np.int32)
        nt = xyz_tess.shape[0]
        res, out = self._periodic_result(int(nreal), nt, descriptors)
        check(self.lib.tm_eval_images(self.ctx, _ptr(xyz_tess), _ptr(Z_tess), nt, int(nreal), self._flags(do_force, has_vdw, descriptors, fold), C.byref(out)),
              "tm_eval_images")
        return res

    def pinned(self, shape, dtype=np.float64):
        """A page-locked numpy array (torch pinned tensor underneath, kept alive by the array).  Coordinates, atomic
        numbers and result arrays kept in such memory go to and from the device without the staging memcpy."""
        import torch
        t = torch.zeros(shape, dtype={np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32,
                                      np.dtype(np.float32): torch.float32}[np.dtype(dtype)]).pin_memory()
        return t.numpy()

    def evaluate_lattice(self, xyz, Z, lattice, ntess, do_force=True, has_vdw=True, descriptors=False, fold=False, outputs=None, into=None):
        """Periodic cell + lattice.  outputs: e.g. ("Etotal", "gradient") = what the reference's periodic callback returns
        (TFMolManage.py:1353-1358); None = every output.  into: see _periodic_result."""
        xyz = np.ascontiguousarray(xyz, np.float64)
        Z = np.ascontiguousarray(Z, np.int32)
        lat = np.ascontiguousarray(lattice, np.float64).reshape(9)
        n = xyz.shape[0]
        res, out = self._periodic_result(n, n, descriptors, outputs, into)   # charges of the real atoms only
        check(self.lib.tm_eval_lattice(self.ctx, _ptr(xyz), _ptr(Z), n, _ptr(lat), int(ntess), self._flags(do_force, has_vdw, descriptors, fold), C.byref(out)),
              "tm_eval_lattice")
        return res

    def bind_lattice(self, xyz, Z, lattice, ntess, do_force=True, has_vdw=True, outputs=None, into=None):
        """evaluate_lattice with everything but the numbers fixed: returns call() -> the same result dict every time,
        re-reading `xyz` / `Z` (the arrays themselves, which must be C-contiguous float64 / int32 — update them in place)
        and overwriting the result arrays.  For MD-style loops: no per-call argument conversion, and with page-locked
        arrays (Engine.pinned) no host-side copy either."""
        if not (isinstance(xyz, np.ndarray) and xyz.dtype == np.float64 and xyz.flags.c_contiguous and
                isinstance(Z, np.ndarray) and Z.dtype == np.int32 and Z.flags.c_contiguous):
            raise ValueError("bind_lattice: xyz must be C-contiguous float64 and Z C-contiguous int32 numpy arrays")
        lat = np.ascontiguousarray(lattice, np.float64).reshape(9).copy()
        n = xyz.shape[0]
        res, out = self._periodic_result(n, n, False, outputs, into)
        args = (self.ctx, _ptr(xyz), _ptr(Z), n, _ptr(lat), int(ntess), self._flags(do_force, has_vdw, False, False), C.byref(out))
        fn, keep = self.lib.tm_eval_lattice, (xyz, Z, lat, out)

        def call():
            rc = fn(*args)
            if rc:
                check(rc, "tm_eval_lattice")
            return res
        call.keep = keep
        return call

    def evaluate_lattice_dev(self, xyz_ptr, Z_ptr, nreal, lattice, ntess, e_ptr, grad_ptr, charge_ptr=None, do_force=True, has_vdw=True,
                             reuse_nlist=False):
        """Device pointers in, device pointers out, no host synchronisation.  reuse_nlist: keep the neighbour rows of the
        previous call (set_skin > 0; positions not re-wrapped and within skin / 2 of those of the building call)."""
        lat = np.ascontiguousarray(lattice, np.float64).reshape(9)
        flags = self._flags(do_force, has_vdw, False) | (TM_F_REUSE_NLIST if reuse_nlist else 0)
        check(self.lib.tm_eval_lattice_dev(self.ctx, xyz_ptr, Z_ptr, int(nreal), _ptr(lat), int(ntess), flags,
                                           e_ptr, grad_ptr, charge_ptr), "tm_eval_lattice_dev")

    def evaluate_dev(self, xyz_ptr, Z_ptr, nmol, maxnatom, e_ptr, grad_ptr, charge_ptr=None, do_force=True, has_vdw=True):
        """Molecule set, device pointers in and out, no host synchronisation (tm_eval_dev): xyz [nmol*maxnatom*3] f64,
        Z [nmol*maxnatom] i32 (zero = padding), e [4*nmol] = Etotal | Ebp | Ecc | Evdw, grad [nmol*maxnatom*3], charge."""
        check(self.lib.tm_eval_dev(self.ctx, xyz_ptr, Z_ptr, int(nmol), int(maxnatom), self._flags(do_force, has_vdw, False),
                                   e_ptr, grad_ptr, charge_ptr), "tm_eval_dev")

    def set_skin(self, skin):
        """Verlet skin in Angstrom (0 = rebuild every call, the reference's behaviour): see tm_set_skin in include/tmolb200.h."""
        check(self.lib.tm_set_skin(self.ctx, float(skin)), "tm_set_skin")

    def set_stream(self, cuda_stream):
        check(self.lib.tm_set_stream(self.ctx, cuda_stream), "tm_set_stream")

    def sync(self):
        check(self.lib.tm_sync(self.ctx), "tm_sync")

    def timings(self):
        t = tm_timings()
        check(self.lib.tm_get_timings(self.ctx, C.byref(t)), "tm_get_timings")
        return {n: getattr(t, n) for n, _ in tm_timings._fields_}
