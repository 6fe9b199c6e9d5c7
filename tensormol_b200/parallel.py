"""Owned-atom slab partition of one periodic evaluation over the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink).  Every rank holds all positions; rank r owns
the centres whose fractional coordinate along the first lattice vector falls in [r/W, (r+1)/W)
(the same rule `is_centre` applies on the device, tensormol_b200/csrc/tm_nlist.cu).  The data path needs
three small all-reduces per step:

    phase A  (device)  neighbour build, descriptors, both nets forward + unit backward for owned centres
    all-reduce  q_raw[nreal]                      (owned entries, zeros elsewhere)
    phase B  (device)  neutralisation, Coulomb/vdW pair kernel for owned centres
    all-reduce  e[6] = (-, Ebp, Ecc, Evdw, sum dE/dq, -)
    phase C  (device)  charge-net chain rule + descriptor-gradient force kernel
    all-reduce  grad[nreal,3]

With `enable_p2p()` the three host all-reduces disappear: every rank owns a symmetric device buffer that its peers can
address over NVLink (torch symmetric memory supplies the peer pointers: plumbing), the producing kernels store q_raw /
the energy partials / the force partials straight into every peer's buffer, bump a per-peer flag, and the consuming
phase begins with a device-side wait (tm_slab_p2p_setup, include/tmolb200.h).  A step is then pure device work and is
captured as ONE CUDA graph.

The reference has no counterpart (it is single-process); the partition reproduces
its periodic force convention because every rank drops the same image-row terms (Q10).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import TM_F_FORCE, TM_F_VDW, check


def slab_owner(xyz, lattice, world):
    """Rank that owns each atom: floor(frac_a * world) clipped to [0, world-1], frac_a = (x . inv(L))[0]."""
    inv = np.linalg.inv(np.asarray(lattice, np.float64))
    frac = np.asarray(xyz, np.float64) @ inv[:, 0]
    return np.clip(np.floor(frac * world).astype(np.int64), 0, world - 1)


class SlabEvaluator:
    """Drives the three device phases and the collectives between them.  `backend` is an Engine (CUDA)
    or any object with slab_phase_a/b/c taking torch tensors (the gloo CPU tests pass a numpy stand-in)."""

    def __init__(self, backend, nreal, rank, world, device, dist=None):
        import torch
        self.torch = torch
        self.backend = backend
        self.rank, self.world = int(rank), int(world)
        self.nreal = int(nreal)
        self.dist = dist
        self.qraw = torch.zeros(self.nreal, dtype=torch.float64, device=device)
        self.e = torch.zeros(6, dtype=torch.float64, device=device)
        self.grad = torch.zeros(self.nreal, 3, dtype=torch.float64, device=device)

    def enable_p2p(self, group=None):
        """Switches the exchange between the phases from NCCL all-reduces to direct peer-memory stores (needs an Engine
        backend, world > 1 and torch symmetric memory on an NVLink-connected node).  Returns True when active."""
        torch = self.torch
        if self.world <= 1 or self.dist is None or not hasattr(self.backend, "p2p_setup"):
            return False
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = self.backend.p2p_bytes(self.world, self.nreal)
            grp = group if group is not None else self.dist.group.WORLD
            if hasattr(symm, "enable_symm_mem_for_group"):      # needed by older torch releases, a no-op / deprecated later
                try:
                    symm.enable_symm_mem_for_group(grp.group_name)
                except Exception:
                    pass
            buf = symm.empty(nbytes, dtype=torch.uint8, device=self.qraw.device)
            hdl = symm.rendezvous(buf, grp)
            buf.zero_()
            torch.cuda.synchronize()
            self.dist.barrier()                    # every rank's flags are zero before anybody signals
            self.backend.p2p_setup(self.world, self.rank, self.nreal, [int(p) for p in hdl.buffer_ptrs])
            self._symm = (buf, hdl)                # keep the mapping alive
            self.p2p = True
        except Exception as ex:                    # no symmetric memory on this system: stay on NCCL
            self.p2p = False
            self.p2p_error = repr(ex)
        return self.p2p

    def _allreduce(self, t):
        if self.world > 1 and self.dist is not None:
            self.dist.all_reduce(t)

    def step(self, xyz, Z, lattice, ntess, do_force=True):
        """xyz [nreal,3] f64 and Z [nreal] i32 tensors on the device.  Returns (e, grad): e[0] = Etotal."""
        b = self.backend
        flags = (TM_F_FORCE if do_force else 0) | TM_F_VDW
        if getattr(self, "p2p", False):            # peer-memory exchange: no host collective between the phases
            b.slab_phase_a(xyz, Z, self.nreal, lattice, ntess, self.rank, self.world, self.qraw)
            b.slab_phase_b(self.qraw, self.e)
            b.slab_phase_c(self.e, flags, self.grad)
            return self.e, self.grad
        b.slab_phase_a(xyz, Z, self.nreal, lattice, ntess, self.rank, self.world, self.qraw)
        self._allreduce(self.qraw)
        b.slab_phase_b(self.qraw, self.e)
        self._allreduce(self.e)
        b.slab_phase_c(self.e, flags, self.grad)
        self._allreduce(self.grad)
        self.e[0] = self.e[1] + self.e[2] + self.e[3]
        return self.e, self.grad

    def capture(self, xyz, Z, lattice, ntess, stream, do_force=True):
        """Captures the three device phases as three CUDA graphs (the collectives between them stay eager NCCL calls),
        for fixed tensors xyz / Z and a fixed lattice: afterwards step_replay() costs 3 graph launches + 3 all-reduces
        instead of ~40 kernel launches.  Positions are updated by writing into `xyz` in place."""
        torch = self.torch
        b = self.backend
        flags = (TM_F_FORCE if do_force else 0) | TM_F_VDW
        with torch.cuda.stream(stream):
            self.step(xyz, Z, lattice, ntess, do_force)     # every library buffer reaches its final size
        stream.synchronize()
        self._graphs = []
        if getattr(self, "p2p", False):            # one graph for the whole step
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                self.step(xyz, Z, lattice, ntess, do_force)
            self._graphs = [g]
            return
        for fn in (lambda: b.slab_phase_a(xyz, Z, self.nreal, lattice, ntess, self.rank, self.world, self.qraw),
                   lambda: b.slab_phase_b(self.qraw, self.e),
                   lambda: b.slab_phase_c(self.e, flags, self.grad)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                fn()
            self._graphs.append(g)

    def step_replay(self):
        if len(self._graphs) == 1:
            self._graphs[0].replay()
            return self.e, self.grad
        ga, gb, gc = self._graphs
        ga.replay()
        self._allreduce(self.qraw)
        gb.replay()
        self._allreduce(self.e)
        gc.replay()
        self._allreduce(self.grad)
        self.e[0] = self.e[1] + self.e[2] + self.e[3]
        return self.e, self.grad


class EngineSlabBackend:
    """Adapter: torch CUDA tensors -> C-ABI slab entry points of libtmolb200."""

    def __init__(self, engine):
        self.eng = engine
        self.lib = engine.lib

    def slab_phase_a(self, xyz, Z, nreal, lattice, ntess, rank, world, qraw):
        lat = np.ascontiguousarray(lattice, np.float64).reshape(9)
        check(self.lib.tm_slab_phase_a(self.eng.ctx, C.c_void_p(xyz.data_ptr()), C.c_void_p(Z.data_ptr()), int(nreal),
                                       lat.ctypes.data_as(C.c_void_p), int(ntess), int(rank), int(world), C.c_void_p(qraw.data_ptr())), "tm_slab_phase_a")

    def p2p_bytes(self, world, nreal):
        return int(self.lib.tm_slab_p2p_bytes(int(world), int(nreal)))

    def p2p_setup(self, world, rank, nreal, ptrs):
        arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(p) for p in ptrs])
        check(self.lib.tm_slab_p2p_setup(self.eng.ctx, int(world), int(rank), int(nreal), arr), "tm_slab_p2p_setup")

    def slab_phase_b(self, qraw, e):
        check(self.lib.tm_slab_phase_b(self.eng.ctx, C.c_void_p(qraw.data_ptr()), C.c_void_p(e.data_ptr())), "tm_slab_phase_b")

    def slab_phase_c(self, e, flags, grad):
        check(self.lib.tm_slab_phase_c(self.eng.ctx, C.c_void_p(e.data_ptr()), int(flags), C.c_void_p(grad.data_ptr())), "tm_slab_phase_c")


# ---- molecule batches: independent units, sharded across the ranks -------------------------
def batch_shard_bounds(natom, world):
    """Contiguous molecule blocks [b[r], b[r+1]) per rank, cut where the running atom count crosses r/world of the total, so a
    ragged batch is balanced by work rather than by molecule count.  Deterministic, every molecule in exactly one block,
    blocks may be empty (fewer molecules than ranks)."""
    natom = np.asarray(natom, np.int64)
    csum = np.concatenate([[0], np.cumsum(natom)])
    total = int(csum[-1])
    b = [0]
    for r in range(1, int(world)):
        target = total * r / float(world)
        cut = int(np.searchsorted(csum, target, side="left"))
        # csum[cut] >= target > csum[cut-1]: take the nearer of the two cuts
        if cut > 0 and target - csum[cut - 1] < csum[min(cut, len(natom))] - target:
            cut -= 1
        b.append(min(max(cut, b[-1]), len(natom)))
    b.append(len(natom))
    return np.asarray(b, np.int64)


class BatchShardEvaluator:
    """Energy + force of a padded molecule batch (the EvalBPDirectEEUpdateSet / training-style contract) with the molecules
    sharded over the ranks: each rank evaluates its block with ONE tm_eval call; there is no data-path collective, only the
    final gather of the per-molecule results (one all_gather of a packed float64 table, blocks padded to the largest).
    `backend.evaluate(xyzs, Zs, natom, do_force=..., has_vdw=...)` is Engine.evaluate (CUDA) or a stand-in with the same
    signature (the gloo CPU test).  The reference has no counterpart (single process)."""

    KEYS = ("Etotal", "Ebp", "Ecc", "Evdw", "dipole", "Ebp_atom", "charge", "gradient")

    def __init__(self, backend, rank, world, dist=None, device="cpu"):
        self.backend, self.rank, self.world, self.dist, self.device = backend, int(rank), int(world), dist, device

    @staticmethod
    def _pack(r, nloc, maxn):
        row = np.zeros((nloc, 7 + 5 * maxn))
        row[:, 0], row[:, 1], row[:, 2], row[:, 3] = r["Etotal"], r["Ebp"], r["Ecc"], r["Evdw"]
        row[:, 4:7] = r["dipole"]
        row[:, 7:7 + maxn] = r["Ebp_atom"]
        row[:, 7 + maxn:7 + 2 * maxn] = r["charge"]
        row[:, 7 + 2 * maxn:] = np.asarray(r["gradient"]).reshape(nloc, 3 * maxn)
        return row

    @staticmethod
    def _unpack(row, maxn):
        n = row.shape[0]
        return dict(Etotal=row[:, 0].copy(), Ebp=row[:, 1].copy(), Ecc=row[:, 2].copy(), Evdw=row[:, 3].copy(), dipole=row[:, 4:7].copy(),
                    Ebp_atom=row[:, 7:7 + maxn].copy(), charge=row[:, 7 + maxn:7 + 2 * maxn].copy(),
                    gradient=row[:, 7 + 2 * maxn:].reshape(n, maxn, 3).copy())

    def evaluate_local(self, xyzs, Zs, natom, do_force=True, has_vdw=True):
        """This rank's block only: (lo, hi, result dict of the block's molecules).  What a training-style step needs (its
        losses are sums over molecules: one scalar all-reduce instead of the gather)."""
        b = batch_shard_bounds(natom, self.world)
        lo, hi = int(b[self.rank]), int(b[self.rank + 1])
        maxn = np.asarray(Zs).shape[1]
        if hi == lo:
            return lo, hi, self._unpack(np.zeros((0, 7 + 5 * maxn)), maxn)
        return lo, hi, self.backend.evaluate(xyzs[lo:hi], Zs[lo:hi], np.asarray(natom)[lo:hi], do_force=do_force, has_vdw=has_vdw)

    def evaluate(self, xyzs, Zs, natom, do_force=True, has_vdw=True):
        """Every rank returns the results of the WHOLE batch, in batch order."""
        import torch
        lo, hi, r = self.evaluate_local(xyzs, Zs, natom, do_force, has_vdw)
        nmol, maxn = np.asarray(Zs).shape
        if self.world == 1 or self.dist is None:
            return r
        b = batch_shard_bounds(natom, self.world)
        cap = int(np.max(np.diff(b)))
        mine = np.zeros((cap, 7 + 5 * maxn))
        mine[:hi - lo] = self._pack(r, hi - lo, maxn)
        send = torch.from_numpy(mine).to(self.device)
        recv = torch.empty((self.world * cap, send.shape[1]), dtype=send.dtype, device=self.device)     # concatenation along dim 0
        self.dist.all_gather_into_tensor(recv, send)
        recv = recv.cpu().numpy().reshape(self.world, cap, send.shape[1])
        rows = np.concatenate([recv[k, :int(b[k + 1] - b[k])] for k in range(self.world)], axis=0)
        assert rows.shape[0] == nmol
        return self._unpack(rows, maxn)


def sharded_batch_losses(backend, batch_data, rank, world, dist=None, device="cpu", EnergyScalar=1.0, GradScalar=1.0 / 20.0, DipoleScalar=1.0):
    """The reference's minibatch losses (BPInstance.batch_losses, TFMolInstanceDirect.py:4860-4901) with the molecules of the
    batch sharded over the ranks: each rank evaluates its block (BatchShardEvaluator.evaluate_local, one tm_eval), forms its
    partial sums of the three squared-residual losses in float64, and ONE all-reduce of three scalars completes them --
    the training-style, batch-sharded form of config C2.  batch_data as TData.GetTrainBatch / GetTestBatch return it
    (WithGrad_=True).  Returns the same six losses on every rank."""
    xyzs, Zs, Elabels, Dlabels, grads = batch_data[:5]
    inv_natom = np.asarray(batch_data[9], np.float64)
    natom = np.rint(1.0 / inv_natom).astype(np.int64)
    lo, hi, r = BatchShardEvaluator(backend, rank, world).evaluate_local(xyzs, Zs, natom)
    w = float(np.asarray(Zs).shape[1]) * inv_natom[lo:hi]
    part = np.array([0.5 * np.sum(((np.asarray(r["Etotal"]) - Elabels[lo:hi]) * w) ** 2),
                     0.5 * np.sum(((np.asarray(r["gradient"]) - grads[lo:hi]) * w[:, None, None]) ** 2),
                     0.5 * np.sum(((np.asarray(r["dipole"]) - Dlabels[lo:hi]) * w[:, None]) ** 2)])
    if world > 1 and dist is not None:
        import torch
        t = torch.from_numpy(part).to(device)
        dist.all_reduce(t)
        part = t.cpu().numpy()
    e_loss, g_loss, d_loss = (float(v) for v in part)
    loss_eg = e_loss * EnergyScalar + g_loss * GradScalar
    return dict(loss=loss_eg + d_loss * DipoleScalar, loss_dipole=d_loss, loss_EandG=loss_eg, energy_loss=e_loss, grads_loss=g_loss, dipole_loss=d_loss)
