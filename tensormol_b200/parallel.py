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
