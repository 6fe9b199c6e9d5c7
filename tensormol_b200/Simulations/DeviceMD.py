"""Periodic velocity-Verlet / Nose-Hoover MD with the whole state on the GPU.

Same equations, units and attribute names as PeriodicVelocityVerlet / PeriodicNoseThermostat (reference:
Simulations/PeriodicMD.py:21-146, SimpleMD.py:90-129), but positions, velocities and accelerations never leave the
device: one MD step = integrator update + Lattice.ModuloLattice + the libtmolb200 energy/force call
(tm_eval_lattice_dev, images made on the GPU, neighbour list rebuilt every step), captured once as a CUDA graph and
replayed.  The host reads the (KE, EPot) log back every `sync_every` steps.

torch is used for device memory, the elementwise integrator arithmetic and the graph capture (plumbing); the
energy/force step is the C-ABI library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ..Containers.Mol import Mol
from ..Util import *   # noqa: F401,F403

_ACC = pow(10.0, -10.0)


class DevicePeriodicVelocityVerlet:
    def __init__(self, manager_, mol_, lattice_, name_="DevPdicMD", v0_=None, rng_=15.0, device_=0, graph_=True, sync_every_=100,
                 skin_=0.0, nl_every_=1):
        """manager_: a TFMolManage (its engine evaluates the forces); mol_: Mol of the primitive cell; lattice_: 3x3 rows.
        PARAMS: MDMaxStep, MDTemp, MDdt, MDV0 (None | "Random"), MDThermostat (None | "Nose").
        skin_ > 0 and nl_every_ > 1: Verlet skin (the idea at ForceModifiers/Periodic.py:224,262-268) -- the neighbour rows
        are built every nl_every_ steps out to cutoff + skin_ and reused in between (tm_set_skin / TM_F_REUSE_NLIST); the
        coordinates are wrapped into the cell only on the building steps.  The library checks on the device that no atom
        moved more than skin_ / 2 between builds and raises at the next synchronisation if one did."""
        from ..ForceModifiers.Periodic import Lattice
        self.lattice = Lattice(np.asarray(lattice_, np.float64))
        self.ntess = self.lattice.NTess(rng_)
        self.mol0 = self.lattice.CenteredInLattice(mol_)          # like PeriodicForce.__init__ (Periodic.py:288)
        self._setup(manager_, mol_, self.mol0.coords, name_, v0_, device_, graph_, sync_every_, skin_, nl_every_)

    def _setup(self, manager_, mol_, coords0, name_, v0_, device_, graph_, sync_every_, skin_=0.0, nl_every_=1):
        import torch
        from ..engine import GraphedCall
        self.torch = torch
        self.name = name_
        self.maxstep = int(PARAMS["MDMaxStep"])   # noqa: F405
        self.T = PARAMS["MDTemp"]                 # noqa: F405
        self.dt = float(PARAMS["MDdt"])           # noqa: F405
        manager_.Instances.refresh()
        self.engine = manager_.Instances.engine
        self.atoms = mol_.atoms.copy()
        self.natoms = len(self.atoms)
        self.m = np.array([ATOMICMASSES[z - 1] for z in self.atoms])   # noqa: F405
        dev = torch.device("cuda", device_)
        self.device = dev
        f64 = dict(dtype=torch.float64, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.engine.set_stream(C.c_void_p(self.stream.cuda_stream))
        if getattr(self, "lattice", None) is not None:
            L = self.lattice.lattice
            self._L = torch.tensor(L, **f64)
            self._toLat = torch.tensor(np.dot(L.T, np.linalg.inv(np.dot(L, L.T))), **f64)   # Lattice.InLat
        self._x = torch.tensor(coords0, **f64)
        v0 = np.zeros((self.natoms, 3)) if v0_ is None else np.asarray(v0_, np.float64)
        if v0_ is None and PARAMS["MDV0"] == "Random":   # noqa: F405
            from .SimpleMD import Thermostat
            v0 = np.random.randn(self.natoms, 3)
            Thermostat(self.m, v0)
        self._v = torch.tensor(v0, **f64)
        self._a = torch.zeros(self.natoms, 3, **f64)
        self._anew = torch.zeros(self.natoms, 3, **f64)
        self._tmp = torch.zeros(self.natoms, 3, **f64)
        self._frac = torch.zeros(self.natoms, 3, **f64)
        self._m = torch.tensor(self.m, **f64)
        # a = 1e-10 * F / m,  F = -JOULEPERHARTREE * dE/dx  (TFMolManage.py:1320, SimpleMD.py:52)
        self._gscale = torch.tensor(-_ACC * JOULEPERHARTREE / self.m, **f64).reshape(-1, 1)   # noqa: F405
        self._Z = torch.tensor(self.atoms.astype(np.int32), dtype=torch.int32, device=dev)
        self._e = torch.zeros(6, **f64)
        self._g = torch.zeros(self.natoms, 3, **f64)
        self._log = torch.zeros(self.maxstep + 1, 2, **f64)       # (KE per atom J/mol, EPot Hartree) per step
        self._row = torch.zeros(1, 2, **f64)
        self._count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.nose = PARAMS["MDThermostat"] == "Nose"   # noqa: F405
        if self.nose:
            from .SimpleMD import Thermostat
            vv = self._v.cpu().numpy()
            Thermostat(self.m, vv)                     # rescale to MDTemp like NoseThermostat.__init__
            self._v.copy_(torch.tensor(vv, **f64))
            kT = IDEALGASR * _ACC * self.T             # noqa: F405
            tau = 20.0 * self.dt * self.natoms
            self._Q = kT * tau * tau
            self._target = ((3.0 * self.natoms + 1) / 2.0) * kT
            self._eta = torch.zeros((), **f64)
            self._ke = torch.zeros((), **f64)
        self.sync_every = int(sync_every_)
        self.skin, self.nl_every = float(skin_), max(1, int(nl_every_))
        if self.nl_every > 1 and not self.skin > 0.0:
            raise ValueError("nl_every_ > 1 needs skin_ > 0")
        self.engine.set_skin(self.skin if self.nl_every > 1 else 0.0)
        self.t = 0.0
        self.md_log = None
        # initial energy (EPot0); like the reference's VelocityVerlet the acceleration starts at zero (SimpleMD.py:355)
        with torch.cuda.stream(self.stream):
            self._force()
        self.stream.synchronize()
        self.EPot0 = float(self._e[0].item())
        self.EPot = self.EPot0
        self.KE = 0.0
        step = self._step_nose if self.nose else self._step_nve
        self._replay = GraphedCall(step, self.stream, warmup=1) if graph_ else step
        self._replay_reuse = self._replay
        if self.nl_every > 1:      # a second step that keeps the neighbour rows (and does not wrap)
            reuse = lambda: step(True)
            self._replay_reuse = GraphedCall(reuse, self.stream, warmup=1) if graph_ else reuse
        if graph_:   # the capture and its warm-up advanced the state: rewind
            with torch.cuda.stream(self.stream):
                self._x.copy_(torch.tensor(coords0, **f64))
                self._v.copy_(torch.tensor(vv if self.nose else v0, **f64))
                if self.nose:
                    self._eta.zero_()
                self._count.zero_()
                self._a.zero_()
            self.stream.synchronize()

    # ---- device pieces (static shapes, in-place: capturable) --------------------------------------------------
    def _force(self, reuse=False):
        self.engine.evaluate_lattice_dev(C.c_void_p(self._x.data_ptr()), C.c_void_p(self._Z.data_ptr()), self.natoms, self.lattice.lattice,
                                         self.ntess, C.c_void_p(self._e.data_ptr()), C.c_void_p(self._g.data_ptr()), reuse_nlist=reuse)

    def _wrap(self):
        torch = self.torch
        torch.matmul(self._x, self._toLat, out=self._frac)
        torch.fmod(self._frac, 1.0, out=self._frac)
        self._frac.add_((self._frac < 0.0).to(self._frac.dtype))
        torch.matmul(self._frac, self._L, out=self._x)

    def _record(self):
        torch = self.torch
        v2 = (self._v * self._v).sum(dim=1)
        self._row[0, 0] = 0.5 * torch.dot(v2, self._m) * 1e10 / self.natoms     # KineticEnergy(), SimpleMD.py:33
        self._row[0, 1] = self._e[0]
        self._log.index_copy_(0, self._count, self._row)
        self._count.add_(1)

    def _step_nve(self, reuse=False):
        dt = self.dt
        self._x.add_(self._v, alpha=dt).add_(self._a, alpha=0.5 * dt * dt)
        if not reuse:
            self._wrap()
        self._force(reuse)
        self.torch.mul(self._g, self._gscale, out=self._anew)
        self._v.add_(self._a, alpha=0.5 * dt).add_(self._anew, alpha=0.5 * dt)
        self._a.copy_(self._anew)
        self._record()

    def _step_nose(self, reuse=False):
        torch = self.torch
        dt = self.dt
        # x += v dt + 1/2 (a - eta v) dt^2 ; v(dt/2) = v + 1/2 (a - eta v) dt      (PeriodicMD.py:28-31)
        torch.mul(self._v, self._eta, out=self._tmp)
        torch.sub(self._a, self._tmp, out=self._tmp)
        ke = 0.5 * torch.dot((self._v * self._v).sum(dim=1), self._m)
        self._x.add_(self._v, alpha=dt).add_(self._tmp, alpha=0.5 * dt * dt)
        if not reuse:
            self._wrap()
        self._v.add_(self._tmp, alpha=0.5 * dt)                       # now v(dt/2)
        self._force(reuse)
        torch.mul(self._g, self._gscale, out=self._a)
        kedto2 = 0.5 * torch.dot((self._v * self._v).sum(dim=1), self._m)
        self._eta.add_((dt / (2.0 * self._Q)) * (ke - self._target))
        self._eta.add_((dt / (2.0 * self._Q)) * (kedto2 - self._target))
        self._v.add_(self._a, alpha=0.5 * dt).div_(1.0 + 0.5 * dt * self._eta)
        self._record()

    # ---- host view ---------------------------------------------------------------------------------------------
    @property
    def x(self):
        return self._x.cpu().numpy()

    @property
    def v(self):
        return self._v.cpu().numpy()

    @property
    def a(self):
        return self._a.cpu().numpy()

    def Density(self):
        m_kg = float(np.sum(self.m))                                  # kg/mol
        vol = abs(float(np.linalg.det(self.lattice.lattice)))        # A^3
        return (m_kg * 1000.0 / AVOCONST) / (vol * 1e-24)             # noqa: F405  g/cm^3

    def _pull_log(self, nsteps):
        lg = self._log[:nsteps].cpu().numpy()
        self.md_log[:nsteps, 0] = np.arange(nsteps) * self.dt
        self.md_log[:nsteps, 4] = lg[:, 0]
        self.md_log[:nsteps, 5] = lg[:, 1]
        self.md_log[:nsteps, 6] = lg[:, 0] + (lg[:, 1] - self.EPot0) * JOULEPERHARTREE   # noqa: F405
        if nsteps:
            self.KE, self.EPot = float(lg[-1, 0]), float(lg[-1, 1])

    def WriteTrajectory(self):
        m = Mol(self.atoms, self.x)
        if getattr(self, "lattice", None) is not None:
            m.properties["Lattice"] = self.lattice.lattice.copy()
        m.properties["Time"] = self.t
        m.properties["KineticEnergy"] = self.KE
        m.properties["PotEnergy"] = self.EPot
        m.WriteXYZfile(PARAMS["results_dir"], "MDTrajectory" + self.name, 'a', True)   # noqa: F405

    def Prop(self, nsteps=None):
        """Runs MDMaxStep (or nsteps) steps; md_log columns as the reference's (time, -, -, -, KE, EPot, Etot-EPot0)."""
        torch = self.torch
        n = self.maxstep if nsteps is None else min(int(nsteps), self.maxstep)
        self.md_log = np.zeros((self.maxstep, 7))
        with torch.cuda.stream(self.stream):
            for step in range(n):
                if step % self.nl_every == 0:
                    self._replay()
                else:
                    self._replay_reuse()
                if (step + 1) % self.sync_every == 0:
                    self.stream.synchronize()
                    self._pull_log(step + 1)
                    self.t = (step + 1) * self.dt
                    if PARAMS["MDLogTrajectory"]:   # noqa: F405
                        self.WriteTrajectory()
                    LOGGER.info("Step: %i time: %.1f(fs) KE(kJ/mol): %.5f EPot(Eh): %.5f Etot(kJ/mol): %.5f",   # noqa: F405
                                step + 1, self.t, self.KE / 1000.0, self.EPot, self.KE / 1000.0 + self.EPot * KJPERHARTREE)   # noqa: F405
        self.stream.synchronize()
        self.engine.sync()      # device flags of the graph replays (e.g. an atom that outran the Verlet skin) surface here
        self._pull_log(n)
        self.t = n * self.dt
        return self.md_log


class DeviceVelocityVerlet(DevicePeriodicVelocityVerlet):
    """Isolated molecule: VelocityVerlet / NoseThermostat of the reference (Simulations/SimpleMD.py:14-38, 90-129, 322-425)
    with the state on the GPU; the force call is tm_eval_dev (device pointers in and out), one CUDA graph per MD step.
    Same PARAMS, attributes and md_log columns as the periodic device driver (column 4 is the kinetic energy of the step
    it is logged with; the reference's host loop logs the previous step's there, SimpleMD.py:406-411)."""

    def __init__(self, manager_, mol_, name_="DevMD", v0_=None, device_=0, graph_=True, sync_every_=100):
        self.lattice = None
        self._setup(manager_, mol_, np.asarray(mol_.coords, np.float64).copy(), name_, v0_, device_, graph_, sync_every_)

    def _force(self, reuse=False):
        self.engine.evaluate_dev(C.c_void_p(self._x.data_ptr()), C.c_void_p(self._Z.data_ptr()), 1, self.natoms,
                                 C.c_void_p(self._e.data_ptr()), C.c_void_p(self._g.data_ptr()))

    def _wrap(self):
        pass

    def Density(self):
        raise AttributeError("an isolated molecule has no cell")
