"""Periodic wrapper with the reference's classes (ForceModifiers/Periodic.py:12-473): Lattice (wrap + tessellate),
LocalForce, PeriodicForce.  General (non-orthorhombic) cells are supported like in the reference.

B200 extension: a force bound with BindLatticeForce(f, rng) receives (z, x_wrapped, lattice, ntess, DoForce) and lets
the library tessellate on the device (tm_eval_lattice) instead of shipping (2 ntess+1)^3 image copies per step."""
from __future__ import annotations

import numpy as np

from ..Containers.Mol import Mol
from ..Math.LinearOperations import MatrixPower, MovingAverage
from ..Util import *   # noqa: F401,F403


class Lattice:
    def __init__(self, latvec_):
        """latvec_: 3x3 tensor of lattice vectors (rows)."""
        self.lattice = np.array(latvec_, dtype=np.float64).copy()
        self.latticeCenter = (self.lattice[0] + self.lattice[1] + self.lattice[2]) / 2.0
        d = np.linalg.norm(self.LatticeFacePoints() - self.latticeCenter[None, :], axis=1)
        self.latticeMinDiameter = 2.0 * np.min(d)
        L = self.lattice
        self.lp = np.array([np.zeros(3), L[0], L[1], L[2], L[0] + L[1], L[0] + L[2], L[1] + L[2], L[0] + L[1] + L[2]])
        self.ntess = 1
        self.facenormals = self.LatticeNormals()

    def LatticeFacePoints(self):
        """vertices, face centres and axis centres (Periodic.py:30-49)."""
        L = self.lattice
        return np.array([L[0], L[1], L[2], L[0] + L[1], L[0] + L[2], L[1] + L[2], L[0] + L[1] + L[2], np.zeros(3),
                         0.5 * (L[0] + L[1]), 0.5 * (L[2] + L[1]), 0.5 * (L[0] + L[2]),
                         0.5 * (L[0] + L[1]) + L[2], 0.5 * (L[2] + L[1]) + L[0], 0.5 * (L[0] + L[2]) + L[1]])

    def LatticeNormals(self):
        lp = self.lp
        fn = np.array([np.cross(lp[1] - lp[0], lp[2] - lp[0]), np.cross(lp[1] - lp[0], lp[3] - lp[0]), np.cross(lp[2] - lp[0], lp[3] - lp[0]),
                       np.cross(lp[4] - lp[-1], lp[5] - lp[-1]), np.cross(lp[6] - lp[-1], lp[5] - lp[-1]), np.cross(lp[6] - lp[-1], lp[4] - lp[-1])])
        return fn / np.sqrt(np.sum(fn * fn, axis=1))[:, np.newaxis]

    def InRangeOfLatNormals(self, pt, rng_):
        for i in range(6):
            ref = self.lp[0] if i < 3 else self.lp[7]
            if np.abs(np.sum(self.facenormals[i] * (pt - ref))) < rng_:
                return True
        return False

    def CenteredInLattice(self, mol):
        m = Mol(mol.atoms, self.ModuloLattice(mol.coords - mol.Center() + self.latticeCenter))
        m.properties["Lattice"] = self.lattice.copy()
        return m

    def InLat(self, crds):
        latmet = MatrixPower(np.dot(self.lattice, self.lattice.T), -1)
        return np.dot(crds, np.dot(self.lattice.T, latmet))

    def FromLat(self, crds):
        return np.dot(crds, self.lattice)

    def ModuloLattice(self, crds):
        """Transports all coordinates into the primitive cell (Periodic.py:87-100)."""
        fpart = np.fmod(self.InLat(crds), 1.0)
        fpart[fpart < 0.0] += 1.0
        return self.FromLat(fpart)

    def TessNTimes(self, atoms_, coords_, ntess_):
        """ntess_^3 positive-octant copies, originals first (Periodic.py:101-130)."""
        natom = atoms_.shape[0]
        newAtoms = np.zeros(ntess_ ** 3 * natom, dtype=np.uint8)
        newCoords = np.zeros((ntess_ ** 3 * natom, 3))
        newAtoms[:natom] = atoms_
        newCoords[:natom] = coords_
        ind = 1
        for i in range(ntess_):
            for j in range(ntess_):
                for k in range(ntess_):
                    if i == 0 and j == 0 and k == 0:
                        continue
                    newAtoms[ind * natom:(ind + 1) * natom] = atoms_
                    newCoords[ind * natom:(ind + 1) * natom] = coords_ + i * self.lattice[0] + j * self.lattice[1] + k * self.lattice[2]
                    ind += 1
        return newAtoms, newCoords

    def NTess(self, rng_):
        """Number of image shells needed for an interaction range (Periodic.py:143-147)."""
        if rng_ > self.latticeMinDiameter:
            return int(rng_ / self.latticeMinDiameter) + 1
        return 1

    def TessLattice(self, atoms_, coords_, rng_):
        """Real atoms first, then the (2 ntess+1)^3 - 1 image blocks in i,j,k loop order (Periodic.py:131-168)."""
        self.ntess = self.NTess(rng_)
        natom = atoms_.shape[0]
        side = 2 * self.ntess + 1
        newAtoms = np.zeros(side ** 3 * natom, dtype=np.uint8)
        newCoords = np.zeros((side ** 3 * natom, 3))
        newAtoms[:natom] = atoms_
        newCoords[:natom] = coords_
        ind = 1
        for i in range(-self.ntess, self.ntess + 1):
            for j in range(-self.ntess, self.ntess + 1):
                for k in range(-self.ntess, self.ntess + 1):
                    if i == 0 and j == 0 and k == 0:
                        continue
                    newAtoms[ind * natom:(ind + 1) * natom] = atoms_
                    newCoords[ind * natom:(ind + 1) * natom] = coords_ + i * self.lattice[0] + j * self.lattice[1] + k * self.lattice[2]
                    ind += 1
        return newAtoms, newCoords


class LocalForce:
    def __init__(self, f_, rng_=5.0, NeedsTriples_=False, lattice_form_=False):
        self.range = rng_
        self.func = f_
        self.NeedsTriples = NeedsTriples_
        self.lattice_form = lattice_form_

    def __call__(self, z, x, NZ, DoForce=True):
        return self.func(z, x, NZ, DoForce)


class PeriodicForce:
    def __init__(self, pm_, lat_):
        """pm_: a molecule; lat_: lattice vectors.  Short-ranged forces evaluated by tessellation."""
        self.lattice = Lattice(lat_)
        self.NL = None
        self.mol0 = self.lattice.CenteredInLattice(pm_)
        self.atoms = self.mol0.atoms.copy()
        self.natoms = self.mol0.NAtoms()
        self.natomsReal = pm_.NAtoms()
        self.maxrng = 0.0
        self.LocalForces = []
        self.lastx = np.zeros(pm_.coords.shape)
        self.nlthresh = 0.05

    def ReLattice(self, lat_):
        self.lattice = Lattice(lat_)

    def Density(self):
        """g/cm**3 of the bulk."""
        m = np.array([ATOMICMASSES[x - 1] for x in self.mol0.atoms]) * 1000.0   # noqa: F405
        latvol = np.linalg.det(self.lattice.lattice)
        return (np.sum(m) / AVOCONST) / (latvol * pow(10, -24))   # noqa: F405

    def AdjustLattice(self, x_, lat0_, latp_):
        latmet = MatrixPower(np.dot(lat0_, lat0_.T), -1)
        return np.dot(np.dot(x_, np.dot(lat0_.T, latmet)), latp_)

    def LatticeStep(self, x_):
        """Coordinate search over the nine lattice components (reference Periodic.py:319-363): each component is tried
        at +/- PARAMS["OptLatticeStep"]; a trial cell is adopted when the LAST bound local force gives an energy lower
        by more than 1e-5; sweeps repeat while any trial was adopted; the step is halved when a call adopts none."""
        xx = x_.copy()
        e, f = self.__call__(xx)
        dlat = PARAMS["OptLatticeStep"]
        moved_once, moved = False, True
        while moved:
            moved = False
            for i in range(3):
                for j in range(3):
                    for sign in (1.0, -1.0):
                        trial = self.lattice.lattice.copy()
                        trial[i, j] += sign * dlat
                        latt = Lattice(trial)
                        xtmp = latt.ModuloLattice(xx)
                        z, x = latt.TessLattice(self.atoms, xtmp, self.maxrng)
                        et, ft = (self.LocalForces[-1])(z, x, self.natomsReal)
                        if et < e and abs(e - et) > 0.00001:
                            e = et
                            self.ReLattice(trial)
                            xx = xtmp
                            moved = moved_once = True
        if not moved_once and PARAMS["OptLatticeStep"] > 0.001:
            PARAMS["OptLatticeStep"] = PARAMS["OptLatticeStep"] / 2.0
        return xx

    def Save(self, x_, name_="PMol"):
        m = Mol(self.atoms, x_)
        m.properties["Lattice"] = self.lattice.lattice.copy()
        m.WriteXYZfile("./results/", name_, 'w', True)

    def BindForce(self, lf_, rng_):
        """lf_(z, x_tess, nreal[, DoForce]) -> energy, force on >= nreal rows (Periodic.py:368-375)."""
        self.LocalForces.append(LocalForce(lf_, rng_))

    def BindLatticeForce(self, lf_, rng_):
        """B200 extension: lf_(z, x_wrapped, lattice, ntess, DoForce) -> energy, force[nreal]; images are made on the GPU."""
        self.LocalForces.append(LocalForce(lf_, rng_, lattice_form_=True))

    def __call__(self, x_, DoForce=True):
        """Energy per unit cell and force on all primitive atoms (Periodic.py:376-400)."""
        etore = 0.0
        ftore = np.zeros((self.natomsReal, 3))
        if self.maxrng == 0.0:
            self.maxrng = max([f.range for f in self.LocalForces])
        xw = self.lattice.ModuloLattice(x_)
        z = x = None
        for f in self.LocalForces:
            if f.lattice_form:
                ntess = self.lattice.NTess(self.maxrng)
                out = f.func(self.atoms, xw, self.lattice.lattice, ntess, DoForce)
            else:
                if z is None:
                    z, x = self.lattice.TessLattice(self.atoms, xw, self.maxrng)
                out = f(z, x, self.natomsReal) if DoForce else f(z, x, self.natomsReal, DoForce)
            if DoForce:
                einc, finc = out
                etore += np.sum(einc)
                ftore += np.asarray(finc)[:self.natomsReal]
            else:
                etore += np.sum(out)
        return etore, ftore      # the reference returns the pair for both values of DoForce (Periodic.py:400)

    def TestGradient(self, x_):
        """Walk along the force and print E vs the projected force (Periodic.py:401-422)."""
        e0, g0 = self.__call__(x_)
        g0 = g0 / JOULEPERHARTREE   # noqa: F405
        es = np.zeros(40)
        for i, d in enumerate(range(-20, 20)):
            dx = d * 0.01 * g0
            es[i], gi = self.__call__(x_ + dx)
            print("es ", es[i], i, np.sqrt(np.sum(dx * dx)), np.sum(gi / JOULEPERHARTREE * g0), np.sum(g0 * g0))   # noqa: F405
        return es

    def RDF(self, x_, z1=8, z2=8, rng=15.0, dx=0.02, name_="RDF.txt"):
        from .. import MolEmb
        zt, xt = self.lattice.TessLattice(self.atoms, x_, rng)
        ni = MolEmb.CountInRange(zt, xt, self.natoms, z1, z2, rng, dx)
        ri = np.arange(0.0, rng, dx)
        density = ni[-1] / (4.18879 * ri[-1] ** 3)
        x2gi = np.gradient(ni / (12.56637 * density), dx)
        gi = np.zeros(ri.shape)
        gi[1:] = x2gi[1:] / (ri[1:] * ri[1:])
        return MovingAverage(gi, 2)

    def RDF_inC(self, x_, z_, lat_, z1=8, z2=8, rng=10.0, dx=0.02, name_="RDF.txt"):
        """Radial distribution function of a CUBIC cell of edge lat_ from MolEmb.GetRDF_Bin (reference Periodic.py:451-472)."""
        from .. import MolEmb
        ri = np.arange(0.0, rng, dx)
        ni = np.zeros(ri.shape)
        for index in MolEmb.GetRDF_Bin(x_, z_, rng, dx, lat_, z1, z2):
            ni[index:] += 1
        ni /= float(np.count_nonzero(np.asarray(z_) == z1))
        density = ni[-1] / (4.18879 * ri[-1] * ri[-1] * ri[-1])
        x2gi = np.gradient(ni / (12.56637 * density), dx)
        with np.errstate(divide="ignore", invalid="ignore"):
            gi = x2gi / (ri * ri)
        gi[0] = 0.0
        return MovingAverage(gi, 2)
