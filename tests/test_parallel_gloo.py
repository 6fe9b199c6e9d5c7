"""world_size-2 gloo test (CPU) of the multi-GPU host logic: the slab ownership rule partitions the atoms, and
SlabEvaluator's three phases + three all-reduces reproduce the single-rank result.  The device phases are replaced
by a numpy stand-in with the same dependency structure (charges need a global mean; the force needs the global sum
of dE/dq), so what is under test is tensormol_b200/parallel.py, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensormol_b200.parallel import SlabEvaluator, slab_owner


class NumpyBackend:
    """q_raw,i = 0.01 * (#neighbours within 3 A); q = q_raw - mean; E = sum_i 0.1 q_raw,i + 1/2 sum_{i!=j} q_i q_j exp(-r_ij);
    every per-atom piece is computed only for OWNED atoms, exactly like the device phases."""

    def __init__(self):
        self.state = {}

    def slab_phase_a(self, xyz, Z, nreal, lattice, ntess, rank, world, qraw):
        x = xyz.numpy()
        own = slab_owner(x, lattice, world) == rank
        d = np.linalg.norm(x[:, None] - x[None], axis=-1)
        np.fill_diagonal(d, 1e9)
        self.state = dict(x=x, own=own, d=d, n=nreal)
        q = np.zeros(nreal)
        q[own] = 0.01 * (d[own] < 3.0).sum(1)
        qraw.copy_(torch.from_numpy(q))

    def slab_phase_b(self, qraw, e):
        s = self.state
        q = qraw.numpy() - qraw.numpy().mean()
        k = np.exp(-s["d"])
        own = s["own"]
        dedq = np.zeros(s["n"])
        dedq[own] = (k[own] * q[None, :]).sum(1)
        s["q"], s["dedq"], s["qraw"] = q, dedq, qraw.numpy().copy()
        out = np.zeros(6)
        out[1] = 0.1 * s["qraw"][own].sum()
        out[2] = 0.5 * (q[own, None] * q[None, :] * k[own]).sum()
        out[4] = dedq[own].sum()
        e.copy_(torch.from_numpy(out))

    def slab_phase_c(self, e, flags, grad):
        s = self.state
        x, own, q, d = s["x"], s["own"], s["q"], s["d"]
        g = np.zeros_like(x)
        # d/dx of the pair term for owned centres i (acting on i and on j), charges held fixed
        for i in np.where(own)[0]:
            w = -0.5 * q[i] * q * np.exp(-d[i]) / d[i]
            vec = x[i] - x
            g[i] += (w[:, None] * vec).sum(0)
            g -= w[:, None] * vec
        # plus a term that needs the GLOBAL sum of dE/dq (stand-in for the neutralisation backward)
        g[own] += (s["dedq"][own] - e.numpy()[4] / s["n"])[:, None] * 1e-3
        grad.copy_(torch.from_numpy(g))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    n = 60
    lat = np.array([[12.0, 0, 0], [1.0, 11.0, 0], [0, 0.5, 10.0]])
    x = rng.uniform(0, 1, (n, 3)) @ lat
    ev = SlabEvaluator(NumpyBackend(), n, rank, world, "cpu", dist)
    e, g = ev.step(torch.from_numpy(x), torch.zeros(n, dtype=torch.int32), lat, 1)
    if rank == 0:
        q.put((e.numpy().copy(), g.numpy().copy()))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_slab_owner_partitions_atoms():
    rng = np.random.default_rng(1)
    lat = np.array([[12.0, 0, 0], [1.0, 11.0, 0], [0, 0.5, 10.0]])
    x = rng.uniform(0, 1, (500, 3)) @ lat
    for w in (1, 2, 4, 8):
        o = slab_owner(x, lat, w)
        assert o.min() >= 0 and o.max() <= w - 1
        frac = (x @ np.linalg.inv(lat))[:, 0]
        assert np.array_equal(o, np.clip(np.floor(frac * w), 0, w - 1).astype(int))
        assert np.bincount(o, minlength=w).sum() == 500
    # orthorhombic: slabs along x
    o = slab_owner(np.array([[0.1, 5, 5], [6.1, 5, 5], [11.9, 0, 0]]), np.eye(3) * 12.0, 2)
    assert o.tolist() == [0, 1, 1]


def test_two_rank_gloo_matches_single_rank():
    ctx = mp.get_context("spawn")
    res = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res[world] = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    e1, g1 = res[1]
    e2, g2 = res[2]
    assert np.allclose(e1[:5], e2[:5], rtol=1e-12, atol=1e-14)
    assert e1[0] == pytest.approx(e1[1] + e1[2] + e1[3])
    assert np.allclose(g1, g2, rtol=1e-12, atol=1e-14)
    assert np.abs(g1).max() > 0


# ---- molecule batches sharded over the ranks (config C2: independent units, no data-path collective, one final gather) ----
from tensormol_b200.parallel import BatchShardEvaluator, batch_shard_bounds   # noqa: E402


class MoleculeBackend:
    """Engine.evaluate's signature on a toy per-molecule model: every output depends on that molecule only."""

    def evaluate(self, xyzs, Zs, natom, do_force=True, has_vdw=True):
        nmol, maxn = Zs.shape
        mask = (np.arange(maxn)[None, :] < np.asarray(natom)[:, None]).astype(float)
        r2 = (xyzs ** 2).sum(-1) * mask
        q = (Zs * 0.01 + r2) * mask
        return dict(Etotal=r2.sum(1) + q.sum(1), Ebp=r2.sum(1), Ecc=q.sum(1), Evdw=np.zeros(nmol), dipole=(q[:, :, None] * xyzs).sum(1),
                    Ebp_atom=r2, charge=q, gradient=2 * xyzs * mask[:, :, None] * (1.0 if do_force else 0.0))


def _ragged_batch(nmol, seed=3):
    rng = np.random.default_rng(seed)
    natom = rng.integers(1, 9, nmol)
    maxn = 8
    xyzs = rng.normal(size=(nmol, maxn, 3))
    Zs = rng.choice([1, 6, 8], size=(nmol, maxn)).astype(np.int32)
    for m in range(nmol):
        xyzs[m, natom[m]:] = 0
        Zs[m, natom[m]:] = 0
    return xyzs, Zs, natom


def _labelled_batch(xyzs, Zs, natom):
    """GetTrainBatch's layout with seeded labels (the four index tables are not read by the loss code)."""
    rng = np.random.default_rng(9)
    n = len(natom)
    return [xyzs, Zs, rng.normal(size=n), rng.normal(size=(n, 3)), rng.normal(size=xyzs.shape), None, None, None, None, 1.0 / natom]


def test_batch_shard_bounds_cover_and_balance():
    for nmol in (0, 1, 2, 7, 100):
        natom = np.random.default_rng(nmol).integers(1, 50, nmol)
        for w in (1, 2, 3, 4, 8):
            b = batch_shard_bounds(natom, w)
            assert b[0] == 0 and b[-1] == nmol and len(b) == w + 1 and np.all(np.diff(b) >= 0)
    natom = np.random.default_rng(5).integers(1, 50, 1000)
    b = batch_shard_bounds(natom, 8)
    work = np.add.reduceat(natom, b[:-1])
    assert work.max() - work.min() <= 2 * natom.max()            # balanced by atoms to within a molecule either side
    assert batch_shard_bounds(np.full(16, 40), 4).tolist() == [0, 4, 8, 12, 16]


def _batch_worker(rank, world, port, q, nmol):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyzs, Zs, natom = _ragged_batch(nmol)
    ev = BatchShardEvaluator(MoleculeBackend(), rank, world, dist, "cpu")
    r = ev.evaluate(xyzs, Zs, natom)
    lo, hi, loc = ev.evaluate_local(xyzs, Zs, natom)
    s = torch.tensor([loc["Etotal"].sum()])
    dist.all_reduce(s)
    from tensormol_b200.parallel import sharded_batch_losses
    L = sharded_batch_losses(MoleculeBackend(), _labelled_batch(xyzs, Zs, natom), rank, world, dist)
    q.put((rank, r, float(s[0]), (lo, hi), L))
    dist.destroy_process_group()


@pytest.mark.parametrize("nmol", [11, 1])
def test_two_rank_gloo_batch_shards_equal_whole_batch(nmol):
    """Both ranks end with the whole batch's results in batch order, equal to one unsharded call (bit-exact: the units are
    independent); a batch smaller than the world leaves a rank with an empty block."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, 2, port, q, nmol)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xyzs, Zs, natom = _ragged_batch(nmol)
    whole = MoleculeBackend().evaluate(xyzs, Zs, natom)
    blocks = sorted(g[3] for g in got)
    assert blocks[0][0] == 0 and blocks[0][1] == blocks[1][0] and blocks[1][1] == nmol
    from tensormol_b200.parallel import sharded_batch_losses
    L1 = sharded_batch_losses(MoleculeBackend(), _labelled_batch(xyzs, Zs, natom), 0, 1)
    w = xyzs.shape[1] / natom
    assert L1["energy_loss"] == pytest.approx(0.5 * np.sum(((whole["Etotal"] - _labelled_batch(xyzs, Zs, natom)[2]) * w) ** 2), rel=1e-13)
    for rank, r, esum, _, L in got:
        for k in BatchShardEvaluator.KEYS:
            assert np.array_equal(r[k], whole[k]), (rank, k)
        assert esum == pytest.approx(whole["Etotal"].sum(), rel=1e-13)
        for k in L1:
            assert L[k] == pytest.approx(L1[k], rel=1e-12), (rank, k)     # the sharded sums equal the single-rank losses
