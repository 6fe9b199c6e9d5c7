"""Config C2 (BASELINE.json): 10,000 random geometries of a morphine-size C,H,N,O molecule, batched energy+force through
tm_eval (host buffers).  Prints molecules/s and checks a size-independent property: a molecule evaluated inside the
batch equals the same molecule evaluated alone.

    python scripts/c2_batch.py [nmol]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/c2_batch.py [nmol]

Under torchrun the molecules are sharded over the ranks (tensormol_b200.parallel.BatchShardEvaluator: independent units, no
data-path collective, one final gather); the time is the max over ranks of the wall time of a whole sharded call incl. the gather."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import load_golden
from tensormol_b200.engine import default_params
from tensormol_b200.SystemBuilders import perturbed_molecule_batch
from tensormol_b200.engine import Engine, random_weights
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
g = load_golden("morphine")
P = default_params()
hidden = [500, 500, 500]
eng = Engine(list(g["eles"]), hidden, P, device=local)
eng.set_weights(random_weights(eng.eles, eng.D, hidden, 0))
Zs, xyzs = perturbed_molecule_batch(g["Z"], g["xyz"], nmol, sigma=0.05, seed=1)
nat = np.full(nmol, Zs.shape[1], np.int64)
if world > 1:
    import torch
    import torch.distributed as dist
    from tensormol_b200.parallel import BatchShardEvaluator
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    ev = BatchShardEvaluator(eng, rank, world, dist, "cuda:%d" % local)
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); r = ev.evaluate(xyzs, Zs, nat); torch.cuda.synchronize(); t1 = time.perf_counter()
    dt = torch.tensor([t1 - t0], device="cuda"); dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt[0])
    if rank == 0:
        print(f"{nmol} molecules x {Zs.shape[1]} atoms sharded over {world} GPUs: {1e3 * dt:.2f} ms per call incl. the gather = "
              f"{nmol / dt:.0f} molecules/s = {nmol * Zs.shape[1] / dt / 1e6:.2f} M atom-steps/s")
else:
    for it in range(3):
        t0 = time.perf_counter(); r = eng.evaluate(xyzs, Zs, nat); t1 = time.perf_counter()
    t = eng.timings()
    print(f"{nmol} molecules x {Zs.shape[1]} atoms (D = {eng.D}): {1e3 * (t1 - t0):.2f} ms per batch call = {nmol / (t1 - t0):.0f} molecules/s = {nmol * Zs.shape[1] / (t1 - t0) / 1e6:.2f} M atom-steps/s;"
          f" device {t['total']:.2f} ms", {k: round(v, 3) for k, v in t.items() if isinstance(v, float)})
if rank == 0:
    for m in (0, 17, nmol - 1):
        r1 = eng.evaluate(xyzs[m:m + 1], Zs[m:m + 1], nat[m:m + 1])
        de = abs(r1["Etotal"][0] - r["Etotal"][m]) / abs(r["Etotal"][m])
        dg = np.abs(r1["gradient"][0] - r["gradient"][m]).max()
        print(f"molecule {m}: batch vs alone rel dE {de:.2e} max|dgrad| {dg:.2e}")
        assert de < 2e-6 and dg < 1e-6
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
