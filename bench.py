#!/usr/bin/env python
"""Benchmark of the BP+EE energy+force step (BASELINE.json metric: atom-steps/s on the 24,000-atom
periodic water box; the other configs of BASELINE.json with --config).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c4|c3|c2|c5]

N>1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
  value      whole-job atom-steps/s with the inputs resident in HBM (tm_eval_lattice_dev / slab phases / tm_eval shards)
  e2e        the same through the host-buffer C-ABI call a reference user makes (H2D + D2H inside the timed region),
             returning what the reference's call returns (periodic: Etotal + force, TFMolManage.py:1353-1358)
  roofline   dominant kernel group (per-element MLP GEMMs) against the measured burst tensor peak
  cpu_baseline / --impl reference: the float64 oracle port of the reference graph (oracle/) with the reference's own
             MolEmb.Make_NListNaive (oracle/_ref, compiled from C_API/MolEmb.cpp) for the neighbour search, timed on
             the host cores on a bounded sample OF THE SAME WORKLOAD: the centres of a sub-volume of the bench system in
             their full environment (TensorFlow is not installable here and the reference package does not import under
             Python 3.12).  Both arms print the same `config`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def hot_params():
    return dict(AN1_r_Rc=4.6, AN1_a_Rc=3.1, AN1_eta=4.0, AN1_zeta=8.0, AN1_num_r_Rs=32, AN1_num_a_Rs=8, AN1_num_a_As=8,
                EECutoffOn=0.0, EECutoffOff=15.0, Elu_Width=4.6, Poly_Width=4.6, DSFAlpha=0.18, AddEcc=True,
                sigmoid_alpha=100.0, NeuronType="sigmoid_with_param")


def mlp_flops_per_atom(D, hidden):
    """fwd MACs of one net; x2 nets, x2 flop/MAC, x2 for the backward-data pass."""
    mac, fan = 0, D
    for h in hidden:
        mac += fan * h
        fan = h
    mac += fan
    return 2 * 2 * 2 * mac


# ------------------------------------------------------------------------------------------------ workloads
def build_workload(args):
    """Geometry + model of a BASELINE.json config.  Returns a dict; `config` is what both arms print."""
    from tensormol_b200.SystemBuilders import perturbed_molecule_batch, water_box, wrap_into_cell
    cfg = args.config
    if cfg in ("c4", "c3"):
        nx = args.nx if cfg == "c4" else 10
        Z, X, lat = water_box(nx, spacing=3.1072, seed=3, jitter=0.05)
        X = wrap_into_cell(X, lat)
        hidden = [500, 500, 500]           # water nets of the reference (samples/test_tensormol01.py:14)
        name = "C4" if cfg == "c4" else "C3"
        what = "single-point energy+force" if cfg == "c4" else "energy+force per MD step (neighbour rebuild every step)"
        return dict(kind="lattice", eles=[1, 8], hidden=hidden, Z=Z, X=X, lat=lat, ntess=1, natom=len(Z),
                    config={"workload": f"{len(Z)}-atom periodic water box ({name}: {nx}^3 waters, L={lat[0, 0]:.3f} A, 27 images; SURVEY's 192-atom-cell x 5^3 recipe replaced by "
                                        f"a jittered simple-cubic box of the same density), BP+EE {what}, nets {hidden}, random-init weights seed 0",
                            "natom": int(len(Z)), "nets": hidden, "cutoffs_A": [4.6, 3.1, 15.0]})
    if cfg == "c5":
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "evq2_periodic.npz")))
        hidden = [2000, 2000, 2000]
        Z, X, lat = g["Z"].astype(np.int32), g["xyz"], g["lattice"]
        return dict(kind="lattice", eles=[1, 6, 7, 8], hidden=hidden, Z=Z, X=wrap_into_cell(X, lat), lat=lat, ntess=int(g["ntess"]), natom=len(Z),
                    config={"workload": f"2evq peptide + explicit water, {len(Z)} atoms C/H/N/O (C5), bounding-box cell, BP+EE energy+force per step, nets {hidden}, random-init weights seed 0",
                            "natom": int(len(Z)), "nets": hidden, "cutoffs_A": [4.6, 3.1, 15.0]})
    if cfg == "c2":
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "morphine.npz")))
        hidden = [2000, 2000, 2000]        # C/H/N/O nets of the reference (samples/test_tensormol01.py:32)
        nmol = args.nmol
        Zs, xyzs = perturbed_molecule_batch(g["Z"], g["xyz"], nmol, sigma=0.05, seed=1)
        return dict(kind="batch", eles=[1, 6, 7, 8], hidden=hidden, Zs=Zs, xyzs=xyzs, natom=int(Zs.size), nmol=nmol,
                    config={"workload": f"{nmol} random geometries (N(0, 0.05 A)) of morphine, {Zs.shape[1]} atoms C/H/N/O (C2), batched BP+EE energy+force, nets {hidden}, random-init weights seed 0",
                            "natom": int(Zs.size), "nets": hidden, "cutoffs_A": [4.6, 3.1, 15.0]})
    raise SystemExit(f"unknown config {cfg}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    """Burst tensor peak (the GEMM group is timed by itself, kernel by kernel, at full clocks) and the HBM copy peak."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json: bf16 burst, HBM copy)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_counters():
    """Per-launch counters of the hot kernels from the committed `ncu --set full` capture of this command (the run itself
    cannot read them): profiles/r2_counters.json, written by scripts/ncu_counters.py.  Absent -> None (reported as null)."""
    p = os.path.join(ROOT, "profiles", "r2_counters.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------ CPU arms
def _ref_molemb():
    so = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "MolEmb*.so")))
    if not so:
        return None
    try:
        spec = importlib.util.spec_from_file_location("MolEmb", so[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


def periodic_sample(Z, X, lat, edge, rc):
    """Centres = the atoms whose fractional coordinates are all below edge / |a_i| (a sub-volume at the origin corner of
    the cell), followed by every atom or image within rc of that sub-volume's bounding box, padded with far-away atoms to
    a multiple of the centre count (the oracle tiles the charges over blocks of nreal, TFMolInstanceDirect.py:5892).
    Every centre sees exactly the environment it has in the full system."""
    L = np.asarray(lat, np.float64)
    f = X @ np.linalg.inv(L)
    fe = edge / np.linalg.norm(L, axis=1)
    cen = np.all(f < fe[None, :], axis=1)
    Xc, Zc = X[cen], Z[cen]
    lo, hi = Xc.min(0) - rc - 0.1, Xc.max(0) + rc + 0.1
    envX, envZ = [], []
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            for k in (-1, 0, 1):
                Y = X + i * L[0] + j * L[1] + k * L[2]
                m = np.all((Y > lo) & (Y < hi), axis=1)
                if i == 0 and j == 0 and k == 0:
                    m &= ~cen
                envX.append(Y[m])
                envZ.append(Z[m])
    envX, envZ = np.concatenate(envX), np.concatenate(envZ)
    M = len(Zc)
    pad = (-len(envZ)) % M
    if pad:
        far = np.stack([1.0e4 + 40.0 * np.arange(pad), np.full(pad, 1.0e4), np.full(pad, 1.0e4)], axis=1)
        envX = np.concatenate([envX, far])
        envZ = np.concatenate([envZ, np.full(pad, Z[0], Z.dtype)])
    return np.concatenate([Xc, envX]), np.concatenate([Zc, envZ]).astype(np.uint8), M, len(envZ) - pad


def cpu_sample_run(wl, reps, threads=None, budget_s=60.0):
    """Oracle port of the reference graph on the bounded sample; returns (centres per step, seconds per rep, cores, description)."""
    import torch
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from tensormol_b200.engine import descriptor_width, random_weights
    if threads:
        torch.set_num_threads(threads)
    P = hot_params()
    W = random_weights(wl["eles"], descriptor_width(len(wl["eles"]), P), wl["hidden"], 0)
    orc = og.Oracle(wl["eles"], W, P)
    molemb = _ref_molemb()
    nl_kind = "oracle_np candidate search (scipy cKDTree) + the reference's accept test"
    saved = onp.nlist_csr
    if molemb is not None:
        def nlist_csr_ref(x, rng, nreal, do_perms):   # the reference's own sweep, consumed as its Python callers do
            rows = molemb.Make_NListNaive(np.ascontiguousarray(x, np.float64), float(rng), int(nreal), int(do_perms))
            cnt = np.fromiter((len(r) for r in rows), np.int64, len(rows))
            off = np.zeros(len(rows) + 1, np.int64)
            np.cumsum(cnt, out=off[1:])
            idx = np.fromiter((j for r in rows for j in sorted(r)), np.int64, int(off[-1]))
            return off, idx
        onp.nlist_csr = nlist_csr_ref
        nl_kind = "the reference's own MolEmb.Make_NListNaive (oracle/_ref, compiled from C_API/MolEmb.cpp)"
    try:
        if wl["kind"] == "lattice":
            Xs, Zs, M, nenv = periodic_sample(wl["Z"], wl["X"], wl["lat"], args_sample_edge(wl), P["EECutoffOff"])
            run = lambda: orc.evaluate_periodic(Xs, Zs, M)
            desc = (f"{M} centres of the bench system (a {args_sample_edge(wl):.1f} A sub-volume) with their complete environment ({nenv} atoms and images within 15 A), "
                    f"float64 oracle port: neighbour search by {nl_kind} + index assembly (numpy) + graph + autograd forces (torch CPU)")
            units = M
        else:
            k = min(16, wl["nmol"])
            xs, zs = wl["xyzs"][:k], wl["Zs"][:k]
            nat = np.full(k, zs.shape[1], np.int64)
            run = lambda: orc.evaluate(xs, zs, nat)
            desc = f"{k} molecules of the batch ({zs.shape[1]} atoms each), float64 oracle port: neighbour search by {nl_kind} + index assembly + graph + autograd forces (torch CPU)"
            units = int(zs.size)
        t0 = time.perf_counter()
        run()                                        # warm-up
        warm = time.perf_counter() - t0
        reps = max(1, min(reps, int(budget_s / max(warm, 1e-3))))
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
    finally:
        onp.nlist_csr = saved
    return units, ts, torch.get_num_threads(), desc


def args_sample_edge(wl):
    """Edge of the sampled sub-volume: ~650 centres for liquid water, everything for the 1,568-atom protein box capped at ~400."""
    if wl["natom"] >= 3000:
        return 18.7
    return 0.62 * float(np.linalg.norm(np.asarray(wl["lat"], np.float64), axis=1).min())


def run_reference(args, rank):
    if rank != 0:
        return
    wl = build_workload(args)
    units, ts, cores, desc = cpu_sample_run(wl, max(args.steps, 1), os.cpu_count(), budget_s=90.0)
    total = sum(ts)
    val = units * len(ts) / total
    line = {"impl": "reference", "metric": "atom-steps/s (energy+force)", "value": val, "unit": "atom-steps/s", "n_gpus": args.gpus,
            "steps": len(ts), "warmup": 1, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": wl["config"],
            "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference arm = float64 CPU restatement of the TF graph with the reference's compiled MolEmb neighbour search (TensorFlow absent offline); "
                    "published by the reference: ~84 atom-steps/s periodic @12k atoms, 273 aperiodic @24k (BASELINE.md)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args, rank, world, local_rank):
    import torch
    from tensormol_b200.engine import Engine, GraphedCall, random_weights
    from tensormol_b200.parallel import BatchShardEvaluator, EngineSlabBackend, SlabEvaluator, batch_shard_bounds
    dist = None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    P = hot_params()
    wl = build_workload(args)
    natom = wl["natom"]
    hidden = wl["hidden"]
    eng = Engine(wl["eles"], hidden, P, device=local_rank)
    eng.set_weights(random_weights(wl["eles"], eng.D, hidden, 0))
    eng.set_gemm_mode(args.gemm_mode)
    # a non-default torch stream shared with the library, so torch.cuda.Event brackets the library's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(C.c_void_p(stream.cuda_stream))
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    extra = {}
    lattice = wl["kind"] == "lattice"
    slab = None
    p2p = False
    if lattice:
        Z, X, lat, ntess = wl["Z"], wl["X"], wl["lat"], wl["ntess"]
        xyz_t = torch.tensor(X, dtype=torch.float64, device=dev)
        Z_t = torch.tensor(Z, dtype=torch.int32, device=dev)
        e_t = torch.zeros(6, dtype=torch.float64, device=dev)
        g_t = torch.zeros(natom, 3, dtype=torch.float64, device=dev)
        slab = SlabEvaluator(EngineSlabBackend(eng), natom, rank, world, dev, dist) if world > 1 else None
        p2p = bool(slab is not None and args.p2p and slab.enable_p2p())

        def step_resident():
            if slab is None:
                eng.evaluate_lattice_dev(C.c_void_p(xyz_t.data_ptr()), C.c_void_p(Z_t.data_ptr()), natom, lat, ntess,
                                         C.c_void_p(e_t.data_ptr()), C.c_void_p(g_t.data_ptr()))
            else:
                slab.step(xyz_t, Z_t, lat, ntess)
        units_per_step = natom
    else:
        # molecule batch: independent units, sharded over the ranks with no data-path collective.
        # tm_eval takes host buffers, so "resident" here is the device time of the rank's call (CUDA events inside the
        # library: copy-in to copy-out excluded) and e2e the wall time of the same call.
        Zs, xyzs = wl["Zs"], wl["xyzs"]
        nat = np.full(Zs.shape[0], Zs.shape[1], np.int64)
        b = batch_shard_bounds(nat, world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        sub = max(1, args.sub_batch)

        def step_resident():
            for s0 in range(lo, hi, sub):
                s1 = min(hi, s0 + sub)
                eng.evaluate(xyzs[s0:s1], Zs[s0:s1], nat[s0:s1])
        units_per_step = natom

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    # ---- stage timings and the launch count: one kernel-by-kernel pass (events inside a graph cannot be read back) ----
    stage = {"mlp": 0.0, "desc": 0.0, "force": 0.0, "pair": 0.0, "nlist": 0.0}
    launches = 0
    nstage = 5
    if lattice and slab is None:
        for _ in range(nstage):
            flush.zero_()
            step_resident()
            t = eng.timings()
            stage["mlp"] += t["mlp_fwd"] + t["mlp_bwd"]
            stage["desc"] += t["desc"]
            stage["force"] += t["force"]
            stage["pair"] += t["pair"]
            stage["nlist"] += t["nlist"]
            launches = t["launches"]
    elif lattice:
        # per-rank phase times of an eager step (phase B / C begin with the wait for the peers: load imbalance shows here)
        ph = torch.zeros(nstage, 6, dtype=torch.float64)
        for it in range(nstage):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            barrier()
            flush.zero_()
            b_ = slab.backend
            evs[0].record()
            b_.slab_phase_a(xyz_t, Z_t, natom, lat, ntess, rank, world, slab.qraw)
            evs[1].record()
            if not p2p:
                slab._allreduce(slab.qraw)
            b_.slab_phase_b(slab.qraw, slab.e)
            evs[2].record()
            if not p2p:
                slab._allreduce(slab.e)
            b_.slab_phase_c(slab.e, 3, slab.grad)
            if not p2p:
                slab._allreduce(slab.grad)
            evs[3].record()
            torch.cuda.synchronize()
            t = eng.timings()
            ph[it] = torch.tensor([evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3]), t["nlist"], t["desc"], t["mlp_fwd"]])
            launches = t["launches"]
        phm = ph.median(0).values.to(dev)
        allph = [torch.zeros_like(phm) for _ in range(world)]
        dist.all_gather(allph, phm)
        extra["rank_phase_ms"] = {"columns": ["phase_a", "phase_b(+wait)", "phase_c(+wait)", "a:nlist", "a:desc", "a:mlp_fwd"],
                                  "rows": [[round(float(v), 4) for v in r.cpu()] for r in allph]}
    else:
        torch.cuda.synchronize()
        launches = eng.timings()["launches"] * max(1, -(-(hi - lo) // sub))

    step_timed = step_resident
    if args.graph and lattice:
        if slab is None:
            step_timed = GraphedCall(step_resident, stream, warmup=1)
        else:                              # one graph (peer-memory exchange) or three with eager NCCL all-reduces between
            slab.capture(xyz_t, Z_t, lat, ntess, stream)
            step_timed = slab.step_replay
        for _ in range(3):
            step_timed()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if lattice:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for a, b_ in evs:
            flush.zero_()                      # L2 flush between timed iterations (not timed)
            a.record()
            step_timed()
            b_.record()
        barrier()
        eng.sync()                             # raises if a device flag was set (capacity, unwrapped input, exchange time-out)
        ms = sum(a.elapsed_time(b_) for a, b_ in evs)
    else:
        nrep = max(1, min(args.steps, 20))
        ms = 0.0
        barrier()
        for _ in range(nrep):
            flush.zero_()
            for s0 in range(lo, hi, sub):
                s1 = min(hi, s0 + sub)
                eng.evaluate(xyzs[s0:s1], Zs[s0:s1], nat[s0:s1])
                t = eng.timings()
                ms += t["total"] - t["h2d"] - t["d2h"]
                for k, kk in (("mlp", ("mlp_fwd", "mlp_bwd")), ("desc", ("desc",)), ("force", ("force",)), ("pair", ("pair",)), ("nlist", ("nlist",))):
                    stage[k] += sum(t[x] for x in kk) / nrep
        barrier()
        args.steps = nrep
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = units_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer API -------------------------------------------------
    if lattice and slab is None:
        h2d = natom * 3 * 8 + natom * 4
        d2h = 7 * 8 + natom * 3 * 8 + 8
        want = ("Etotal", "gradient")          # what EvalBPDirectEEUpdateSinglePeriodic returns (TFMolManage.py:1353-1358)
        # (i) the caller's arrays in ordinary pageable memory: the library stages them through its own pinned buffer
        for _ in range(4):
            eng.evaluate_lattice(X, Z, lat, ntess, outputs=want)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = eng.evaluate_lattice(X, Z, lat, ntess, outputs=want)
        torch.cuda.synchronize()
        e2e_pageable_s = time.perf_counter() - t0
        # (ii) the headline e2e: coordinates, atomic numbers and the gradient in page-locked host arrays (Engine.pinned),
        # which the copy engine reads and writes directly; same call, same bytes over the bus every step
        Xp, Zp = eng.pinned(X.shape), eng.pinned(Z.shape, np.int32)
        Xp[:] = X
        Zp[:] = Z
        into = {"gradient": eng.pinned((1, natom, 3))}
        call = eng.bind_lattice(Xp, Zp, lat, ntess, outputs=want, into=into)   # = evaluate_lattice with the arguments bound once
        for _ in range(4):
            call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = call()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e_tot = float(r["Etotal"][0])
    elif lattice:
        xh = torch.tensor(X, dtype=torch.float64).pin_memory()
        gh = torch.zeros(natom, 3, dtype=torch.float64).pin_memory()
        h2d = natom * 3 * 8
        d2h = natom * 3 * 8 + 48
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            xyz_t.copy_(xh, non_blocking=True)
            step_timed()
            gh.copy_(slab.grad, non_blocking=True)
            e_tot = float(slab.e[0].item())
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        # force parity on the real ranks: the slab result against ONE rank evaluating the whole box
        if rank == 0:
            eng1 = Engine(wl["eles"], hidden, P, device=local_rank)
            eng1.set_weights(random_weights(wl["eles"], eng.D, hidden, 0))
            eng1.set_gemm_mode(args.gemm_mode)
            r1 = eng1.evaluate_lattice(X, Z, lat, ntess, outputs=("Etotal", "gradient"))
            gN = slab.grad.cpu().numpy()
            extra["vs_single_rank"] = {"max_abs_dgrad_Ha_per_A": float(np.abs(gN - r1["gradient"][0]).max()),
                                       "max_abs_grad_Ha_per_A": float(np.abs(r1["gradient"][0]).max()),
                                       "rel_dE": float(abs(e_tot - r1["Etotal"][0]) / abs(r1["Etotal"][0]))}
            eng1.close()
    else:
        ev = BatchShardEvaluator(eng, rank, world, dist, dev)
        nm, maxn = Zs.shape
        h2d = (hi - lo) * maxn * 28 + (hi - lo) * 8
        d2h = (hi - lo) * (7 + 5 * maxn) * 8
        nrep = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(nrep):
            for s0 in range(lo, hi, sub):
                s1 = min(hi, s0 + sub)
                r = eng.evaluate(xyzs[s0:s1], Zs[s0:s1], nat[s0:s1])
        barrier()
        e2e_s = (time.perf_counter() - t0) * args.steps / nrep
        e_tot = float(r["Etotal"][0])
        if dist is not None:
            te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_s = float(te.item())
    e2e_val = units_per_step * args.steps / e2e_s

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    tc_peak, hbm_peak, which = measured_peak()
    flops = mlp_flops_per_atom(eng.D, hidden) * natom
    if lattice:
        par = f"slab{world}" if world > 1 else "single"
        exch = ("peer-memory stores + device flags" if p2p else "3 NCCL all-reduces" + (f" (peer memory unavailable: {getattr(slab, 'p2p_error', '')[:120]})" if args.p2p and not p2p else "")) if world > 1 else None
        launch = ("CUDA graph replay of the step" if (world == 1 or p2p) else "one CUDA graph per phase, eager NCCL all-reduces between") if args.graph else "kernel by kernel"
    else:
        par = f"batch-sharded over {world} ranks, no data-path collective" if world > 1 else "single"
        exch = None
        launch = f"tm_eval per sub-batch of {sub} molecules; value = device time inside the calls (copies excluded), e2e = wall time"
    line = {"metric": "atom-steps/s (energy+force)", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,   # the workload is the whole job at every N (BASELINE.json metric)
            "dtype": {0: "f32"}.get(args.gemm_mode, "f16x2-split (fp32 accumulate)"), "data": "synthetic",
            "config": wl["config"],
            "measurement": {"l2": "flushed between timed iterations (512 MiB write)", "parallelism": par, "exchange": exch, "gemm_mode": args.gemm_mode, "launch": launch,
                            "stage_ms_note": "stage_ms / rank_phase_ms come from a kernel-by-kernel pass with events between the stages; value from the graph replay, so the stages sum to more than ms_per_step"},
            "e2e": {"value": e2e_val, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches) * args.steps if launches else None,
            "clocks": sampler.summary(), "Etotal": e_tot}
    line.update(extra)
    if lattice and slab is None:
        line["e2e"]["host_memory"] = "page-locked caller arrays (Engine.pinned) copied to and from the device directly; tm_eval_lattice through Engine.bind_lattice"
        line["e2e"]["pageable_value"] = units_per_step * args.steps / e2e_pageable_s   # ordinary numpy arrays, staged by the library
    if world == 1:
        mlp_ms = stage["mlp"] / nstage if lattice else stage["mlp"]
        df_ms = (stage["desc"] + stage["force"]) / nstage if lattice else stage["desc"] + stage["force"]
        ach = flops / (mlp_ms * 1e-3) / 1e12
        ctr = ncu_counters() if (args.config == "c4" and args.nx == 20 and args.gemm_mode == 1) else None
        gem = (ctr or {}).get("k_gemm_tc")
        overlapped = lattice and natom <= 8192 and not os.environ.get("TM_NO_OVERLAP")
        if overlapped:
            # small cells: the library runs the backward nets on a side stream concurrently with the pair kernel, so the
            # eager pass's "mlp" is the forward nets plus whatever of the backward pass was still exposed, and "pair"
            # carries the contention: no separable GEMM time, hence no GEMM roofline for this configuration
            line["measurement"]["stage_ms_note"] += ("; this cell is small enough (<= 8192 centres) that the backward nets overlap the pair kernel on a side stream: "
                                                     "mlp = forward + exposed rest, pair includes the contention; roofline.frac is not defined here")
            ach = None
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": (ach / tc_peak) if ach else None,
                            "traffic": gem["dram_bytes_per_step"] if gem else None,
                            "traffic_source": (ctr or {}).get("source") if gem else None,
                            "kernel": "grouped per-element MLP GEMMs (fwd + bwd-data, both nets); split precision issues 3 MMAs per product, so the tensor-pipe ceiling is 1/3 in these units",
                            "peak_source": which, "ms_per_step": mlp_ms, "algorithmic_flops_per_step": flops}
        if lattice:
            n_r, n_a = (40.3, 10.4)
            bytes_df = (44 + 8 * (n_r + n_a) + 8 * eng.D) * natom
            line["descriptor_roofline"] = {"bound": "hbm", "achieved": bytes_df / (df_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                           "frac": bytes_df / (df_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_step": df_ms,
                                           "kernels": "k_desc + k_force", "algorithmic_bytes_per_step": bytes_df}
            df = [(ctr or {}).get(k) for k in ("k_desc_fast", "k_force_fast")]
            if all(df):
                # issue bound: warp instructions of the two kernels (ncu sm__inst_executed.sum) at 4 per clock and SM
                sm_mhz = (line["clocks"].get("sm_mhz") or 1965.0)
                inst = sum(d["warp_inst_per_launch"] for d in df)
                t_issue = inst / (148 * 4 * sm_mhz * 1e6) * 1e3
                line["issue_roofline"] = {"bound": "issue", "warp_inst_per_step": inst, "ms_at_4_ipc_per_sm": t_issue, "ms_per_step": df_ms,
                                          "frac": t_issue / df_ms, "kernels": "k_desc + k_force", "source": ctr.get("source")}
        line["stage_ms"] = {k: (v / nstage if lattice else v) for k, v in stage.items()}
        try:
            n, ts, cores, desc = cpu_sample_run(wl, 3, os.cpu_count(), budget_s=25.0)
            line["cpu_baseline"] = {"value": n * len(ts) / sum(ts), "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as ex:   # the oracle is optional at bench time; never fail the GPU line for it
            line["cpu_baseline"] = {"value": None, "unit": "atom-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=["c4", "c3", "c2", "c5"],
                    help="BASELINE.json config: c4 = 24,000-atom water box (default, the metric's), c3 = 3,000-atom box, c2 = molecule batch (nets 2000^3), c5 = 2evq box")
    ap.add_argument("--nx", type=int, default=20, help="c4: waters per box edge (20 -> 24,000 atoms)")
    ap.add_argument("--nmol", type=int, default=10000, help="c2: molecules in the batch")
    ap.add_argument("--sub-batch", type=int, default=2500, help="c2: molecules per tm_eval call")
    ap.add_argument("--gemm-mode", type=int, default=1, help="0 = fp32 FFMA, 1 = tcgen05 split-fp16 (default)")
    ap.add_argument("--p2p", type=int, default=1, help="N > 1: 1 = exchange between the slab phases by peer-memory stores over NVLink (default), 0 = NCCL all-reduces")
    ap.add_argument("--graph", type=int, default=1, help="1 = the resident step is replayed from a CUDA graph (default), 0 = launched kernel by kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
