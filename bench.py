#!/usr/bin/env python
"""Benchmark of the BP+EE energy+force step (BASELINE.json metric: atom-steps/s on the 24,000-atom
periodic water box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N>1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
  value      whole-job atom-steps/s with the positions resident in HBM (tm_eval_lattice_dev / slab phases)
  e2e        the same through the host-buffer C-ABI call (tm_eval_lattice; H2D + D2H inside the timed region)
  roofline   dominant kernel group (per-element MLP GEMMs) against the measured tensor peak
  cpu_baseline / --impl reference: the float64 oracle port of the reference graph (oracle/), timed on the
             host cores on a bounded sample of the same workload (TensorFlow is not installable here and
             the reference package does not import under Python 3.12).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN = [500, 500, 500]          # water nets of the reference (samples/test_tensormol01.py:14)
SAMPLE_NX = 6                     # CPU sample: 216 waters = 648 atoms, same density, 27 images


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
                "reasons": reasons}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json: bf16 sustained, HBM copy)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_sample_run(reps, threads=None):
    """Oracle port of the reference graph on the bounded sample; returns (atom-steps/s, cores, description)."""
    import torch
    from oracle import oracle_graph as og
    from oracle import oracle_np as onp
    from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
    from tensormol_b200.engine import descriptor_width, random_weights
    if threads:
        torch.set_num_threads(threads)
    P = hot_params()
    Z, X, lat = water_box(SAMPLE_NX)
    X = wrap_into_cell(X, lat)
    W = random_weights([1, 8], descriptor_width(2, P), HIDDEN, 0)
    orc = og.Oracle([1, 8], W, P)
    Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])
    orc.evaluate_periodic(Xt, Zt, len(Z))          # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        Zt, Xt = onp.tess_lattice(lat, Z.astype(np.uint8), X, P["EECutoffOff"])
        orc.evaluate_periodic(Xt, Zt, len(Z))
        ts.append(time.perf_counter() - t0)
    desc = f"{len(Z)}-atom periodic water box (same density, 27 images, nets {HIDDEN}), float64 oracle port: tessellation + neighbour tables + graph + autograd forces"
    return len(Z), ts, torch.get_num_threads(), desc


def run_reference(args, rank):
    if rank != 0:
        return
    natom, ts, cores, desc = cpu_sample_run(args.steps, os.cpu_count())
    total = sum(ts)
    val = natom * len(ts) / total
    line = {"impl": "reference", "metric": "atom-steps/s (energy+force)", "value": val, "unit": "atom-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "24k-atom periodic water box BP+EE energy+force (C4); each step = bounded sample: " + desc},
            "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = float64 CPU restatement of the TF graph + MolEmb-equivalent neighbour search (TensorFlow absent offline); published: ~84 atom-steps/s periodic @24k extrapolated, 273 aperiodic (BASELINE.md)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args, rank, world, local_rank):
    import torch
    from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
    from tensormol_b200.engine import Engine, random_weights
    from tensormol_b200.parallel import EngineSlabBackend, SlabEvaluator
    dist = None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    P = hot_params()
    Z, X, lat = water_box(args.nx, spacing=3.1072, seed=3, jitter=0.05)
    X = wrap_into_cell(X, lat)
    natom = len(Z)
    eng = Engine([1, 8], HIDDEN, P, device=local_rank)
    eng.set_weights(random_weights([1, 8], eng.D, HIDDEN, 0))
    eng.set_gemm_mode(args.gemm_mode)
    # a non-default torch stream shared with the library, so torch.cuda.Event brackets the library's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(C.c_void_p(stream.cuda_stream))
    xyz_t = torch.tensor(X, dtype=torch.float64, device=dev)
    Z_t = torch.tensor(Z, dtype=torch.int32, device=dev)
    e_t = torch.zeros(6, dtype=torch.float64, device=dev)
    g_t = torch.zeros(natom, 3, dtype=torch.float64, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    slab = SlabEvaluator(EngineSlabBackend(eng), natom, rank, world, dev, dist) if world > 1 else None
    p2p = bool(slab is not None and args.p2p and slab.enable_p2p())

    def step_resident():
        if slab is None:
            eng.evaluate_lattice_dev(C.c_void_p(xyz_t.data_ptr()), C.c_void_p(Z_t.data_ptr()), natom, lat, 1,
                                     C.c_void_p(e_t.data_ptr()), C.c_void_p(g_t.data_ptr()))
        else:
            slab.step(xyz_t, Z_t, lat, 1)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    # stage timings and the launch count come from one kernel-by-kernel pass (events inside a graph cannot be read back)
    stage = {"mlp": 0.0, "desc": 0.0, "force": 0.0, "pair": 0.0, "nlist": 0.0}
    launches = 0
    nstage = 5
    if slab is None:
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
    else:
        torch.cuda.synchronize()
        launches = eng.timings()["launches"]      # kernels of the three phases of this rank's last step
    step_timed = step_resident
    if args.graph:
        from tensormol_b200.engine import GraphedCall
        if slab is None:
            step_timed = GraphedCall(step_resident, stream, warmup=1)
        else:                              # three graphs, the NCCL all-reduces between them stay eager
            slab.capture(xyz_t, Z_t, lat, 1, stream)
            step_timed = slab.step_replay
        for _ in range(3):
            step_timed()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in evs:
        flush.zero_()                      # L2 flush between timed iterations (not timed)
        a.record()
        step_timed()
        b.record()
    barrier()
    eng.sync()                             # raises if a device flag was set (capacity, unwrapped input, exchange time-out)
    ms = sum(a.elapsed_time(b) for a, b in evs)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = natom * args.steps / (ms * 1e-3)

    # ---- end to end through the host-buffer API -------------------------------------------------
    h2d = natom * 3 * 8 + natom * 4 + 80
    d2h = (7 + 5 * natom) * 8
    if slab is None:
        eng.evaluate_lattice(X, Z, lat, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = eng.evaluate_lattice(X, Z, lat, 1)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e_tot = float(r["Etotal"][0])
    else:
        xh = torch.tensor(X, dtype=torch.float64).pin_memory()
        gh = torch.zeros(natom, 3, dtype=torch.float64).pin_memory()
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
    e2e_val = natom * args.steps / e2e_s

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    tc_peak, hbm_peak, which = measured_peak()
    flops = mlp_flops_per_atom(eng.D, HIDDEN) * natom
    line = {"metric": "atom-steps/s (energy+force)", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,   # the 24,000-atom box is the whole job at every N (BASELINE.json metric)
            "dtype": {0: "f32", 1: "f16x2-split (fp32 accumulate)"}[args.gemm_mode], "data": "synthetic",
            "config": {"workload": f"{natom}-atom periodic water box (C4: {args.nx}^3 waters, L={lat[0, 0]:.3f} A, 27 images), BP+EE single-point energy+force, nets {HIDDEN}, random-init weights seed 0",
                       "l2": "flushed between timed iterations (512 MiB write)", "parallelism": f"slab{world}" if world > 1 else "single", "exchange": ("peer-memory stores + device flags" if p2p else "3 NCCL all-reduces" + (f" (peer memory unavailable: {getattr(slab, 'p2p_error', '')[:120]})" if args.p2p and not p2p else "")) if world > 1 else None,
                       "gemm_mode": args.gemm_mode, "launch": ("CUDA graph replay of the step" if (world == 1 or p2p) else "one CUDA graph per phase, eager NCCL all-reduces between") if args.graph else "kernel by kernel"},
            "e2e": {"value": e2e_val, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches) * args.steps if launches else None,
            "clocks": sampler.summary(), "Etotal": e_tot}
    if slab is None:
        mlp_ms = stage["mlp"] / nstage
        df_ms = (stage["desc"] + stage["force"]) / nstage
        ach = flops / (mlp_ms * 1e-3) / 1e12
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of the six k_gemm_tc launches of one 24k-atom step
        # (profiles/r02_ncu_hot_kernels.csv, ncu --set full), only meaningful for that workload
        traffic = 1.04e9 if (args.nx == 20 and args.gemm_mode == 1) else None
        line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak, "traffic": traffic,
                            "kernel": "grouped per-element MLP GEMMs (fwd + bwd-data, both nets)", "peak_source": which,
                            "ms_per_step": mlp_ms, "algorithmic_flops_per_step": flops}
        bytes_df = (44 + 8 * (40.3 + 10.4) + 8 * eng.D) * natom
        line["descriptor_roofline"] = {"bound": "hbm", "achieved": bytes_df / (df_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                       "frac": bytes_df / (df_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_step": df_ms,
                                       "kernels": "k_desc + k_force", "algorithmic_bytes_per_step": bytes_df}
        line["stage_ms"] = {k: v / nstage for k, v in stage.items()}
        try:
            n, ts, cores, desc = cpu_sample_run(2, os.cpu_count())
            line["cpu_baseline"] = {"value": n * len(ts) / sum(ts), "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": desc}
        except Exception as ex:   # the oracle is optional at bench time; never fail the GPU line for it
            line["cpu_baseline"] = {"value": None, "unit": "atom-steps/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=20, help="waters per box edge (20 -> 24,000 atoms)")
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
