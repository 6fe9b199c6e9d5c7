"""Host-side pieces of bench.py that decide what the numbers mean (no GPU): the same-config CPU sample of the reference arm,
the FLOP count behind `roofline`, the workloads of the BASELINE.json configs."""
import argparse
import os
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    d = dict(config="c4", nx=20, nmol=10000)
    d.update(kw)
    return argparse.Namespace(**d)


def test_periodic_sample_gives_every_centre_its_full_environment():
    """bench.periodic_sample (the bounded sample both the cpu_baseline leg and `--impl reference` time): the centres of the
    sub-volume see exactly the neighbours, by element and distance, that they have in the fully tessellated cell."""
    from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
    Z, X, lat = water_box(6, spacing=3.1072, seed=3, jitter=0.05)      # 648 atoms, L = 18.6 A
    X = wrap_into_cell(X, lat)
    rc = 7.5
    Xs, Zs, M, nenv = bench.periodic_sample(Z, X, lat, 8.0, rc)
    assert 0 < M < len(Z) and len(Xs) == len(Zs) and (len(Zs) - M) % M == 0
    # the full environment: 27 images (rc < L)
    img = np.concatenate([X + i * lat[0] + j * lat[1] + k * lat[2] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)])
    zimg = np.tile(Z, 27)
    # centres are the first M rows of the sample and are atoms of the cell
    for c in range(M):
        d_full = np.linalg.norm(img - Xs[c], axis=1)
        d_samp = np.linalg.norm(Xs - Xs[c], axis=1)
        a = sorted((int(z), round(float(d), 9)) for z, d in zip(zimg[(d_full < rc) & (d_full > 1e-9)], d_full[(d_full < rc) & (d_full > 1e-9)]))
        b = sorted((int(z), round(float(d), 9)) for z, d in zip(Zs[(d_samp < rc) & (d_samp > 1e-9)], d_samp[(d_samp < rc) & (d_samp > 1e-9)]))
        assert a == b, c
    # the padding atoms are far from everything
    assert np.all(np.linalg.norm(Xs[M + nenv:] - Xs[0], axis=1) > 1e3) or len(Xs) == M + nenv


def test_mlp_flop_count_is_the_layer_sum():
    D, hidden = 256, [500, 500, 500]
    macs = D * 500 + 500 * 500 + 500 * 500 + 500              # one net, forward
    assert bench.mlp_flops_per_atom(D, hidden) == 8 * macs     # x2 nets, x2 flop per MAC, x2 backward-data
    assert abs(bench.mlp_flops_per_atom(D, hidden) * 24000 / 1e12 - 0.1207) < 1e-3


def test_workloads_are_the_baseline_configs():
    import json
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "north_star" in base
    w = bench.build_workload(_args(config="c4"))
    assert w["kind"] == "lattice" and w["natom"] == 24000 and w["hidden"] == [500, 500, 500] and w["config"]["cutoffs_A"] == [4.6, 3.1, 15.0]
    f = w["X"] @ np.linalg.inv(w["lat"])
    assert f.min() >= 0.0 and f.max() < 1.0                    # wrapped, as tm_eval_lattice expects
    w3 = bench.build_workload(_args(config="c3"))
    assert w3["natom"] == 3000
    w5 = bench.build_workload(_args(config="c5"))
    assert w5["natom"] == 1568 and w5["eles"] == [1, 6, 7, 8] and w5["hidden"] == [2000] * 3
    w2 = bench.build_workload(_args(config="c2", nmol=16))
    assert w2["kind"] == "batch" and w2["Zs"].shape == (16, 40) and w2["xyzs"].shape == (16, 40, 3) and w2["natom"] == 640
    # both arms print the same `config` object
    assert bench.build_workload(_args(config="c4"))["config"] == w["config"]


def _json_lines(out):
    import json
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line_alone_and_under_torchrun():
    """`bench.py --impl reference`: the CPU arm (oracle port + the reference's compiled neighbour search when oracle/_ref is
    built) on the bench's own config; under torchrun only rank 0 works and prints, the other ranks exit 0."""
    import subprocess
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    base = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c3", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(base, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "atom-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"] == bench.build_workload(_args(config="c3"))["config"]          # same config object as the B200 arm
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    tr = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577"]
    r2 = subprocess.run(tr + base[1:] + ["--gpus", "2"], capture_output=True, text=True, timeout=400, env=env, cwd=ROOT)
    assert r2.returncode == 0, r2.stderr[-2000:]
    lines2 = _json_lines(r2.stdout)
    assert len(lines2) == 1 and lines2[0]["impl"] == "reference" and lines2[0]["n_gpus"] == 2
