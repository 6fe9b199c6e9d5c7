"""Per-kernel counters of one bench step from an `ncu --set full` report -> profiles/r2_counters.json (read by bench.py).

    ncu --set full --clock-control none --import-source on -s <skip> -c <n> -o gpurun_out/step python bench.py --steps 2 --warmup 1 --graph 0
    python scripts/ncu_counters.py gpurun_out/step.ncu-rep profiles/r2_counters.json "<command that was profiled>"
"""
import csv, io, json, subprocess, sys

rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = {"dur_ns": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum", "inst": "smsp__inst_executed.sum",
        "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread"}
units = rows[1]


def val(r, key):
    i = col.get(want[key])
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    if key in ("dram_rd", "dram_wr"):
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    if key == "dur_ns":
        v *= {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(u, 1)
    return v


kn = col["Kernel Name"]
acc = {}
body = [r for r in rows[2:] if len(r) > kn]
PER_STEP = {"k_gemm_tc": 6}        # launches of one step (every other kernel: one); the capture window may straddle two steps,
for r in body:                     # so per-launch means are formed first and scaled to one step
    name = r[kn].split("(")[0].replace("void ", "").split("<")[0]
    a = acc.setdefault(name, dict(launches=0, dur_ns=0.0, dram=0.0, inst=0.0, issue=[], tensor=[], regs=None))
    a["launches"] += 1
    a["dur_ns"] += val(r, "dur_ns") or 0.0
    a["dram"] += (val(r, "dram_rd") or 0.0) + (val(r, "dram_wr") or 0.0)
    a["inst"] += val(r, "inst") or 0.0
    for k, kk in (("issue", "issue_pct"), ("tensor", "tensor_pct")):
        v = val(r, kk)
        if v is not None:
            a[k].append(v)
    a["regs"] = val(r, "regs")
res = {"source": f"{rep} ({cmd}); ncu --set full --clock-control none, one step, cold-cache serialised launches"}
for name, a in acc.items():
    per = PER_STEP.get(name, 1)
    f = per / a["launches"]
    res[name] = {"launches_per_step": per, "launches_captured": a["launches"], "ncu_us_per_step": f * a["dur_ns"] / 1e3, "dram_bytes_per_step": f * a["dram"],
                 "warp_inst_per_launch": a["inst"] / a["launches"], "warp_inst_per_step": f * a["inst"],
                 "issue_active_pct": sum(a["issue"]) / len(a["issue"]) if a["issue"] else None,
                 "tensor_pipe_active_pct": sum(a["tensor"]) / len(a["tensor"]) if a["tensor"] else None, "registers": a["regs"]}
json.dump(res, open(out, "w"), indent=1)
for k, v in res.items():
    if k != "source":
        print(f"{k:24s} x{v['launches_per_step']:2d} {v['ncu_us_per_step']:8.1f} us  dram {v['dram_bytes_per_step'] / 1e6:8.1f} MB  inst {v['warp_inst_per_step'] / 1e6:7.1f} M  issue {v['issue_active_pct']}  tensor {v['tensor_pipe_active_pct']}")
