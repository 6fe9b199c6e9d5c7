"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, mean ns, share of one step.

    python scripts/launch_table.py gpurun_out/launches.csv [first_kernel_substring]

The step is taken as the launches between two consecutive occurrences of the first kernel of a step."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    kn, mv, gs, bs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    out = []
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        try:
            out.append((r[kn].split("(")[0].replace("void ", ""), float(r[mv].replace(",", "")), r[gs], r[bs]))
        except ValueError:
            pass
    return out


def main():
    L = load(sys.argv[1])
    first = sys.argv[2] if len(sys.argv) > 2 else L[0][0]
    idx = [i for i, l in enumerate(L) if first in l[0]]
    if len(idx) >= 3:
        a, b = idx[-2], idx[-1]
    else:
        a, b = 0, len(L)
    step = L[a:b]
    tot = sum(l[1] for l in step)
    print(f"step: {len(step)} launches, {tot / 1000:.1f} us serialised")
    for name, ns, g, bsz in step:
        print(f"{name[:44]:44s} {ns / 1000:8.2f} us {100 * ns / tot:5.1f} %  grid {g} block {bsz}")


if __name__ == "__main__":
    main()
