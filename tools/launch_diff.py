#!/usr/bin/env python
"""Per-launch comparison of two launch lists of tools/one_step.py (raw ncu csv of tools/launch_list.sh, or the *_step.csv
kept under profiles/): prints the conv launches side by side and the totals.

    python tools/launch_diff.py profiles/r2d_launches_step.csv gpurun_out/r2e_launches.csv
"""
import csv
import sys


def short(name):
    return name.split("(")[0].replace("void ", "").replace("mvs::", "")


def load(path):
    lines = [l for l in open(path)]
    if lines and lines[0].startswith("#,kernel"):
        return [(short(r[1]), r[2], float(r[4])) for r in csv.reader(lines[1:])]
    rows = [r for r in csv.reader(l for l in lines if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, g, v = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    launches = [(short(r[k]), r[g], float(r[v]) / 1e3) for r in rows]
    starts = [i for i, l in enumerate(launches) if "to_c8h_kernel" in l[0]]
    return launches[starts[-1]:]                      # the last whole-model step (earlier ones are warm-up)


def main():
    a, b = load(sys.argv[1]), load(sys.argv[2])
    ca, cb = [x for x in a if "conv3d_umma" in x[0]], [x for x in b if "conv3d_umma" in x[0]]
    for x, y in zip(ca, cb):
        print(f"{x[0]:26s} {x[1]:16s} {x[2]:8.1f} -> {y[2]:8.1f} us")
    print(f"conv launches: {sum(x[2] for x in ca):.0f} -> {sum(x[2] for x in cb):.0f} us;  "
          f"step: {len(a)} launches {sum(x[2] for x in a):.0f} us -> {len(b)} launches {sum(x[2] for x in b):.0f} us")
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write("#,kernel,grid,block,duration_us\n")
            for i, (n, g, us) in enumerate(b):
                f.write(f'{i},"{n}","{g}","",{us:.2f}\n')


if __name__ == "__main__":
    main()
