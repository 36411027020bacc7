#!/usr/bin/env python
"""Turns the ncu launch list of `bench.py --steps 1 --warmup 3` (gpu__time_duration.sum, --csv) into the per-step
launch table and the per-kernel share summary kept under profiles/.

    python tools/launch_summary.py gpurun_out/x_launches.csv profiles/<tag>_launches_fast
The bench runs 3 warm-up steps + 1 timed step + 2 e2e warm-ups + 1 e2e step: the FIRST quarter-sized block of launches that
starts with the batched pose algebra is one step; we take the 4th step (the timed one).
"""
import collections
import csv
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, g, b, v = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Value")
    launches = [(r[k], r[g], r[b], float(r[v]) / 1e3) for r in rows]
    # a step starts at the first torch kernel after a softargmin launch (or at the beginning)
    ends = [i for i, l in enumerate(launches) if "softargmin_conf_kernel" in l[0]]
    per_step = 3                                     # three stages => three softargmin launches per step
    steps = [ends[i:i + per_step] for i in range(0, len(ends), per_step)]
    timed = 3                                        # steps 0..2 are warm-up
    first = steps[timed - 1][-1] + 1
    last = steps[timed][-1]
    step = launches[first:last + 1]
    with open(out + "_step.csv", "w") as f:
        f.write("#,kernel,grid,block,duration_us\n")
        for i, (n, gg, bb, us) in enumerate(step):
            f.write(f'{i},"{n[:110]}","{gg}","{bb}",{us:.2f}\n')
    agg = collections.OrderedDict()
    for n, _, _, us in step:
        key = n.split("(")[0][-78:]
        a = agg.setdefault(key, [0.0, 0])
        a[0] += us; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    with open(out + "_summary.md", "w") as f:
        f.write(f"{len(step)} launches, {tot / 1e3:.3f} ms summed (cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write("| ms | share | launches | kernel |\n|---|---|---|---|\n")
        for key, (us, n) in sorted(agg.items(), key=lambda t: -t[1][0]):
            f.write(f"| {us / 1e3:.3f} | {100 * us / tot:.1f}% | {n} | `{key}` |\n")
    print(open(out + "_summary.md").read())


if __name__ == "__main__":
    main()
