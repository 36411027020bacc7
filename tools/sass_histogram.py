#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of mvs_b200/libmvs_b200.so (runs without a GPU): the evidence that the kernels use the
Blackwell paths they claim (B200_PROFILING.md "What proves a Blackwell-native kernel").

    python tools/sass_histogram.py [out.md]          # default: profiles/r2_sass_opcodes.md
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UTMALDG = tensor-map TMA load
(cp.async.bulk.tensor), UBLKCP = 1-D bulk TMA copy (cp.async.bulk), LDGSTS = cp.async, SYNCS = mbarrier ops.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mvs_b200", "libmvs_b200.so")
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "HMMA", "HFMA2", "FFMA2", "LDS", "LDG", "STG", "LDL", "STL"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcodes.md")
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    with open(out, "w") as f:
        f.write("# SASS opcode histogram per kernel -- `python tools/sass_histogram.py` (cuobjdump -sass mvs_b200/libmvs_b200.so)\n\n")
        f.write("| kernel | instr | " + " | ".join(WATCH) + " |\n|---|---|" + "---|" * len(WATCH) + "\n")
        tot = collections.Counter()
        for k, c in kernels.items():
            f.write(f"| `{k[:110]}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |\n")
            tot.update(c)
        f.write(f"| **library total** | {tot['_total']} | " + " | ".join(str(tot[w]) for w in WATCH) + " |\n")
    print(out, "kernels:", len(kernels), {w: tot[w] for w in WATCH[:9]})


if __name__ == "__main__":
    main()
