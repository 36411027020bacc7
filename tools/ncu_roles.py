#!/usr/bin/env python
"""Attribute the warp-state samples of a conv3d_umma_kernel capture to pipeline roles / barrier waits.
    python tools/ncu_roles.py gpurun_out/x.ncu-rep
Barrier offsets in the CTA's barrier block: +0x00 full[8], +0x40 empty[8], +0x80 tfull[4], +0xa0 tempty[4]."""
import collections
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[h]; cs = hdr.index("Source"); samp = hdr.index("# Samples")
body = rows[h + 1:]
tot = sum(float(r[samp] or 0) for r in body)
names = {0x00: "wait full   (issuer  <- loads)", 0x40: "wait empty  (producer <- free slot)",
         0x80: "wait tfull  (epilogue <- MMAs)", 0xa0: "wait tempty (issuer  <- epilogue)"}
agg = collections.Counter(); last = None
for r in body:
    s = r[cs]; n = float(r[samp] or 0)
    m = re.search(r"TRYWAIT[^\[]*\[[^\]]*?(?:\+0x([0-9a-f]+))?\]", s)
    if m:
        off = int(m.group(1) or "0", 16)
        key = 0xa0 if off >= 0xa0 else 0x80 if off >= 0x80 else 0x40 if off >= 0x40 else 0
        last = names[key]; agg[last] += n; continue
    if last and ("BRA" in s or "YIELD" in s or "NOP" in s):
        agg[last] += n; continue
    last = None
    if "EXIT" in s: agg["exit barrier (CTA tail)"] += n
    elif "LDTM" in s: agg["LDTM (tcgen05.ld)"] += n
    elif "UTCHMMA" in s or "UTCBAR" in s: agg["UTCHMMA/UTCBAR issue"] += n
    elif "LDGSTS" in s: agg["LDGSTS (cp.async issue)"] += n
    elif "DEPBAR" in s: agg["cp.async wait_group"] += n
    elif "STG" in s: agg["STG (epilogue store)"] += n
    elif "LDG" in s: agg["LDG (weights / skip / scale)"] += n
    elif re.search(r"\bU[A-Z0-9]+\b", s.split()[0] if s.split() else "") or s.strip().startswith(("UIADD", "ULOP", "UMOV", "UIMAD", "USEL", "ULEA", "UISETP", "LDCU", "USHF")):
        agg["uniform-datapath (issuer descriptor math)"] += n
    else: agg["other ALU/control"] += n
print(f"{sys.argv[1]}: {tot:.0f} samples")
for k, v in agg.most_common():
    print(f"   {100 * v / tot:5.1f}%  {k}")
