#!/usr/bin/env python
"""profiles/warp_variance_traffic.json from the three per-stage `ncu --set full` captures of the fused builder
(dram__bytes_read.sum + dram__bytes_write.sum per launch) -- the `roofline.traffic` figure bench.py reports.

    python tools/traffic_json.py gpurun_out/<tag>_warp_c8h_s{1,2,3}.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvs_b200 import synth


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u)
        return v * scale if scale else v
    return {"read": get("dram__bytes_read.sum"), "write": get("dram__bytes_write.sum"),
            "time_us": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                       ({"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[hdr.index("gpu__time_duration.sum")].lower().replace("second", "s").replace("usecond", "us")]
                        if units[hdr.index("gpu__time_duration.sum")].lower() in ("ns", "us", "ms") else 1.0),
            "issue": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1": get("l1tex__t_sector_hit_rate.pct"), "l2": get("lts__t_sector_hit_rate.pct")}


def main():
    cfg = synth.CONFIGS["cfg3"]
    launches = []
    for si, rep in enumerate(sys.argv[1:4]):
        c, d, h, w = cfg["stages"][si]
        m = metrics(rep)
        alg = synth.warp_variance_bytes(cfg["n_views"], 1, c, d, h, w, 2, 2, per_pixel_depth=(si > 0))
        launches.append({"stage": si + 1, "dram_read_MB": round(m["read"] / 1e6, 1), "dram_write_MB": round(m["write"] / 1e6, 1),
                         "algorithmic_MB": round(alg / 1e6, 1), "time_us": round(m["time_us"], 1),
                         "issue_active_pct": round(m["issue"], 1), "l1_hit_pct": round(m["l1"], 1), "l2_hit_pct": round(m["l2"], 1)})
    mean = sum((l["dram_read_MB"] + l["dram_write_MB"]) * 1e6 for l in launches) / len(launches)
    out = {"fast": mean,
           "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), mean over the 3 cfg3 stage launches",
           "source": "ncu --set full --clock-control none, tools/ncu_capture.sh (tools/prof_warp.py --mode c8h --stages N --reps 1); "
                     "summaries in profiles/<tag>_warp_c8h_sN.txt",
           "launches": launches,
           "algorithmic_bytes_mean": sum(l["algorithmic_MB"] for l in launches) * 1e6 / len(launches)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "warp_variance_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
