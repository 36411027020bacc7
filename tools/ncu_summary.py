#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics (raw page) and the top stall sites (source page).
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 25]
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__ctas_launched.sum",
        "smsp__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tmem.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    print("== raw metrics:", vals[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"  {h:82s} {vals[i]:>16s} {units[i]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h = None
    for r_i, r in enumerate(src):
        if "Source" in r and any("Sampl" in c for c in r):
            h = r; body = src[r_i + 1:]; break
    if h is None:
        print("no source page"); return
    col_src = h.index("Source")
    col_samp = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_")]
    rows = []
    for r in body:
        try:
            s = float(r[col_samp])
        except (ValueError, IndexError):
            continue
        rows.append((s, r))
    tot = sum(s for s, _ in rows) or 1
    print(f"== top stall sites (of {tot:.0f} samples)")
    for s, r in sorted(rows, key=lambda t: -t[0])[:top]:
        st = sorted(((float(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
        st = ", ".join(f"{n[6:]}={v:.0f}" for v, n in st if v > 0)
        print(f"  {100 * s / tot:5.1f}%  {r[col_src][:70]:70s} {st}")


if __name__ == "__main__":
    main()
