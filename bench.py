#!/usr/bin/env python
"""bench.py -- depth-maps/sec of the MVSNet-family cost-volume hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode strict|fast]

Workload (config.workload): BASELINE.json configs[2] = "cfg3": CasMVSNet 3-stage hot path at DTU
1600x1184, N=5 views, D=(48,32,8) -- the configuration the metric is quoted on; it fits one GPU.
One step = one reference view through warp+variance -> CostRegNet -> softmax/regression/confidence
for all three stages (incl. the inter-stage hypothesis resampling), from feature maps to the final
depth + confidence maps.  Synthetic DTU-shaped inputs (mvs_b200/synth.py), seeded weights.

N > 1: one process per GPU (torchrun), every rank runs its own reference views (the path shards over
independent reference views; inference has no collective) => weak scaling; the only communication
is the barrier and the max-over-ranks of the device time.

`--impl reference`: the reference's own CPU implementation of the path (its PyTorch op sequence,
restated in oracle/torch_port.py because /root/reference cannot travel to the GPU box), with all
host threads, on a bounded row-crop of the same workload per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# rank 0 prints ONE JSON line on stdout: NCCL's "NCCL version ..." banner (any NCCL_DEBUG level >= VERSION) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from mvs_b200 import synth

METRIC = "depth-maps/sec (ref-views/sec) at DTU 1600x1184, N=5"
UNIT = "depth-maps/s"
CFG = synth.CONFIGS["cfg3"]
NDEPTHS = (48, 32, 8)
IMG_HW = (1184, 1600)
CPU_SAMPLE_ROWS = 320          # full-res rows per CPU-baseline step (of 1184): keeps every stage /8-divisible


# ------------------------------------------------------------------------------------------------
def host_inputs(seed=0, rows=IMG_HW[0]):
    """Feature maps (per view, per stage), Cas projection matrices, depth values -- NumPy, fp32."""
    n = CFG["n_views"]
    feats, projs = [dict() for _ in range(n)], {}
    for i, (c, d, h, w) in enumerate(CFG["stages"]):
        key = f"stage{i + 1}"
        hh = h * rows // IMG_HW[0]
        f = synth.features(n, c, hh, w, seed + i, 1)
        for v in range(n):
            feats[v][key] = f[v]
        projs[key] = synth.cas_proj_matrices(n, w, seed, 1)
    return feats, projs, synth.depth_planes(192, 1)


def weights():
    import cases
    return [cases.costreg_state("cas", cin=c, base=8, seed=50 + i) for i, (c, _, _, _) in enumerate(CFG["stages"])]


def algorithmic_bytes(mode):
    s = 4 if mode == "strict" else 2
    per_stage = [synth.warp_variance_bytes(CFG["n_views"], 1, c, d, h, w, s, s, per_pixel_depth=(i > 0))
                 for i, (c, d, h, w) in enumerate(CFG["stages"])]
    return per_stage


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_cpu_port(steps, warmup, rows):
    """The reference's op sequence on the host cores (oracle/torch_port.py), row-cropped sample."""
    from oracle import torch_port as TP
    torch.set_num_threads(os.cpu_count() or 1)
    feats, projs, dv = host_inputs(rows=rows)
    tf = [{k: torch.from_numpy(a) for k, a in f.items()} for f in feats]
    tp = {k: torch.from_numpy(a) for k, a in projs.items()}
    sds = [{k: torch.from_numpy(np.asarray(a)) for k, a in sd.items()} for sd in weights()]
    tdv = torch.from_numpy(dv)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            TP.cas_cascade(tf, tp, tdv, sds, ndepths=NDEPTHS, img_hw=(rows, IMG_HW[1]))
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    frac = rows / IMG_HW[0]
    total = sum(times)
    return {"value": frac * len(times) / total, "ms_per_step": 1e3 * total / len(times), "frac": frac,
            "cores": torch.get_num_threads(),
            "sample": f"rows 0..{rows - 1} of {IMG_HW[0]} ({100 * frac:.0f}% of one cfg3 ref view: all 3 stages, full "
                      f"width/D/C/N) per step, {len(times)} step(s), torch {torch.__version__} CPU, fp32"}


def run_gpu_port(dev, steps, warmup, autocast):
    """SURVEY.md §8(d) "reference GPU baseline beside it": the reference's op sequence
    (oracle/torch_port.py = grid_sample + cuDNN conv3d + BN/ReLU + softmax, its stock code path) on
    the same B200, full cfg3 ref view, cudnn.benchmark=True as the reference's train.py:25 sets it.
    A reported baseline (checker code timed as the thing to beat), never on the product path."""
    from oracle import torch_port as TP
    feats, projs, dv = host_inputs()
    tf = [{k: torch.from_numpy(a).to(dev) for k, a in f.items()} for f in feats]
    tp = {k: torch.from_numpy(a).to(dev) for k, a in projs.items()}
    sds = [{k: torch.from_numpy(np.asarray(a)).to(dev) for k, a in sd.items()} for sd in weights()]
    tdv = torch.from_numpy(dv).to(dev)
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    out = {}
    try:
        for name, ac in (("fp32_tf32", False), ("autocast_bf16", True)):
            if ac and not autocast:
                continue
            def one():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    return TP.cas_cascade(tf, tp, tdv, sds, ndepths=NDEPTHS, img_hw=IMG_HW)
            for _ in range(warmup):
                one()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                one()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            out[name] = {"value": 1e3 / ms, "ms_per_step": ms}
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del tf, sds
    torch.cuda.empty_cache()
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu_port(args.steps, args.warmup, CPU_SAMPLE_ROWS)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: CasMVSNet 3-stage hot path 1600x1184 N=5 D=(48,32,8)",
                       "sample_fraction_per_step": r["frac"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch.distributed as dist
    from mvs_b200 import modules, cascade, ops, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mvs_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    feats_np, projs_np, dv_np = host_inputs(seed=rank)
    regs = []
    for (c, _, _, _), sd in zip(CFG["stages"], weights()):
        net = modules.CostRegNet(c, 8, mode=args.mode)
        net.load_state_dict({k: torch.from_numpy(np.asarray(a)) for k, a in sd.items()}, strict=True)
        regs.append(net.to(dev).eval())
    # hand-off dtype of the 2D FeatureNet: bf16 in fast mode (cfg3 is "bf16 inference": what the reference's
    # autocast FeatureNet emits), fp32 in strict mode.  NCHW either way; packed to C8 inside the step.
    fdt = torch.bfloat16 if args.mode == "fast" else torch.float32
    pinned = [{k: torch.from_numpy(a).to(fdt).pin_memory() for k, a in f.items()} for f in feats_np]
    feats = [{k: t.to(dev) for k, t in f.items()} for f in pinned]
    projs = {k: torch.from_numpy(a).to(dev) for k, a in projs_np.items()}
    dv = torch.from_numpy(dv_np).to(dev)
    dmin, dmax = float(dv_np[0, 0]), float(dv_np[0, -1])
    h2d_bytes = sum(t.numel() * t.element_size() for f in pinned for t in f.values())
    host_out = [torch.empty(1, *IMG_HW, dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h_bytes = sum(t.numel() * 4 for t in host_out)

    def step(fs):
        with torch.no_grad():
            return cascade.cascade_hot_path(fs, projs, dv, regs, ndepths=NDEPTHS, img_hw=IMG_HW, depth_min=dmin,
                                            depth_max=dmax)

    def timed(fn, k, tail=None):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        if tail is not None:
            tail()                                       # e.g. make the timing stream wait for side-stream work
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- pass 1 (eager launches, per-kernel CUDA events): roofline of the fused builder + the eager step time ----
    for _ in range(max(args.warmup, 3)):
        step(feats)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ops.KERNEL_TIMERS = {}
    ms_eager = timed(lambda: step(feats), args.steps)
    timers, ops.KERNEL_TIMERS = ops.KERNEL_TIMERS, None
    launches_per_step = (_lib.launch_count() - n0) // args.steps

    # ---- pass 2 (the deployment form): the same step captured once in a CUDA graph (mvs_b200.GraphedStep) and
    # replayed -- ~85 launches per ref view cost more host time than GPU time once issued one by one from Python ----
    use_graph = not args.no_graph
    if use_graph:
        from mvs_b200.graph import GraphedStep
        g_res = GraphedStep(lambda: step(feats))
        for _ in range(2):
            g_res()
        ms = timed(g_res, args.steps)
    else:
        ms = ms_eager
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps

    # ---- e2e: every step copies ITS inputs from pinned host memory and reads its result back.  The copy of
    # step i+1 runs on a copy stream into the other half of a double buffer while step i computes, stage by stage
    # (coarse stage first; the compute stream waits per STAGE through cascade_hot_path's stage_hook, so stage 1 starts
    # after 14 % of the bytes have landed); the read-back of step i-1 runs on its own stream (PCIe is full duplex).
    # Nothing is reused across steps.  Two launch forms are measured, the better one is reported:
    #   "graph"  whole step replayed from a CUDA graph after ALL of its inputs have landed,
    #   "staged" eager launches with the per-stage waits.
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    NBUF = 3            # device-side input buffers: the copy of step i+1 only waits for the compute of step i-2
    dbuf = [[{k: torch.empty_like(t, device=dev) for k, t in f.items()} for f in pinned] for _ in range(NBUF)]
    stage_keys = [f"stage{i + 1}" for i in range(len(NDEPTHS))]
    ev_stage = [[torch.cuda.Event() for _ in stage_keys] for _ in range(NBUF)]
    ev_free = [torch.cuda.Event() for _ in range(NBUF)]
    ev_d2h = [torch.cuda.Event() for _ in range(NBUF)]
    host_outs = [[torch.empty(1, *IMG_HW, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(NBUF)]
    g_e2e = [GraphedStep(lambda j=j: step(dbuf[j])) for j in range(NBUF)] if use_graph else None

    def make_step_e2e(form):
        counter = [0]

        def step_e2e():
            i = counter[0]; counter[0] += 1
            j = i % NBUF
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(copy_stream):
                if i >= NBUF:
                    copy_stream.wait_event(ev_free[j])       # the step that last read this buffer has finished
                for si, key in enumerate(stage_keys):
                    for fd, fh in zip(dbuf[j], pinned):
                        fd[key].copy_(fh[key], non_blocking=True)
                    ev_stage[j][si].record(copy_stream)
            if i >= NBUF:
                cur.wait_event(ev_d2h[j])                    # the read-back of this half's previous result has finished
            if form == "graph":
                cur.wait_event(ev_stage[j][-1])
                out = g_e2e[j]()
            else:
                with torch.no_grad():
                    out = cascade.cascade_hot_path(dbuf[j], projs, dv, regs, ndepths=NDEPTHS, img_hw=IMG_HW, depth_min=dmin,
                                                   depth_max=dmax, stage_hook=lambda si: cur.wait_event(ev_stage[j][si]))
            ev_free[j].record(cur)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev_free[j])
                host_outs[j][0].copy_(out["depth"], non_blocking=True)
                host_outs[j][1].copy_(out["photometric_confidence"], non_blocking=True)
                ev_d2h[j].record(d2h_stream)
        return step_e2e

    def e2e_tail():                                      # the timed region ends when the LAST result has landed on the host
        cur = torch.cuda.current_stream()
        for j in range(NBUF):
            cur.wait_event(ev_d2h[j])

    e2e_forms = {}
    for form in (["graph"] if use_graph else []) + ["staged"]:
        fn = make_step_e2e(form)
        for _ in range(NBUF):
            fn()
        e2e_forms[form] = timed(fn, args.steps, e2e_tail)
    e2e_form = min(e2e_forms, key=e2e_forms.get)
    ms_e2e = e2e_forms[e2e_form]

    # the same host->device copies alone (no compute): shows how much of the e2e step is the PCIe transfer
    def h2d_only():
        for fd, fh in zip(dbuf[0], pinned):
            for k, t in fh.items():
                fd[k].copy_(t, non_blocking=True)
    h2d_only()
    ms_h2d = timed(h2d_only, args.steps)

    # roofline of the fused warp+variance kernel from the events recorded inside the timed region
    ev = timers.get("warp_variance", [])
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev)
    n_launch = len(ev)
    per_stage = algorithmic_bytes(args.mode)
    alg_bytes = sum(per_stage) * (n_launch / len(per_stage))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = run_cpu_port(1, 0, CPU_SAMPLE_ROWS) if (world == 1 and not args.no_cpu_baseline) else None
    ref_gpu = None
    if world == 1 and not args.no_ref_gpu:
        try:
            ref_gpu = run_gpu_port(dev, 5, 3, True)
        except Exception as e:   # a reported side number must never take the bench line down
            ref_gpu = {"error": f"{type(e).__name__}: {e}"[:200]}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "warp_variance_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.mode)
    line = {
        "metric": METRIC, "value": world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "strict" else "bf16",
        "data": "synthetic", "eager_ms_per_step": ms_eager / args.steps,
        "launch": ("CUDA graph replay of the captured step (mvs_b200.GraphedStep); eager_ms_per_step = the same step issued "
                   "launch by launch from Python, the pass the per-kernel events of `roofline` come from") if use_graph
                  else "eager launches",
        "config": {"workload": "cfg3: CasMVSNet 3-stage hot path 1600x1184 N=5 D=(48,32,8), 1 ref view per GPU per step",
                   "mode": args.mode, "l2": "inputs+intermediates per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                   "features": ("bf16" if args.mode == "fast" else "fp32") + " NCHW feature maps (packed to fp16 C8H inside the step in fast mode)"},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "h2d_only_ms_per_step": ms_h2d / args.steps, "form": e2e_form,
                "ms_per_step_by_form": {k: v / args.steps for k, v in e2e_forms.items()},
                "note": "input copy (per stage, coarse first) overlaps compute on a copy stream, read-back on a third stream; "
                        "the step is PCIe-bound when h2d_only_ms_per_step ~ ms_per_step"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "warp_variance (fused homography warp + variance, 3 launches/step)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                     "avg_launch_ms": kernel_ms / max(n_launch, 1), "launches_timed": n_launch,
                     "share_of_step": kernel_ms / ms_eager if ms_eager > 0 else None},
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                                "sample": cpu["sample"]}
    if ref_gpu is not None:
        line["ref_gpu_baseline"] = {"unit": UNIT, "what": "reference op sequence (grid_sample + cuDNN conv3d, oracle/torch_port.py) "
                                    "on the same GPU, full cfg3 ref view, cudnn.benchmark=True, features resident", **ref_gpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["strict", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches only (no CUDA graph replay)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-cuDNN side measurement")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
