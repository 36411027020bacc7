#!/usr/bin/env python
"""bench.py -- depth-maps/sec of the MVSNet-family cost-volume hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode fast|strict]
                    [--config cfg3|cfg5|cfg2]

Default workload (config.workload): BASELINE.json configs[2] = "cfg3": CasMVSNet 3-stage, DTU 1600x1184, N=5 views,
D=(48,32,8) -- the configuration the metric is quoted on; it fits one GPU.  `--config cfg5` (Tanks&Temples shape
1920x1056, N=7, D=64/32/8, 4 reference views per GPU per step) and `--config cfg2` (MVSNet 640x512, N=5, D=192, batch 4)
run BASELINE.json configs[4] / [1] through the same code.

What is timed (one "step" = one batch of reference views through the path):
  value        hot path only, inputs resident in HBM in the hand-off format the repo's FeatureNet mirror emits (fp16 C8H
               feature maps): warp+variance -> CostRegNet -> softmax/regression/confidence for every stage incl. the
               inter-stage hypothesis resampling.  CUDA-graph replay; `eager_ms_per_step` = launch by launch, the pass
               the per-kernel CUDA events of `roofline` come from.
  from_images  the whole model (FeatureNet mirror + hot path) from device-resident uint8 images.
  e2e          what the reference's driver does per sample (CasMVSNet/test.py:176-181: tocuda -> model(imgs, proj_matrices,
               depth_values) -> tensor2numpy): uint8 images in PINNED HOST memory -> H2D -> whole model -> D2H of depth +
               confidence, copies inside the timed region and overlapped with the neighbouring steps on side streams.
  strict       the same hot path in strict fp32 mode (the 1e-4 parity mode), so that "speed at which parity" is a
               driver-run number.
  parity       measured in-run on this workload at full size against the reference's op sequence executed on the same GPU in
               fp32 with TF32 off (oracle/torch_port.py: checker, outside every timed region).
N > 1: one process per GPU (torchrun), every rank runs its own reference views (independent units, no data-path
collective at inference) => weak scaling; the only communication is the barrier and the max-over-ranks of the device time.

`--impl reference`: the reference's own CPU implementation of the path from images (its PyTorch op sequence incl. the
FeatureNet, restated in oracle/torch_port.py because /root/reference cannot travel to the GPU box), all host threads,
one FULL reference view per step.
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

# Exactly ONE line on the real stdout: everything else that writes to fd 1 from here on (NCCL's version banner comes from C
# code and ignored NCCL_DEBUG_FILE on the 2-GPU box, library chatter, warnings) is sent to stderr; emit() writes the JSON
# line to the saved descriptor.
_REAL_STDOUT = None          # set in __main__ only: importing bench (tests, tools) must not touch the process's descriptors


def _claim_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


UNIT = "depth-maps/s"
WORKLOADS = {
    "cfg3": dict(kind="cas", key="cfg3", img_hw=(1184, 1600), ndepths=(48, 32, 8),
                 metric="depth-maps/sec (ref-views/sec) at DTU 1600x1184, N=5",
                 name="cfg3: CasMVSNet 3-stage 1600x1184 N=5 D=(48,32,8), 1 ref view per GPU per step"),
    "cfg5": dict(kind="cas", key="cfg5", img_hw=(1056, 1920), ndepths=(64, 32, 8),
                 metric="depth-maps/sec (ref-views/sec) at Tanks&Temples 1920x1056, N=7",
                 name="cfg5: CasMVSNet 3-stage 1920x1056 N=7 D=(64,32,8), 4 ref views per GPU per step"),
    "cfg4": dict(kind="cvp_train", key="cfg4", img_hw=(864, 1152), ndepths=(48, 8), nsrc=4, nscale=2, batch=2,
                 metric="training samples/sec (ref-views/sec), CVP-MVSNet 1152x864, N=5, nscale=2",
                 name="cfg4: CVP-MVSNet training step 1152x864 N=5 nscale=2 (D=48 coarse, 8 refine), 2 samples per GPU per step, "
                      "Adam, single flat NCCL gradient bucket"),
    "cfg2": dict(kind="mvsnet", key="cfg2", img_hw=(512, 640), ndepths=(192,),
                 metric="depth-maps/sec (ref-views/sec) at 640x512, N=5, D=192",
                 name="cfg2: MVSNet 640x512 N=5 D=192, 4 ref views per GPU per step"),
}
SEED_MODEL = 40


# ------------------------------------------------------------------------------------------------------------------
def host_inputs(wl, seed=0, batch=None):
    """uint8 images [B,N,3,H,W], projection matrices, depth values -- NumPy, as the loader hands them."""
    cfg = synth.CONFIGS[wl["key"]]
    n, B = cfg["n_views"], batch or cfg["batch"]
    H, W = wl["img_hw"]
    if wl["kind"] == "cas":
        projs = {f"stage{i + 1}": synth.cas_proj_matrices(n, W // s, seed, B) for i, s in enumerate((4, 2, 1))}
        return dict(imgs=synth.images_u8(n, H, W, seed, B), projs=projs, depth_values=synth.depth_planes(192, B))
    c, d, h, w = cfg["stages"][0]
    return dict(imgs=synth.images_u8(n, H, W, seed, B), projs=synth.proj_matrices(n, w, seed, B),
                depth_values=synth.depth_planes(d, B, hi=synth.DTU_DEPTH_MIN + 2.65 * d))


def model_state(wl):
    import cases
    if wl["kind"] == "cas":
        return cases.full_model_state(SEED_MODEL)
    return cases.mvsnet_model_state(SEED_MODEL)


def algorithmic_bytes(wl, mode, batch):
    cfg = synth.CONFIGS[wl["key"]]
    s = 4 if mode == "strict" else 2
    return [synth.warp_variance_bytes(cfg["n_views"], batch, c, d, h, w, s, s, per_pixel_depth=(wl["kind"] == "cas" and i > 0))
            for i, (c, d, h, w) in enumerate(cfg["stages"])]


def pin_to_gpu_numa(index):
    """Bind this process to the CPUs next to its GPU before any pinned allocation (first touch places the pinned pages on
    that NUMA node); eight ranks streaming from one node was the 8-GPU e2e limiter of round 1.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


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


# ------------------------------------------------------------------------------------------------------------------
# checker / baseline legs (oracle/torch_port.py; never inside a timed region of the repo arm)
def _port_inputs(wl, dev, batch=None):
    hi = host_inputs(wl, batch=batch)
    sd = {k: torch.from_numpy(np.asarray(a)).to(dev) for k, a in model_state(wl).items()}
    dv = torch.from_numpy(hi["depth_values"]).to(dev)
    imgs = torch.from_numpy(hi["imgs"]).to(dev).float() / 255.0              # general_eval.py:81-86
    if wl["kind"] == "cas":
        projs = {k: torch.from_numpy(a).to(dev) for k, a in hi["projs"].items()}
        return dict(imgs=imgs, projs=projs, dv=dv, sd=sd)
    return dict(imgs=imgs, projs=torch.from_numpy(hi["projs"]).to(dev), dv=dv, sd=sd)


def _port_step(wl, pi, features=None):
    from oracle import torch_port as TP
    if wl["kind"] == "cas":
        if features is not None:
            sds = [{k[len(f"cost_regularization.{i}."):]: v for k, v in pi["sd"].items() if k.startswith(f"cost_regularization.{i}.")}
                   for i in range(len(wl["ndepths"]))]
            return TP.cas_cascade(features, pi["projs"], pi["dv"], sds, ndepths=wl["ndepths"], img_hw=wl["img_hw"])
        return TP.cas_model(pi["imgs"], pi["projs"], pi["dv"], pi["sd"], ndepths=wl["ndepths"])
    if features is None:
        return TP.mvsnet_model(pi["imgs"], pi["projs"], pi["dv"], pi["sd"])
    projs = torch.unbind(pi["projs"], 1)
    var = TP.variance_volume(features[0], features[1:], projs[0], projs[1:], pi["dv"])
    csd = {k[len("cost_regularization."):]: v for k, v in pi["sd"].items() if k.startswith("cost_regularization.")}
    depth, conf = TP.regress(TP.costreg(var, csd, "mvsnet"), pi["dv"], clamp_index=False)
    return {"depth": depth, "photometric_confidence": conf}


def run_cpu_port(wl, steps, warmup):
    """The reference's op sequence on the host cores, ONE full reference view per step (batch 1 of the config's batch)."""
    torch.set_num_threads(os.cpu_count() or 1)
    pi = _port_inputs(wl, "cpu", batch=1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            _port_step(wl, pi)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    what = "images -> FeatureNet x N views -> 3-stage cascade" if wl["kind"] == "cas" else "images -> FeatureNet x N views -> cost volume -> CostRegNet -> regression"
    return {"value": len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": torch.get_num_threads(),
            "sample": f"1 full reference view per step ({what}; full H/W/D/C/N of {wl['key']}), {len(times)} step(s), "
                      f"torch {torch.__version__} CPU, fp32, oracle/torch_port.py"}


def run_gpu_port(wl, dev, steps, warmup):
    """SURVEY.md 8(d) "reference GPU baseline beside it": the reference's op sequence (grid_sample + cuDNN conv3d + BN/ReLU +
    softmax, incl. its FeatureNet) on the same B200, cudnn.benchmark=True as the reference's train.py:25 sets it.  A reported
    baseline (checker code timed as the thing to beat), never on the product path."""
    pi = _port_inputs(wl, dev, batch=1)
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    out = {}
    try:
        for name, ac in (("fp32_tf32", False), ("autocast_bf16", True)):
            def one():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    return _port_step(wl, pi)
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
    del pi
    torch.cuda.empty_cache()
    return out


def depth_errors(ours, ref, keys):
    """Per stage: relative L-inf and relative L1 of the depth map, fraction of pixels within 1e-4 / 1e-3 relative."""
    out = {}
    for k in keys:
        a, b = ours[k]["depth"].double(), ref[k]["depth"].double()
        rel = (a - b).abs() / b.abs()
        out[k] = {"depth_rel_linf": float(rel.max()), "depth_rel_l1": float((a - b).abs().mean() / b.abs().mean()),
                  "frac_within_1e-4": float((rel <= 1e-4).double().mean()), "frac_within_1e-3": float((rel <= 1e-3).double().mean())}
    return out


# ------------------------------------------------------------------------------------------------------------------
def main_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu_port(wl, args.steps, args.warmup)
    line = {"impl": "reference", "metric": wl["metric"], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample_fraction_per_step": 1.0,
                       "inputs": "uint8 images scaled by 1/255 (the loader's read_img), full view"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
def main_ours(args, wl):
    import torch.distributed as dist
    from mvs_b200 import modules, cascade, ops, _lib
    from mvs_b200.featurenet import CascadeMVSNet
    from mvs_b200.mvsnet import MVSNet
    from mvs_b200.graph import GraphedStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; mvs_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    numa_cpus = pin_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    cfg = synth.CONFIGS[wl["key"]]
    B, cas = cfg["batch"], wl["kind"] == "cas"
    hi = host_inputs(wl, seed=rank)
    sd = {k: torch.from_numpy(np.asarray(a)) for k, a in model_state(wl).items()}
    dv = torch.from_numpy(hi["depth_values"]).to(dev)
    dmin, dmax = float(hi["depth_values"][0, 0]), float(hi["depth_values"][0, -1])
    keys = [f"stage{i + 1}" for i in range(len(wl["ndepths"]))] if cas else None

    def build_model(mode):
        if cas:
            m = CascadeMVSNet(ndepths=wl["ndepths"], mode=mode)
            m.load_state_dict(sd, strict=True)
        else:
            m = MVSNet(mode=mode)
            m.load_state_dict(sd, strict=True)
        return m.to(dev).eval()

    model = build_model(args.mode)
    host_in = [torch.from_numpy(hi["imgs"]).pin_memory()]                                       # uint8 [B,N,3,H,W]
    imgs_dev = host_in[0].to(dev)
    with torch.no_grad():
        feats = model.extract(imgs_dev)                 # hand-off format of the mode: C8H fp16 (fast) / NCHW fp32 (strict)
    if cas:
        projs = {k: torch.from_numpy(a).to(dev) for k, a in hi["projs"].items()}

        def hot_step(fs=feats, m=model):
            with torch.no_grad():
                return cascade.cascade_hot_path(fs, projs, dv, m.cost_regularization, ndepths=wl["ndepths"], img_hw=wl["img_hw"],
                                                depth_min=dmin, depth_max=dmax)

        def full_step(im, m=model):
            with torch.no_grad():
                return m(im, projs, dv, depth_min=dmin, depth_max=dmax)
        out_shape = (B, *wl["img_hw"])
    else:
        projs = torch.from_numpy(hi["projs"]).to(dev)

        def hot_step(fs=feats, m=model):
            with torch.no_grad():
                return modules.mvsnet_hot_path(list(fs), projs, dv, m.cost_regularization)

        def full_step(im, m=model):
            with torch.no_grad():
                return m(im, projs, dv)
        out_shape = (B, *cfg["stages"][0][2:])
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_in)
    d2h_bytes = 2 * int(np.prod(out_shape)) * 4

    # ---- pass 1 (eager launches, per-kernel CUDA events): roofline of the fused builder + the eager step time ----
    for _ in range(max(args.warmup, 3)):
        hot_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ops.KERNEL_TIMERS = {}
    ms_eager = timed(hot_step, args.steps)
    timers, ops.KERNEL_TIMERS = ops.KERNEL_TIMERS, None
    launches_per_step = (_lib.launch_count() - n0) // args.steps

    # ---- pass 2 (the deployment form): the same step captured once in a CUDA graph and replayed ----
    use_graph = not args.no_graph
    if use_graph:
        g_hot = GraphedStep(hot_step)
        for _ in range(2):
            g_hot()
        ms = timed(g_hot, args.steps)
        value_forms = {"serial": ms}
        # Second launch form: consecutive reference views are independent, so two captured steps alternate on two streams and
        # view i+1 fills the SMs that the low-resolution layers / kernel tails of view i leave idle.  Same kernels, same
        # results, every step still a whole reference-view batch; the better form is reported (`value_form`).
        g_hot_b = GraphedStep(hot_step)
        vs = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        graphs, ev_done, vcount = [g_hot, g_hot_b], [torch.cuda.Event(), torch.cuda.Event()], [0]
        ev_start = torch.cuda.Event()

        def step_2s():
            i = vcount[0]; vcount[0] += 1
            j = i & 1
            if i < 2:                                        # the side streams start after whatever precedes on this stream
                ev_start.record(torch.cuda.current_stream())
                vs[j].wait_event(ev_start)
            with torch.cuda.stream(vs[j]):
                graphs[j]()
                ev_done[j].record(vs[j])

        def tail_2s():                                       # the timed region ends when both streams have drained
            cur = torch.cuda.current_stream()
            for j in range(2):
                cur.wait_event(ev_done[j])
            vcount[0] = 0

        for _ in range(4):
            step_2s()
        tail_2s()
        value_forms["two_stream"] = timed(step_2s, args.steps, tail_2s)
        value_form = min(value_forms, key=value_forms.get)
        ms = value_forms[value_form]
    else:
        ms = ms_eager
        value_forms, value_form = {"eager": ms}, "eager"
    clocks = sampler.stop() if rank == 0 else None

    # ---- whole model from device-resident images (FeatureNet mirror + hot path) ----
    NBUF = 2
    dbuf = [torch.empty_like(host_in[0], device=dev) for _ in range(NBUF)]
    for t in dbuf:
        t.copy_(host_in[0])
    g_full = [GraphedStep(lambda j=j: full_step(dbuf[j])) for j in range(NBUF)] if use_graph else None
    run_full = (lambda j: g_full[j]()) if use_graph else (lambda j: full_step(dbuf[j]))
    ms_full = timed(lambda: run_full(0), args.steps)
    from_images = {"value": world * B * args.steps / (ms_full * 1e-3), "unit": UNIT, "ms_per_step": ms_full / args.steps,
                   "what": "FeatureNet mirror (native fp16 C8 engine: tcgen05 3x3 layers with the images folded onto the row axis, "
                           "space-to-depth 5x5/s2 layers, fused FPN laterals; all N views batched, C8H out) + hot path, uint8 images resident in HBM"}

    # ---- e2e: every step copies ITS inputs from pinned host memory and reads its result back; the copy of step i+1 (copy
    # stream) and the read-back of step i-1 (its own stream; PCIe is full duplex) overlap the compute of step i ----
    copy_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]
    ev_free = [torch.cuda.Event() for _ in range(NBUF)]
    ev_d2h = [torch.cuda.Event() for _ in range(NBUF)]
    host_outs = [[torch.empty(out_shape, dtype=torch.float32).pin_memory() for _ in range(2)] for _ in range(NBUF)]
    counter = [0]

    def step_e2e():
        i = counter[0]; counter[0] += 1
        j = i % NBUF
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            if i >= NBUF:
                copy_stream.wait_event(ev_free[j])           # the step that last read this input buffer has finished
            dbuf[j].copy_(host_in[0], non_blocking=True)
            ev_in[j].record(copy_stream)
        cur.wait_event(ev_in[j])
        if i >= NBUF:
            cur.wait_event(ev_d2h[j])                        # this buffer's previous result has left the device
        out = run_full(j)
        ev_free[j].record(cur)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(ev_free[j])
            host_outs[j][0].copy_(out["depth"], non_blocking=True)
            host_outs[j][1].copy_(out["photometric_confidence"], non_blocking=True)
            out["depth"].record_stream(d2h_stream); out["photometric_confidence"].record_stream(d2h_stream)
            ev_d2h[j].record(d2h_stream)

    def e2e_tail():                                          # the timed region ends when the LAST result has landed on the host
        cur = torch.cuda.current_stream()
        for j in range(NBUF):
            cur.wait_event(ev_d2h[j])

    for _ in range(NBUF + 1):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, e2e_tail)
    e2e_forms = {"serial": ms_e2e}

    # Second launch form (graph mode): the model's two phases as two graphs on two streams -- the feature
    # extractor of step i+1 runs while the hot path of step i does (its low-resolution layers leave most SMs idle); same
    # kernels, same results, same copies inside the timed region.  The better form is reported (`e2e.form`).
    if use_graph:
        ext_stream = torch.cuda.Stream(device=dev)
        with torch.no_grad():
            g_ext = [GraphedStep(lambda j=j: model.extract(dbuf[j])) for j in range(NBUF)]
        g_hot2 = [GraphedStep(lambda j=j: hot_step(g_ext[j].out)) for j in range(NBUF)]
        ev_ext = [torch.cuda.Event() for _ in range(NBUF)]
        ev_hot = [torch.cuda.Event() for _ in range(NBUF)]
        ev_in2 = [torch.cuda.Event() for _ in range(NBUF)]
        ev_d2h2 = [torch.cuda.Event() for _ in range(NBUF)]
        counter2 = [0]

        def step_e2e_2s():
            i = counter2[0]; counter2[0] += 1
            j = i % NBUF
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(copy_stream):
                if i >= NBUF:
                    copy_stream.wait_event(ev_ext[j])        # the extractor that last read this image buffer has finished
                dbuf[j].copy_(host_in[0], non_blocking=True)
                ev_in2[j].record(copy_stream)
            with torch.cuda.stream(ext_stream):
                ext_stream.wait_event(ev_in2[j])
                if i >= NBUF:
                    ext_stream.wait_event(ev_hot[j])         # the hot path that last read these feature maps has finished
                g_ext[j]()
                ev_ext[j].record(ext_stream)
            cur.wait_event(ev_ext[j])
            if i >= NBUF:
                cur.wait_event(ev_d2h2[j])
            out = g_hot2[j]()
            ev_hot[j].record(cur)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev_hot[j])
                host_outs[j][0].copy_(out["depth"], non_blocking=True)
                host_outs[j][1].copy_(out["photometric_confidence"], non_blocking=True)
                ev_d2h2[j].record(d2h_stream)

        def e2e_tail_2s():
            cur = torch.cuda.current_stream()
            for j in range(NBUF):
                cur.wait_event(ev_d2h2[j])

        for _ in range(NBUF + 1):
            step_e2e_2s()
        e2e_forms["two_stream"] = timed(step_e2e_2s, args.steps, e2e_tail_2s)
    # Third form: every step lives on ONE stream (H2D copy -> whole-model graph -> D2H copies) and consecutive steps alternate
    # between two streams: the copies, the extractor and the hot paths of two independent reference views overlap freely.
    if use_graph:
        fs = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
        ev_f = [torch.cuda.Event(), torch.cuda.Event()]
        ev_fs = torch.cuda.Event()
        counter3 = [0]

        def step_e2e_alt():
            i = counter3[0]; counter3[0] += 1
            j = i & 1
            if i < 2:
                ev_fs.record(torch.cuda.current_stream())
                fs[j].wait_event(ev_fs)
            with torch.cuda.stream(fs[j]):
                dbuf[j].copy_(host_in[0], non_blocking=True)
                out = g_full[j]()
                host_outs[j][0].copy_(out["depth"], non_blocking=True)
                host_outs[j][1].copy_(out["photometric_confidence"], non_blocking=True)
                ev_f[j].record(fs[j])

        def e2e_tail_alt():
            cur = torch.cuda.current_stream()
            for j in range(2):
                cur.wait_event(ev_f[j])
            counter3[0] = 0

        torch.cuda.synchronize()
        for _ in range(4):
            step_e2e_alt()
        e2e_tail_alt()
        e2e_forms["alternating"] = timed(step_e2e_alt, args.steps, e2e_tail_alt)
    e2e_form = min(e2e_forms, key=e2e_forms.get)
    ms_e2e = e2e_forms[e2e_form]

    def h2d_only():
        dbuf[0].copy_(host_in[0], non_blocking=True)
    h2d_only()
    ms_h2d = timed(h2d_only, args.steps)

    # ---- roofline of the fused warp+variance kernel from the events recorded inside the timed eager pass ----
    ev = timers.get("warp_variance", [])
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev)
    n_launch = len(ev)
    per_stage = algorithmic_bytes(wl, args.mode, B)
    alg_bytes = sum(per_stage) * (n_launch / len(per_stage))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    stage_ms = [sum(a.elapsed_time(b) for a, b in ev[i::len(per_stage)]) / max(len(ev[i::len(per_stage)]), 1) for i in range(len(per_stage))]
    per_gpu = [achieved]
    if world > 1:                # every rank times its own builder launches: BASELINE configs[4] asks for the per-GPU figure
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = achieved
        dist.all_reduce(t)
        per_gpu = [float(v) for v in t.tolist()]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- single-GPU extras (rank 0, N = 1): strict mode, in-run parity, CPU and cuDNN baselines ----
    extras = {}
    if world == 1 and not args.no_parity:
        try:
            extras.update(parity_and_strict(args, wl, dev, model, build_model, hi, projs, dv, dmin, dmax, keys))
        except Exception as e:   # reported side numbers must never take the bench line down
            extras["parity"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    cpu = run_cpu_port(wl, 1, 0) if (world == 1 and not args.no_cpu_baseline) else None
    ref_gpu = None
    if world == 1 and not args.no_ref_gpu:
        try:
            ref_gpu = run_gpu_port(wl, dev, 5, 3)
        except Exception as e:
            ref_gpu = {"error": f"{type(e).__name__}: {e}"[:200]}
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "warp_variance_traffic.json")
    if os.path.exists(tpath) and wl["key"] == "cfg3":
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get(args.mode), "static: " + tj.get("source", "ncu capture committed under profiles/")
    line = {
        "metric": wl["metric"], "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "strict" else "bf16",
        "data": "synthetic", "eager_ms_per_step": ms_eager / args.steps, "value_form": value_form,
        "ms_per_step_by_form": {k: v / args.steps for k, v in value_forms.items()},
        "launch": ("CUDA graph replay of the captured step (mvs_b200.GraphedStep); value_form 'two_stream' = consecutive steps "
                   "(independent reference views) alternate on two streams; eager_ms_per_step = the same step issued "
                   "launch by launch from Python, the pass the per-kernel events of `roofline` come from") if use_graph
                  else "eager launches",
        "config": {"workload": wl["name"], "mode": args.mode, "ref_views_per_step": B,
                   "l2": "inputs+intermediates per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                   "features": "value: feature maps resident in HBM in the hand-off format of the FeatureNet mirror "
                               + ("(fp16 C8H)" if args.mode == "fast" else "(fp32 NCHW)"),
                   "numa_cpus": numa_cpus},
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "h2d_only_ms_per_step": ms_h2d / args.steps, "form": e2e_form,
                "ms_per_step_by_form": {k: v / args.steps for k, v in e2e_forms.items()},
                "what": "pinned uint8 images [B,N,3,H,W] -> H2D -> /255 + FeatureNet mirror + hot path -> D2H depth + confidence "
                        "(the reference's model(imgs, proj_matrices, depth_values) boundary)",
                "note": "forms: serial = copy stream + one whole-model graph per step + read-back stream; two_stream = the extractor "
                        "of step i+1 (own graph / stream) under the hot path of step i; alternating = every step (copy in, whole-model "
                        "graph, copies out) on one of two streams in turn; the copies are inside the timed region in all of them"},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {"kernel": f"warp_variance (fused homography warp + variance, {len(per_stage)} launch(es)/step)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                     "avg_launch_ms": kernel_ms / max(n_launch, 1), "launches_timed": n_launch,
                     "per_stage_ms": stage_ms, "per_stage_frac": [b / (t * 1e-3) / 1e9 / peak if t > 0 else None for b, t in zip(per_stage, stage_ms)],
                     "share_of_step": kernel_ms / ms_eager if ms_eager > 0 else None,
                     "per_gpu_achieved": per_gpu, "per_gpu_frac": [v / peak for v in per_gpu]},
        "clocks": clocks,
    }
    line["from_images"] = from_images
    line.update(extras)
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": "port", "sample": cpu["sample"]}
    if ref_gpu is not None:
        line["ref_gpu_baseline"] = {"unit": UNIT, "what": "reference op sequence (FeatureNet + grid_sample + cuDNN conv3d, oracle/torch_port.py) "
                                    "on the same GPU from images, 1 ref view per step, cudnn.benchmark=True", **ref_gpu}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def parity_and_strict(args, wl, dev, model, build_model, hi, projs, dv, dmin, dmax, keys):
    """In-run parity at the benchmarked size (checker = the reference's op sequence on the same GPU, fp32, TF32 off) and the
    strict-mode step time.  Not timed as part of `value`."""
    from mvs_b200 import modules, cascade, ops
    from oracle import torch_port as TP
    out = {}
    cas = wl["kind"] == "cas"
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    try:
        pi = _port_inputs(wl, dev)
        with torch.no_grad():
            if cas:
                feats32 = [TP.featurenet(pi["imgs"][:, v], pi["sd"], "feature.") for v in range(pi["imgs"].shape[1])]
                ref_hot = _port_step(wl, pi, features=feats32)
                ks = keys
            else:
                feats32 = [TP.mvsnet_featurenet(pi["imgs"][:, v], pi["sd"]) for v in range(pi["imgs"].shape[1])]
                ref_hot = {"stage1": _port_step(wl, pi, features=feats32)}
                ks = ["stage1"]
        strict = model if args.mode == "strict" else build_model("strict")
        fast = model if args.mode == "fast" else build_model("fast")

        def hot(m, fs):
            with torch.no_grad():
                if cas:
                    return cascade.cascade_hot_path(fs, projs, dv, m.cost_regularization, ndepths=wl["ndepths"], img_hw=wl["img_hw"],
                                                    depth_min=dmin, depth_max=dmax)
                return {"stage1": modules.mvsnet_hot_path(list(fs), projs, dv, m.cost_regularization)}

        res = {}
        for name, m in (("strict", strict), ("fast", fast)):
            res[name] = depth_errors(hot(m, feats32), ref_hot, ks)
        last = ks[-1]
        cur = res[args.mode][last]
        out["parity"] = {"mode": args.mode, "depth_rel_linf": cur["depth_rel_linf"], "depth_l1": cur["depth_rel_l1"],
                         "vs": "reference op sequence (oracle/torch_port.py) on the same GPU, fp32, TF32 off, same fp32 feature maps in; "
                               "final-stage depth at the benchmarked size; errors compound through the cascade",
                         "hot_path": res}
        with torch.no_grad():
            imgs_u8 = torch.from_numpy(hi["imgs"]).to(dev)
            ref_full = _port_step(wl, pi)
            if cas:
                full = {name: depth_errors(m(imgs_u8, projs, dv, depth_min=dmin, depth_max=dmax), ref_full, ks)
                        for name, m in (("strict", strict), ("fast", fast))}
            else:
                full = {name: depth_errors({"stage1": m(imgs_u8, projs, dv)}, {"stage1": ref_full}, ks)
                        for name, m in (("strict", strict), ("fast", fast))}
        out["parity"]["from_images"] = full
        # strict-mode step time (hot path, features resident): the "speed at 1e-4" number
        if args.mode == "fast" and not args.no_strict:
            with torch.no_grad():
                fs = strict.extract(torch.from_numpy(hi["imgs"]).to(dev))
            for _ in range(2):
                hot(strict, fs)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            n = 3
            for _ in range(n):
                hot(strict, fs)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n
            B = synth.CONFIGS[wl["key"]]["batch"]
            out["strict"] = {"ms_per_step": ms, "value": B * 1e3 / ms, "unit": UNIT, "dtype": "f32",
                             "parity": res["strict"][last], "what": "same hot path, strict fp32 kernels (parity mode), eager launches"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
# cfg4: CVP-MVSNet training step (BASELINE.json configs[3]) -- the only place the path has a collective
def cvp_train_inputs(wl, seed, batch, dev):
    H, W = wl["img_hw"]
    nsrc = wl["nsrc"]
    ref_in, src_in, ref_ex, src_ex = [torch.from_numpy(a).to(dev) for a in synth.cvp_cameras(nsrc, W, seed=seed, batch=batch)]
    imgs = torch.from_numpy(synth.images_u8(nsrc + 1, H, W, seed=seed, batch=batch))
    dmin = torch.full((batch,), synth.DTU_DEPTH_MIN, dtype=torch.float64, device=dev)
    dmax = torch.full((batch,), synth.DTU_DEPTH_MAX, dtype=torch.float64, device=dev)
    gts = [torch.from_numpy(synth.depth_surface(H >> i, W >> i, batch)).to(dev) for i in range(wl["nscale"])]
    return dict(imgs_u8=imgs, cams=(ref_in, src_in, ref_ex, src_ex), dmin=dmin, dmax=dmax, gts=gts)


def main_train(args, wl):
    import types
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_args = types.SimpleNamespace(nsrc=wl["nsrc"], nscale=wl["nscale"], mode="train")
    B = wl["batch"]
    if args.impl == "reference":
        # the reference's training step on the host cores: forward + backward + Adam through the torch port, ONE sample per step
        if rank != 0:
            return
        from oracle import torch_port as TP
        from mvs_b200 import pyramid
        import torch.nn.functional as F
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(SEED_MODEL)
        sd = {k: v.detach().clone().requires_grad_(v.dtype == torch.float32 and "running" not in k)
              for k, v in pyramid.network(cfg_args).state_dict().items()}
        inp = cvp_train_inputs(wl, 0, 1, "cpu")
        imgs = inp["imgs_u8"].float() / 255.0
        opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-3)
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            d = TP.cvp_network(imgs[:, 0], imgs[:, 1:], *inp["cams"], inp["dmin"], inp["dmax"], sd, wl["nscale"], True)
            loss = sum(F.smooth_l1_loss(a[g > 425], g[g > 425], reduction="mean") for a, g in zip(d, inp["gts"]))
            loss.backward()
            opt.step()
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        v = len(times) / sum(times)
        line = {"impl": "reference", "metric": wl["metric"], "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": wl["name"], "sample_fraction_per_step": 1.0 / B},
                "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": "1 training sample per step (forward + backward + Adam), oracle/torch_port.py cvp_network, fp32"},
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    from mvs_b200 import pyramid, ops, _lib
    from mvs_b200.train import GradBucket, masked_smooth_l1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    pin_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(SEED_MODEL)                       # identical initial weights on every rank (what DDP's broadcast ensures)
    net = pyramid.network(cfg_args, mode="strict").to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    bucket = GradBucket(net.parameters())
    inp = cvp_train_inputs(wl, rank, B, dev)
    host_imgs = inp["imgs_u8"].pin_memory()
    dimgs = host_imgs.to(dev)
    host_loss = torch.empty(1).pin_memory()

    def step(imgs_u8):
        imgs = imgs_u8.float() / 255.0
        opt.zero_grad(set_to_none=False)
        out = net(imgs[:, 0], imgs[:, 1:], *inp["cams"], inp["dmin"], inp["dmax"])
        loss = sum(masked_smooth_l1(d, g, g > 425) for d, g in zip(out["depth_est_list"], inp["gts"]))   # CVP-MVSNet/train.py:205-209
        loss.backward()
        bucket.reduce()
        bucket.wait()
        opt.step()
        return loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step(dimgs)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ops.KERNEL_TIMERS = {}
    ms = timed(lambda: step(dimgs), args.steps)
    timers, ops.KERNEL_TIMERS = ops.KERNEL_TIMERS, None
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None

    def step_e2e():
        d = host_imgs.to(dev, non_blocking=True)
        host_loss.copy_(step(d).reshape(1), non_blocking=True)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    ev = timers.get("warp_variance", [])
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev)
    H, W = wl["img_hw"]
    per_level = [synth.warp_variance_bytes(wl["nsrc"] + 1, B, 16, d, H >> l, W >> l, 4, 4, per_pixel_depth=(l == 0))
                 for d, l in zip(wl["ndepths"], (wl["nscale"] - 1, 0))]
    alg = sum(per_level) * (len(ev) / len(per_level))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    achieved = alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    line = {"metric": wl["metric"], "value": world * B * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "mode": "strict (training always runs the fp32 kernels)", "samples_per_step_per_gpu": B,
                       "collective": f"one flat bucket of {bucket.numel} fp32 gradients ({bucket.numel * 4 / 1e6:.2f} MB), NCCL all-reduce SUM on a side stream, averaged",
                       "l2": "activations per step exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(host_imgs.numel()), "d2h_bytes_per_step": 4,
                    "what": "pinned uint8 images -> H2D -> /255 -> forward + backward + all-reduce + Adam -> D2H loss"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "warp_variance forward (strict fp32 builder, 2 launches/step: coarse planes + per-pixel refine)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": alg / max(len(ev), 1), "avg_launch_ms": kernel_ms / max(len(ev), 1),
                         "launches_timed": len(ev), "share_of_step": kernel_ms / ms if ms > 0 else None,
                         "note": "the strict NCHW builder is the parity kernel (5 % of HBM peak); the training step is dominated by the fp32 "
                                 "SIMT convolutions (forward, data gradient, weight gradient)"},
            "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline_note"] = "run `bench.py --config cfg4 --impl reference` for the CPU arm (one training sample per step takes minutes)"
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fast", choices=["strict", "fast"])
    ap.add_argument("--config", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches only (no CUDA graph replay)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-cuDNN side measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity check and the strict-mode timing")
    ap.add_argument("--no-strict", action="store_true", help="skip the strict-mode timing")
    a = ap.parse_args()
    _claim_stdout()
    w = WORKLOADS[a.config]
    if w["kind"] == "cvp_train":
        main_train(a, w)
    elif a.impl == "reference":
        main_reference(a, w)
    else:
        main_ours(a, w)
