#!/usr/bin/env python
"""bench.py — frames/sec of the FSF sparse forward hot path on synthetic nuScenes-shaped frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--points P] [--sweeps S]

One "step" = one frame (P points x 6 cameras, default the 10-sweep 300k-point configuration the
BASELINE.json metric is quoted on) through every stage of fullysparsefusion_b200.frame.  Frames are
sharded one per GPU per step (weak scaling, no data-path collective — SURVEY.md §8e).

`value`  frames/s with the frame's inputs already resident in HBM (device-timed, max over ranks).
`e2e`    frames/s through the same public API starting from pinned HOST buffers: the H2D copy of
         points / id planes / lidar2img and a D2H read of the result are inside the timed region.
`roofline`  the dominant HBM-bound stage: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json.
`cpu_baseline`  the torch-CPU port of the reference path (oracle/fsf_torch_cpu.py) on this box's cores.
`--impl reference` times that CPU port alone (the reference cannot be installed: DESIGN.md).
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

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "nuScenes 10-sweep frames/sec (FSF sparse forward hot path)"
UNIT = "frames/s"
FALLBACK_HBM_GBS = 6650.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=300000)
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args):
    return {"workload": f"FSF_nuScenes_config {args.sweeps}-sweep frame: {args.points} pts x 6 cams @1600x900, "
                        "10 class id planes (BASELINE configs[2] shape, one frame per GPU per step)",
            "points": args.points, "sweeps": args.sweeps, "cams": 6, "classes": 10, "frames_per_step_per_gpu": 1}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_cpu_port(args, steps: int, warmup: int):
    """Time the torch-CPU port of the reference path; one step = one full frame on all host cores."""
    import torch

    from fullysparsefusion_b200 import frame
    from oracle import fsf_torch_cpu as P

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    host = frame.synth_frame_host(args.points, args.sweeps, seed=0, pin=False)
    stages, _ = P.build_stages(host)
    per_stage = {name: 0.0 for name, _ in stages}
    for _ in range(warmup):
        for _, fn in stages:
            fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        for name, fn in stages:
            t = time.perf_counter()
            fn()
            per_stage[name] += time.perf_counter() - t
    dt = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} full frame(s) of the same workload after {warmup} warm-up, torch {torch.__version__} CPU ops, "
                      f"{cores} threads", "ms_per_frame": dt * 1e3,
            "stage_ms": {k: v / steps * 1e3 for k, v in per_stage.items()}}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    base = run_cpu_port(args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_frame"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "stage_ms": base["stage_ms"],
            "note": "reference not installable here (mmcv/mmdet3d fork/spconv/torch_scatter absent); CPU port of its path"}
    print(json.dumps(line))
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return main_reference(args)

    import torch
    import torch.distributed as dist

    from fullysparsefusion_b200 import _capi, frame

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU port)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _capi.load()

    # three distinct frames rotate through the steps: 3 x 96 MB of inputs (+ ~1 GB of intermediates per
    # frame) exceed the 126 MB L2, so no step starts on a warm cache
    n_frames = 3
    hosts = [frame.synth_frame_host(args.points, args.sweeps, seed=rank * 16 + i) for i in range(n_frames)]
    frames = [frame.frame_to_device(h, dev, seed=i) for i, h in enumerate(hosts)]
    pipes = [frame.build_stages(f) for f in frames]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, events=None):
        stages, _ = pipes[i % n_frames]
        for name, fn in stages:
            if events is not None:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                events.append((name, a, b))
            else:
                fn()

    # ---- device-resident throughput ------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs ~0.2 s to start: begin before the warm-up so samples cover the timed region
    t_w = time.perf_counter()
    for i in range(args.warmup):
        step(i)
    while time.perf_counter() - t_w < 0.5:  # keep the GPU under load until the sampler is running
        step(0)
    barrier()
    events = []
    launches0 = _capi.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        step(i, events)
    t_end.record()
    barrier()
    launches = _capi.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_end)

    # ---- end to end from pinned host buffers ------------------------------------------------------
    def e2e_step(i):
        h = hosts[i % n_frames]
        f = frames[i % n_frames]
        f.points.copy_(h["points"], non_blocking=True)
        f.mask.copy_(h["mask"], non_blocking=True)
        f.lidar2img.copy_(h["lidar2img"], non_blocking=True)
        step(i)
        _, st = pipes[i % n_frames]
        return st["fg"].cpu(), st["pre_coors"].size(0)

    for i in range(max(3, args.warmup // 2)):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        fg_host, _ = e2e_step(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop()

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()

    if rank == 0:
        per_stage = {}
        for name, a, b in events:
            per_stage.setdefault(name, []).append(a.elapsed_time(b))
        stage_ms = {k: sum(v) / len(v) for k, v in per_stage.items()}
        _, st = pipes[0]
        alg = frame.algorithmic_bytes(frames[0], st)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (FALLBACK_HBM_GBS, "fallback")
        stage_gbs = {k: alg[k] / (stage_ms[k] * 1e-3) / 1e9 for k in stage_ms if k in alg}
        # dominant = the HBM-bound stage with the most time (scatter + projection are the named targets)
        dom = max(("vfe_scatter", "pre_voxelize", "project", "neck"), key=lambda k: stage_ms.get(k, 0.0))
        roofline = {"bound": "hbm", "kernel": dom, "achieved": stage_gbs[dom], "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": stage_gbs[dom] / peak, "traffic": None,
                    "stage_gbs": {k: round(v, 1) for k, v in stage_gbs.items()},
                    "stage_frac": {k: round(v / peak, 4) for k, v in stage_gbs.items()}}
        h2d = frames[0].nbytes()
        line = {"metric": METRIC, "value": world * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {**workload(args), "l2": "3 distinct frames rotate (288 MB of inputs > 126 MB L2)",
                           "stages": [n for n, _ in pipes[0][0]], "voxels": int(st["voxel_coors"].size(0)),
                           "pre_voxels": int(st["pre_coors"].size(0))},
                "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": int(fg_host.numel())},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}}
        if world == 1 and not args.no_cpu_baseline:
            base = run_cpu_port(args, steps=3, warmup=1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
