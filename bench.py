#!/usr/bin/env python
"""bench.py — frames/sec of the FSF sparse forward hot path on synthetic nuScenes-shaped frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--points P] [--sweeps S]

One "step" = one frame (P points x 6 cameras, default the 10-sweep 300k-point configuration the
BASELINE.json metric is quoted on) through every stage of fullysparsefusion_b200.fsf.FSF (segment, enhance, frustum, fsd, combine).  Frames are
sharded one per GPU per step (weak scaling, no data-path collective — SURVEY.md §8e).

`value`  frames/s with the frame's inputs already resident in HBM (device-timed, max over ranks).
`e2e`    frames/s through the same public API starting from pinned HOST buffers: the H2D copy of
         points / id planes / lidar2img and a D2H read of the result are inside the timed region.
`roofline`  the dominant kernel of the frame (the tcgen05 gather-GEMM of the sparse convolutions): useful flops per
            launch / its CUDA-event time inside the timed region vs the measured bf16 tensor peak (MEASURED_PEAKS.json).
`roofline_hbm`  the scatter + projection family BASELINE.json's metric names: algorithmic bytes / event time vs the
            measured HBM peak, per shape in the frame (profile pass) and at op level for 300 k and 1 M points.
`kernels`   every timed op family from a separate profile pass (per-op events cost ~3 ms per frame, so they stay
            out of the timed region except around the dominant kernel).
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=300000)
    ap.add_argument("--sweeps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="nuscenes", choices=["nuscenes", "av2"],
                    help="nuscenes: FSF_nuScenes_config (the BASELINE metric's configuration); av2: FSF_AV2_config (BASELINE configs[3]: "
                         "~107 k 4-d points, 7 ring cameras with one int32 id plane, 26 classes, the 32 x 2048 x 2048 grid)")
    ap.add_argument("--scope", default="full", choices=["hot", "full"],
                    help="full (default): FSF.simple_test = segment → combine + query refinement + final boxes (decode + rotated "
                         "NMS) on both arms; hot: segment → combine only (the scope of the round-1 numbers)")
    ap.add_argument("--train", action="store_true",
                    help="training step of the segmentation stage (fullysparsefusion_b200/train.py: forward, focal + vote loss, backward "
                         "through the tcgen05 gather-GEMM / fsfb_conv_wgrad, bucketed NCCL gradient all-reduce overlapped with backward, "
                         "AdamW) instead of the inference frame; one frame per GPU per step, synthetic labels")
    return ap.parse_args()


def workload(args):
    """The `config` of the JSON line: byte-identical for both arms (static description of the workload, nothing measured)."""
    if args.config == "av2":
        head = {"workload": f"FSF_AV2_config frame: {args.points} 4-d pts x 7 ring cams @2048x1550, one int32 id plane per camera, "
                            "26 classes, +-204.8 m range (BASELINE configs[3] shape, one frame per GPU per step)",
                "points": args.points, "sweeps": 1, "cams": 7, "classes": 26}
    else:
        head = {"workload": f"FSF_nuScenes_config {args.sweeps}-sweep frame: {args.points} pts x 6 cams @1600x900, "
                            "10 class id planes (BASELINE configs[2] shape, one frame per GPU per step)",
                "points": args.points, "sweeps": args.sweeps, "cams": 6, "classes": 10}
    return {**head, "frames_per_step_per_gpu": 1,
            "scope": "FSF.simple_test: segment, enhance, frustum, fsd, combine" + (", refine, boxes (decode + rotated NMS)"
                                                                                    if args.scope == "full" else ""),
            "l2": "3 distinct frames rotate (288 MB of inputs > 126 MB L2)"}


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


def make_model(seed: int = 0, config: str = "nuscenes"):
    """Random-init FSF (stock nuScenes architecture).  The segmentation logits are calibrated so that the
    synthetic scene behaves like a busy real one: ~5 % of voxels per class pass the 0.1 group-score
    threshold (a constant-logit random head would pass none or all of them)."""
    import torch

    from fullysparsefusion_b200 import fsf as FSFM

    torch.manual_seed(seed)
    model = FSFM.FSF(config)
    g = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():  # non-trivial eval-mode BN statistics
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
    return model


def calibrate_seg_head(model, logits) -> None:
    """Rescale conv_seg so foreground logits have sigma 1.5 around 0 and background sits at +4.7."""
    import torch

    with torch.no_grad():
        head = model.segmentation_head.conv_seg
        fgl = logits[:, :-1]
        std, mean = fgl.std(0).clamp(min=1e-3).cpu(), fgl.mean(0).cpu()
        scale = torch.cat([1.5 / std, torch.ones(1)])
        head.weight.mul_(scale[:, None].to(head.weight.device))
        bias = head.bias.cpu() * scale
        bias[:-1] -= mean * scale[:-1]
        bias[-1] = 4.7 - float(logits[:, -1].mean())
        head.bias.copy_(bias.to(head.bias.device))
    model.segmentation_head.refresh()


def synth_frame(points: int, sweeps: int, seed: int, config: str = "nuscenes"):
    import torch

    from fullysparsefusion_b200 import synth

    if config == "av2":   # 7 ring cameras, one int32 id plane each (loading.py:169-186), 4-d points + un-augmented xyz
        mask = synth.mask_planes(cams=7, classes=1, H=1550, W=2048, seed=seed, dtype="int32")
        return dict(points=torch.from_numpy(synth.av2_points(points, seed=seed)), mask=torch.from_numpy(mask),
                    anno=torch.from_numpy(synth.mask_anno(mask, seed=seed, categories=26)),
                    lidar2img=torch.from_numpy(synth.lidar2img(7, 1550, 2048)))
    mask = synth.mask_planes(seed=seed)
    return dict(points=torch.from_numpy(synth.ring_points(points, sweeps=sweeps, seed=seed)), mask=torch.from_numpy(mask),
                anno=torch.from_numpy(synth.mask_anno(mask, seed=seed)), lidar2img=torch.from_numpy(synth.lidar2img()))


def run_cpu_port(args, steps: int, warmup: int, model=None, frame=None):
    """Time the torch-CPU port of the reference path; one step = one full frame on all host cores."""
    import torch

    from oracle import fsf_torch_cpu as P

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if model is None:
        model = make_model(config=args.config)
    model = model.cpu()
    frame = frame or synth_frame(args.points, args.sweeps, seed=0, config=args.config)
    cpu = P.CpuFSF(model)
    per_stage = {}
    n_run = 0
    t_total = 0.0
    with torch.no_grad():
        for it in range(warmup + steps):
            stages, st = cpu.stages(frame["points"], frame["mask"], frame["anno"], frame["lidar2img"])
            if args.scope == "full":
                stages = stages + cpu.extra_stages
            t0 = time.perf_counter()
            for name, fn in stages:
                t = time.perf_counter()
                fn()
                if it >= warmup:
                    per_stage[name] = per_stage.get(name, 0.0) + time.perf_counter() - t
            if it == 0 and not getattr(model, "_calibrated", False):
                calibrate_seg_head(model, st["seg_logits"])
                model._calibrated = True
            if it >= warmup:
                t_total += time.perf_counter() - t0
                n_run += 1
    dt = t_total / n_run
    run_cpu_port.last_state = st          # bench.py's parity block compares the GPU frame against it
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_run} full frame(s) of the same workload after {warmup} warm-up, torch {torch.__version__} CPU ops "
                      f"+ scipy CCL, {cores} threads", "ms_per_frame": dt * 1e3,
            "stage_ms": {k: v / n_run * 1e3 for k, v in per_stage.items()}}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # one full frame costs ~8-10 s on 16 host cores: the requested --steps / --warmup are honoured up to a ~2 minute budget
    # (the run must end "within a few minutes"), the numbers actually run are the ones reported
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    base = run_cpu_port(args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": base["ms_per_frame"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "stage_ms": base["stage_ms"],
            "note": "reference not installable here (mmcv / mmdet3d fork / spconv / torch_scatter absent): the torch-CPU port of its "
                    "path (oracle/fsf_torch_cpu.py) on all host cores of ONE process, whatever --gpus says; steps capped at 8 full "
                    "frames (+ <= 2 warm-up) to stay within minutes"}
    print(json.dumps(line))
    return 0


def frame_parity(gpu_st, cpu_st):
    """The GPU frame against the CPU port on the SAME frame and weights: feature errors relative to the tensor's largest
    magnitude (north-star tolerance 1e-4 per op; a whole frame chains ~60 of them) and exact comparison of every count / index
    the two arms share.  Selection steps (score thresholds, NMS) may legitimately flip on a near-tie once features differ in the
    6th digit, so index tensors are compared by size and by mismatch count, and breaches are decided on the features."""
    import numpy as np

    def rel(a, b):
        """(largest error, 99.9th percentile of the per-row largest error), both relative to the tensor's largest magnitude: a
        row whose upstream selection flipped on a near-tie shows up in the first, the arithmetic in the second"""
        a, b = a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy()
        if a.shape != b.shape:
            return None
        scale = max(float(np.abs(b).max()), 1e-12)
        err = np.abs(a - b).reshape(a.shape[0], -1).max(1) if a.ndim > 1 else np.abs(a - b)
        return float(err.max() / scale), float(np.quantile(err, 0.999) / scale)

    pairs = {"voxel_feats": ("voxel_feats", "voxel_feats"), "seg_logits": ("seg_logits", "seg_logits"),
             "seg_feats": ("seg_feats", "seg_feats"), "frustum_obj_feats": ("frustum_obj_feats", "frustum_obj_feats"),
             "fsd_obj_feats": ("fsd_obj_feats", "fsd_obj_feats"), "obj_feats": ("obj_feats", "obj_feats"),
             "refine_cls": ("refine0_cls", "refine0_cls"), "refine_reg": ("refine0_reg", "refine0_reg"),
             "det_scores": ("det_scores", "det_scores"), "det_boxes": ("det_boxes", "det_boxes")}
    feats, shape_mismatch = {}, []
    for name, (kg, kc) in pairs.items():
        if kg in gpu_st and kc in cpu_st:
            r = rel(gpu_st[kg], cpu_st[kc])
            if r is None:
                shape_mismatch.append(f"{name}: {tuple(gpu_st[kg].shape)} vs {tuple(cpu_st[kc].shape)}")
            else:
                feats[name] = r
    # index tensors: compared as MULTISETS (one extra / missing cluster shifts every later row, so a position-wise comparison says
    # nothing); counts of queries; detections matched by label and centre
    def multiset_diff(a, b):
        a, b = np.sort(a.detach().cpu().numpy().astype(np.int64).ravel()), np.sort(b.detach().cpu().numpy().astype(np.int64).ravel())
        ua, ca = np.unique(a, return_counts=True)
        ub, cb = np.unique(b, return_counts=True)
        keys = np.union1d(ua, ub)
        fa = np.zeros(keys.size, np.int64); fa[np.searchsorted(keys, ua)] = ca
        fb = np.zeros(keys.size, np.int64); fb[np.searchsorted(keys, ub)] = cb
        return int(np.abs(fa - fb).sum())

    idx, counts = {}, {}
    for name, (kg, kc) in {"fsd_rows": ("fsd_rows", "fsd_rows"), "refine_pooled_points": ("refine0_pts_inds", "refine0_pts_inds")}.items():
        if kg in gpu_st and kc in cpu_st:
            idx[name] = {"multiset_diff": multiset_diff(gpu_st[kg], cpu_st[kc]), "of": int(cpu_st[kc].numel())}
    for name, (kg, kc) in {"voxels": ("voxel_feats", "voxel_feats"), "queries": ("obj_feats", "obj_feats"),
                           "detections": ("det_boxes", "det_boxes")}.items():
        if kg in gpu_st and kc in cpu_st:
            counts[name] = [int(gpu_st[kg].size(0)), int(cpu_st[kc].size(0))]
    det_match = None
    if "det_boxes" in gpu_st and "det_boxes" in cpu_st and cpu_st["det_boxes"].size(0) > 0:
        gb, cb = gpu_st["det_boxes"].detach().cpu().numpy(), cpu_st["det_boxes"].detach().cpu().numpy()
        gl, cl = gpu_st["det_labels"].detach().cpu().numpy(), cpu_st["det_labels"].detach().cpu().numpy()
        d = np.linalg.norm(gb[:, None, :3] - cb[None, :, :3], axis=2) + 1e3 * (gl[:, None] != cl[None, :])
        det_match = float((d.min(1) < 0.05).mean()) if gb.shape[0] else 0.0
    max_rel = max(v[0] for v in feats.values()) if feats else None
    p999 = {k: v[1] for k, v in feats.items()}
    feats = {k: v[0] for k, v in feats.items()}
    # breach: 99.9 % of the per-point / per-voxel rows beyond 5e-4 of scale (60 chained ops at <= 1e-4 each, added in
    # quadrature, stay far below), a different voxel set, a selection step that lost / gained more than 0.5 % of its rows, or
    # fewer than 97 % of the final boxes found (same label, centre within 5 cm) in the port's output
    breach = [k for k in ("voxel_feats", "seg_logits", "seg_feats") if p999.get(k, 1.0 if k in ("voxel_feats",) and counts.get("voxels", [0, 0])[0] != counts.get("voxels", [0, 0])[1] else 0) > 5e-4]
    for k, v in idx.items():
        # (the pooled points are cut at the extractor's 50 000-row capacity in roi order: one query more or less upstream moves
        #  the cut, so that multiset is held to 5 %, the uncapped selections to 0.5 %)
        if v["multiset_diff"] > (0.05 if k == "refine_pooled_points" else 0.005) * max(1, v["of"]) + 2:
            breach.append(k)
    if "queries" in counts and abs(counts["queries"][0] - counts["queries"][1]) > 0.005 * counts["queries"][1] + 2:
        breach.append("queries")
    if det_match is not None and det_match < 0.97:
        breach.append("detections")
    return {"max_rel": max_rel, "rel_by_tensor": {k: float(f"{v:.3g}") for k, v in feats.items()},
            "rel_p999_by_tensor": {k: float(f"{v:.3g}") for k, v in p999.items()}, "index_mismatches": idx, "counts_gpu_cpu": counts,
            "detections_matched": det_match, "shape_mismatch": shape_mismatch, "breach": breach,
            "against": "oracle/fsf_torch_cpu.py (torch CPU port) on frame 0 with the GPU arm's weights"}


def ncu_traffic(kernel_file: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture (profiles/r2_ncu_traffic.json,
    written by tools/ncu_table.py from the .ncu-rep); tied to the kernel source by its sha256 — a stale capture reports null."""
    import hashlib
    try:
        rec = json.load(open(os.path.join(REPO, "profiles", "r2_ncu_traffic.json")))
        digest = hashlib.sha256(open(os.path.join(REPO, "fullysparsefusion_b200", "csrc", kernel_file), "rb").read()).hexdigest()
        fresh = rec.get("source_sha256") == digest
        return (rec["dram_bytes"] if fresh else None), (rec["of"] + ("" if fresh else " — STALE: kernel source changed since the capture "
                                                                 f"({rec['dram_bytes'] / 1e6:.1f} MB then)"))
    except (OSError, KeyError, ValueError):
        return None, "no ncu capture committed for this kernel source"


def resolve(v):
    """Profiler byte/flop entries: int, or (device scalar, multiplier, constant)."""
    if isinstance(v, tuple):
        return int(v[0].item()) * v[1] + v[2]
    return v


def synth_labels(points, num_classes: int = 10):
    """Deterministic point labels for the training bench: returns above the ground inside 40 m take the class of their azimuth
    sector, everything else is background (= num_classes); vote target = offset to the sector's anchor at 20 m."""
    import math

    import torch

    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    az = torch.atan2(y, x)
    sector = ((az + math.pi) / (2 * math.pi) * num_classes).long().clamp(0, num_classes - 1)
    fg = (z > -1.3) & (x * x + y * y < 1600.0)
    labels = torch.where(fg, sector, torch.full_like(sector, num_classes))
    ang = (sector.float() + 0.5) / num_classes * 2 * math.pi - math.pi
    anchor = torch.stack([20 * torch.cos(ang), 20 * torch.sin(ang), torch.zeros_like(ang)], 1)
    return labels, (anchor - points[:, :3]).clamp(-3, 3)


def main_train(args):
    """`--train`: segmentation-stage training step, one frame per GPU per step."""
    import torch
    import torch.distributed as dist

    from fullysparsefusion_b200 import _capi
    from fullysparsefusion_b200 import dist as fdist
    from fullysparsefusion_b200.train import SegmentorTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --train: no CUDA device")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _capi.load()
    from fullysparsefusion_b200 import synth
    n_frames = 3
    frames = []
    for i in range(n_frames):
        pts = torch.from_numpy(synth.ring_points(args.points, sweeps=args.sweeps, seed=fdist.frame_seed(rank, i))).to(dev)
        lab, vote = synth_labels(pts)
        frames.append((pts, lab, vote))
    torch.manual_seed(0)   # identical initial weights on every rank
    model = make_model(config=args.config).to(dev)
    trainer = SegmentorTrainer(model)
    n_params = sum(p.numel() for p in trainer.params)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses, exposed = [], []

    def step(i):
        out = trainer.step(*frames[i % n_frames])
        return out

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step(i)
        losses.append(out["loss"])
        if trainer.reducer is not None:
            exposed.append(trainer.reducer._events)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _capi.launch_count() - launches0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    exp_ms = [a.elapsed_time(b) for a, b in exposed]
    # phase split of one more (untimed) step: forward+loss / backward / reducer wait / optimizer
    ph = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    trainer.opt.zero_grad(set_to_none=True)
    if trainer.reducer is not None:
        trainer.reducer.begin()
    ph[0].record()
    loss, _ = trainer.forward_loss(*frames[0])
    ph[1].record()
    loss.backward()
    ph[2].record()
    if trainer.reducer is not None:
        trainer.reducer.finish()
    ph[3].record()
    trainer.opt.step()
    ph[4].record()
    torch.cuda.synchronize()
    phases = {n: round(ph[j].elapsed_time(ph[j + 1]), 3) for j, n in enumerate(["forward_loss", "backward", "allreduce_wait", "optimizer"])}
    if rank == 0:
        ls = [float(v) for v in losses]
        line = {"metric": "nuScenes 10-sweep TRAIN frames/sec, segmentation stage (VoteSegmentor: VFE + sparse U-Net + seg head)",
                "value": world * args.steps / (ms / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "mode": "train",
                "config": {"workload": f"FSF_nuScenes_config {args.sweeps}-sweep frame: {args.points} pts, segmentation stage only, synthetic labels, "
                                       "one frame per GPU per step", "points": args.points, "sweeps": args.sweeps,
                           "optimizer": "AdamW", "norm": "naiveSyncBN1d (one differentiable [2C] all-reduce per layer when n_gpus > 1)"},
                "parameters": n_params, "grad_bytes_per_step": 4 * n_params,
                "buckets": len(trainer.reducer.buckets) if trainer.reducer is not None else 0,
                "allreduce_exposed_ms": round(sum(exp_ms) / len(exp_ms), 3) if exp_ms else 0.0,
                "allreduce_share_of_step": round(sum(exp_ms) / ms, 4) if exp_ms else 0.0,
                "phases_ms": phases, "loss_first_last": [ls[0], ls[-1]], "gpu_launches": int(launches), "clocks": clocks}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return main_reference(args)
    if args.train:
        return main_train(args)

    import torch
    import torch.distributed as dist

    from fullysparsefusion_b200 import _capi, ops
    from fullysparsefusion_b200 import dist as fdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU port)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # rank 0 prints exactly ONE JSON line on stdout: native libraries (NCCL's version banner at NCCL_DEBUG=VERSION/WARN, ...)
    # write to file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the line goes to the saved fd
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _capi.load()

    # three distinct frames rotate through the steps: 3 x 96 MB of inputs plus > 1 GB of per-frame
    # intermediates exceed the 126 MB L2, so no step starts on a warm cache
    n_frames = 3
    hosts = [{k: v.pin_memory() for k, v in synth_frame(args.points, args.sweeps, seed=fdist.frame_seed(rank, i), config=args.config).items()}
             for i in range(n_frames)]
    frames = [{k: v.to(dev) for k, v in h.items()} for h in hosts]
    model = make_model(config=args.config).to(dev)
    with torch.no_grad():
        st0 = model(frames[0]["points"], frames[0]["mask"], frames[0]["anno"], frames[0]["lidar2img"])
        calibrate_seg_head(model, st0["seg_logits"])
        del st0
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}
    rc_final = 0

    def overflow_launches():
        import ctypes
        cnt = ctypes.c_uint(0)
        _capi.check(_capi.load().fsfb_gemm_f16_overflows(ctypes.byref(cnt)), "fsfb_gemm_f16_overflows")
        return int(cnt.value)

    def step(i, events=None, scope=None, frame=None):
        f = frame if frame is not None else frames[i % n_frames]
        stages, st = model.stages(f["points"], f["mask"], f["anno"], f["lidar2img"])
        if (scope or args.scope) == "full":   # FSF.simple_test's tail (FSF.py:1158-1171; frustum_cluster_head.py:595-698)
            stages = stages + [("refine", lambda: model.refine(st, f["points"])), ("boxes", lambda: model.get_bboxes(st))]
        for name, fn in stages:
            if events is not None:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                events.append((name, a, b))
            else:
                fn()
        last["st"] = st
        return st

    # ---- device-resident throughput ------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi needs ~0.2 s to start: begin before the warm-up so samples cover the timed region
    t_w = time.perf_counter()
    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        while time.perf_counter() - t_w < 0.5:
            step(0)
        barrier()
        ops.PROFILER, ops.PROFILE_ONLY = [], {"gather_gemm_conv"}   # events only around the dominant kernel
        launches0 = _capi.launch_count()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_start.record()
        for i in range(args.steps):
            step(i)
        t_end.record()
        barrier()
        prof_dom, ops.PROFILER, ops.PROFILE_ONLY = ops.PROFILER, None, None
        launches = _capi.launch_count() - launches0
        ms_total = t_start.elapsed_time(t_end)
        # the round-1 scope (segment → combine) timed the same way, for continuity with BENCH_r01 (not the headline)
        ms_hot = None
        if args.scope == "full":
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            h0.record()
            for i in range(args.steps):
                step(i, scope="hot")
            h1.record()
            barrier()
            ms_hot = h0.elapsed_time(h1)

        # ---- profile pass (not part of `value`): per-stage and per-op events ---------------------------
        events = []
        ops.PROFILER = []
        n_prof = min(args.steps, 10)
        for i in range(n_prof):
            step(i, events)
        torch.cuda.synchronize()
        prof, ops.PROFILER = ops.PROFILER, None

        # ---- end to end from pinned host buffers --------------------------------------------------
        # through the library's own loader front end (loading.FrameStager): frame i + 1 is uploaded from its pinned slot on
        # the copy stream while frame i computes; every step still uploads its own inputs and reads its result back
        from fullysparsefusion_b200.loading import FrameStager
        stager = FrameStager(dev, slots=n_frames)
        slot_host = []
        for s_i in range(n_frames):     # the decoded frames live in the stager's pinned slots (a loader decodes straight into them)
            bufs = {k: stager.host_buffer(k, tuple(v.shape), v.dtype) for k, v in hosts[s_i].items()}
            for k, v in hosts[s_i].items():
                bufs[k].copy_(v)
            slot_host.append(bufs)
            stager.put(bufs["points"], bufs["mask"], bufs["anno"], bufs["lidar2img"])
            stager.release(stager.get())
        torch.cuda.synchronize()
        state = {"next": 0}

        def e2e_put():
            h = slot_host[state["next"] % n_frames]
            stager.put(h["points"], h["mask"], h["anno"], h["lidar2img"])
            state["next"] += 1

        e2e_put()

        def e2e_step(i):
            e2e_put()                    # upload of the NEXT frame: overlaps this frame's kernels
            f = stager.get()
            st = step(i, frame=f)
            if args.scope == "full":     # the frame's result: final boxes, scores, labels (D2H)
                res = st["det_boxes"].cpu(), st["det_scores"].cpu(), st["det_labels"].cpu()
            else:
                res = st["obj_cls"].cpu(), st["obj_reg"].cpu(), st["obj_centers"].cpu()
            stager.release(f)
            return res

        for i in range(3):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            res = e2e_step(i)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop()

    ms_total, ms_e2e, ms_hot = fdist.max_over_ranks(torch.tensor([ms_total, ms_e2e, ms_hot if ms_hot is not None else 0.0],
                                                                 dtype=torch.float64, device=dev)).tolist()

    if rank == 0:
        per_stage = {}
        for name, a, b in events:
            per_stage.setdefault(name, []).append(a.elapsed_time(b))
        stage_ms = {k: statistics.median(v) for k, v in per_stage.items()}   # median: one slow frame must not skew a stage
        def fold(records, n_steps):
            kern, shapes = {}, {}
            for name, a, b, nb, fl in records:
                ms, nb, fl = a.elapsed_time(b), resolve(nb), resolve(fl)
                for table, key in ((kern, name.split("[")[0]), (shapes, name)):
                    k = table.setdefault(key, dict(ms=0.0, bytes=0, flops=0, calls=0))
                    k["ms"] += ms
                    k["bytes"] += nb
                    k["flops"] += fl
                    k["calls"] += 1
            return kern, shapes

        kern, shapes = fold(prof, n_prof)
        dom_kern, _ = fold(prof_dom, args.steps)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (FALLBACK_HBM_GBS, "fallback")
        tpeak, tpeak_src = (peaks["bf16_tflops_sustained"], "measured (bf16 dense, sustained)") if "bf16_tflops_sustained" in peaks \
            else (1400.0, "fallback")
        table = {}
        for name, k in kern.items():
            sec = k["ms"] * 1e-3
            table[name] = {"ms_per_frame": round(k["ms"] / n_prof, 4), "calls_per_frame": k["calls"] // n_prof,
                           "GB/s": round(k["bytes"] / sec / 1e9, 1), "hbm_frac": round(k["bytes"] / sec / 1e9 / peak, 4)}
            if k["flops"]:
                table[name]["TFLOP/s"] = round(k["flops"] / sec / 1e12, 2)
                table[name]["tensor_frac_of_bf16_sustained"] = round(k["flops"] / sec / 1e12 / tpeak, 4)
        # dominant kernel: the gather-GEMM launches of the sparse convolutions, timed inside the timed region
        d = dom_kern["gather_gemm_conv"]
        ach_t = d["flops"] / (d["ms"] * 1e-3) / 1e12
        traffic, traffic_of = ncu_traffic("gemm_ss.cu")
        roofline = {"bound": "tensor", "kernel": "k_gather_gemm_ss (sparse-convolution gather-GEMM, fp16-split kind::f16 on tcgen05)",
                    "achieved": ach_t, "peak": tpeak, "peak_source": tpeak_src, "unit": "TFLOP/s", "frac": ach_t / tpeak,
                    # dram bytes of ONE launch (SubM 27x128->128 on the frame's level-0 voxels) from the committed ncu capture,
                    # null when the kernel source changed since; `achieved` above averages all 34 convolution shapes of the frame
                    "traffic": traffic, "traffic_of": traffic_of,
                    "launches_timed": d["calls"], "launches_per_frame": d["calls"] // args.steps,
                    "flops_per_launch": d["flops"] // d["calls"], "us_per_launch": round(d["ms"] / d["calls"] * 1e3, 2),
                    "share_of_step": round(d["ms"] / ms_total, 3),
                    "note": "achieved = USEFUL flops (2*Cin*Cout per rulebook pair) / CUDA-event time of every launch in the timed "
                            "region; the kernel executes 3 fp16 MMAs per useful product (the fp16 split keeps fp32 parity) plus zero "
                            "rows of partially filled tiles, so executed tensor work is >= 3x this figure (DESIGN.md section 4)"}
        # the scatter + projection family (BASELINE metric): per shape in the frame + op level at 300 k / 1 M points
        cand = {n: v for n, v in shapes.items() if n.split("[")[0] in ("segment_reduce", "project_sample_select", "gather_rows")}
        dom = max(cand, key=lambda n: cand[n]["ms"])
        dd = cand[dom]
        ach = dd["bytes"] / (dd["ms"] * 1e-3) / 1e9
        roofline_hbm = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                        "frac": ach / peak, "bytes_per_launch": dd["bytes"] // dd["calls"],
                        "traffic": None,
                        "us_per_launch": round(dd["ms"] / dd["calls"] * 1e3, 2),
                        "shapes": {n: {"us_per_launch": round(v["ms"] / v["calls"] * 1e3, 1), "launches_per_frame": v["calls"] // n_prof,
                                       "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 3)}
                                   for n, v in sorted(cand.items(), key=lambda kv: -kv[1]["ms"])[:8]},
                        "family_frac": {f: round(kern[f]["bytes"] / (kern[f]["ms"] * 1e-3) / 1e9 / peak, 4)
                                        for f in ("segment_reduce", "gather_rows", "project_sample_select") if f in kern},
                        "note": "algorithmic bytes (SURVEY.md 8d) / CUDA-event time from the profile pass; inside the frame the "
                                "events also see launch gaps of small shapes, `ops` below times each op back to back"}
        if world == 1:
            sys.path.insert(0, os.path.join(REPO, "tools"))
            import op_bench
            roofline_hbm["ops"] = [op_bench.run(p, peak, dev) for p in (300000, 1000000)]
        st = last["st"]
        h2d = sum(v.numel() * v.element_size() for v in hosts[0].values())
        d2h = sum(r.numel() * r.element_size() for r in res)
        line = {"metric": METRIC, "value": fdist.throughput(args.steps, world, ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (tensor-core GEMMs: fp16-split operands, 22 mantissa bits, fp32 accumulate)",
                "data": "synthetic",
                "config": workload(args),
                "frame_stats": {"stages": list(stage_ms), "voxels": int(st["voxel_coors"].size(0)),
                                "pre_voxels": int(st["pre_coors"].size(0)), "frustum_rows": int(st["frustum_rows"].numel()),
                                "frustum_queries": int(st["frustum_obj_coors"].size(0)), "fsd_rows": int(st["fsd_rows"].numel()),
                                "fsd_queries": int(st["fsd_obj_coors"].size(0)),
                                "detections": int(st["det_boxes"].size(0)) if "det_boxes" in st else None,
                                "f16_overflow_launches": overflow_launches()},
                "e2e": {"value": fdist.throughput(args.steps, world, ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h,
                        "note": "through loading.FrameStager (the library's loader front end): every step uploads its frame from a pinned "
                                "slot on a copy stream — overlapping the previous frame's kernels — and reads its boxes back; "
                                "starts from DECODED pinned host buffers (points, uint8 id planes, annotation table, matrices); the "
                                "wire format in front of it — 60 PNG planes per frame — costs ~0.38 core-seconds of inflate per frame "
                                "(DESIGN.md section 6: ~12 frames/s on 8 decode threads), so from disk the decode, not this path, "
                                "bounds the rate"},
                "hot_scope": ({"value": fdist.throughput(args.steps, world, ms_hot * 1e-3), "ms_per_step": ms_hot / args.steps,
                               "scope": "segment..combine (the scope of the round-1 bench lines)"} if ms_hot else None),
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm, "kernels": table,
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}}
        if world == 1 and not args.no_cpu_baseline:
            import copy

            cpu_model = copy.deepcopy(model).cpu()
            cpu_model._calibrated = True  # same calibrated weights as the GPU arm
            base = run_cpu_port(args, steps=1, warmup=0, model=cpu_model, frame={k: v.clone() for k, v in hosts[0].items()})
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # ---- does the timed frame compute what the reference path computes?  frame 0 on both arms, same weights ----
            with torch.no_grad():
                gst = step(0)
                torch.cuda.synchronize()
            line["parity"] = frame_parity(gst, run_cpu_port.last_state)
        if world == 1:
            try:
                sys.path.insert(0, os.path.join(REPO, "tools"))
                import library_bar
                line["library_baseline"] = library_bar.run(args.points, dev)
            except Exception as e:  # the bar is context, never a reason to lose the bench line
                line["library_baseline"] = {"error": repr(e)[:300]}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
        if line.get("parity", {}).get("breach"):
            sys.stderr.write("bench.py: PARITY BREACH against the CPU port: %s\n" % line["parity"])
            rc_final = 3
    if world > 1:
        dist.destroy_process_group()
    return rc_final


if __name__ == "__main__":
    sys.exit(main())
