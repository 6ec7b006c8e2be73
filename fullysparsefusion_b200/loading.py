"""Input side of the hot path: the instance-mask sample on disk → the tensors the frame consumes.

Mirrors `LoadMaskFromFiles` (projects/mmdet3d_plugin/datasets/pipelines/loading.py:22-339): same constructor, same
`__call__(results)` contract (`results['mask_data']`, `results['mask_anno']`, the in-place `results['lidar2img']` rescale of the
resized cameras), same on-disk format (`{cam}_{class}.png` + `anno.json`, written by tools/mask_tools/save_mask_nusc.py:138-171).
What differs is the memory plan:

* the planes are decoded by a thread pool (cv2 releases the GIL) straight into ONE preallocated host buffer
  `[cams, classes, H, W]` — pinned when a `FrameStager` owns it — instead of 60 tensors + `torch.stack`;
* they stay uint8 end to end (upstream casts them to float32 on the device, 345.6 MB written per projection call; the
  projection kernel of this repo samples the u8 planes directly);
* nearest-neighbour resizing of the odd cameras (AV2 front, Waymo back pair) is an index gather during that copy, with ATen's
  `upsample_nearest` source index (`min(floor(dst * in/out), in - 1)` in float32) so the planes are bit-identical;
* `FrameStager` rotates pinned host slots and device slots and issues the H2D copies on its own stream, so frame i+1 uploads
  while frame i computes.

`write_mask_sample` is the writer half of the format (what save_result_format stores), used by the tests and tools.
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

NUSC_CLASSES = ["car", "truck", "trailer", "bus", "construction_vehicle", "bicycle", "motorcycle", "pedestrian", "traffic_cone",
                "barrier"]
WAYMO_CLASSES = ["vehicle", "pedestrian", "cyclist"]


def _imread(path: str) -> np.ndarray:
    import cv2

    img = cv2.imread(path, -1)                      # IMREAD_UNCHANGED, as upstream (:153, :172, :219)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim != 2:
        raise ValueError(f"{path}: expected a single-channel id plane, got shape {img.shape}")
    return img


def nearest_index(n_in: int, n_out: int) -> np.ndarray:
    """Source index of every output row/column for a nearest-neighbour resize, as ATen's upsample_nearest computes it
    (what torchvision's `resize(..., InterpolationMode.NEAREST)` runs, loading.py:62, :96, :125)."""
    dst = np.arange(n_out, dtype=np.int64)
    if n_out == n_in:
        return dst
    if n_out == 2 * n_in:
        return dst >> 1
    scale = np.float32(n_in) / np.float32(n_out)
    return np.minimum(np.floor(dst.astype(np.float32) * scale).astype(np.int64), n_in - 1)


def _place(dst: np.ndarray, img: np.ndarray) -> None:
    """dst[...] = img, resized with nearest neighbour when the shapes differ."""
    if img.shape == dst.shape:
        np.copyto(dst, img, casting="unsafe")
    else:
        iy, ix = nearest_index(img.shape[0], dst.shape[0]), nearest_index(img.shape[1], dst.shape[1])
        np.copyto(dst, img[iy[:, None], ix[None, :]], casting="unsafe")


class LoadMaskFromFiles:
    """Drop-in for the pipeline step of the same name (loading.py:22-339).

    results in : 'sample_idx' (nuScenes) | 'pts_filename' (Waymo) | 'img_info'['uuid'] (AV2); 'lidar2img' (list of 4x4, only
                 touched for the resized cameras).
    results out: 'mask_data' torch [cams, classes, H, W] (u8; AV2 int32 [7,1,H,W] as upstream's astype(np.int32)),
                 'mask_anno' torch f32 [obj_max_num, 9] = (x1,y1,x2,y2,score,category,cam_id,obj_id,valid) sorted by obj_id.
    `out`: optional preallocated destination for mask_data (a numpy view of a pinned tensor); `workers`: decode threads.
    `layout="hwc16"` (u8 samples only, EXPERIMENTAL consumer `ops.project_sample_select_hwc`): the planes are written
    class-interleaved, `mask_data[cam, y, x, k]` = plane of class k, padded to 16 bytes per texel (pad bytes zero) — the sampling
    kernel then reads the ids of a texel with one 16-byte load instead of one sector per class."""

    def __init__(self, data_path, class_names=None, obj_max_num=250, is_argo=False, is_waymo=False, workers: int = 8,
                 layout: str = "chw"):
        if layout not in ("chw", "hwc16"):
            raise ValueError("layout must be 'chw' (upstream's [cams, classes, H, W]) or 'hwc16' ([cams, H, W, 16] u8)")
        self.layout = layout
        self.data_path = data_path
        self.obj_max_num = obj_max_num
        self.class_names = list(NUSC_CLASSES if class_names is None else class_names)
        self.is_argo = is_argo
        self.is_waymo = is_waymo
        self.workers = max(1, int(workers))
        self._pool: Optional[ThreadPoolExecutor] = None
        self._cleared = set()

    # ---- layout of one sample ------------------------------------------------------------------------------------------
    def _plan(self, results) -> Tuple[str, List[str], Tuple[int, int], Dict[int, Tuple[int, int]], np.dtype]:
        """(sample_dir, plane file names in [cam, class] order, (cams, classes), {cam: resize shape}, dtype)"""
        if self.is_argo:       # load_argo (:167-185): one plane per camera, front camera resized to 1550 x 2048
            sample = results["img_info"]["uuid"]
            return sample, [f"{c}.png" for c in range(7)], (7, 1), {0: (1550, 2048)}, np.dtype(np.int32)
        if self.is_waymo:      # load_waymo (:140-165): 5 cameras x 3 classes, the two back cameras resized to 1280 x 1920
            sample = results["pts_filename"].split("/")[-1].replace(".bin", "")
            names = [f"{c}_{n}.png" for c in range(5) for n in WAYMO_CLASSES]
            return sample, names, (5, len(WAYMO_CLASSES)), {3: (1280, 1920), 4: (1280, 1920)}, np.dtype(np.uint8)
        sample = results["sample_idx"]   # load_nusc (:208-230)
        names = [f"{c}_{n}.png" for c in range(6) for n in self.class_names]
        return sample, names, (6, len(self.class_names)), {}, np.dtype(np.uint8)

    def __call__(self, results, out: Optional[np.ndarray] = None):
        sample, names, (cams, classes), resized, dtype = self._plan(results)
        sample_dir = os.path.join(self.data_path, sample)
        paths = [os.path.join(sample_dir, n) for n in names]
        with open(os.path.join(sample_dir, "anno.json"), "r") as f:
            anno = json.load(f)

        # the frame's plane shape: a camera that is not resized keeps its own; resized cameras take the target
        probe_cam = next(c for c in range(cams) if c not in resized)
        first = _imread(paths[probe_cam * classes])
        H, W = first.shape
        for cam, shape in resized.items():
            if tuple(shape) != (H, W):
                raise ValueError(f"camera {cam} resizes to {shape} but camera {probe_cam} is {(H, W)}: planes cannot be stacked")
        hwc = self.layout == "hwc16"
        if hwc and (dtype != np.uint8 or classes > 16):
            raise ValueError("layout='hwc16' needs uint8 planes and at most 16 classes")
        shape = (cams, H, W, 16) if hwc else (cams, classes, H, W)
        if out is None:
            out = np.zeros(shape, dtype=dtype) if hwc else np.empty(shape, dtype=dtype)
        elif out.shape != shape or out.dtype != dtype:
            raise ValueError(f"out must be {dtype} {shape}, got {out.dtype} {out.shape}")
        elif hwc and classes < 16:
            addr = (out.__array_interface__["data"][0], shape)
            if addr not in self._cleared:        # pad bytes of a caller's buffer are cleared once: nothing here writes them again
                out[..., classes:] = 0
                self._cleared.add(addr)
        if dtype == np.uint8 and first.dtype != np.uint8:
            raise ValueError(f"{paths[probe_cam * classes]}: {first.dtype} plane where uint8 ids are expected")

        ori_shape: Dict[int, Tuple[int, int]] = {}

        def job(p: int):
            cam, cls = divmod(p, classes)
            img = first if p == probe_cam * classes else _imread(paths[p])
            if cam in resized and cls == 0:
                ori_shape[cam] = img.shape
            elif cam not in resized and img.shape != (H, W):
                raise ValueError(f"{paths[p]}: shape {img.shape} differs from {(H, W)}")
            _place(out[cam, :, :, cls] if hwc else out[cam, cls], img)

        if self.workers == 1:
            for p in range(len(paths)):
                job(p)
        else:
            if self._pool is None:
                self._pool = ThreadPoolExecutor(self.workers, thread_name_prefix="fsfb-mask")
            list(self._pool.map(job, range(len(paths))))

        # resized cameras: projection rows and boxes follow the planes (resize_img :47-75, resize_img_waymo :109-137)
        for cam in sorted(resized):
            hf, wf = H / ori_shape[cam][0], W / ori_shape[cam][1]
            l2i = np.asarray(results["lidar2img"][cam])      # in place when it already is an array, as upstream
            l2i[0] *= wf
            l2i[1] *= hf
            results["lidar2img"][cam] = l2i
            objs = anno[cam] if self.is_argo else [o for v in anno[cam].values() for o in v]
            for o in objs:
                b = o["bbox"]
                o["bbox"] = [b[0] * wf, b[1] * hf, b[2] * wf, b[3] * hf]

        results["mask_anno"] = self.reorg_anno_single_cls(anno) if self.is_argo else self.reorg_anno_multi_cls(anno)
        results["mask_data"] = torch.from_numpy(out)
        return results

    # ---- annotations ------------------------------------------------------------------------------------------------------
    def _table(self, rows: List[List[float]]) -> torch.Tensor:
        """rows of (x1,y1,x2,y2,score,category,cam_id,obj_id) → [obj_max_num, 9] f32, zero padded, last column = valid."""
        if len(rows) > self.obj_max_num:
            raise ValueError(f"{len(rows)} objects exceed obj_max_num={self.obj_max_num}")
        table = torch.zeros((self.obj_max_num, 9), dtype=torch.float32)
        if rows:
            table[: len(rows), :8] = torch.tensor(rows, dtype=torch.float64).to(torch.float32)
            table[: len(rows), 8] = 1.0
        return table

    @staticmethod
    def _row(o) -> List[float]:
        return list(o["bbox"][:4]) + [o["score"], o["category"], o["cam_id"], o["obj_id"]]

    def reorg_anno_single_cls(self, annos) -> torch.Tensor:
        """loading.py:273-299: per camera a list of objects, kept in file order."""
        return self._table([self._row(o) for cam in annos for o in cam])

    def reorg_anno_multi_cls(self, annos) -> torch.Tensor:
        """loading.py:301-339: per camera {class: [objects]}, rows sorted by obj_id."""
        rows = [self._row(o) for cam in annos for objs in cam.values() for o in objs]
        rows.sort(key=lambda r: r[7])
        return self._table(rows)


class SaveNoAugPoints:
    """Pipeline step of the same name (loading.py:342-354): appends a copy of xyz to every point, `[N,C]` → `[N,C+3]`, so the
    un-augmented coordinates reach `FSF.split_points_last_3dim` (FSF.py:1123); with ground truth present it also keeps
    `no_aug_gt_bboxes_3d` / `no_aug_gt_labels_3d`.  `results['points']` may be a tensor or an object with a `.tensor`."""

    def __call__(self, results):
        pts = results["points"]
        t = pts.tensor if hasattr(pts, "tensor") else pts
        out = torch.cat([t, t[:, :3].clone()], dim=-1)
        if hasattr(pts, "tensor"):
            pts.tensor = out
        else:
            results["points"] = out
        if "gt_bboxes_3d" in results:
            results["no_aug_gt_bboxes_3d"] = results["gt_bboxes_3d"].clone()
            results["no_aug_gt_labels_3d"] = torch.from_numpy(np.asarray(results["gt_labels_3d"]))
        return results


def write_mask_sample(sample_dir: str, mask: np.ndarray, anno_rows: np.ndarray, class_names: Optional[Sequence[str]] = None,
                      single_cls: bool = False) -> None:
    """Store id planes [cams, classes, H, W] and annotation rows [(x1,y1,x2,y2,score,category,cam_id,obj_id,valid)] in the
    sample format save_result_format writes (tools/mask_tools/save_mask_nusc.py:138-171): `{cam}_{class}.png` (u8; `{cam}.png`
    16-bit when single_cls, the AV2 layout) + `anno.json` = per camera {class name: [object dicts]} (a flat list when
    single_cls)."""
    import cv2

    cams, classes, H, W = mask.shape
    names = list(NUSC_CLASSES if class_names is None else class_names)
    os.makedirs(sample_dir, exist_ok=True)
    anno: list = [[] if single_cls else {n: [] for n in names[:classes]} for _ in range(cams)]
    for r in np.asarray(anno_rows):
        if len(r) > 8 and not r[8]:
            continue
        obj = {"bbox": [float(v) for v in r[:4]], "score": float(r[4]), "category": int(r[5]), "cam_id": int(r[6]), "obj_id": int(r[7])}
        if single_cls:
            anno[obj["cam_id"]].append(obj)
        else:
            anno[obj["cam_id"]][names[obj["category"]]].append(obj)
    with open(os.path.join(sample_dir, "anno.json"), "w") as f:
        json.dump(anno, f, indent=2)
    for cam in range(cams):
        for cls in range(classes):
            if single_cls:
                ok = cv2.imwrite(os.path.join(sample_dir, f"{cam}.png"), mask[cam, cls].astype(np.uint16))
            else:
                assert mask[cam, cls].max() < 255, "for uint8"
                ok = cv2.imwrite(os.path.join(sample_dir, f"{cam}_{names[cls]}.png"), mask[cam, cls].astype(np.uint8))
            if not ok:
                raise OSError(f"could not write plane {cam}/{cls} under {sample_dir}")


class FrameStager:
    """Rotating pinned-host / device slots for the per-frame inputs (points, mask planes, annotation table, lidar2img) with
    the uploads on a dedicated stream.

        stager = FrameStager(device, slots=2)
        stager.put(points, mask_u8, anno, lidar2img)      # host tensors / arrays → pinned slot → async H2D
        frame = stager.get()                               # device tensors; the compute stream waits on the copy event

    `lidar2img` is assembled ONCE per frame into a [cams,4,4] f32 tensor (upstream rebuilds it from python lists inside every
    frustum_gather call, FSF.py:248-252, three times per frame).  On a CPU device the class degrades to plain copies (used by
    the CPU tests); nothing here touches the oracle."""

    KEYS = ("points", "mask", "anno", "lidar2img")

    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.slots = max(1, int(slots))
        self._host: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.slots)]
        self._dev: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.slots)]
        self._ready: List[Optional["torch.cuda.Event"]] = [None] * self.slots
        self._free: List[Optional["torch.cuda.Event"]] = [None] * self.slots
        self._stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._unreleased: List[bool] = [False] * self.slots
        self._fetched_stream = None
        self._head = 0          # next slot to fill
        self._queued: List[int] = []
        self._shapes: List[Optional[dict]] = [None] * self.slots
        self.h2d_bytes = 0

    def _buf(self, table, slot, key, shape, dtype, pinned):
        t = table[slot].get(key)
        if t is None or t.dtype != dtype or t.numel() < int(np.prod(shape)):
            n = int(np.prod(shape))
            if pinned:
                t = torch.empty(n, dtype=dtype, pin_memory=self.cuda)
            else:
                t = torch.empty(n, dtype=dtype, device=self.device)
            table[slot][key] = t
        return t[: int(np.prod(shape))].view(*shape)

    def host_buffer(self, key: str, shape, dtype) -> torch.Tensor:
        """The pinned destination of the NEXT put() for `key`: decode straight into it (`LoadMaskFromFiles(..., out=
        stager.host_buffer('mask', shape, torch.uint8).numpy())`) and pass the same tensor to put() — no staging copy."""
        self._wait_free(self._head)
        return self._buf(self._host, self._head, key, tuple(shape), dtype, pinned=True)

    def _wait_free(self, slot):
        if self.cuda and self._unreleased[slot]:
            # the frame fetched from this slot was never release()d: the only safe free point is "everything queued so far"
            self._fetched_stream.synchronize()
            self._unreleased[slot] = False
        ev = self._free[slot]
        if ev is not None:
            ev.synchronize()      # the frame that last used this slot has been consumed
            self._free[slot] = None
        ev = self._ready[slot]
        if ev is not None:
            ev.synchronize()      # and its upload no longer reads the pinned slot
            self._ready[slot] = None

    def put(self, points, mask, anno, lidar2img) -> None:
        if len(self._queued) == self.slots:
            raise RuntimeError("all slots hold frames that were not fetched with get()")
        slot = self._head
        self._wait_free(slot)
        src = {"points": points, "mask": mask, "anno": anno,
               "lidar2img": np.asarray([np.asarray(m, dtype=np.float32) for m in lidar2img], dtype=np.float32)
               if not torch.is_tensor(lidar2img) else lidar2img.to(torch.float32)}
        staged = {}
        for key in self.KEYS:
            t = src[key]
            t = torch.from_numpy(np.ascontiguousarray(t)) if not torch.is_tensor(t) else t.contiguous()
            h = self._buf(self._host, slot, key, tuple(t.shape), t.dtype, pinned=True)
            if h.data_ptr() != t.data_ptr():
                h.copy_(t)
            staged[key] = h
        # device slots are allocated on the caller's stream (never inside the copy stream's context: the caching allocator
        # would tie the blocks to that stream)
        devs = {key: self._buf(self._dev, slot, key, tuple(h.shape), h.dtype, pinned=False) for key, h in staged.items()}
        if self.cuda:
            # a (re)allocated block may still be in use by work queued on the caller's stream: the copy stream starts after it
            self._stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._stream):
                for key, h in staged.items():
                    devs[key].copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            self._ready[slot] = ev
        else:
            for key, h in staged.items():
                devs[key].copy_(h)
        self.h2d_bytes += sum(h.numel() * h.element_size() for h in staged.values())
        self._shapes[slot] = {k: (tuple(h.shape), h.dtype) for k, h in staged.items()}
        self._queued.append(slot)
        self._head = (slot + 1) % self.slots

    def get(self) -> Dict[str, torch.Tensor]:
        """Oldest staged frame as device tensors (views of the slot: valid until `slots` further put() calls)."""
        if not self._queued:
            raise RuntimeError("get() without a staged frame")
        slot = self._queued.pop(0)
        if self.cuda and self._ready[slot] is not None:        # None: the host already waited for this upload
            torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        frame = {k: self._buf(self._dev, slot, k, shape, dtype, pinned=False) for k, (shape, dtype) in self._shapes[slot].items()}
        frame["_slot"] = slot
        if self.cuda:
            # default free point for callers that never call release(): everything queued on the current stream up to this
            # get() — release() moves it behind the frame's own consumers (call it; without it the slot is only protected
            # against work queued BEFORE the frame was fetched, and the next put() into it waits on the stream below)
            self._fetched_stream = torch.cuda.current_stream(self.device)
            self._unreleased[slot] = True
        return frame

    def release(self, frame: Dict[str, torch.Tensor]) -> None:
        """Mark the frame's slot reusable once the work queued so far on the current stream has finished."""
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._free[frame["_slot"]] = ev
            self._unreleased[frame["_slot"]] = False
