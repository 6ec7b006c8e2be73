"""Seeded synthetic nuScenes-/AV2-shaped inputs (numpy only; no CUDA, no oracle).

Shapes and value ranges follow the reference's data layer:
  points  [N,8] f32 = x,y,z,intensity,dt + un-augmented xyz (SaveNoAugPoints,
          projects/mmdet3d_plugin/datasets/pipelines/loading.py:342-354), clipped to the
          PointsRangeFilter box of projects/configs/_base_/datasets/nuscenes_dataloader.py:15;
  mask_data u8 [cams,classes,H,W] instance-id planes, ids 1..250 increasing over (cam, class)
          as tools/mask_tools/save_mask_nusc.py:142-156 paints them (assert max < 255, :169);
  mask_anno [250,9] f32 = x1,y1,x2,y2,score,category,cam_id,obj_id,valid (loading.py:301-339);
  lidar2img [cams,4,4] f32 = K [R|t].
Used by tests/, tools/make_golden.py and bench.py so that every leg sees identical inputs.
"""
from __future__ import annotations

import numpy as np

NUSC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)
NUSC_CLIP = (-50.0, -50.0, -4.99, 50.0, 50.0, 2.99)
NUSC_VOXEL = (0.2, 0.2, 0.2)
AV2_RANGE = (-204.8, -204.8, -3.2, 204.8, 204.8, 3.2)
AV2_VOXEL = (0.2, 0.2, 0.2)


def ring_points(n: int, sweeps: int = 1, seed: int = 0, clip=NUSC_CLIP, beams: int = 32,
                n_boxes: int = 48, max_range: float = 60.0) -> np.ndarray:
    """[n,8] f32 spinning-LiDAR-like sweep(s): ground returns, far walls and box surfaces."""
    rng = np.random.default_rng(seed)
    per = int(np.ceil(n / sweeps))
    az_steps = int(np.ceil(per / beams))
    elev = np.deg2rad(np.linspace(-30.0, 10.0, beams)).astype(np.float64)
    # objects: car-sized boxes on the ground, shared by all sweeps
    bx = rng.uniform(-40, 40, n_boxes)
    by = rng.uniform(-40, 40, n_boxes)
    bw = rng.uniform(1.6, 2.4, n_boxes)
    bl = rng.uniform(3.5, 6.0, n_boxes)
    bh = rng.uniform(1.4, 2.2, n_boxes)
    byaw = rng.uniform(-np.pi, np.pi, n_boxes)
    out = []
    for s in range(sweeps):
        az = (np.arange(az_steps) + rng.uniform(0, 1)) * (2 * np.pi / az_steps)
        el, azg = np.meshgrid(elev, az, indexing="ij")
        el = el.ravel()
        azg = azg.ravel()
        with np.errstate(divide="ignore"):
            r_ground = np.where(el < -1e-3, 1.8 / -np.sin(el), np.inf)
        r_wall = rng.uniform(25.0, max_range, el.shape)
        r = np.minimum(r_ground, r_wall) * (1 + rng.normal(0, 0.002, el.shape))
        x = r * np.cos(el) * np.cos(azg)
        y = r * np.cos(el) * np.sin(azg)
        z = r * np.sin(el)
        # ~18 % of returns are replaced by points on box surfaces (dense object clusters)
        k = rng.random(el.shape) < 0.18
        nb = int(k.sum())
        bi = rng.integers(0, n_boxes, nb)
        u = rng.uniform(-0.5, 0.5, (nb, 3))
        face = rng.integers(0, 3, nb)
        u[np.arange(nb), face] = np.where(rng.random(nb) < 0.5, -0.5, 0.5)
        lx, ly, lz = u[:, 0] * bl[bi], u[:, 1] * bw[bi], (u[:, 2] + 0.5) * bh[bi] - 1.8
        c, sn = np.cos(byaw[bi]), np.sin(byaw[bi])
        x[k] = bx[bi] + c * lx - sn * ly
        y[k] = by[bi] + sn * lx + c * ly
        z[k] = lz
        x = x + 0.5 * s  # ego motion between sweeps
        pts = np.stack([x, y, z, rng.uniform(0, 1, el.shape), np.full(el.shape, 0.05 * s)], 1)
        out.append(pts)
    pts = np.concatenate(out, 0)
    lo, hi = np.array(clip[:3]), np.array(clip[3:])
    pts[:, :3] = np.clip(pts[:, :3], lo + 1e-3, hi - 1e-3)
    pts = pts[:n]
    if pts.shape[0] < n:  # pad by resampling (tiny n)
        extra = pts[rng.integers(0, pts.shape[0], n - pts.shape[0])]
        pts = np.concatenate([pts, extra], 0)
    pts = pts.astype(np.float32)
    return np.concatenate([pts, pts[:, :3]], 1).astype(np.float32)


def uniform_points(n: int, seed: int = 0, clip=NUSC_CLIP) -> np.ndarray:
    """Worst case for voxel ranking: M ≈ N distinct voxels."""
    rng = np.random.default_rng(seed)
    lo, hi = np.array(clip[:3]), np.array(clip[3:])
    xyz = rng.uniform(lo, hi, (n, 3))
    pts = np.concatenate([xyz, rng.uniform(0, 1, (n, 1)), np.zeros((n, 1))], 1).astype(np.float32)
    return np.concatenate([pts, pts[:, :3]], 1).astype(np.float32)


def lidar2img(cams: int = 6, H: int = 900, W: int = 1600) -> np.ndarray:
    """[cams,4,4] f32 pinhole cameras around the ego vehicle (nuScenes-like intrinsics)."""
    if cams == 6:
        yaws = np.deg2rad([0.0, -55.0, 55.0, 180.0, 110.0, -110.0])
    else:
        yaws = np.linspace(0, 2 * np.pi, cams, endpoint=False)
    fx = 1266.0 * W / 1600.0
    fy = 1266.0 * H / 900.0
    K = np.array([[fx, 0, 816.0 * W / 1600.0, 0], [0, fy, 491.0 * H / 900.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    out = np.zeros((cams, 4, 4))
    for i, yaw in enumerate(yaws):
        fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
        right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd], 0)
        c = np.array([0.5 * np.cos(yaw), 0.5 * np.sin(yaw), -0.3])  # camera centre in the lidar frame
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = -R @ c
        out[i] = K @ T
    return out.astype(np.float32)


def mask_planes(cams: int = 6, classes: int = 10, H: int = 900, W: int = 1600, seed: int = 0,
                n_obj: int = 250, overlap: bool = False, dtype=np.uint8):
    """[cams,classes,H,W] id planes with ≤ n_obj axis-aligned ellipses, ids increasing over (cam, class)."""
    rng = np.random.default_rng(seed + 1000)
    mask = np.zeros((cams, classes, H, W), dtype=dtype)
    planes = cams * classes
    per_plane = np.full(planes, n_obj // planes)
    per_plane[: n_obj - per_plane.sum()] += 1
    obj = 0
    yy, xx = None, None
    for p in range(planes):
        cam, cls = divmod(p, classes)
        for _ in range(int(per_plane[p])):
            obj += 1
            scale = 3.0 if overlap else 1.0
            a = rng.uniform(0.02, 0.08) * W * scale
            b = rng.uniform(0.03, 0.10) * H * scale
            cx = rng.uniform(0.05, 0.95) * W
            cy = rng.uniform(0.40, 0.85) * H
            x0, x1 = int(max(0, cx - a)), int(min(W, cx + a + 1))
            y0, y1 = int(max(0, cy - b)), int(min(H, cy + b + 1))
            yy, xx = np.mgrid[y0:y1, x0:x1]
            inside = ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0
            sub = mask[cam, cls, y0:y1, x0:x1]
            sub[inside] = obj
    return mask


def mask_anno(mask: np.ndarray, seed: int = 0, n_obj: int = 250, categories=None) -> np.ndarray:
    """[n_obj,9] f32 rows (x1,y1,x2,y2,score,category,cam_id,obj_id,valid) sorted by obj_id.  categories: number of classes to
    draw the category from when the planes are not per class (AV2: one id plane per camera, 26 classes)."""
    rng = np.random.default_rng(seed + 2000)
    cams, classes, H, W = mask.shape
    anno = np.zeros((n_obj, 9), dtype=np.float32)
    for cam in range(cams):
        for cls in range(classes):
            ids = np.unique(mask[cam, cls])
            for oid in ids[ids > 0]:
                ys, xs = np.nonzero(mask[cam, cls] == oid)
                cat = cls if categories is None else int(rng.integers(0, categories))
                anno[int(oid) - 1] = (xs.min(), ys.min(), xs.max(), ys.max(), rng.uniform(0.3, 1.0), cat, cam, oid, 1)
    return anno


def cluster_points(m: int, seed: int = 0, batches: int = 1, n_clusters: int = 40):
    """Voted-centre-like blobs for CCL: ([m,3] f32, [m] i32 batch idx)."""
    rng = np.random.default_rng(seed + 3000)
    centres = rng.uniform(-40, 40, (n_clusters, 2))
    which = rng.integers(0, n_clusters, m)
    xy = centres[which] + rng.normal(0, 0.35, (m, 2))
    noise = rng.random(m) < 0.1
    xy[noise] = rng.uniform(-45, 45, (int(noise.sum()), 2))
    z = rng.uniform(-2, 1, (m, 1))
    b = np.sort(rng.integers(0, batches, m)).astype(np.int32)
    if batches > 1:  # interleave so batch members are not contiguous
        perm = rng.permutation(m)
        b = b[perm]
    return np.concatenate([xy, z], 1).astype(np.float32), b


def av2_points(n: int, seed: int = 0) -> np.ndarray:
    """[n,7] f32 Argoverse-2-shaped sweep: x,y,z,intensity + un-augmented xyz (4-d points, FSF_AV2_config.py:70), spread over the
    +-204.8 m / +-3.2 m range of that config (the spinning-LiDAR pattern of ring_points stretched to a 200 m horizon)."""
    p = ring_points(n, sweeps=max(1, round(n / 50000)), seed=seed, clip=(-200.0, -200.0, -3.19, 200.0, 200.0, 3.19), beams=64,
                    n_boxes=96, max_range=190.0)
    out = np.concatenate([p[:, :4], p[:, :3]], axis=1)
    return np.ascontiguousarray(out, dtype=np.float32)
