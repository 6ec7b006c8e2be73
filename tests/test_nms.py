"""Multi-class rotated BEV NMS: oracle sanity on CPU, fsfb_nms_* parity on the GPU."""
import numpy as np
import pytest

from oracle import fsf_oracle as O


def _scene(k, c, seed):
    rng = np.random.default_rng(seed)
    centers = rng.uniform(-30, 30, (max(3, k // 6), 2))
    xy = centers[rng.integers(0, len(centers), k)] + rng.normal(0, 1.2, (k, 2))      # clumps: heavy overlap inside a clump
    dims = rng.uniform([1.5, 3.5, 1.4], [2.2, 5.0, 2.0], (k, 3))
    boxes = np.concatenate([xy, rng.uniform(-2, 0, (k, 1)), dims, rng.uniform(-np.pi, np.pi, (k, 1)), rng.normal(0, 1, (k, 2))], 1)
    logits = rng.normal(-1.0, 2.0, (k, c))
    return boxes.astype(np.float32), logits.astype(np.float32)


def test_rotated_iou_oracle():
    a = np.array([0, 0, 0, 4, 2, 1, 0.0])
    assert abs(O.rotated_iou_bev(a, np.array([1, 0, 0, 4, 2, 1, 0.0])) - 0.6) < 1e-12
    assert abs(O.rotated_iou_bev(a, np.array([0, 0, 0, 2, 4, 1, np.pi / 2])) - 1.0) < 1e-12
    assert O.rotated_iou_bev(a, np.array([10, 0, 0, 4, 2, 1, 0.3])) == 0.0
    b = np.array([0.3, -0.2, 0, 3, 1.5, 1, 0.7])
    assert abs(O.rotated_iou_bev(a, b) - O.rotated_iou_bev(b, a)) < 1e-12   # symmetric
    boxes, logits = _scene(60, 2, 0)
    out = O.multiclass_nms(boxes, logits, 0.1, 0.35, 500)
    assert len(out[0]) == len(out[1]) == len(out[2]) == len(out[3]) > 0
    for c in range(2):   # survivors of a class do not overlap each other beyond the threshold, in descending score
        rows = out[3][out[2] == c]
        assert (np.diff(out[1][out[2] == c]) <= 0).all()
        for i in range(len(rows)):
            for j in range(i):
                assert O.rotated_iou_bev(boxes[rows[j]], boxes[rows[i]]) <= 0.35


def test_rotated_iou_rotation_direction():
    """mmdet3d 0.x iou3d turns the (dx along x, dy along y) rectangle CLOCKWISE by yaw.  Two 4 x 1 boxes at yaw = pi/4 have their long
    axis along (1, -1); shifted by 0.5 along (1, 1) — across the long axis — they overlap in a 4 x 0.5 strip: IoU = 2 / 6.  The
    counter-clockwise reading would slide them along their length instead (IoU 3.5 / 4.5)."""
    d = 0.5 / np.sqrt(2)
    a = np.array([0, 0, 0, 4, 1, 1, np.pi / 4])
    b = np.array([d, d, 0, 4, 1, 1, np.pi / 4])
    assert abs(O.rotated_iou_bev(a, b) - 1 / 3) < 1e-9
    from oracle import fsf_torch_cpu as P
    assert abs(P.rotated_iou_pairs(a[None, [0, 1, 3, 4, 6]], b[None, [0, 1, 3, 4, 6]])[0] - 1 / 3) < 1e-9


def test_pool_box_and_nms_rectangle_share_one_footprint():
    """The refine stage hands the SAME decoded boxes to the point pooling and to the NMS: a point the pooling puts inside a box
    (no margin) must lie inside the rectangle the NMS intersects, for an elongated box at a generic yaw."""
    rng = np.random.default_rng(5)
    box = np.array([1.0, -2.0, 0.0, 1.2, 6.0, 2.0, 0.7], np.float32)      # w (x size) 1.2, l (y size) 6.0
    pts = (rng.uniform(-5, 5, (4000, 3)) + box[:3]).astype(np.float32)
    pi, ri, f = O.dynamic_point_pool(box[None], pts, [0.0, 0.0, 0.0], 4096, 100000)
    assert 50 < pi.size < 4000
    corners = O._rect_corners(box)
    inside = np.ones(len(pts), bool)
    for i in range(4):      # counter-clockwise polygon: inside = left of every edge
        p0, p1 = corners[i], corners[(i + 1) % 4]
        inside &= (p1[0] - p0[0]) * (pts[:, 1] - p0[1]) - (p1[1] - p0[1]) * (pts[:, 0] - p0[0]) >= -1e-4
    assert inside[pi].all()
    near = np.abs(pts[:, 2] - box[2]) <= 0.999      # pooled set == rectangle set wherever the z test passes
    assert set(np.flatnonzero(inside & near & (np.abs(pts[:, 2] - box[2]) < 1.0))) >= set(pi.tolist()) - set(np.flatnonzero(~near).tolist())


def test_rotated_iou_three_independent_ways():
    """The NMS oracle is 'parity unpinned' (mmdet3d's iou3d is un-vendored), so its IoU is cross-checked against two independent
    computations: the vectorised fixed-buffer clipper of the CPU port (different code, same published algorithm) and a
    rasterised area estimate (different algorithm altogether), plus a closed form."""
    from oracle import fsf_torch_cpu as P

    # unit square against its 45-degree rotation: the intersection is a regular octagon of area 2 (sqrt 2 - 1)
    sq, rot = np.array([0, 0, 0, 1, 1, 1, 0.0]), np.array([0, 0, 0, 1, 1, 1, np.pi / 4])
    inter = 2 * (np.sqrt(2) - 1)
    assert abs(O.rotated_iou_bev(sq, rot) - inter / (2 - inter)) < 1e-12
    boxes, _ = _scene(90, 1, 21)
    ii, jj = np.triu_indices(90, 1)
    want = np.array([O.rotated_iou_bev(boxes[i], boxes[j]) for i, j in zip(ii, jj)])
    bev = boxes[:, [0, 1, 3, 4, 6]].astype(np.float64)
    got = P.rotated_iou_pairs(bev[ii], bev[jj])
    assert (want > 0).sum() > 200
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)

    def inside(b, x, y):
        c, s = np.cos(b[6]), np.sin(b[6])      # box frame = axes turned clockwise by yaw (mmdet3d 0.x): local = R(+yaw) shift
        lx, ly = (x - b[0]) * c - (y - b[1]) * s, (x - b[0]) * s + (y - b[1]) * c
        return (np.abs(lx) <= b[3] / 2) & (np.abs(ly) <= b[4] / 2)

    pairs = np.flatnonzero(want > 0.05)[:40]
    for p in pairs:
        a, b = boxes[ii[p]].astype(np.float64), boxes[jj[p]].astype(np.float64)
        r = 0.5 * max(np.hypot(a[3], a[4]), np.hypot(b[3], b[4])) + np.hypot(a[0] - b[0], a[1] - b[1])
        g = np.linspace(-r, r, 700)
        x, y = np.meshgrid(a[0] + g, a[1] + g)
        ia, ib = inside(a, x, y), inside(b, x, y)
        est = (ia & ib).sum() / max((ia | ib).sum(), 1)
        assert abs(est - want[p]) < 0.01, (est, want[p])


@pytest.mark.gpu
@pytest.mark.parametrize("k,c,score_thr,max_num,seed", [(150, 3, 0.1, 500, 1), (200, 10, 0.01, 500, 2), (120, 2, 0.05, 20, 3),
                                                        (50, 4, 0.999, 500, 4), (1, 1, 0.0, 5, 5), (300, 1, 0.2, 500, 6)])
def test_multiclass_nms_parity(cuda, k, c, score_thr, max_num, seed):
    import torch
    from fullysparsefusion_b200 import ops
    for attempt in range(12):   # a scene in which no deciding IoU sits within 2e-4 of the threshold (fp32 vs fp64)
        boxes, logits = _scene(k, c, seed + 100 * attempt)
        wb, ws, wl, wr, close = O.multiclass_nms(boxes, logits, score_thr, 0.35, max_num, margin=2e-4)
        if not close:
            break
    assert not close
    gb, gs, gl, gr = ops.multiclass_nms(torch.from_numpy(boxes).to(cuda), torch.from_numpy(logits).to(cuda), score_thr, 0.35, max_num)
    assert np.array_equal(gr.cpu().numpy(), wr) and np.array_equal(gl.cpu().numpy(), wl)     # indices and labels bit-exact
    np.testing.assert_allclose(gs.cpu().numpy(), ws, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(gb.cpu().numpy(), wb)
