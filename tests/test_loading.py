"""Input side: fullysparsefusion_b200.loading against the reference's own LoadMaskFromFiles.

tests/golden/mask_samples/ holds sample directories in the on-disk format of tools/mask_tools/save_mask_nusc.py;
tests/golden/mask_loading.npz holds what the reference's loader (datasets/pipelines/loading.py) returned for them
(tools/make_golden.py loading).  Planes and lidar2img are bit-exact; the annotation table is exact (same float32 rounding)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import loading as L
from fullysparsefusion_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.join(GOLD, "mask_samples")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "mask_loading.npz"))


@pytest.mark.parametrize("workers", [1, 4])
def test_nusc_sample_matches_reference(gold, workers):
    loader = L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"), workers=workers)
    for tok in ("tok0", "empty"):
        res = loader(dict(sample_idx=tok))
        assert res["mask_data"].dtype == torch.uint8 and res["mask_anno"].dtype == torch.float32
        assert np.array_equal(res["mask_data"].numpy(), gold[f"nusc_{tok}_mask"])
        assert np.array_equal(res["mask_anno"].numpy(), gold[f"nusc_{tok}_anno"])
    assert gold["nusc_tok0_anno"][:, 8].sum() > 50 and gold["nusc_empty_anno"].sum() == 0


def test_nusc_decodes_into_caller_buffer(gold):
    buf = np.full((6, 10, 90, 160), 7, np.uint8)
    res = L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"))(dict(sample_idx="tok0"), out=buf)
    assert res["mask_data"].data_ptr() == torch.from_numpy(buf).data_ptr()
    assert np.array_equal(buf, gold["nusc_tok0_mask"])
    with pytest.raises(ValueError):
        L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"))(dict(sample_idx="tok0"), out=np.zeros((6, 10, 90, 161), np.uint8))
    with pytest.raises(FileNotFoundError):
        L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"))(dict(sample_idx="missing"))


def _sha(t):
    return np.frombuffer(hashlib.sha256(t.contiguous().numpy().tobytes()).digest(), dtype=np.uint8)


def test_argo_sample_matches_reference(gold):
    l2i = [m.copy() for m in gold["argo_l2i_in"]]
    res = L.LoadMaskFromFiles(os.path.join(ROOT, "argo"), is_argo=True)(dict(img_info=dict(uuid="uuid0"), lidar2img=l2i))
    m = res["mask_data"]
    assert m.dtype == torch.int32 and tuple(m.shape) == tuple(gold["argo_shape"])
    assert np.array_equal(m[0, 0, ::97].numpy(), gold["argo_rows"]) and np.array_equal(m[0, 0, :, ::89].numpy(), gold["argo_cols"])
    assert np.array_equal(_sha(m), gold["argo_sha"])                              # the whole [7,1,1550,2048] stack, resized camera included
    assert np.array_equal(np.stack(res["lidar2img"]), gold["argo_l2i_out"])
    assert not np.array_equal(gold["argo_l2i_out"][0], gold["argo_l2i_in"][0])     # the front camera's rows were rescaled
    assert np.array_equal(res["mask_anno"].numpy(), gold["argo_anno"])


def test_waymo_sample_matches_reference(gold):
    l2i = [m.copy() for m in gold["waymo_l2i_in"]]
    res = L.LoadMaskFromFiles(os.path.join(ROOT, "waymo"), is_waymo=True)(
        dict(pts_filename="data/waymo/training/velodyne/0001234.bin", lidar2img=l2i))
    m = res["mask_data"]
    assert m.dtype == torch.uint8 and tuple(m.shape) == tuple(gold["waymo_shape"])
    assert np.array_equal(m[3:, :, ::61].numpy(), gold["waymo_rows"])
    assert np.array_equal(_sha(m), gold["waymo_sha"])
    assert np.array_equal(np.stack(res["lidar2img"]), gold["waymo_l2i_out"])
    assert np.array_equal(res["mask_anno"].numpy(), gold["waymo_anno"])


@pytest.mark.parametrize("n_in,n_out", [(886, 1280), (2048, 1550), (1550, 2048), (7, 7), (5, 10), (10, 3), (3, 1000)])
def test_nearest_index_is_atens(n_in, n_out):
    src = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1)
    want = torch.nn.functional.interpolate(src, size=(n_out, 1), mode="nearest").view(-1).long().numpy()
    assert np.array_equal(L.nearest_index(n_in, n_out), want)


def test_writer_reader_round_trip(tmp_path):
    mask = synth.mask_planes(6, 10, 60, 100, seed=5, n_obj=90)
    anno = synth.mask_anno(mask, seed=5)
    L.write_mask_sample(str(tmp_path / "s"), mask, anno)
    res = L.LoadMaskFromFiles(str(tmp_path))(dict(sample_idx="s"))
    assert np.array_equal(res["mask_data"].numpy(), mask)
    assert np.array_equal(res["mask_anno"].numpy(), anno)
    with pytest.raises(ValueError):
        L.LoadMaskFromFiles(str(tmp_path), obj_max_num=10)(dict(sample_idx="s"))


def test_frame_stager_cpu_rotation():
    """Slot rotation and contents on the CPU device (the CUDA path differs only in pinned buffers, the copy stream and events)."""
    st = L.FrameStager("cpu", slots=2)
    frames = []
    for i in range(5):
        pts = synth.ring_points(200 + 10 * i, seed=i)
        mask = synth.mask_planes(6, 10, 20, 30, seed=i, n_obj=20)
        anno = synth.mask_anno(mask, seed=i)
        l2i = [m for m in synth.lidar2img(6, 20, 30).astype(np.float64)]
        frames.append((pts, mask, anno, l2i))
    st.put(*frames[0])
    for i in range(1, 5):
        st.put(*frames[i])                      # upload i while i-1 is in use
        with pytest.raises(RuntimeError):
            st.put(*frames[i])                  # both slots hold unfetched frames
        f = st.get()
        assert np.array_equal(f["points"].numpy(), frames[i - 1][0]) and np.array_equal(f["mask"].numpy(), frames[i - 1][1])
        assert np.array_equal(f["anno"].numpy(), frames[i - 1][2])
        assert f["lidar2img"].dtype == torch.float32 and tuple(f["lidar2img"].shape) == (6, 4, 4)
        assert np.array_equal(f["lidar2img"].numpy(), np.stack(frames[i - 1][3]).astype(np.float32))
        st.release(f)
    f = st.get()
    assert np.array_equal(f["points"].numpy(), frames[4][0])
    with pytest.raises(RuntimeError):
        st.get()
    buf = st.host_buffer("mask", (6, 10, 20, 30), torch.uint8)
    buf.copy_(torch.from_numpy(frames[2][1]))
    st.put(frames[2][0], buf, frames[2][2], frames[2][3])     # decode-in-place path: no staging copy
    assert np.array_equal(st.get()["mask"].numpy(), frames[2][1])


@pytest.mark.gpu
def test_disk_to_ids_on_device(cuda, gold):
    """sample directory → pinned slot → device → projection kernel: ids equal the oracle's on the reference-loaded planes."""
    from fullysparsefusion_b200 import ops
    from oracle import fsf_oracle as O

    dev = cuda
    st = L.FrameStager(dev, slots=2)
    loader = L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"))
    l2i = synth.lidar2img(6, 90, 160)
    want_mask = gold["nusc_tok0_mask"]
    for i in range(4):                                   # more frames than slots: slot reuse after release()
        pts = synth.ring_points(3000 + 100 * i, seed=70 + i)
        buf = st.host_buffer("mask", (6, 10, 90, 160), torch.uint8)
        res = loader(dict(sample_idx="tok0"), out=buf.numpy())
        st.put(pts, buf, res["mask_anno"], [m for m in l2i])
        f = st.get()
        assert f["mask"].is_cuda and f["mask"].dtype == torch.uint8 and f["lidar2img"].dtype == torch.float32
        ids = ops.project_sample(f["points"][:, 5:8].contiguous(), f["lidar2img"], f["mask"])
        st.release(f)
        assert np.array_equal(ids.cpu().numpy(), O.points_in_mask(pts[:, 5:8], want_mask, l2i))
        assert np.array_equal(f["anno"].cpu().numpy(), gold["nusc_tok0_anno"])
    assert st.h2d_bytes == 4 * (6 * 10 * 90 * 160 + 250 * 9 * 4 + 6 * 16 * 4) + sum((3000 + 100 * i) * 8 * 4 for i in range(4))


def test_save_no_aug_points():
    import types

    pts = torch.arange(20, dtype=torch.float32).view(4, 5)
    res = L.SaveNoAugPoints()(dict(points=pts.clone()))
    assert torch.equal(res["points"], torch.cat([pts, pts[:, :3]], 1))
    holder = types.SimpleNamespace(tensor=pts.clone())                       # mmdet3d's LiDARPoints carries `.tensor`
    res = L.SaveNoAugPoints()(dict(points=holder, gt_bboxes_3d=torch.ones(2, 9), gt_labels_3d=np.array([1, 3])))
    assert torch.equal(holder.tensor, torch.cat([pts, pts[:, :3]], 1)) and res["points"] is holder
    assert torch.equal(res["no_aug_gt_bboxes_3d"], torch.ones(2, 9)) and res["no_aug_gt_labels_3d"].tolist() == [1, 3]


def test_hwc16_layout_is_the_planes_interleaved(gold):
    res = L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"), layout="hwc16", workers=3)(dict(sample_idx="tok0"))
    m = res["mask_data"].numpy()
    assert m.shape == (6, 90, 160, 16) and m.dtype == np.uint8
    assert np.array_equal(m[..., :10].transpose(0, 3, 1, 2), gold["nusc_tok0_mask"]) and not m[..., 10:].any()
    buf = np.full((6, 90, 160, 16), 9, np.uint8)                       # a reused destination: pad bytes are cleared
    L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"), layout="hwc16")(dict(sample_idx="tok0"), out=buf)
    assert np.array_equal(buf, m)
    with pytest.raises(ValueError):
        L.LoadMaskFromFiles(os.path.join(ROOT, "argo"), is_argo=True, layout="hwc16")(
            dict(img_info=dict(uuid="uuid0"), lidar2img=[x.copy() for x in gold["argo_l2i_in"]]))
    with pytest.raises(ValueError):
        L.LoadMaskFromFiles(ROOT, layout="nhwc")


@pytest.mark.gpu
def test_hwc16_projection_matches_planar(cuda, gold):
    """fsfb_project_sample_select_hwc (written without GPU access; gated until brought up) against the validated planar kernel."""
    from fullysparsefusion_b200 import ops

    dev = cuda
    planar = torch.from_numpy(gold["nusc_tok0_mask"]).to(dev)
    hwc = L.LoadMaskFromFiles(os.path.join(ROOT, "nusc"), layout="hwc16")(dict(sample_idx="tok0"))["mask_data"].to(dev)
    anno = torch.from_numpy(gold["nusc_tok0_anno"]).to(dev)
    l2i = torch.from_numpy(synth.lidar2img(6, 90, 160)).to(dev)
    for n in (1, 777, 20000):
        xyz = torch.from_numpy(synth.ring_points(n, seed=n)[:, :3].copy()).to(dev)
        want = ops.project_sample_select(xyz, l2i, planar, want_overlap=True, anno=anno)
        got = ops.project_sample_select_hwc(xyz, l2i, hwc, 10, want_overlap=True, anno=anno)
        for w, g in zip(want, got):
            assert torch.equal(w, g)


def test_split_heuristics_stay_inside_what_the_kernels_accept():
    """Host-side split selection (ops._pick_splits / ops.linear_k_splits): the C entry points reject splits outside
    1..koff (convolutions) and 1..K chunks (Linear), and a K range must not be empty."""
    from fullysparsefusion_b200 import ops

    for rows in (1, 127, 248, 1024, 1400, 3782, 11000, 16384, 16385, 75000, 300000):
        for cin in (3, 16, 32, 128, 256, 512, 768, 896, 1000, 1024):
            for cout in (2, 11, 33, 128, 256, 384, 1024, 1025):
                s = ops.linear_k_splits(rows, cin, cout)
                kc = (cin + 31) // 32
                assert 1 <= s <= max(1, kc // 4) and s <= 8
                if s > 1:   # only shapes the row-tile kernel serves with a split epilogue
                    assert cout <= 1024 and rows <= 16384 and (rows >= 1024 or cout <= 256)
        for koff in (1, 8, 27):
            for kc in (1, 2, 4, 8, 16, 32):
                for cpad in (128, 256, 512, 1024, 2048):
                    s = ops._pick_splits(rows, cpad, koff, kc)
                    assert 1 <= s <= max(1, koff // 3)
    assert ops.linear_k_splits(3782, 1024, 128) == 8 and ops.linear_k_splits(300000, 128, 128) == 1
